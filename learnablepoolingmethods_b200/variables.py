"""Variable store: the eager stand-in for TF-1.x variable scopes on the hot path.

Parameters are fp32 torch tensors keyed by the reference's TF variable names, so a
TF checkpoint -> state-dict conversion is a pure rename.  The fp16 operand "shadows" the kernels
consume (concatenated / transposed / padded layouts) are derived from them by the C-ABI cast kernels
and are rebuilt when the store is marked dirty (after an optimiser step or load_state_dict).
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, Optional

import torch

NON_TRAINABLE_SUFFIXES = ("moving_mean", "moving_variance")


class VariableStore:
    def __init__(self, device=None, seed: int = 1810):
        self.device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        self.vars: Dict[str, torch.Tensor] = {}
        self._gen = torch.Generator().manual_seed(seed)   # CPU generator: identical on every rank
        self._scope = []
        self.version = 0          # bumped whenever parameter values change
        self.layout_version = 0   # bumped whenever a variable or shadow buffer is (re)allocated: captured CUDA graphs hold
                                  # raw addresses and are re-captured only then (values are refreshed in place)
        self.shadows: Dict[str, torch.Tensor] = {}
        self.shadow_version = -1

    # -- scopes ------------------------------------------------------------------------------
    @contextlib.contextmanager
    def variable_scope(self, name: str):
        self._scope.append(name)
        try:
            yield
        finally:
            self._scope.pop()

    def _full(self, name: str) -> str:
        return "/".join(self._scope + [name])

    # -- creation ----------------------------------------------------------------------------
    def get_variable(self, name: str, shape, init: str, arg: Optional[float] = None) -> torch.Tensor:
        """init: 'normal' (stddev=arg) | 'glorot' (tf.layers.dense / slim.fully_connected default)
        | 'zeros' | 'ones'."""
        full = self._full(name)
        if full in self.vars:
            v = self.vars[full]
            if tuple(v.shape) != tuple(shape):
                raise ValueError(f"variable {full} exists with shape {tuple(v.shape)}, requested {tuple(shape)}")
            return v
        if init == "normal":
            t = torch.randn(shape, generator=self._gen, dtype=torch.float32) * arg
        elif init == "glorot":
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=self._gen, dtype=torch.float32) * 2 - 1) * lim
        elif init == "zeros":
            t = torch.zeros(shape, dtype=torch.float32)
        elif init == "ones":
            t = torch.ones(shape, dtype=torch.float32)
        else:
            raise ValueError(init)
        t = t.to(self.device)
        self.vars[full] = t
        self.version += 1
        self.layout_version += 1
        return t

    def batch_norm_vars(self, scope: str, c: int):
        with self.variable_scope(scope):
            return (self.get_variable("beta", (c,), "zeros"), self.get_variable("gamma", (c,), "ones"),
                    self.get_variable("moving_mean", (c,), "zeros"), self.get_variable("moving_variance", (c,), "ones"))

    # -- bookkeeping -------------------------------------------------------------------------
    def trainable(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in self.vars.items() if not k.endswith(NON_TRAINABLE_SUFFIXES)}

    def mark_dirty(self):
        self.version += 1

    def state_dict(self, prefix: str = "") -> Dict[str, torch.Tensor]:
        return {prefix + k: v.detach().clone() for k, v in self.vars.items()}

    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        for k, v in sd.items():
            k = k[len(prefix):] if prefix and k.startswith(prefix) else k
            t = torch.as_tensor(v, dtype=torch.float32)
            if k in self.vars:
                if tuple(self.vars[k].shape) != tuple(t.shape):
                    raise ValueError(f"{k}: shape {tuple(t.shape)} != {tuple(self.vars[k].shape)}")
                self.vars[k].copy_(t)
            else:
                self.vars[k] = t.to(self.device).clone()
                self.layout_version += 1
        self.mark_dirty()

    def num_parameters(self) -> int:
        return sum(v.numel() for v in self.trainable().values())


_default_store: Optional[VariableStore] = None


def default_store() -> VariableStore:
    global _default_store
    if _default_store is None:
        _default_store = VariableStore()
    return _default_store


def reset_default_store(device=None, seed: int = 1810) -> VariableStore:
    global _default_store
    _default_store = VariableStore(device, seed)
    return _default_store
