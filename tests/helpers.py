"""Shared helpers for the GPU parity tests."""
import torch


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def oracle_params(store, dtype=torch.float32):
    """Copy the product's variables into the oracle's (params, state) dictionaries (same TF names)."""
    from oracle import netvlad_oracle as O
    P, S = {}, {}
    for k, v in store.vars.items():
        t = v.detach().cpu().to(dtype).clone()
        (S if k.endswith(O.TRAINABLE_EXCLUDE) else P)[k] = t
    return P, S


def perturb(store, seed=0):
    """Move BN/LN affine, biases and moving statistics off their 0/1 initial values."""
    g = torch.Generator().manual_seed(seed)
    for k, v in store.vars.items():
        if k.endswith(("beta", "bias", "biases", "moving_mean")):
            v.copy_((torch.randn(v.shape, generator=g) * 0.1).to(v.device))
        elif k.endswith(("gamma", "moving_variance")):
            v.copy_((1 + 0.2 * torch.rand(v.shape, generator=g)).to(v.device))
    store.mark_dirty()
