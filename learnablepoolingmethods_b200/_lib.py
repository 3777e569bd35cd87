"""ctypes binding of liblpm_b200.so (the C-ABI declared in include/lpm_b200.h).

The library is built in-tree by `learnablepoolingmethods_b200.build`.  There is no fallback: if the
shared object is missing or a call fails, a `LpmError` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblpm_b200.so")


class LpmError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("a_mn", C.c_int), ("lda", C.c_longlong), ("a_batch_stride", C.c_longlong),
        ("B", C.c_void_p), ("b_mn", C.c_int), ("ldb", C.c_longlong), ("b_batch_stride", C.c_longlong),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int), ("splits", C.c_int),
        ("force_bn", C.c_int),
        ("out", C.c_void_p), ("out_f32", C.c_int), ("ldc", C.c_longlong),
        ("out_batch_stride", C.c_longlong), ("out_split_stride", C.c_longlong),
        ("bias", C.c_void_p), ("row_scale", C.c_void_p), ("row_scale_batch_stride", C.c_longlong),
        ("relu", C.c_int), ("accumulate", C.c_int), ("alpha", C.c_float),
        ("stat_sum", C.c_void_p), ("stat_sq", C.c_void_p),
        ("mask", C.c_void_p), ("ld_mask", C.c_longlong),
        ("add1", C.c_void_p), ("add2", C.c_void_p), ("ld_add", C.c_longlong),
        ("no_tma_store", C.c_int),
    ]


class GatingTail(C.Structure):
    """lpm_gating_tail (include/lpm_b200.h): the context-gating tail of lpm_gemm_splitk_gated_fwd."""
    _fields_ = [
        ("counters", C.c_void_p),
        ("part2", C.c_void_p), ("splits2", C.c_int), ("split_stride2", C.c_longlong),
        ("bias", C.c_void_p),
        ("act32", C.c_void_p), ("act16", C.c_void_p), ("act_split3", C.c_int),
        ("wg", C.c_void_p), ("ldwg", C.c_longlong),
        ("wg_diag", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("moving_mean", C.c_void_p), ("moving_var", C.c_void_p),
        ("decay", C.c_float), ("eps", C.c_float), ("training", C.c_int),
        ("g_sum", C.c_void_p),
        ("out32", C.c_void_p), ("out16", C.c_void_p), ("out_split3", C.c_int),
        ("save_mean", C.c_void_p), ("save_rstd", C.c_void_p),
    ]


_lib = None


def load() -> C.CDLL:
    """Load liblpm_b200.so (building is a separate, explicit step)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LpmError(
            f"{LIB_PATH} not found: build it with `python -m learnablepoolingmethods_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    lib.lpm_last_error.restype = C.c_char_p
    lib.lpm_version.restype = C.c_int
    lib.lpm_ortho_reg_workspace_bytes.restype = C.c_ulonglong
    _lib = lib
    return lib


launch_count = 0   # kernels launched through the C-ABI by this process (bench.py's gpu_launches)

# kernels per C-ABI call (everything not listed launches exactly one)
_LAUNCHES = {"lpm_layernorm_joint_fwd": 2, "lpm_layernorm_joint_bwd": 2, "lpm_colsum": 2, "lpm_xent_fwd": 2,
             "lpm_adam_clip_step": 3, "lpm_rank_adam_step": 2, "lpm_shard_sqnorm": 2, "lpm_shard_adam": 2, "lpm_ortho_reg": 5}


def check(rc: int, what: str = "") -> None:
    global launch_count
    launch_count += _LAUNCHES.get(what, 1)
    if rc != 0:
        msg = load().lpm_last_error().decode("utf-8", "replace")
        raise LpmError(f"{what} failed (code {rc}): {msg}")


_DEVICE_INDEX = None


def stream_ptr() -> C.c_void_p:
    """cudaStream_t of torch's current stream on this process's device (one process per GPU: the device index is
    read once).  torch.cuda.current_stream() costs ~14 us per call, a third of the host time of a training step;
    the raw getter is ~0.3 us."""
    global _DEVICE_INDEX
    import torch
    if _DEVICE_INDEX is None:
        _DEVICE_INDEX = torch.cuda.current_device()
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(_DEVICE_INDEX))


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())
