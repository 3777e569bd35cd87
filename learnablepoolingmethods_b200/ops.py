"""Thin Python wrappers over the C-ABI kernels (device pointers in, nothing allocated by the library).

Every function here launches hand-written sm_100a kernels from liblpm_b200.so on the current
torch CUDA stream.  torch is used for memory ownership only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmDesc, check, ptr, stream_ptr


def _lda(t: torch.Tensor) -> int:
    assert t.stride(-1) == 1, "innermost dimension must be contiguous"
    return t.stride(-2)


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = True,
         out: Optional[torch.Tensor] = None, out_dtype=torch.float16, bias=None, row_scale=None,
         relu: bool = False, alpha: float = 1.0, accumulate: bool = False, splits: int = 1,
         stats: bool = False, force_bn: int = 0, M=None, N=None, K=None):
    """D = epilogue(alpha * A x B).

    a: [.., M, K] (a_mn=False) or [.., K, M] (a_mn=True), fp16, last dim contiguous.
    b: [.., N, K] (b_mn=False) or [.., K, N] (b_mn=True, i.e. a row-major weight [in, out]).
    A leading batch dimension on a and/or b makes the product batched (a 2-D operand is shared).
    Returns out (and (sum, sumsq) row-statistic partials [batch, n_tiles, M] when stats=True).
    With splits > 1 returns the fp32 partials [splits, batch?, M, N].
    """
    lib = _lib.load()
    assert a.dtype == torch.float16 and b.dtype == torch.float16
    batch = 1
    if a.dim() == 3:
        batch = a.shape[0]
    if b.dim() == 3:
        batch = max(batch, b.shape[0])
    am, ak = (a.shape[-1], a.shape[-2]) if a_mn else (a.shape[-2], a.shape[-1])
    bn, bk = (b.shape[-1], b.shape[-2]) if b_mn else (b.shape[-2], b.shape[-1])
    M = M or am
    N = N or bn
    K = K or min(ak, bk)
    d = GemmDesc()
    d.A, d.a_mn, d.lda = ptr(a), int(a_mn), _lda(a)
    d.a_batch_stride = a.stride(0) if a.dim() == 3 else 0
    d.B, d.b_mn, d.ldb = ptr(b), int(b_mn), _lda(b)
    d.b_batch_stride = b.stride(0) if b.dim() == 3 else 0
    d.M, d.N, d.K, d.batch, d.force_bn = M, N, K, batch, force_bn
    eff_splits = lib.lpm_gemm_splits(K, splits)
    d.splits = eff_splits
    shape = (batch, M, N) if (a.dim() == 3 or b.dim() == 3) else (M, N)
    if eff_splits > 1:
        assert out is None and bias is None and not relu and not stats
        out = torch.empty((eff_splits,) + shape, dtype=torch.float32, device=a.device)
        d.out_split_stride = out.stride(0)
        view = out[0]
    else:
        if out is None:
            out = torch.empty(shape, dtype=out_dtype, device=a.device)
        view = out
    d.out = ptr(out)
    d.out_f32 = int(out.dtype == torch.float32)
    d.ldc = _lda(view)
    d.out_batch_stride = view.stride(0) if view.dim() == 3 else 0
    d.bias = ptr(bias)
    d.row_scale = ptr(row_scale)
    d.row_scale_batch_stride = M if row_scale is not None else 0
    d.relu, d.accumulate, d.alpha = int(relu), int(accumulate), float(alpha)
    st = None
    if stats:
        n_tiles = -(-N // (force_bn or lib.lpm_gemm_tile_n(N)))
        st = torch.empty((2, batch, n_tiles, M), dtype=torch.float32, device=a.device)
        d.stat_sum, d.stat_sq = ptr(st[0]), ptr(st[1])
    check(lib.lpm_gemm_f16(C.byref(d), stream_ptr()), "lpm_gemm_f16")
    if stats:
        return out, st
    return out
