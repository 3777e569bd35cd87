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
         stats: bool = False, force_bn: int = 0, M=None, N=None, K=None, mask=None, add1=None, add2=None,
         no_tma_store: bool = False):
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
    # automatic split-K: fp32-output products with a plain epilogue that would occupy only a few of the 148 SMs
    # (weight gradients with a long reduction, skinny dX products) are split over K and reduced deterministically
    want_f32 = (out.dtype == torch.float32) if torch.is_tensor(out) else (out_dtype == torch.float32)
    plain = bias is None and row_scale is None and not relu and not stats and mask is None and add1 is None
    auto = None
    if splits == 1 and want_f32 and plain and not accumulate and not isinstance(out, str):
        bn_t = force_bn or lib.lpm_gemm_tile_n(N)
        tiles = -(-M // 128) * -(-N // bn_t) * batch
        kb = -(-K // 64)
        if tiles <= 64 and kb >= 16:
            auto = max(2, min(148 // tiles, kb // 4))
    if auto is not None:
        parts = gemm(a, b, a_mn=a_mn, b_mn=b_mn, splits=auto, force_bn=force_bn, M=M, N=N, K=K)
        shape_ = (batch, M, N) if (a.dim() == 3 or b.dim() == 3) else (M, N)
        if out is None:
            out = torch.empty(shape_, dtype=torch.float32, device=a.device)
        assert out.is_contiguous()
        splitk_reduce(parts, alpha=alpha, out32=out)
        return out
    eff_splits = lib.lpm_gemm_splits(K, splits)
    d.splits = eff_splits
    shape = (batch, M, N) if (a.dim() == 3 or b.dim() == 3) else (M, N)
    if splits > 1:   # fp32 partials [eff_splits, ...] (eff_splits may be 1 for a short K)
        assert out is None and bias is None and not relu and not stats
        out = torch.empty((eff_splits,) + shape, dtype=torch.float32, device=a.device)
        d.out_split_stride = out.stride(0)
        view = out[0]
    else:
        if out is None:
            out = torch.empty(shape, dtype=out_dtype, device=a.device)
        view = out
    if isinstance(out, str):   # out="none": statistics only
        out = None
        d.out = None
    else:
        d.out = ptr(out)
        d.out_f32 = int(out.dtype == torch.float32)
        d.ldc = _lda(view)
        d.out_batch_stride = view.stride(0) if view.dim() == 3 else 0
    d.bias = ptr(bias)
    d.row_scale = ptr(row_scale)
    d.row_scale_batch_stride = M if row_scale is not None else 0
    d.relu, d.accumulate, d.alpha = int(relu), int(accumulate), float(alpha)
    d.no_tma_store = int(no_tma_store)
    if mask is not None:
        d.mask, d.ld_mask = ptr(mask), _lda(mask)
    if add1 is not None:
        d.add1, d.ld_add = ptr(add1), _lda(add1)
        if add2 is not None:
            assert _lda(add2) == _lda(add1)
            d.add2 = ptr(add2)
    st = None
    if stats:
        n_tiles = -(-N // (force_bn or lib.lpm_gemm_tile_n(N)))
        st = torch.zeros((2, batch, 2 * n_tiles, M), dtype=torch.float32, device=a.device)   # x2: epilogue groups
        d.stat_sum, d.stat_sq = ptr(st[0]), ptr(st[1])
    check(lib.lpm_gemm_f16(C.byref(d), stream_ptr()), "lpm_gemm_f16")
    if stats:
        return out, st
    return out


_gated_counters = {}


def gemm_splitk_gated(a, w, *, splits, bias, wg, gamma, beta, moving_mean, moving_var, training, wg_diag=None, save=False,
                      parts2=None, act_split3=True, out_split3=True, decay=0.999, eps=1e-3):
    """K3 (frame_level_models.py:2314-2368) as ONE launch: hidden = a @ w (+ the partials `parts2` of an earlier split-K
    launch) + bias; gates = BN(hidden @ wg [- diag(wg) * hidden]); out = hidden * sigmoid(gates).  a: fp16 [B <= 128, Kd];
    w: fp16 [Kd, H] (row-major weight); wg: fp32 [H, H].  The reduction of the split-K partials and the whole gating run in
    the tail of the GEMM kernel (lpm_gemm_splitk_gated_fwd), with the gate product in exact fp32.
    Returns (act32, act16 or [hi|lo|hi], out32, out16 or [hi|lo|hi], (mean, rstd) or None, g_sum, partials)."""
    lib = _lib.load()
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and wg.dtype == torch.float32
    B, Kd = a.shape
    H = w.shape[1]
    dev = a.device
    eff = lib.lpm_gemm_splits(Kd, splits)
    parts = torch.empty((eff, B, H), dtype=torch.float32, device=dev)
    d = GemmDesc()
    d.A, d.a_mn, d.lda, d.a_batch_stride = ptr(a), 0, _lda(a), 0
    d.B, d.b_mn, d.ldb, d.b_batch_stride = ptr(w), 1, _lda(w), 0
    d.M, d.N, d.K, d.batch, d.splits, d.force_bn = B, H, Kd, 1, eff, 0
    d.out, d.out_f32, d.ldc, d.out_batch_stride, d.out_split_stride = ptr(parts), 1, H, 0, parts.stride(0)
    d.alpha = 1.0
    # grid-barrier scratch, one buffer per device: the kernel leaves the counters at zero, so consecutive launches can share
    # them; two gated products must not run CONCURRENTLY on one device (the engine runs one head at a time; the buffer is
    # created by the first eager call, i.e. before any graph capture)
    key = dev.index
    if key not in _gated_counters:
        _gated_counters[key] = torch.zeros(4, dtype=torch.int32, device=dev)
    act32, out32, g_sum = _f32((B, H), dev), _f32((B, H), dev), _f32((B, H), dev)
    act16 = _f16((B, 3 * H if act_split3 else H), dev)
    out16 = _f16((B, 3 * H if out_split3 else H), dev)
    sm = _f32((2, H), dev) if save else None
    t = _lib.GatingTail()
    t.counters = ptr(_gated_counters[key])
    if parts2 is not None:
        assert parts2.dtype == torch.float32 and parts2.is_contiguous() and parts2.shape[-2:] == (B, H)
        p2 = parts2.reshape(-1, B, H)
        t.part2, t.splits2, t.split_stride2 = ptr(p2), p2.shape[0], p2.stride(0)
    t.bias = ptr(bias)
    t.act32, t.act16, t.act_split3 = ptr(act32), ptr(act16), int(act_split3)
    t.wg, t.ldwg, t.wg_diag = ptr(wg), wg.stride(0), ptr(wg_diag)
    t.gamma, t.beta, t.moving_mean, t.moving_var = ptr(gamma), ptr(beta), ptr(moving_mean), ptr(moving_var)
    t.decay, t.eps, t.training = float(decay), float(eps), int(training)
    t.g_sum = ptr(g_sum)
    t.out32, t.out16, t.out_split3 = ptr(out32), ptr(out16), int(out_split3)
    if save:
        t.save_mean, t.save_rstd = ptr(sm[0]), ptr(sm[1])
    check(lib.lpm_gemm_splitk_gated_fwd(C.byref(d), C.byref(t), stream_ptr()), "lpm_gemm_splitk_gated_fwd")
    return act32, act16, out32, out16, sm, g_sum, parts


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
BN_EPS = 1e-3      # slim.batch_norm default
BN_DECAY = 0.999   # slim.batch_norm default
LN_EPS = 1e-12     # tf.contrib.layers.layer_norm


def _f32(shape, dev):
    return torch.empty(shape, dtype=torch.float32, device=dev)


def _f16(shape, dev):
    return torch.empty(shape, dtype=torch.float16, device=dev)


def splitk_reduce(parts, *, bias=None, relu=False, alpha=1.0, out32=None, out16=None, accumulate=False, parts2=None,
                  split3=False):
    """Deterministic sum of split-K partials [splits, ...] (+ a second set `parts2`), bias / ReLU, fp32 and fp16 outputs;
    split3: out16 is [rows, 3*cols] = [hi | lo | hi] (split-precision operand, see split_hi_lo)."""
    lib = _lib.load()
    splits = parts.shape[0]
    n = parts[0].numel()
    cols = parts.shape[-1]
    if parts2 is None and not split3:
        check(lib.lpm_splitk_reduce(ptr(parts), splits, C.c_longlong(parts.stride(0)), C.c_longlong(n), cols, ptr(bias),
                                    int(relu), C.c_float(alpha), int(accumulate), ptr(out32), ptr(out16), stream_ptr()),
              "lpm_splitk_reduce")
        return
    assert parts2 is None or parts2[0].numel() == n
    check(lib.lpm_splitk_reduce_ex(ptr(parts), splits, C.c_longlong(parts.stride(0)), ptr(parts2),
                                   0 if parts2 is None else parts2.shape[0],
                                   C.c_longlong(0 if parts2 is None else parts2.stride(0)), C.c_longlong(n), cols, ptr(bias),
                                   int(relu), C.c_float(alpha), int(accumulate), ptr(out32), ptr(out16), int(split3),
                                   stream_ptr()), "lpm_splitk_reduce")


def cast_f16(src: torch.Tensor, dst: Optional[torch.Tensor] = None, cols_dst: Optional[int] = None):
    """fp32 [rows, cols] -> fp16 [rows, cols_dst >= cols] (zero padded)."""
    lib = _lib.load()
    src2 = src.reshape(-1, src.shape[-1])
    rows, cols = src2.shape
    if dst is None:
        cols_dst = cols_dst or cols
        dst = torch.empty((rows, cols_dst), dtype=torch.float16, device=src.device)
    cols_dst = cols_dst or dst.shape[-1]
    check(lib.lpm_cast_f32_to_f16(ptr(src2), C.c_longlong(src2.stride(0)), rows, cols, ptr(dst),
                                  C.c_longlong(dst.stride(0)), cols_dst, stream_ptr()), "lpm_cast_f32_to_f16")
    return dst


def split_hi_lo(src: torch.Tensor, dst: Optional[torch.Tensor] = None, along_rows=False):
    """fp32 [rows, cols] -> split-precision fp16 operand: [rows, 3*cols] = [hi | lo | hi] (activations); along_rows=True:
    the column block `dst` of a [3*rows, .] weight operand = [hi ; hi ; lo]; along_rows=2: dst [rows, cols] = lo only."""
    lib = _lib.load()
    rows, cols = src.shape
    assert src.dtype == torch.float32 and src.stride(1) == 1
    if dst is None:
        assert not along_rows
        dst = torch.empty((rows, 3 * cols), dtype=torch.float16, device=src.device)
    check(lib.lpm_split_hi_lo_f16(ptr(src), _ll(src.stride(0)), rows, cols, ptr(dst), _ll(dst.stride(0)), int(along_rows),
                                  stream_ptr()), "lpm_split_hi_lo_f16")
    return dst


def transpose_f32(src: torch.Tensor):
    lib = _lib.load()
    rows, cols = src.shape
    dst = _f32((cols, rows), src.device)
    check(lib.lpm_transpose_f32(ptr(src), rows, cols, ptr(dst), stream_ptr()), "lpm_transpose_f32")
    return dst


def transpose_f32_dual(src: torch.Tensor, want32=True, out32=None, out16=None):
    """(src^T as fp32 or None, src^T as fp16): cluster-major centre layouts of the pooling kernel / backward.
    out32 / out16: refresh existing buffers in place (their addresses may be baked into a captured CUDA graph)."""
    lib = _lib.load()
    rows, cols = src.shape
    dst = (out32 if out32 is not None else _f32((cols, rows), src.device)) if want32 else None
    dst16 = out16 if out16 is not None else _f16((cols, rows), src.device)
    check(lib.lpm_transpose_f32_dual(ptr(src), rows, cols, ptr(dst), ptr(dst16), stream_ptr()), "lpm_transpose_f32_dual")
    return dst, dst16


def bn_finalize(psum, psq, count, gamma, beta, moving_mean, moving_var, *, training, bessel, save=False,
                decay=BN_DECAY, eps=BN_EPS, psum_stride=None):
    """Reduce (sum, sumsq) partial rows [P, C] (row stride psum_stride) -> folded affine (scale, shift) [C]."""
    lib = _lib.load()
    Cn = gamma.numel()
    dev = gamma.device
    scale, shift = _f32((Cn,), dev), _f32((Cn,), dev)
    sm = _f32((2, Cn), dev) if save else None
    P = 0 if psum is None else psum.shape[0]
    pstride = psum_stride or Cn
    check(lib.lpm_batchnorm_finalize(ptr(psum), ptr(psq), P, C.c_longlong(pstride), Cn, C.c_double(count), ptr(gamma),
                                     ptr(beta), ptr(moving_mean), ptr(moving_var), C.c_float(decay), C.c_float(eps),
                                     int(bessel), int(training), ptr(scale), ptr(shift),
                                     ptr(sm[0]) if save else None, ptr(sm[1]) if save else None, stream_ptr()),
          "lpm_batchnorm_finalize")
    return (scale, shift, sm) if save else (scale, shift)


QUANT_MAX, QUANT_MIN = 2.0, -2.0     # readers.py:185-193 / utils.py:28 defaults


def _sample_call(name, x, args_after_x):
    """fp32 frames (already L2-normalised) or uint8 codes (dequantised + normalised inside the kernel)."""
    lib = _lib.load()
    if x.dtype == torch.uint8:
        fn = getattr(lib, name + "_u8")
        check(fn(ptr(x), C.c_float(QUANT_MAX), C.c_float(QUANT_MIN), *args_after_x), name + "_u8")
    else:
        assert x.dtype == torch.float32
        check(getattr(lib, name)(ptr(x), *args_after_x), name)


def sample_bn_stats(x, num_frames, T):
    lib = _lib.load()
    B, Fmax, F = x.shape
    blocks = lib.lpm_sample_stats_blocks()
    partial = _f32((blocks, 2, F), x.device)
    _sample_call("lpm_sample_bn_stats", x, (ptr(num_frames), B, Fmax, F, T, ptr(partial), stream_ptr()))
    return partial


def sample_bn_apply(x, num_frames, T, scale, shift, out=None, split_col=None):
    """Sampled + batch-normed frames as fp16: one [B*T, F] matrix, or (split_col given) two contiguous
    per-modality matrices ([B*T, split_col], [B*T, F - split_col]).  x: fp32 frames or uint8 codes."""
    B, Fmax, F = x.shape
    if split_col is None:
        if out is None:
            out = _f16((B * T, F), x.device)
        _sample_call("lpm_sample_bn_apply", x, (ptr(num_frames), B, Fmax, F, T, ptr(scale), ptr(shift), ptr(out), 0, None,
                                                stream_ptr()))
        return out
    ya, yb = _f16((B * T, split_col), x.device), _f16((B * T, F - split_col), x.device)
    _sample_call("lpm_sample_bn_apply", x, (ptr(num_frames), B, Fmax, F, T, ptr(scale), ptr(shift), ptr(ya), split_col,
                                            ptr(yb), stream_ptr()))
    return ya, yb


def random_frame_index(num_frames, T, max_frames, *, mode=0, uniform=None, seed=0):
    """int32 [B, T] gather indices of SampleRandomFrames (mode 0) / SampleRandomSequence (mode 1), model_utils.py:26-73.
    uniform: optional fp32 [B, T] / [B] draws in [0,1) (parity tests inject tf.random_uniform's values)."""
    lib = _lib.load()
    B = num_frames.shape[0]
    idx = torch.empty((B, T), dtype=torch.int32, device=num_frames.device)
    if uniform is not None:
        uniform = uniform.to(device=num_frames.device, dtype=torch.float32).contiguous()
        assert uniform.numel() == (B * T if mode == 0 else B)
    check(lib.lpm_random_frame_index(ptr(num_frames), ptr(uniform), C.c_ulonglong(seed & (2 ** 64 - 1)), B, T, max_frames,
                                     mode, ptr(idx), stream_ptr()), "lpm_random_frame_index")
    return idx


def gather_bn_stats(x, frame_index, T):
    """sample_bn_stats with explicit gather indices (int32 [B, T])."""
    lib = _lib.load()
    B, Fmax, F = x.shape
    blocks = lib.lpm_sample_stats_blocks()
    partial = _f32((blocks, 2, F), x.device)
    check(lib.lpm_gather_bn_stats(ptr(x), int(x.dtype == torch.uint8), C.c_float(QUANT_MAX), C.c_float(QUANT_MIN),
                                  ptr(frame_index), B, Fmax, F, T, ptr(partial), stream_ptr()), "lpm_gather_bn_stats")
    return partial


def gather_bn_apply(x, frame_index, T, scale, shift):
    """sample_bn_apply with explicit gather indices: fp16 [B*T, F]."""
    lib = _lib.load()
    B, Fmax, F = x.shape
    out = _f16((B * T, F), x.device)
    check(lib.lpm_gather_bn_apply(ptr(x), int(x.dtype == torch.uint8), C.c_float(QUANT_MAX), C.c_float(QUANT_MIN),
                                  ptr(frame_index), B, Fmax, F, T, ptr(scale), ptr(shift), ptr(out), 0, None, stream_ptr()),
          "lpm_gather_bn_apply")
    return out


def ortho_reg(w, scale, *, dw=None, grad_scale=1.0, accumulate=True, want_value=True):
    """Orthogonal regulariser of module_utils.py:55-90 on w fp32 [D, K]: returns the value (device scalar) and adds
    grad_scale * gradient into dw (fp32 [D, K]) when given."""
    lib = _lib.load()
    D, K = w.shape
    assert w.dtype == torch.float32 and w.is_contiguous()
    nbytes = lib.lpm_ortho_reg_workspace_bytes(D, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=w.device)    # not the shared scratch: modalities run on two streams
    value = _f32((1,), w.device) if want_value else None
    check(lib.lpm_ortho_reg(ptr(w), D, K, C.c_float(scale), C.c_float(grad_scale), int(accumulate), ptr(value), ptr(dw),
                            ptr(ws), C.c_ulonglong(nbytes), stream_ptr()), "lpm_ortho_reg")
    return value


def netvlad_pool_fwd(x16, B, T, wc16, logit_scale, logit_shift, centers, *, valid_frames=None,
                     save_assign=False, assign_in=None):
    """x16: fp16 view [B*T, D] (row stride may exceed D); centers: fp16 [K, D] (cluster-major shadow made by
    transpose_f32_dual) or fp32 [D, K] (the TF layout; transposed + rounded here, one extra launch).
    Returns z [B,K,D] fp16, rscale [B,K], a_sum, assign.
    assign_in (fp16 [B*T, K] contiguous): NetVladV2 mode, the soft-assignment phase is skipped."""
    lib = _lib.load()
    D = x16.shape[1]
    K = wc16.shape[1] if assign_in is None else assign_in.shape[-1]
    dev = x16.device
    if centers.dtype == torch.float32:
        centers = transpose_f32_dual(centers.contiguous(), want32=False)[1]
    assert centers.dtype == torch.float16 and tuple(centers.shape) == (K, D) and centers.is_contiguous()
    z = _f16((B, K, D), dev)
    rscale = _f32((B, K), dev)
    a_sum = _f32((B, K), dev)
    assign = _f16((B, T, K), dev) if save_assign else None
    ldx = x16.stride(0)
    check(lib.lpm_netvlad_pool_fwd(ptr(x16), C.c_longlong(ldx), C.c_longlong(ldx * T), ptr(wc16),
                                   C.c_longlong(wc16.stride(0) if wc16 is not None else 0), ptr(logit_scale),
                                   ptr(logit_shift), ptr(centers), ptr(valid_frames), B, T, D, K, ptr(z), ptr(rscale),
                                   ptr(a_sum), ptr(assign), ptr(assign_in), stream_ptr()), "lpm_netvlad_pool_fwd")
    return z, rscale, a_sum, assign


def netvlad_finalize(z, rscale, d_major=True):
    lib = _lib.load()
    B, K, D = z.shape
    out = _f32((B, D * K) if d_major else (B, K, D), z.device)
    check(lib.lpm_netvlad_finalize(ptr(z), ptr(rscale), B, K, D, int(d_major), ptr(out), stream_ptr()),
          "lpm_netvlad_finalize")
    return out


def mha_core_fwd(qkv, B, L, Dm, H, *, scale, key_scale=None, key_shift=None, want_lse=False):
    lib = _lib.load()
    out = _f16((B * L, Dm), qkv.device)
    lse = _f32((B, H, L), qkv.device) if want_lse else None
    check(lib.lpm_mha_core_fwd(ptr(qkv), C.c_longlong(qkv.stride(0)), B, L, Dm, H, C.c_float(scale), ptr(key_scale),
                               ptr(key_shift), ptr(out), C.c_longlong(out.stride(0)), ptr(lse), stream_ptr()),
          "lpm_mha_core_fwd")
    return (out, lse) if want_lse else out


def scale_rows_f16(x, rs):
    lib = _lib.load()
    y = torch.empty_like(x)
    check(lib.lpm_scale_rows_f16(ptr(x), ptr(rs), C.c_longlong(x.shape[0]), x.shape[1], ptr(y), stream_ptr()),
          "lpm_scale_rows_f16")
    return y


def layernorm_joint_fwd(a, b, b_row_scale, B, rows, D, gamma, beta, *, out=None, out_stride=None, save=False,
                        eps=LN_EPS, u_out=None):
    """u = a + b*row_scale (stored over a, or into u_out); returns y = LN_joint(u) (fp16) [, (mean, rstd) [B,2]]."""
    lib = _lib.load()
    dev = a.device
    if out is None:
        out = _f16((B, rows, D), dev)
        out_stride = rows * D
    sm = _f32((B, 2), dev) if save else None
    import os
    if b is not None and lib.lpm_layernorm_chain_supported(rows, D) and not os.environ.get("LPM_NO_LN_FUSED"):
        # one pass: a thread-block cluster per sample, u kept in shared memory, moments over DSMEM
        u_dst = (u_out if u_out is not None else a) if save or u_out is not None else None
        check(lib.lpm_layernorm_chain_fwd(ptr(a), _ll(rows * D), ptr(b), _ll(rows * D), ptr(b_row_scale), B, rows, D,
                                          C.c_float(eps), ptr(gamma), ptr(beta), ptr(u_dst), _ll(rows * D), ptr(sm),
                                          None, None, None, _ll(0), None, ptr(out), _ll(out_stride), stream_ptr()),
              "lpm_layernorm_chain_fwd")
        return (out, sm) if save else out
    partial = _f32((B, 64), dev)
    check(lib.lpm_layernorm_joint_fwd(ptr(a), ptr(b), ptr(b_row_scale), ptr(u_out), B, rows, D, C.c_longlong(rows * D),
                                      C.c_longlong(rows * D), ptr(gamma), ptr(beta), C.c_float(eps), ptr(out),
                                      C.c_longlong(out_stride), ptr(partial), ptr(sm), stream_ptr()),
          "lpm_layernorm_joint_fwd")
    return (out, sm) if save else out


def layernorm_chain_supported(rows, D) -> bool:
    import os
    if os.environ.get("LPM_NO_LN_CHAIN"):
        return False
    return bool(_lib.load().lpm_layernorm_chain_supported(int(rows), int(D)))


def layernorm_chain_fwd(a, b, B, rows, D, gamma1, beta1, gamma2, beta2, *, out, out_stride, save=False, eps=LN_EPS,
                        out_lo=None):
    """y = LN(LN(a + b; gamma1, beta1) + b; gamma2, beta2) in one pass (transformer_utils.py:712-713 + :410-411).
    save=True also returns (u1, stats1, u2, stats2) for the backward; a is left untouched.
    out_lo: optional fp16 tensor with out's layout that receives fp16(y - fp16(y)) (split-precision operand)."""
    lib = _lib.load()
    dev = a.device
    u1 = _f16((B, rows, D), dev) if save else None
    u2 = _f16((B, rows, D), dev) if save else None
    s1 = _f32((B, 2), dev) if save else None
    s2 = _f32((B, 2), dev) if save else None
    if out_lo is not None:
        assert out_lo.stride(0) == out_stride
        check(lib.lpm_layernorm_chain_fwd_split(ptr(a), _ll(rows * D), ptr(b), _ll(rows * D), None, B, rows, D, C.c_float(eps),
                                                ptr(gamma1), ptr(beta1), ptr(u1), _ll(rows * D), ptr(s1), ptr(gamma2),
                                                ptr(beta2), ptr(u2), _ll(rows * D), ptr(s2), ptr(out), _ll(out_stride),
                                                ptr(out_lo), stream_ptr()), "lpm_layernorm_chain_fwd")
        return (out, u1, s1, u2, s2) if save else out
    check(lib.lpm_layernorm_chain_fwd(ptr(a), _ll(rows * D), ptr(b), _ll(rows * D), None, B, rows, D, C.c_float(eps),
                                      ptr(gamma1), ptr(beta1), ptr(u1), _ll(rows * D), ptr(s1), ptr(gamma2), ptr(beta2),
                                      ptr(u2), _ll(rows * D), ptr(s2), ptr(out), _ll(out_stride), stream_ptr()),
          "lpm_layernorm_chain_fwd")
    return (out, u1, s1, u2, s2) if save else out


def gating_fwd(act, g, gamma, beta, moving_mean, moving_var, *, training, wg_diag=None, save=False,
               decay=BN_DECAY, eps=BN_EPS, split3=False):
    """Context gating.  g: the gate pre-activations [B, H], or their split-K partials [S, B, H] (summed inside; the sum is
    returned as 4th result).  split3: the fp16 output is the split-precision operand [B, 3H] = [hi | lo | hi]."""
    lib = _lib.load()
    B, H = act.shape
    dev = act.device
    out32, out16 = _f32((B, H), dev), _f16((B, 3 * H if split3 else H), dev)
    sm = _f32((2, H), dev) if save else None
    if g.dim() == 2 and not split3:
        check(lib.lpm_gating_fwd(ptr(act), ptr(g), B, H, ptr(wg_diag), ptr(gamma), ptr(beta), ptr(moving_mean),
                                 ptr(moving_var), C.c_float(decay), C.c_float(eps), int(training), ptr(out32), ptr(out16),
                                 ptr(sm[0]) if save else None, ptr(sm[1]) if save else None, stream_ptr()),
              "lpm_gating_fwd")
        return (out32, out16, sm) if save else (out32, out16)
    splits = g.shape[0] if g.dim() == 3 else 1
    g_sum = _f32((B, H), dev) if splits > 1 else g
    check(lib.lpm_gating_fwd_ex(ptr(act), ptr(g), splits, _ll(g.stride(0) if splits > 1 else 0), ptr(g_sum) if splits > 1 else None,
                                B, H, ptr(wg_diag), ptr(gamma), ptr(beta), ptr(moving_mean), ptr(moving_var), C.c_float(decay),
                                C.c_float(eps), int(training), ptr(out32), ptr(out16), int(split3),
                                ptr(sm[0]) if save else None, ptr(sm[1]) if save else None, stream_ptr()),
          "lpm_gating_fwd")
    return (out32, out16, sm, g_sum) if save else (out32, out16, None, g_sum)


def moe_mix_fwd(logits, V, M, expert_off=None):
    lib = _lib.load()
    B = logits.shape[0]
    pred = _f32((B, V), logits.device)
    expert_off = V * (M + 1) if expert_off is None else expert_off
    check(lib.lpm_moe_mix_fwd(ptr(logits), C.c_longlong(logits.stride(0)), B, V, M, expert_off, ptr(pred), stream_ptr()),
          "lpm_moe_mix_fwd")
    return pred


def xent_fwd(pred, labels_u8):
    lib = _lib.load()
    B, V = pred.shape
    row = _f32((B,), pred.device)
    loss = _f32((1,), pred.device)
    check(lib.lpm_xent_fwd(ptr(pred), ptr(labels_u8), B, V, ptr(row), ptr(loss), stream_ptr()), "lpm_xent_fwd")
    return loss, row


# ------------------------------------------------------------------------------------------------
# backward wrappers
# ------------------------------------------------------------------------------------------------
def _ll(v):
    return C.c_longlong(int(v))


def xent_bwd(pred, labels_u8, gscale, upstream=None):
    """dLoss/dpred; `upstream` (optional fp32 device scalar) multiplies gscale on the device."""
    lib = _lib.load()
    dpred = torch.empty_like(pred)
    if upstream is not None:
        assert upstream.dtype == torch.float32 and upstream.numel() == 1 and upstream.is_cuda
        check(lib.lpm_xent_bwd_dev(ptr(pred), ptr(labels_u8), _ll(pred.numel()), C.c_float(gscale), ptr(upstream), ptr(dpred),
                                   stream_ptr()), "lpm_xent_bwd")
        return dpred
    check(lib.lpm_xent_bwd(ptr(pred), ptr(labels_u8), _ll(pred.numel()), C.c_float(gscale), ptr(dpred), stream_ptr()),
          "lpm_xent_bwd")
    return dpred


def moe_mix_bwd(logits, dpred, V, M, expert_off, loss_scale):
    lib = _lib.load()
    B, ncols = logits.shape
    dl = _f16((B, ncols), logits.device)
    check(lib.lpm_moe_mix_bwd(ptr(logits), _ll(logits.stride(0)), B, V, M, expert_off, ptr(dpred), C.c_float(loss_scale),
                              ptr(dl), _ll(dl.stride(0)), ncols, stream_ptr()), "lpm_moe_mix_bwd")
    return dl


def colsum(x, *, alpha=1.0, out=None, accumulate=False, rows=None, cols=None):
    """out[c] (+)= alpha * sum_r x[r, c]; x fp16 or fp32 2-D (row stride may exceed cols)."""
    lib = _lib.load()
    rows = rows or x.shape[0]
    cols = cols or x.shape[1]
    if out is None:
        out = _f32((cols,), x.device)
    partial = _f32((lib.lpm_colsum_chunks(_ll(rows)), cols), x.device)
    check(lib.lpm_colsum(ptr(x), int(x.dtype == torch.float32), _ll(x.stride(0)), _ll(rows), cols, C.c_float(alpha),
                         int(accumulate), ptr(partial), ptr(out), stream_ptr()), "lpm_colsum")
    return out


def colsum_final(partial, chunks, pstride, cols, *, alpha=1.0, out=None, accumulate=False):
    lib = _lib.load()
    if out is None:
        out = _f32((cols,), partial.device)
    check(lib.lpm_colsum_final(ptr(partial), chunks, _ll(pstride), cols, C.c_float(alpha), int(accumulate), ptr(out),
                               stream_ptr()), "lpm_colsum_final")
    return out


def gating_bwd(act, g, gamma, beta, stats, dout, inv_scale, wg_diag=None):
    """Returns dact, dg16, dgamma, dbeta[, ddiag] (ddiag: unscaled extra gradient of diag(gating_weights_2) when the
    forward removed the diagonal, frame_level_models.py:2349-2352)."""
    lib = _lib.load()
    B, H = act.shape
    dev = act.device
    dact, dg = _f32((B, H), dev), _f16((B, H), dev)
    dgamma, dbeta = _f32((H,), dev), _f32((H,), dev)
    ddiag = _f32((H,), dev) if wg_diag is not None else None
    check(lib.lpm_gating_bwd(ptr(act), ptr(g), B, H, ptr(gamma), ptr(beta), ptr(stats[0]), ptr(stats[1]), ptr(dout),
                             C.c_float(inv_scale), ptr(dact), ptr(dg), ptr(dgamma), ptr(dbeta), ptr(wg_diag), ptr(ddiag),
                             stream_ptr()), "lpm_gating_bwd")
    return (dact, dg, dgamma, dbeta) if wg_diag is None else (dact, dg, dgamma, dbeta, ddiag)


def l2_normalize_frames(x, out=None):
    """tf.nn.l2_normalize(model_input, 2) (train.py:262-264) on fp32 frames [B, max_frames, F]; out may be x itself."""
    lib = _lib.load()
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    F = x.shape[-1]
    check(lib.lpm_l2_normalize_rows(ptr(x), _ll(x.numel() // F), F, ptr(out), stream_ptr()), "lpm_l2_normalize_rows")
    return out


def add_diag(m, d, alpha=1.0):
    lib = _lib.load()
    n = d.numel()
    check(lib.lpm_add_diag(ptr(m), n, _ll(m.stride(0)), ptr(d), C.c_float(alpha), stream_ptr()), "lpm_add_diag")
    return m


def hidden_bn_relu6_fwd(x, gamma, beta, moving_mean, moving_var, *, training, relu6=True, save=False,
                        decay=BN_DECAY, eps=BN_EPS):
    """relu6(slim.batch_norm(x)) over the batch rows of x fp32 [B, H] (--netvlad_relu, frame_level_models.py:2321-2340).
    Returns (out32, out16[, (mean, rstd)])."""
    lib = _lib.load()
    B, H = x.shape
    dev = x.device
    out32, out16 = _f32((B, H), dev), _f16((B, H), dev)
    sm = _f32((2, H), dev) if save else None
    check(lib.lpm_hidden_bn_relu6_fwd(ptr(x), B, H, ptr(gamma), ptr(beta), ptr(moving_mean), ptr(moving_var), C.c_float(decay),
                                      C.c_float(eps), int(training), int(relu6), ptr(out32), ptr(out16),
                                      ptr(sm[0]) if save else None, ptr(sm[1]) if save else None, stream_ptr()),
          "lpm_hidden_bn_relu6_fwd")
    return (out32, out16, sm) if save else (out32, out16)


def hidden_bn_relu6_bwd(x, y, dy, gamma, stats, *, inv_scale, relu6=True):
    """dy (fp32, loss-scaled, gradient at y) -> dx in place; returns (dx, dgamma, dbeta) (parameter gradients unscaled)."""
    lib = _lib.load()
    B, H = x.shape
    dgamma, dbeta = _f32((H,), x.device), _f32((H,), x.device)
    check(lib.lpm_hidden_bn_relu6_bwd(ptr(x), ptr(y), ptr(dy), B, H, ptr(gamma), ptr(stats[0]), ptr(stats[1]), int(relu6),
                                      C.c_float(inv_scale), ptr(dy), ptr(dgamma), ptr(dbeta), stream_ptr()),
          "lpm_hidden_bn_relu6_bwd")
    return dy, dgamma, dbeta


def layernorm_joint_bwd(u, dy, dy_stride, B, rows, D, mean_rstd, gamma, *, inv_scale, mask=None, want_du_colsum=False):
    """Returns du (fp16 [B, rows, D]) or (du, du_masked) with a mask, dgamma, dbeta (fp32, unscaled)
    [, colsum(du_masked or du) unscaled]."""
    lib = _lib.load()
    dev = u.device
    ch = lib.lpm_layernorm_bwd_chunks()
    du = _f16((B, rows, D), dev)
    dum = _f16((B, rows, D), dev) if mask is not None else None
    ps = _f32((B * ch * 2,), dev)
    pc = _f32((B * ch, 2, D), dev)
    pdu = _f32((B * ch, D), dev) if want_du_colsum else None
    check(lib.lpm_layernorm_joint_bwd(ptr(u), ptr(dy), _ll(dy_stride), B, rows, D, ptr(mean_rstd), ptr(gamma), ptr(mask),
                                      ptr(du), ptr(dum), ptr(ps), ptr(pc), ptr(pdu), stream_ptr()), "lpm_layernorm_joint_bwd")
    if mask is not None:
        du = (du, dum)
    gb = colsum_final(pc, B * ch, 2 * D, 2 * D, alpha=inv_scale)      # [dgamma | dbeta] in one launch
    dgamma, dbeta = gb[:D], gb[D:]
    if want_du_colsum:
        return du, dgamma, dbeta, colsum_final(pdu, B * ch, D, D, alpha=inv_scale)
    return du, dgamma, dbeta


def netvlad_norm_bwd(z, rscale, dvhat, centers_t):
    lib = _lib.load()
    B, K, D = z.shape
    dz = torch.empty_like(z)
    q = _f32((B, K), z.device)
    check(lib.lpm_netvlad_norm_bwd(ptr(z), ptr(rscale), ptr(dvhat), _ll(B * K), K, D, ptr(centers_t), ptr(dz), ptr(q),
                                   stream_ptr()), "lpm_netvlad_norm_bwd")
    return dz, q


def assign_bwd(G, assign, q, S, stats, gamma, T, *, inv_scale):
    """Soft-assignment + cluster_bn backward.  Returns dS (fp16 [rows, K]), dgamma, dbeta (unscaled)."""
    lib = _lib.load()
    rows, K = G.shape
    dev = G.device
    nb = lib.lpm_assign_bwd_blocks()
    dsh = _f16((rows, K), dev)
    partial = _f32((nb, 2, K), dev)
    check(lib.lpm_assign_bwd1(ptr(G), ptr(assign), ptr(q), ptr(S), ptr(stats[0]), ptr(stats[1]), _ll(rows), T, K,
                              ptr(dsh), ptr(partial), stream_ptr()), "lpm_assign_bwd1")
    csum = _f32((2, K), dev)
    colsum_final(partial, nb, 2 * K, 2 * K, out=csum.view(-1))
    check(lib.lpm_assign_bwd2(ptr(dsh), ptr(S), ptr(stats[0]), ptr(stats[1]), ptr(gamma), ptr(csum), _ll(rows), K,
                              stream_ptr()), "lpm_assign_bwd2")
    dgamma = colsum_final(partial[:, 1], nb, 2 * K, K, alpha=inv_scale)
    dbeta = colsum_final(partial, nb, 2 * K, K, alpha=inv_scale)
    return dsh, dgamma, dbeta


def center_bwd(dV, Z, a_sum, centers_t, beta_in, inv_scale):
    lib = _lib.load()
    B, K, D = Z.shape
    dCt, E = _f32((K, D), Z.device), _f32((K, D), Z.device)
    check(lib.lpm_center_bwd(ptr(dV), ptr(Z), ptr(a_sum), B, K, D, ptr(centers_t), ptr(beta_in), C.c_float(inv_scale),
                             ptr(dCt), ptr(E), stream_ptr()), "lpm_center_bwd")
    return dCt, E


def input_bn_grad(Wc, dWc, dCt, E, gamma_in, dgamma_in, dbeta_in):
    lib = _lib.load()
    D, K = Wc.shape
    check(lib.lpm_input_bn_grad(ptr(Wc), ptr(dWc), ptr(dCt), ptr(E), D, K, ptr(gamma_in), ptr(dgamma_in), ptr(dbeta_in),
                                stream_ptr()), "lpm_input_bn_grad")


def cast_scaled_f16(x, alpha=1.0):
    lib = _lib.load()
    y = _f16(x.shape, x.device)
    check(lib.lpm_cast_scaled_f16(ptr(x), _ll(x.numel()), C.c_float(alpha), ptr(y), stream_ptr()), "lpm_cast_scaled_f16")
    return y


def mha_core_bwd(qkv, o, dout, lse, B, L, Dm, H, *, scale):
    lib = _lib.load()
    dqkv = torch.empty_like(qkv)
    check(lib.lpm_mha_core_bwd(ptr(qkv), _ll(qkv.stride(0)), ptr(o), ptr(dout), _ll(o.stride(0)), ptr(lse), B, L, Dm, H,
                               C.c_float(scale), ptr(dqkv), _ll(dqkv.stride(0)), stream_ptr()), "lpm_mha_core_bwd")
    return dqkv


def step_begin(flag, skipped):
    """Start-of-step latch of the optimiser's overflow flag (skipped += flag; flag = 0)."""
    check(_lib.load().lpm_step_begin(ptr(flag), ptr(skipped), stream_ptr()), "lpm_step_begin")


def adam_clip_step(flat_p, flat_g, flat_m, flat_v, table, chunk_begin, wd, *, clip, lr_t, scratch,
                   shadow=None, b1=0.9, b2=0.999, eps=1e-8, lr_dev=None, tensor_range=None, chunk_range=None):
    """Per-tensor (grad + wd*p) -> clip_by_norm -> Adam over the flat buffers (three launches).
    lr_dev: fp32 [1] device tensor holding the bias-corrected step size (read at execution time: graph replays).
    tensor_range / chunk_range: (first, count) of the tensors to update and of their chunks in `table` (default: all)."""
    lib = _lib.load()
    n_chunks, n_tensors = table.shape[0], wd.numel()
    partial, factor, norms, flag = scratch[:4]
    sp, sc, sl = shadow if shadow is not None else (None, None, None)
    if tensor_range is not None:
        (t0, nt), (c0, nc) = tensor_range, chunk_range
        if nt <= 0 or nc <= 0:
            return
        check(lib.lpm_adam_clip_step_range(ptr(flat_p), ptr(flat_g), ptr(flat_m), ptr(flat_v), ptr(table), c0, nc, ptr(chunk_begin),
                                           t0, nt, ptr(wd), ptr(sp), ptr(sc), ptr(sl), C.c_float(clip), C.c_float(lr_t), ptr(lr_dev),
                                           C.c_float(b1), C.c_float(b2), C.c_float(eps), ptr(partial), ptr(factor), ptr(norms),
                                           ptr(flag), stream_ptr()), "lpm_adam_clip_step")
        return
    if lr_dev is not None:
        check(lib.lpm_adam_clip_step_dev(ptr(flat_p), ptr(flat_g), ptr(flat_m), ptr(flat_v), ptr(table), n_chunks,
                                         ptr(chunk_begin), n_tensors, ptr(wd), ptr(sp), ptr(sc), ptr(sl), C.c_float(clip), ptr(lr_dev),
                                         C.c_float(b1), C.c_float(b2), C.c_float(eps), ptr(partial), ptr(factor), ptr(norms),
                                         ptr(flag), stream_ptr()), "lpm_adam_clip_step")
        return
    check(lib.lpm_adam_clip_step(ptr(flat_p), ptr(flat_g), ptr(flat_m), ptr(flat_v), ptr(table), n_chunks,
                                 ptr(chunk_begin), n_tensors, ptr(wd), ptr(sp), ptr(sc), ptr(sl), C.c_float(clip), C.c_float(lr_t), C.c_float(b1),
                                 C.c_float(b2), C.c_float(eps), ptr(partial), ptr(factor), ptr(norms), ptr(flag),
                                 stream_ptr()), "lpm_adam_clip_step")


def shard_sqnorm(p, g, table, wd1, partial, sumsq):
    lib = _lib.load()
    check(lib.lpm_shard_sqnorm(ptr(g), ptr(p), ptr(table), table.shape[0], ptr(wd1), ptr(partial), ptr(sumsq), stream_ptr()),
          "lpm_shard_sqnorm")


def shard_adam(p, g, m, v, table, wd1, sumsq, *, clip, lr_t, factor, norm, flag, shadow=None, b1=0.9, b2=0.999, eps=1e-8):
    lib = _lib.load()
    sp, sc, sl = shadow if shadow is not None else (None, None, None)
    check(lib.lpm_shard_adam(ptr(p), ptr(g), ptr(m), ptr(v), ptr(table), table.shape[0], ptr(wd1), ptr(sumsq), C.c_float(clip),
                             ptr(factor), ptr(norm), ptr(flag), ptr(sp), ptr(sc), ptr(sl), C.c_float(lr_t), C.c_float(b1),
                             C.c_float(b2), C.c_float(eps), stream_ptr()), "lpm_shard_adam")


def eval_topk(pred, labels_u8, k=20):
    """(top_val [B,k] fp32, top_idx [B,k] int32, top_lab [B,k] uint8, row_stats [B,3] = hit@1 | PERR | #labels)."""
    lib = _lib.load()
    B, V = pred.shape
    dev = pred.device
    assert pred.dtype == torch.float32 and labels_u8.dtype == torch.uint8 and pred.stride(1) == 1 and labels_u8.stride(1) == 1
    tv, ti = _f32((B, k), dev), torch.empty((B, k), dtype=torch.int32, device=dev)
    tl, rs = torch.empty((B, k), dtype=torch.uint8, device=dev), _f32((B, 3), dev)
    check(lib.lpm_eval_topk(ptr(pred), _ll(pred.stride(0)), ptr(labels_u8), _ll(labels_u8.stride(0)), B, V, k, ptr(tv), ptr(ti),
                            ptr(tl), ptr(rs), stream_ptr()), "lpm_eval_topk")
    return tv, ti, tl, rs


def eval_metrics(top_val, top_lab, row_stats):
    """Device tensor [3] = (hit@1, PERR, GAP) of the batch."""
    lib = _lib.load()
    B, k = top_val.shape
    out = _f32((3,), top_val.device)
    check(lib.lpm_eval_metrics(ptr(top_val), ptr(top_lab), B, k, ptr(row_stats), ptr(out), stream_ptr()), "lpm_eval_metrics")
    return out


_WS = {}


def _workspace(nbytes, dev):
    """Caller-owned scratch for the C-ABI calls that ask for one (grown on demand, reused across steps)."""
    t = _WS.get(dev)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _WS[dev] = t
    return t


def rank_adam_supported(R, N) -> bool:
    lib = _lib.load()
    lib.lpm_rank_adam_workspace_bytes.restype = C.c_ulonglong
    return int(lib.lpm_rank_adam_workspace_bytes(int(R), int(N))) > 0


def rank_grad_clip(gram_a, gram_g, alpha, clip, factor, norm, flag):
    """clip_by_norm factor of the rank-R gradient alpha*A^T G from the Gram matrices A A^T and G G^T."""
    lib = _lib.load()
    R = gram_a.shape[0]
    assert gram_a.shape == gram_g.shape == (R, R) and gram_a.is_contiguous() and gram_g.is_contiguous()
    check(lib.lpm_rank_grad_clip(ptr(gram_a), ptr(gram_g), R, C.c_float(alpha), C.c_float(clip), ptr(factor), ptr(norm),
                                 ptr(flag), stream_ptr()), "lpm_rank_grad_clip")


def rank_adam_workspace_bytes(R, N) -> int:
    lib = _lib.load()
    lib.lpm_rank_adam_workspace_bytes.restype = C.c_ulonglong
    return int(lib.lpm_rank_adam_workspace_bytes(int(R), int(N)))


def rank_adam_step(a16, g16, alpha, factor, flag, w, m, v, w16, *, lr_t=0.0, b1=0.9, b2=0.999, eps=1e-8, lr_dev=None,
                   tiled=False, workspace=None):
    """Adam on w [Kd, N] (fp32, with moments m, v) for the never-materialised gradient alpha * a16^T g16.
    lr_dev: fp32 [1] device tensor with the step size (instead of lr_t); tiled: the small-CTA kernel meant for a
    side stream underneath the backward (same results; an int > 1 also splits the columns over that many CTAs); workspace: caller-owned scratch (uint8, at least
    rank_adam_workspace_bytes) -- required when the call runs concurrently with other users of the shared one."""
    lib = _lib.load()
    R, Kd = a16.shape
    N = g16.shape[1]
    assert g16.shape[0] == R and tuple(w.shape) == (Kd, N) and w.is_contiguous() and m.is_contiguous() and v.is_contiguous()
    ws_bytes = rank_adam_workspace_bytes(R, N)
    ws = workspace if workspace is not None else _workspace(ws_bytes, a16.device)
    assert ws.numel() * ws.element_size() >= ws_bytes
    check(lib.lpm_rank_adam_step_ex(ptr(a16), _ll(a16.stride(0)), ptr(g16), _ll(g16.stride(0)), R, _ll(Kd), N, C.c_float(alpha),
                                    ptr(factor), ptr(flag), ptr(w), ptr(m), ptr(v), ptr(w16),
                                    _ll(w16.stride(0) if w16 is not None else 0), C.c_float(lr_t), ptr(lr_dev),
                                    int(tiled), C.c_float(b1), C.c_float(b2), C.c_float(eps), ptr(ws),
                                    C.c_ulonglong(ws_bytes), stream_ptr()), "lpm_rank_adam_step")


# ------------------------------------------------------------------------------------------------
# NetVladV2 helpers
# ------------------------------------------------------------------------------------------------
def mha_logit_stats(qkv, B, L, Dm, H):
    """(sum | sumsq) partials [B*H, 2, L] of the un-scaled attention logits per key channel."""
    lib = _lib.load()
    partial = _f32((B * H, 2, L), qkv.device)
    check(lib.lpm_mha_logit_stats(ptr(qkv), _ll(qkv.stride(0)), B, L, Dm, H, ptr(partial), stream_ptr()),
          "lpm_mha_logit_stats")
    return partial


def colstats_f16(x, rows=None, cols=None):
    lib = _lib.load()
    rows = rows or x.shape[0]
    cols = cols or x.shape[1]
    partial = _f32((lib.lpm_colstats_chunks(_ll(rows), int(cols)), 2, cols), x.device)
    check(lib.lpm_colstats_f16(ptr(x), _ll(x.stride(0)), _ll(rows), cols, ptr(partial), stream_ptr()), "lpm_colstats_f16")
    return partial


def affine_cols_f16(x, scale, shift, out=None):
    lib = _lib.load()
    assert x.is_contiguous()
    out = x if out is None else out
    check(lib.lpm_affine_cols_f16(ptr(x), ptr(out), _ll(x.shape[0]), x.shape[1], ptr(scale), ptr(shift), stream_ptr()),
          "lpm_affine_cols_f16")
    return out


def dropout_f16(x, rate, *, mask_in=None, mask_out=None, seed=0, seed_dev=None, out=None):
    lib = _lib.load()
    check(lib.lpm_dropout_f16(ptr(x), _ll(x.numel()), ptr(mask_in), ptr(mask_out), C.c_ulonglong(seed), ptr(seed_dev),
                              C.c_float(rate), ptr(out), stream_ptr()), "lpm_dropout_f16")
    return x


def netvlad_finalize_f16(z, rscale, out, out_stride):
    lib = _lib.load()
    B, K, D = z.shape
    check(lib.lpm_netvlad_finalize_f16(ptr(z), ptr(rscale), B, K, D, ptr(out), _ll(out_stride), stream_ptr()),
          "lpm_netvlad_finalize_f16")
    return out


def batch_norm_cols_f16(x, gamma, beta, moving_mean, moving_var, *, training, bessel, save=False, out=None):
    """slim.batch_norm over the rows of an fp16 matrix, applied in place (or into `out`).
    Returns (scale, shift[, (mean, rstd)])."""
    rows, cols = x.shape
    if training:
        part = colstats_f16(x)
        r = bn_finalize(part[:, 0], part[:, 1], rows, gamma, beta, moving_mean, moving_var, training=True,
                        bessel=bessel, save=save, psum_stride=2 * cols)
    else:
        r = bn_finalize(None, None, 1, gamma, beta, moving_mean, moving_var, training=False, bessel=bessel, save=save)
    affine_cols_f16(x, r[0], r[1], out=out)
    return r


def batch_norm_cols_bwd(dy, x_pre, stats, gamma, *, inv_scale, relu, q=None, T=1):
    """Backward of batch_norm_cols_f16 in training mode (x_pre = BN input).  dy: fp16 or fp32 [rows, C]; q (fp32
    [rows/T, C], optional) is subtracted from dy on the fly.  Returns dx (fp16; masked by x_pre > 0 when `relu`),
    dgamma, dbeta (fp32, unscaled)."""
    lib = _lib.load()
    rows, cols = x_pre.shape
    f32 = int(dy.dtype == torch.float32)
    assert dy.is_contiguous() and x_pre.is_contiguous()
    ch = lib.lpm_colstats_chunks(_ll(rows), int(cols))
    part = _f32((ch, 2, cols), dy.device)
    check(lib.lpm_batchnorm_bwd_stats(ptr(dy), f32, _ll(dy.stride(0)), ptr(q), T, ptr(x_pre), _ll(x_pre.stride(0)),
                                      _ll(rows), cols, ptr(stats[0]), ptr(stats[1]), 0, ptr(part), stream_ptr()),
          "lpm_batchnorm_bwd_stats")
    csum = _f32((2, cols), dy.device)
    colsum_final(part, ch, 2 * cols, 2 * cols, out=csum.view(-1))
    dbeta = colsum_final(part, ch, 2 * cols, cols, alpha=inv_scale)
    dgamma = colsum_final(part[:, 1], ch, 2 * cols, cols, alpha=inv_scale)
    dx = torch.empty_like(x_pre)
    check(lib.lpm_batchnorm_bwd_apply(ptr(dy), f32, ptr(q), T, ptr(dx), ptr(x_pre), _ll(rows), cols, ptr(stats[0]),
                                      ptr(stats[1]), ptr(gamma), ptr(csum), int(relu), stream_ptr()),
          "lpm_batchnorm_bwd_apply")
    return dx, dgamma, dbeta


def bn_output_param_grads(dy, y, beta, gamma, *, inv_scale):
    """dgamma / dbeta of a batch norm from its OUTPUT y: xhat = (y - beta)/gamma (input_bn, whose input needs no grad)."""
    lib = _lib.load()
    rows, cols = y.shape
    ch = lib.lpm_colstats_chunks(_ll(rows), int(cols))
    part = _f32((ch, 2, cols), dy.device)
    check(lib.lpm_batchnorm_bwd_stats(ptr(dy), int(dy.dtype == torch.float32), _ll(dy.stride(0)), None, 1, ptr(y),
                                      _ll(y.stride(0)), _ll(rows), cols, ptr(beta), ptr(gamma), 1, ptr(part),
                                      stream_ptr()), "lpm_batchnorm_bwd_stats")
    dbeta = colsum_final(part, ch, 2 * cols, cols, alpha=inv_scale)
    dgamma = colsum_final(part[:, 1], ch, 2 * cols, cols, alpha=inv_scale)
    return dgamma, dbeta


def sub_q_cast_f16(G, q, T):
    lib = _lib.load()
    rows, K = G.shape
    out = _f16((rows, K), G.device)
    check(lib.lpm_sub_q_cast_f16(ptr(G), ptr(q), _ll(rows), T, K, ptr(out), stream_ptr()), "lpm_sub_q_cast_f16")
    return out


def dmajor_to_kmajor_f16(dv, B, K, D):
    lib = _lib.load()
    out = _f16((B, K, D), dv.device)
    check(lib.lpm_dmajor_to_kmajor_f16(ptr(dv), _ll(dv.stride(0)), B, K, D, ptr(out), stream_ptr()),
          "lpm_dmajor_to_kmajor_f16")
    return out


def mha_core_bwd_bn(mode, qkv, o, dout, lse, B, L, Dm, H, ks, kb, mean, rstd, m1=None, m2=None):
    lib = _lib.load()
    part = _f32((B * H, 2, L), qkv.device) if mode == 1 else None
    dqkv = torch.empty_like(qkv) if mode == 2 else None
    check(lib.lpm_mha_core_bwd_bn(mode, ptr(qkv), _ll(qkv.stride(0)), ptr(o), ptr(dout), _ll(o.stride(0)), ptr(lse), B, L,
                                  Dm, H, ptr(ks), ptr(kb), ptr(mean), ptr(rstd), ptr(m1), ptr(m2), ptr(part), ptr(dqkv),
                                  _ll(qkv.stride(0)), stream_ptr()), "lpm_mha_core_bwd_bn")
    return part if mode == 1 else dqkv
