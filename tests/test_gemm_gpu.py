"""tcgen05 GEMM kernel vs a torch fp32 product of the same fp16-rounded operands (floating-point
kernel => torch fp32 reference; model-level parity against the oracle lives in test_parity_gpu.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, dev, scale=1.0):
    return (torch.randn(shape, device=dev) * scale).half()


def _check(out, ref, tol=2e-3):
    err = (out.float() - ref).norm() / ref.norm().clamp_min(1e-20)
    assert err < tol, f"rel L2 err {err:.3e}"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 320, 200), (20, 64, 1000), (513, 1024, 136), (256, 96, 72)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (False, False), (True, True), (True, False)])
def test_gemm_layouts(cuda, M, N, K, a_mn, b_mn):
    from learnablepoolingmethods_b200 import ops
    torch.manual_seed(M * 7 + N + K)
    # pad leading dims to multiples of 8 (TMA stride rule) but keep logical sizes ragged
    def alloc(r, c):
        cp = (c + 7) // 8 * 8
        return _rand((r, cp), cuda)[:, :c]
    A = alloc(K, M) if a_mn else alloc(M, K)
    B = alloc(K, N) if b_mn else alloc(N, K)
    Af = (A.t() if a_mn else A).float()
    Bf = (B if b_mn else B.t()).float()
    ref = Af @ Bf
    out = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    torch.cuda.synchronize()
    _check(out, ref, 1e-5)
    out16 = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float16)
    _check(out16, ref, 1e-3)


@pytest.mark.parametrize("M,N,K", [(4900, 1024, 200), (9728, 512, 136), (2560, 1000, 72)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (False, False), (True, True), (True, False)])
def test_gemm_cta_pairs(cuda, M, N, K, a_mn, b_mn):
    """Products with at least 74 tiles of 256 x 256 run on 2-CTA clusters (tcgen05.mma.cta_group::2): odd numbers of
    128-row tiles (4900 -> 39, the peer CTA's last tile is all padding), ragged N / K, every operand layout, fused
    epilogue; identical to the 1-CTA kernel's result (the accumulation order per element is the same)."""
    import ctypes as C
    from learnablepoolingmethods_b200 import ops, _lib
    lib = _lib.load()
    torch.manual_seed(M + N + K)
    def alloc(r, c):
        cp = (c + 7) // 8 * 8
        return _rand((r, cp), cuda)[:, :c]
    A = alloc(K, M) if a_mn else alloc(M, K)
    B = alloc(K, N) if b_mn else alloc(N, K)
    bias = torch.randn(N, device=cuda)
    ref = torch.relu((A.t() if a_mn else A).float() @ (B if b_mn else B.t()).float() + bias)
    out = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=True, out_dtype=torch.float32)
    out16 = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=True)
    lib.lpm_debug_set_gemm_pair_mode(0)
    try:
        single = ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=True, out_dtype=torch.float32)
    finally:
        lib.lpm_debug_set_gemm_pair_mode(1)
    torch.cuda.synchronize()
    _check(out, ref, 1e-5)
    _check(out16, ref, 1e-3)
    assert torch.equal(out, single)


def test_gemm_cta_pairs_batched_stats(cuda):
    from learnablepoolingmethods_b200 import ops
    torch.manual_seed(3)
    Bt, M, N, K = 20, 512, 512, 192
    A, B = _rand((Bt, M, K), cuda), _rand((K, N), cuda, 0.1)
    rs = torch.rand(Bt, M, device=cuda) + 0.5
    ref = rs[:, :, None] * torch.matmul(A.float(), B.float())
    out, st = ops.gemm(A, B, row_scale=rs, out_dtype=torch.float32, stats=True)
    _check(out, ref, 1e-5)
    _check(st[0].sum(dim=1), ref.sum(dim=2), 1e-4)


def test_gemm_epilogue_bias_relu_rowscale_stats(cuda):
    from learnablepoolingmethods_b200 import ops
    torch.manual_seed(0)
    M, N, K = 200, 384, 320
    A, B = _rand((M, K), cuda), _rand((K, N), cuda, 0.1)
    bias = torch.randn(N, device=cuda)
    rs = torch.rand(M, device=cuda) + 0.5
    ref = torch.relu(0.5 * rs[:, None] * (A.float() @ B.float()) + bias)
    out, st = ops.gemm(A, B, bias=bias, row_scale=rs, relu=True, alpha=0.5, out_dtype=torch.float32, stats=True)
    _check(out, ref, 1e-5)
    _check(st[0].sum(dim=(0, 1)), ref.sum(dim=1), 1e-4)
    _check(st[1].sum(dim=(0, 1)), (ref * ref).sum(dim=1), 1e-4)


def test_gemm_batched_and_splitk(cuda):
    from learnablepoolingmethods_b200 import ops
    torch.manual_seed(1)
    Bt, M, N, K = 5, 256, 256, 1024
    A, B = _rand((Bt, M, K), cuda), _rand((Bt, N, K), cuda, 0.1)
    ref = torch.matmul(A.float(), B.float().transpose(1, 2))
    out = ops.gemm(A, B, b_mn=False, out_dtype=torch.float32)
    _check(out, ref, 1e-5)
    # shared B
    out2 = ops.gemm(A, B[0], b_mn=False, out_dtype=torch.float32)
    _check(out2, torch.matmul(A.float(), B[0].float().t()), 1e-5)
    # split-K on a skinny problem (hidden-projection shape class)
    M, N, K = 80, 512, 8192 + 64
    A, W = _rand((M, K), cuda), _rand((K, N), cuda, 0.05)
    parts = ops.gemm(A, W, splits=37)
    _check(parts.sum(dim=0), A.float() @ W.float(), 1e-5)
    # accumulate
    acc = torch.ones(M, N, device=cuda)
    ops.gemm(A[:, :256].contiguous(), W[:256].contiguous(), out=acc, accumulate=True)
    _check(acc, 1 + A[:, :256].float() @ W[:256].float(), 1e-5)


def test_gemm_large_perf_shape(cuda):
    from learnablepoolingmethods_b200 import ops
    torch.manual_seed(2)
    M, N, K = 20480, 1024, 1024
    A, W = _rand((M, K), cuda), _rand((K, N), cuda, 0.03)
    out = ops.gemm(A, W)
    ref = (A.float() @ W.float())
    _check(out, ref, 1e-3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.gemm(A, W, out=out)
    ev0.record()
    for _ in range(10):
        ops.gemm(A, W, out=out)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 10
    print(f"\n[gemm 20480x1024x1024] {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
