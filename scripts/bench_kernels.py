"""CUDA-event timings of the individual hot kernels at config-1 shapes (B=80).  A/B aid, not a bench line."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops, _lib
_lib.load()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


which = set(sys.argv[1:]) or {"mha", "pool", "ln", "gemm"}
B, L, D, H = 80, 256, 1024, 64
if "mha" in which:
    qkv = torch.randn(B * L, 3 * D, device=dev).half()
    dout = torch.randn(B * L, D, device=dev).half()
    o, lse = ops.mha_core_fwd(qkv, B, L, D, H, scale=0.25, want_lse=True)
    us = t(lambda: ops.mha_core_fwd(qkv, B, L, D, H, scale=0.25, want_lse=True))
    print(f"mha_fwd  rgb  {us:8.1f} us  ({4.0 * L * L * D * B / us / 1e6:6.1f} TFLOP/s)")
    us = t(lambda: ops.mha_core_bwd(qkv, o, dout, lse, B, L, D, H, scale=0.25))
    print(f"mha_bwd  rgb  {us:8.1f} us  ({10.0 * L * L * D * B / us / 1e6:6.1f} TFLOP/s)")
if "pool" in which:
    T, Kc = 256, 256
    wc = (torch.randn(D, Kc, device=dev) / 32).half(); ct = torch.randn(D, Kc, device=dev) / 32
    one, zero = torch.ones(Kc, device=dev), torch.zeros(Kc, device=dev)
    for Bp in (80, 148, 1184):
        xb = torch.randn(Bp * T, D, device=dev).half()
        us = t(lambda: ops.netvlad_pool_fwd(xb, Bp, T, wc, one, zero, ct), iters=5)
        print(f"pool_fwd B={Bp:5d} {us:8.1f} us  ({4.0 * T * D * Kc * Bp / us / 1e6:6.1f} TFLOP/s)")
        us = t(lambda: ops.netvlad_pool_fwd(xb, Bp, T, wc, one, zero, ct, save_assign=True), iters=5)
        print(f"pool_fwd B={Bp:5d} {us:8.1f} us  (training: assignment saved)")
if "gemm" in which:
    for (M, N, K) in ((20480, 4096, 1024), (20480, 1024, 4096), (20480, 3072, 1024), (20480, 1024, 1024)):
        a = torch.randn(M, K, device=dev).half(); w = (torch.randn(K, N, device=dev) * 0.03).half()
        out = torch.empty(M, N, dtype=torch.float16, device=dev)
        us = t(lambda: ops.gemm(a, w, out=out))
        print(f"gemm {M}x{N}x{K} {us:8.1f} us  ({2.0 * M * N * K / us / 1e6:6.1f} TFLOP/s)")
if "adam" in which:
    R, Kd, N = 80, 270336, 512
    A = torch.randn(R, Kd, device=dev).half(); G = torch.randn(R, N, device=dev).half()
    w = torch.randn(Kd, N, device=dev) * 0.05; m = torch.zeros_like(w); v = torch.zeros_like(w)
    w16 = torch.empty(Kd, N, dtype=torch.float16, device=dev)
    fac = torch.ones(1, device=dev); flag = torch.zeros(1, dtype=torch.int32, device=dev)
    us = t(lambda: ops.rank_adam_step(A, G, 1.0 / 640, fac, flag, w, m, v, w16, lr_t=2e-4), iters=5)
    gb = Kd * N * 26 / 1e9
    print(f"rank_adam_step {us:8.1f} us  ({gb / us * 1e6:6.0f} GB/s of {gb:.2f} GB algorithmic)")
    us = t(lambda: (ops.gemm(A, A, b_mn=False, out_dtype=torch.float32), ops.gemm(G, G, b_mn=False, out_dtype=torch.float32)), iters=5)
    print(f"gram GEMMs     {us:8.1f} us")
if "ln" in which:
    a = torch.randn(B, L, D, device=dev).half(); b = torch.randn(B, L, D, device=dev).half()
    gm, bt = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    out = torch.empty(B, L * D, dtype=torch.float16, device=dev)
    us = t(lambda: ops.layernorm_joint_fwd(a, b, None, B, L, D, gm, bt, out=out, out_stride=L * D))
    print(f"ln single (infer)  {us:8.1f} us  ({3 * a.numel() * 2 / us / 1e3:6.0f} GB/s algorithmic)")
    us = t(lambda: ops.layernorm_chain_fwd(a, b, B, L, D, gm, bt, gm, bt, out=out, out_stride=L * D))
    print(f"ln chain  (infer)  {us:8.1f} us  ({3 * a.numel() * 2 / us / 1e3:6.0f} GB/s algorithmic)")
    us = t(lambda: ops.layernorm_chain_fwd(a, b, B, L, D, gm, bt, gm, bt, out=out, out_stride=L * D, save=True))
    print(f"ln chain  (train)  {us:8.1f} us  ({5 * a.numel() * 2 / us / 1e3:6.0f} GB/s algorithmic)")
