"""The tcgen05 GEMM against cuBLAS (torch.matmul) at the attention-block shapes of config 1, same timing loop.
A/B aid: cuBLAS is a yardstick here, never on the product path."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
dev = torch.device("cuda:0")


def t(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for (M, N, K) in ((20480, 4096, 1024), (20480, 1024, 4096), (20480, 3072, 1024), (20480, 1024, 1024), (8192, 8192, 8192)):
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(K, N, device=dev) * 0.03).half()
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    us = t(lambda: ops.gemm(a, w, out=out))
    ub = t(lambda: torch.matmul(a, w, out=out))
    wt = w.t().contiguous()
    ut = t(lambda: torch.matmul(a, wt.t(), out=out))
    fl = 2.0 * M * N * K / 1e6
    print(f"{M}x{N}x{K}: lpm {us:7.1f} us {fl / us:7.1f} TF/s | cuBLAS NN {ub:7.1f} us {fl / ub:7.1f} TF/s | cuBLAS NT {ut:7.1f} us {fl / ut:7.1f} TF/s | lpm/cuBLAS {min(ub, ut) / us:.3f}")
