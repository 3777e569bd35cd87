import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
from bench import time_cuda
dev = torch.device("cuda:0")
def run(M, N, K, **kw):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(K, N, device=dev) * 0.03).half()
    o = torch.empty(M, N, dtype=torch.float16, device=dev)
    t1 = time_cuda(lambda: ops.gemm(a, w, out=o, **kw), 20)
    t2 = time_cuda(lambda: ops.gemm(a, w, out="none", **kw), 20)
    f = 2.0 * M * N * K / 1e9
    print(f"M={M} N={N} K={K} {kw}: store {t1*1e3:.1f} us {f/t1:.0f} TF | no-store {t2*1e3:.1f} us {f/t2:.0f} TF")
run(20480, 4096, 1024)
run(20480, 1024, 1024)
run(20480, 3072, 1024)
run(20480, 1024, 4096)
run(20480, 4096, 1024, force_bn=128)
run(8192, 8192, 8192)
print("--- epilogue-bound probes (K=64)")
run(20480, 4096, 64)
run(20480, 4096, 64, no_tma_store=True)
run(20480, 4096, 128)
run(20480, 4096, 256)
run(20480, 4096, 512)
