import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops, _lib
dev = torch.device("cuda:0")
M, N, K = 256, 256, 64
a = torch.randn(M, K, device=dev).half(); w = torch.randn(K, N, device=dev).half(); o = torch.empty(M, N, dtype=torch.float16, device=dev)
for _ in range(10): ops.gemm(a, w, out=o)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(500): ops.gemm(a, w, out=o)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"ops.gemm CPU issue {1e6*(t1-t0)/500:.1f} us/call, drained after {1e6*(t2-t1):.0f} us")
x = torch.randn(1024, 1024, device=dev).half(); rs = torch.ones(1024, device=dev)
t0 = time.perf_counter()
for _ in range(500): ops.scale_rows_f16(x, rs)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"ops.scale_rows_f16 CPU issue {1e6*(t1-t0)/500:.1f} us/call")
t0 = time.perf_counter()
for _ in range(500): torch.empty(1024, 1024, device=dev, dtype=torch.float16)
t1 = time.perf_counter()
print(f"torch.empty {1e6*(t1-t0)/500:.1f} us/call")
import ctypes as C
lib = _lib.load()
t0 = time.perf_counter()
for _ in range(2000): lib.lpm_gemm_tile_n(256)
t1 = time.perf_counter()
print(f"trivial ctypes call {1e6*(t1-t0)/2000:.2f} us/call")
# big gemm GPU time with events over many back-to-back launches
from bench import time_cuda
for (M, N, K) in ((20480, 4096, 1024), (20480, 4096, 64), (20480, 1024, 1024)):
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(K, N, device=dev) * 0.03).half(); o = torch.empty(M, N, dtype=torch.float16, device=dev)
    ms = time_cuda(lambda: ops.gemm(a, w, out=o), 50)
    print(f"gemm {M}x{N}x{K}: {ms*1e3:.1f} us ({2.0*M*N*K/ms/1e9:.0f} TF)")
