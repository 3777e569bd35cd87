import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
dev = torch.device("cuda:0")
M, N, K = 20480, 4096, 64
a = torch.randn(M, K, device=dev).half(); w = (torch.randn(K, N, device=dev) * 0.03).half(); o = torch.empty(M, N, dtype=torch.float16, device=dev)
for _ in range(3): ops.gemm(a, w, out=o)
torch.cuda.synchronize()
