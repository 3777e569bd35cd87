import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
dev = torch.device("cuda:0")
T, D, Kc, B = 256, 1024, 256, 148
xb = torch.randn(B * T, D, device=dev).half()
wc = (torch.randn(D, Kc, device=dev) / 32).half(); ct = ops.transpose_f32_dual(torch.randn(D, Kc, device=dev) / 32, want32=False)[1]
one, zero = torch.ones(Kc, device=dev), torch.zeros(Kc, device=dev)
for _ in range(3):
    ops.netvlad_pool_fwd(xb, B, T, wc, one, zero, ct)
torch.cuda.synchronize()
