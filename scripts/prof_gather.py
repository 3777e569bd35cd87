"""Launches for an ncu --set full capture of the warp-per-frame gather kernels (a2 + a3) and the K = 256 soft-assignment
backward pass at the config-1 shapes."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
dev = torch.device("cuda:0")
B, T, K, F = 80, 256, 256, 1152
codes = torch.randint(0, 256, (B, 300, F), dtype=torch.uint8, device=dev)
nf = torch.full((B,), 300, dtype=torch.int32, device=dev)
one, zero = torch.ones(F, device=dev), torch.zeros(F, device=dev)
G = torch.randn(B * T, K, device=dev); A = torch.softmax(torch.randn(B * T, K, device=dev), -1).half()
S = torch.randn(B * T, K, device=dev).half(); q = torch.randn(B, K, device=dev)
stats = (torch.zeros(K, device=dev), torch.ones(K, device=dev)); gamma = torch.ones(K, device=dev)
for _ in range(2):
    ops.sample_bn_stats(codes, nf, T)
    ops.sample_bn_apply(codes, nf, T, one, zero)
    ops.assign_bwd(G, A, q, S, stats, gamma, T, inv_scale=1.0)
torch.cuda.synchronize()
