"""Two launches of each tcgen05 attention kernel at the config-1 shape, for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
dev = torch.device("cuda:0")
B, L, Dm, H = 80, 256, 1024, 64
qkv = (torch.randn(B * L, 3 * Dm, device=dev) * 0.5).half()
do = (torch.randn(B * L, Dm, device=dev) * 0.1).half()
for _ in range(2):
    o, lse = ops.mha_core_fwd(qkv, B, L, Dm, H, scale=0.25, want_lse=True)
    ops.mha_core_bwd(qkv, o, do, lse, B, L, Dm, H, scale=0.25)
torch.cuda.synchronize()
