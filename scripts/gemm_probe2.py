"""Stand-alone timing of the attention-block GEMMs of one training step in their three operand layouts:
forward Y = X W (NN), backward dX = dY W^T (NT), weight gradient dW = X^T dY (TN)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
from bench import time_cuda
dev = torch.device("cuda:0")
R = 20480
tot = {"fwd": 0.0, "dx": 0.0, "dw": 0.0}
for name, i, o in (("qkv", 1024, 3072), ("out", 1024, 1024), ("ffn1", 1024, 4096), ("ffn2", 4096, 1024)):
    x = torch.randn(R, i, device=dev).half(); w = (torch.randn(i, o, device=dev) * 0.03).half()
    dy = torch.randn(R, o, device=dev).half()
    y = torch.empty(R, o, dtype=torch.float16, device=dev); dx = torch.empty(R, i, dtype=torch.float16, device=dev)
    dw = torch.empty(i, o, dtype=torch.float32, device=dev)
    f = 2.0 * R * i * o / 1e9
    t = time_cuda(lambda: ops.gemm(x, w, out=y), 20); tot["fwd"] += t
    t2 = time_cuda(lambda: ops.gemm(dy, w, b_mn=False, out=dx), 20); tot["dx"] += t2
    t3 = time_cuda(lambda: ops.gemm(x, dy, a_mn=True, b_mn=True, out_dtype=torch.float32, out=dw), 20); tot["dw"] += t3
    print(f"{name:5s} in={i} out={o}: fwd {t*1e3:6.1f} us {f/t:5.0f} TF | dX {t2*1e3:6.1f} us {f/t2:5.0f} TF | dW {t3*1e3:6.1f} us {f/t3:5.0f} TF")
print({k: round(v * 1e3, 1) for k, v in tot.items()}, "us; ideal at 1409 TF:", round(2.0 * R * (3 + 1 + 4 + 4) * 1024 * 1024 / 1409e9 * 1e3, 1), "us each")
