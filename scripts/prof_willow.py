"""Launches for one ncu --set full capture of the kernels added for WillowModelReg: the 64x64 fp16 tile transposes
(d-major flatten and its gradient), the gather with explicit indices, and the orthogonal regulariser."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
dev = torch.device("cuda:0")
B, T, K, D, F = 80, 256, 256, 1024, 1152
z = torch.randn(B, K, D, device=dev).half(); rs = torch.rand(B, K, device=dev)
out = torch.empty(B, K * D + 64 * 128, dtype=torch.float16, device=dev)
codes = torch.randint(0, 256, (B, 300, F), dtype=torch.uint8, device=dev)
nf = torch.full((B,), 300, dtype=torch.int32, device=dev)
one, zero = torch.ones(F, device=dev), torch.zeros(F, device=dev)
w = torch.randn(D, K, device=dev) / 32
dw = torch.zeros(D, K, device=dev)
for _ in range(2):
    ops.netvlad_finalize_f16(z, rs, out[:, :K * D], out.stride(0))
    ops.dmajor_to_kmajor_f16(out[:, :K * D], B, K, D)
    idx = ops.random_frame_index(nf, T, 300, seed=1)
    ops.gather_bn_stats(codes, idx, T)
    ops.gather_bn_apply(codes, idx, T, one, zero)
    ops.ortho_reg(w, 1e-4, dw=dw)
torch.cuda.synchronize()
