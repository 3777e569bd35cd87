import os, sys, torch
sys.path.insert(0, "/root/repo")
from learnablepoolingmethods_b200 import ops
dev = torch.device("cuda:0")
B, Kd, H = 80, 270336, 512
a = (torch.randn(B, Kd, device=dev) * 0.05).half(); a2 = torch.cat([a, a * 0.001]).contiguous()
w = (torch.randn(Kd, H, device=dev) * 0.05).half()
bias = torch.randn(H, device=dev); wg = torch.randn(H, H, device=dev) / 22
gam, bet, mm, mv = torch.ones(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.ones(H, device=dev)
for _ in range(3):
    p1 = ops.gemm(a2, w, splits=74); p1 = p1.view(p1.shape[0] * 2, B, H)
    ops.gemm_splitk_gated(a, w, splits=74, bias=bias, wg=wg, gamma=gam, beta=bet, moving_mean=mm, moving_var=mv, training=False, parts2=p1)
torch.cuda.synchronize()
