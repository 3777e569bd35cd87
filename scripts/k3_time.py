"""K3 as one launch (ops.gemm_splitk_gated) against the four-launch path at the config-1 head shape, training-like (one
set of 74 partial slabs) and inference-like (a second pass leaves 148 more)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
dev = torch.device("cuda:0")
B, Kd, H = 80, 270336, 512
a = (torch.randn(B, Kd, device=dev) * 0.05).half(); a2 = torch.cat([a, a * 0.001]).contiguous()
w = (torch.randn(Kd, H, device=dev) * 0.05).half(); wlo = (w.float() * 0.001).half()
bias = torch.randn(H, device=dev); wg = torch.randn(H, H, device=dev) / 22
gam, bet, mm, mv = torch.ones(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.ones(H, device=dev)


def t(fn, iters=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def fused(train):
    p1 = None
    if not train:
        p1 = ops.gemm(a2, w, splits=74); p1 = p1.view(p1.shape[0] * 2, B, H)
    ops.gemm_splitk_gated(a, w if train else wlo, splits=74, bias=bias, wg=wg, gamma=gam, beta=bet, moving_mean=mm, moving_var=mv,
                          training=train, parts2=p1)


def plain(train):
    act32 = torch.empty(B, H, device=dev); a3 = torch.empty(B, 3 * H, dtype=torch.float16, device=dev)
    if not train:
        p1 = ops.gemm(a2, w, splits=74); p1 = p1.view(p1.shape[0] * 2, B, H)
        plo = ops.gemm(a, wlo, splits=74)
        ops.splitk_reduce(p1, bias=bias, out32=act32, out16=a3, parts2=plo, split3=True)
    else:
        ops.splitk_reduce(ops.gemm(a, w, splits=74), bias=bias, out32=act32, out16=a3, split3=True)


for train in (True, False):
    print("training-like" if train else "inference-like", f"fused (GEMM + reduce + gate product + gating) {t(lambda: fused(train)):.1f} us | plain GEMM(s) + reduce only {t(lambda: plain(train)):.1f} us")
