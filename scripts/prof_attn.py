"""Attention core at the config-1 shape [B=80, H=64, L=256, dh=16] (+ the audio shape) for ncu / timing:
prof_attn.py [iters]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for (B, L, Dm, H) in ((80, 256, 1024, 64), (80, 64, 128, 16)):
    qkv = (torch.randn(B * L, 3 * Dm, device=dev) * 0.5).half()
    do = (torch.randn(B * L, Dm, device=dev) * 0.1).half()
    scale = (Dm // H) ** -0.5
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(2):
        o, lse = ops.mha_core_fwd(qkv, B, L, Dm, H, scale=scale, want_lse=True)
        ops.mha_core_bwd(qkv, o, do, lse, B, L, Dm, H, scale=scale)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(iters):
        o, lse = ops.mha_core_fwd(qkv, B, L, Dm, H, scale=scale, want_lse=True)
    ev[1].record()
    for _ in range(iters):
        ops.mha_core_bwd(qkv, o, do, lse, B, L, Dm, H, scale=scale)
    ev[2].record()
    torch.cuda.synchronize()
    print(f"mha [B={B} H={H} L={L} dh={Dm // H}] fwd {ev[0].elapsed_time(ev[1]) / iters * 1e3:.1f} us  bwd {ev[1].elapsed_time(ev[2]) / iters * 1e3:.1f} us")
