"""tcgen05 attention core (lpm_attn_tc.cu) against the warp-level kernels and an fp64 torch reference, with an error
breakdown per head-in-group / row tile (diagnostic), then timings at the config-1 shape.  check_attn_tc.py [iters]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
from learnablepoolingmethods_b200._lib import load
lib = load()
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def breakdown(name, a, b, B, L, Dm, H):
    a = a.float().reshape(B, L, H, 16); b = b.float().reshape(B, L, H, 16)
    rows = []
    for hp in range(4):
        for ft in range(2):
            rows.append(f"h%4={hp} ft={ft}: {rel(a[:, ft * 128:(ft + 1) * 128, hp::4], b[:, ft * 128:(ft + 1) * 128, hp::4]):.2e}")
    print(f"  {name}: total {rel(a, b):.2e} | " + " | ".join(rows))


B, L, Dm, H = 2, 256, 1024, 64
g = torch.Generator().manual_seed(1)
qkv = torch.randn(B * L, 3 * Dm, generator=g).half().to(dev)
do = (torch.randn(B * L, Dm, generator=g) * 0.5).half().to(dev)
scale = 16 ** -0.5
t = qkv.double().requires_grad_(True)
q, k, v = [u.reshape(B, L, H, 16).permute(0, 2, 1, 3) for u in t.split(Dm, dim=1)]
logits = (q * scale) @ k.transpose(-1, -2)
ref = (torch.softmax(logits, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, Dm)
ref.backward(do.double())
lse_ref = torch.logsumexp(logits, -1).reshape(-1)
res = {}
for mode in (0, 1, 2):
    lib.lpm_debug_set_mha_tc_mode(mode)
    o, lse = ops.mha_core_fwd(qkv, B, L, Dm, H, scale=scale, want_lse=True)
    torch.cuda.synchronize()
    print(f"mode {mode} fwd: out rel {rel(o, ref):.2e}  lse rel {rel(lse.reshape(-1), lse_ref):.2e}")
    if mode == 3:
        breakdown("out vs fp64", o, ref.detach(), B, L, Dm, H)
    dqkv = ops.mha_core_bwd(qkv, o, do, lse, B, L, Dm, H, scale=scale)
    torch.cuda.synchronize()
    for i, nm in enumerate("qkv"):
        e = rel(dqkv[:, i * Dm:(i + 1) * Dm], t.grad[:, i * Dm:(i + 1) * Dm])
        print(f"mode {mode} bwd d{nm} rel {e:.2e}")
        if mode == 2 and e > 1e-3:
            breakdown(f"d{nm} vs fp64", dqkv[:, i * Dm:(i + 1) * Dm], t.grad[:, i * Dm:(i + 1) * Dm], B, L, Dm, H)
    res[mode] = (o, lse, dqkv)
for fm in (1, 2):
    lib.lpm_debug_set_mha_tc_mode((fm << 2) | 2)
    o, lse = ops.mha_core_fwd(qkv, B, L, Dm, H, scale=scale, want_lse=True)
    torch.cuda.synchronize()
    print(f"forward tc v{fm}: out rel {rel(o, ref):.2e}  lse rel {rel(lse.reshape(-1), lse_ref):.2e}  vs warp-level {rel(o, res[0][0]):.2e}")
    if rel(o, ref) > 1e-3:
        breakdown("out vs fp64", o, ref.detach(), B, L, Dm, H)
print("tc vs legacy: dqkv v1", rel(res[1][2], res[0][2]), "v2", rel(res[2][2], res[0][2]))

B = 80
qkv = (torch.randn(B * L, 3 * Dm, device=dev) * 0.5).half()
do = (torch.randn(B * L, Dm, device=dev) * 0.1).half()
for mode in (0, 1, 2, 6, 10):
    lib.lpm_debug_set_mha_tc_mode(mode)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(3):
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
    print(f"mode {mode} [B=80 H=64 L=256 dh=16] fwd {ev[0].elapsed_time(ev[1]) / iters * 1e3:.1f} us  bwd {ev[1].elapsed_time(ev[2]) / iters * 1e3:.1f} us")
lib.lpm_debug_set_mha_tc_mode(2)
