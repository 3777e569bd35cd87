"""Per-unit clock stamps of CTA 0 of the tcgen05 attention backward (lpm_debug_set_mha_clock)."""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops, _lib
dev = torch.device("cuda:0")
lib = _lib.load()
B, L, Dm, H = 80, 256, 1024, 64
qkv = (torch.randn(B * L, 3 * Dm, device=dev) * 0.5).half()
do = (torch.randn(B * L, Dm, device=dev) * 0.1).half()
o, lse = ops.mha_core_fwd(qkv, B, L, Dm, H, scale=0.25, want_lse=True)
for _ in range(2): ops.mha_core_bwd(qkv, o, do, lse, B, L, Dm, H, scale=0.25)
dbg = torch.zeros(512, dtype=torch.int64, device=dev)
lib.lpm_debug_set_mha_clock(C.c_void_p(dbg.data_ptr()))
ops.mha_core_bwd(qkv, o, do, lse, B, L, Dm, H, scale=0.25)
torch.cuda.synchronize()
lib.lpm_debug_set_mha_clock(None)
d = dbg.cpu()
t0 = int(d[0])
wg = d[:512].reshape(4, 16, 8) - t0
print("v2 unit | wg0 (pair 0): begin sdp_ok math slot_ok stored fenced arrived | wg3 (pair 1): same")
for u in range(16):
    print(u, "|", " ".join(f"{int(x):6d}" for x in wg[0, u, :7]), "|", " ".join(f"{int(x):6d}" for x in wg[3, u, :7]))
