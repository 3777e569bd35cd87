import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops, _lib
dev = torch.device("cuda:0")
lib = _lib.load()
T, D, Kc = 256, 1024, 256
wc = (torch.randn(D, Kc, device=dev) / 32).half(); ct = ops.transpose_f32_dual(torch.randn(D, Kc, device=dev) / 32, want32=False)[1]
one, zero = torch.ones(Kc, device=dev), torch.zeros(Kc, device=dev)
for B in (80, 148):
    xb = torch.randn(B * T, D, device=dev).half()
    for _ in range(2): ops.netvlad_pool_fwd(xb, B, T, wc, one, zero, ct)
    dbg = torch.zeros(B, 8, dtype=torch.int64, device=dev)
    lib.lpm_debug_set_pool_clock(C.c_void_p(dbg.data_ptr()))
    ops.netvlad_pool_fwd(xb, B, T, wc, one, zero, ct)
    torch.cuda.synchronize()
    lib.lpm_debug_set_pool_clock(None)
    d = dbg.cpu().double()
    dt = (d[:, 1:5] - d[:, 0:4]).mean(0)
    print(f"B={B}: phase1 {dt[0]:.0f} | softmax {dt[1]:.0f} | a_sum {dt[2]:.0f} | phase2+epilogue {dt[3]:.0f} | total {(d[:,4]-d[:,0]).mean():.0f} cycles")
