"""Small invocations of the hand-rolled mbarrier / TMEM / cluster protocols for compute-sanitizer (racecheck, memcheck):
the fused pooling kernel (1-CTA and the 2-CTA K=512 cluster path), the tcgen05 GEMM (1-CTA persistent and 2-CTA
cta_group::2 pairs, split-K), the cluster layer-norm chain, the attention core and the tiled factored Adam."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "pool"):
    for (B, T, D, K) in ((2, 256, 128, 64), (2, 200, 128, 256), (1, 256, 128, 512)):
        xb = r(B * T, D).half()
        wc = (r(D, K) / 8).half()
        ct = ops.transpose_f32_dual(r(D, K) / 8, want32=False)[1]
        z, rs, a_sum, assign = ops.netvlad_pool_fwd(xb, B, T, wc, torch.ones(K, device=dev), torch.zeros(K, device=dev), ct, save_assign=True)
        torch.cuda.synchronize()
        print("pool", B, T, D, K, float(z.float().abs().mean()))
if which in ("all", "gemm"):
    for (M, N, K, kw) in ((256, 256, 256, {}), (4864, 1024, 128, {}), (80, 512, 4096, {"splits": 8})):
        a, b = r(M, K).half(), (r(K, N) * 0.05).half()
        o = ops.gemm(a, b, **kw)
        torch.cuda.synchronize()
        print("gemm", M, N, K, kw, float(o.float().abs().mean()))
if which in ("all", "k3"):
    # K3: split-K GEMM with the reduction + context-gating tail (two grid-wide barriers inside the kernel)
    B, Kd, H = 16, 2048, 128
    a, w = (r(B, Kd) * 0.05).half(), (r(Kd, H) * 0.05).half()
    p2 = ops.gemm((r(B, Kd) * 0.01).half(), w, splits=4)
    for training in (True, False):
        res = ops.gemm_splitk_gated(a, w, splits=16, bias=r(H), wg=r(H, H) / 11, gamma=torch.ones(H, device=dev), beta=torch.zeros(H, device=dev),
                                    moving_mean=torch.zeros(H, device=dev), moving_var=torch.ones(H, device=dev), training=training,
                                    save=True, parts2=p2)
        torch.cuda.synchronize()
        print("k3", training, float(res[2].abs().mean()))
if which in ("all", "attn_tc"):
    # tcgen05 / TMEM attention (lpm_attn_tc.cu): one sample, 8 heads of depth 16 at 256 positions = two CTAs per kernel;
    # forward opt-in kernel, both backward kernels
    from learnablepoolingmethods_b200._lib import load
    lib = load()
    B, L, Dm, H = 1, 256, 128, 8
    qkv = (r(B * L, 3 * Dm) * 0.5).half()
    do = (r(B * L, Dm) * 0.1).half()
    for mode in (4 | 2, 8 | 1):
        lib.lpm_debug_set_mha_tc_mode(mode)
        o, lse = ops.mha_core_fwd(qkv, B, L, Dm, H, scale=0.25, want_lse=True)
        d = ops.mha_core_bwd(qkv, o, do, lse, B, L, Dm, H, scale=0.25)
        torch.cuda.synchronize()
        print("attn_tc mode", mode, float(o.float().abs().mean()), float(d.float().abs().mean()))
    lib.lpm_debug_set_mha_tc_mode(2)
if which in ("all", "misc"):
    B, R, D = 2, 64, 128
    a, b = r(B, R, D).half(), r(B, R, D).half()
    out = torch.empty(B, R * D, dtype=torch.float16, device=dev)
    ops.layernorm_chain_fwd(a, b, B, R, D, torch.ones(D, device=dev), torch.zeros(D, device=dev), torch.ones(D, device=dev),
                            torch.zeros(D, device=dev), out=out, out_stride=R * D)
    qkv = (r(2 * 64, 3 * 128) * 0.5).half()
    o, lse = ops.mha_core_fwd(qkv, 2, 64, 128, 8, scale=0.25, want_lse=True)
    ops.mha_core_bwd(qkv, o, (r(2 * 64, 128) * 0.1).half(), lse, 2, 64, 128, 8, scale=0.25)
    Rk, Kd, N = 16, 256, 64
    w, m, v = r(Kd, N), r(Kd, N) * 1e-3, torch.rand(Kd, N, device=dev) * 1e-5
    ops.rank_adam_step(r(Rk, Kd).half(), r(Rk, N).half(), 1.0, torch.ones(1, device=dev), torch.zeros(1, dtype=torch.int32, device=dev),
                       w, m, v, torch.zeros(Kd, N, dtype=torch.float16, device=dev), lr_dev=torch.full((1,), 1e-3, device=dev), tiled=True,
                       workspace=torch.empty(ops.rank_adam_workspace_bytes(Rk, N), dtype=torch.uint8, device=dev))
    torch.cuda.synchronize()
    print("misc ok")
