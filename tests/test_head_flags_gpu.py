"""Head variants behind the reference's flags: --gating_remove_diag (frame_level_models.py:2349-2352) and
--netvlad_relu (:2321-2327, 2339-2340), forward and backward, kernels vs torch-fp64 autograd and models vs the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.helpers import oracle_params, perturb, rel

BN_EPS = 1e-3


@pytest.mark.parametrize("B,H", [(80, 512), (7, 96)])
def test_gating_remove_diag_kernels(cuda, B, H):
    from learnablepoolingmethods_b200 import ops
    g = torch.Generator().manual_seed(1)
    act = torch.randn(B, H, generator=g, dtype=torch.float64, requires_grad=True)
    gates = torch.randn(B, H, generator=g, dtype=torch.float64, requires_grad=True)
    diag = (torch.randn(H, generator=g, dtype=torch.float64) * 0.5).requires_grad_(True)
    gamma = (1 + 0.2 * torch.rand(H, generator=g, dtype=torch.float64)).requires_grad_(True)
    beta = (0.1 * torch.randn(H, generator=g, dtype=torch.float64)).requires_grad_(True)
    dout = torch.randn(B, H, generator=g, dtype=torch.float64)
    v = gates - diag * act
    xh = (v - v.mean(0)) / torch.sqrt(v.var(0, unbiased=False) + BN_EPS)
    out = act * torch.sigmoid(xh * gamma + beta)
    out.backward(dout)
    f = lambda t: t.detach().float().to(cuda).contiguous()
    mm, mv = torch.zeros(H, device=cuda), torch.ones(H, device=cuda)
    o32, o16, st = ops.gating_fwd(f(act), f(gates), f(gamma), f(beta), mm, mv, training=True, wg_diag=f(diag), save=True)
    assert rel(o32, out.detach()) < 1e-5
    S = 64.0
    dact, dg16, dgam, dbet, ddiag = ops.gating_bwd(f(act), f(gates), f(gamma), f(beta), st, f(dout * S), 1.0 / S, wg_diag=f(diag))
    torch.cuda.synchronize()
    assert rel(dact / S, act.grad) < 1e-3            # direct path incl. -diag*dv (dv passes through fp16 only in dg16)
    assert rel(dg16.float() / S, gates.grad) < 2e-3
    assert rel(dgam, gamma.grad) < 1e-4 and rel(dbet, beta.grad) < 1e-4
    assert rel(ddiag, diag.grad) < 1e-4
    m = torch.zeros(H, H, device=cuda)
    ops.add_diag(m, ddiag, 2.0)
    assert torch.equal(torch.diagonal(m), 2.0 * ddiag) and float((m - torch.diag(torch.diagonal(m))).abs().max()) == 0.0


@pytest.mark.parametrize("B,H,training", [(80, 512, True), (5, 40, True), (9, 64, False)])
def test_hidden_bn_relu6_kernels(cuda, B, H, training):
    from learnablepoolingmethods_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = (3.0 * torch.randn(B, H, generator=g, dtype=torch.float64) + 1.0).requires_grad_(True)
    gamma = (2 + torch.rand(H, generator=g, dtype=torch.float64)).requires_grad_(True)     # spreads y over (-inf, 0), (0, 6), (6, inf)
    beta = (2 + torch.randn(H, generator=g, dtype=torch.float64)).requires_grad_(True)
    mm0, mv0 = torch.randn(H, generator=g, dtype=torch.float64), 1 + torch.rand(H, generator=g, dtype=torch.float64)
    dy = torch.randn(B, H, generator=g, dtype=torch.float64)
    if training:
        mean, var = x.mean(0), x.var(0, unbiased=False)
    else:
        mean, var = mm0, mv0
    y = torch.clamp((x - mean) / torch.sqrt(var + BN_EPS) * gamma + beta, 0.0, 6.0)
    f = lambda t: t.detach().float().to(cuda).contiguous()
    mm, mv = f(mm0), f(mv0)
    o32, o16, st = ops.hidden_bn_relu6_fwd(f(x), f(gamma), f(beta), mm, mv, training=training, save=True)
    assert float((o32.cpu().double() - y.detach()).abs().max()) < 1e-4
    frac = float(((y > 0) & (y < 6)).double().mean())
    assert 0.2 < frac < 0.9
    if training:
        assert rel(mm, 0.999 * mm0 + 0.001 * x.detach().mean(0)) < 1e-5
        assert rel(mv, 0.999 * mv0 + 0.001 * x.detach().var(0, unbiased=True)) < 1e-5     # Bessel-corrected (fused rank-2 path)
        y.backward(dy)
        S = 32.0
        dx, dgam, dbet = ops.hidden_bn_relu6_bwd(f(x), o32, f(dy * S), f(gamma), st, inv_scale=1.0 / S)
        torch.cuda.synchronize()
        # elements that sit within fp32 rounding of the relu6 kinks may flip: none at these sizes, but keep a small margin
        assert rel(dx / S, x.grad) < 1e-3 and rel(dgam, gamma.grad) < 1e-4 and rel(dbet, beta.grad) < 1e-4
    else:
        assert torch.equal(mm.cpu(), mm0.float()) and torch.equal(mv.cpu(), mv0.float())


@pytest.mark.parametrize("model", ["NetVladV1", "WillowModelReg"])
@pytest.mark.parametrize("flags", [dict(remove_diag=True), dict(netvlad_relu=True, gating=False), dict(netvlad_relu=True, remove_diag=True)])
def test_models_with_head_flags(cuda, model, flags):
    """Forward parity (inference + training) and gradients with the flag combinations; 16 videos so that the batch-norms
    over the batch (gating_bn, hidden1_bn) are reasonably conditioned (numerics)."""
    import numpy as np
    from learnablepoolingmethods_b200 import ops, variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from oracle import netvlad_oracle as O
    B, K, Hd, V, T = 16, 64, 64, 60, 64
    store = variables.VariableStore(cuda, seed=21)
    # det_reg 0.0 disables the regulariser (module_utils.py:67-69): the oracle loss below is the label loss alone
    eng = NetVladEngine(NetVladConfig(model=model, iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V, rgb_det_reg=0.0,
                                      audio_det_reg=0.0, **flags), store)
    perturb(store, seed=4)
    assert ("hidden1_bn/gamma" in store.vars) == bool(flags.get("netvlad_relu")) and \
           ("hidden1_biases" in store.vars) != bool(flags.get("netvlad_relu"))
    x, nf, labels = O.synthetic_batch(B, seed=20181007, vocab=V, video_scale=1.0)
    idx = O.sample_random_frame_indices(nf.numpy(), np.random.RandomState(1).rand(B, T).astype(np.float32))
    okw = dict(vocab_size=V, iterations=T, cluster_size=K, remove_diag=bool(flags.get("remove_diag")),
               gating=flags.get("gating", True), relu=bool(flags.get("netvlad_relu")))
    if model == "WillowModelReg":
        fn = lambda P, S, tr, ri=False: O.willow_model_reg(x, nf, P, S, is_training=tr, frame_index=idx, return_intermediates=ri, **okw)
        fkw = dict(frame_index=torch.from_numpy(idx))
    else:
        fn = lambda P, S, tr, ri=False: O.netvlad_v1(x, nf, P, S, is_training=tr, return_intermediates=ri, **okw)
        fkw = {}
    P, S = oracle_params(store)
    with torch.no_grad():
        ref_inf, inter = fn(P, {k: v.clone() for k, v in S.items()}, False, True)
    pred, ictx = eng.forward(x.to(cuda), nf.to(cuda), False, return_intermediates=True, **fkw)
    e_inf = float((pred.cpu() - ref_inf).abs().max())
    e_med = float((pred.cpu() - ref_inf).abs().median())
    e_h, e_g = rel(ictx["inter"]["hidden"], inter["hidden"]), rel(ictx["inter"]["gated"], inter["gated"])
    for p in P.values():
        p.requires_grad_(True)
    ref = fn(P, S, True)
    O.cross_entropy_loss(ref, labels).backward()
    pred, ctx = eng.forward(x.to(cuda), nf.to(cuda), True, save_for_backward=True, **fkw)
    e_tr = float((pred.cpu() - ref.detach()).abs().max())
    lab = labels.to(torch.uint8).to(cuda)
    grads = eng.backward(ctx, ops.xent_bwd(pred, lab, 1.0 / B))
    torch.cuda.synchronize()
    print(f"\n[{model} {flags}] infer: hidden {e_h:.2e} gated {e_g:.2e} pred max-abs {e_inf:.2e} median {e_med:.2e} | train pred max-abs {e_tr:.2e}")
    # random-init |hidden| ~ 30 feeds un-normalised sigmoid gates (minus diag*hidden with remove_diag): the maximum over
    # B*V predictions is ill-conditioned (numerics), so bound the activations and the bulk of the predictions
    # with --netvlad_relu `hidden` is relu6(BN(.)): the batch norm removes the common mode of the ~30-magnitude
    # projection, which amplifies its 5e-4 relative error a few times on the un-clamped units
    relu = bool(flags.get("netvlad_relu"))
    assert e_h < (5e-3 if relu else 1e-3) and e_g < 2e-2 and e_med < 2e-3 and e_inf < 5e-2 and e_tr < 1e-1
    bad = []
    gmax = max(float(p.grad.norm()) for p in P.values() if p.grad is not None)
    for name in sorted(P):
        if P[name].grad is None:
            continue
        assert name in grads, f"missing gradient for {name}"
        e, gn = rel(grads[name].reshape(P[name].shape), P[name].grad), float(P[name].grad.norm())
        print(f"  {name:60s} rel-L2 {e:.2e}  |g| {gn:.2e}")
        # NetVladV1's descriptor is a LayerNorm output of norm sqrt(K*D): |hidden| ~ 30 with a small spread over the
        # batch, so the batch norms over the batch (hidden1_bn, gating_bn) amplify the fp16 forward error ~10x into every
        # upstream gradient (same conditioning as test_backward_gpu's gating case; the kernels are exact in isolation
        # above).  WillowModelReg's unit-norm descriptor keeps the tight bound.  Vanishing gradients are skipped.
        if gn <= 1e-4 * gmax:
            continue
        # ... and the small ones (|g| < 1 % of the largest) sit on top of that amplified noise: gross-error guard only
        if not e < ((3e-1 if gn > 1e-2 * gmax else 1.0) if model == "NetVladV1" else 1e-1):
            bad.append((name, e))
    assert not bad, bad
    for k in (("hidden1_bn/moving_variance",) if flags.get("netvlad_relu") else ()) + ("input_bn/moving_mean",):
        assert rel(store.vars[k], S[k]) < 2e-3, k


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("remove_diag", [False, True])
def test_fused_k3_head_matches_four_launch_head(cuda, training, remove_diag):
    """NetVladConfig.fused_gating: the hidden projection with the context-gating epilogue inside the split-K GEMM kernel
    (lpm_gemm_splitk_gated_fwd) against the four-launch head on the same model -- predictions, saved activations, moving
    statistics and (training) every gradient."""
    from learnablepoolingmethods_b200 import ops, variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from oracle import netvlad_oracle as O
    B, K, Hd, V, T = 12, 64, 128, 300, 128
    x, nf, labels = O.synthetic_batch(B, seed=11, vocab=V, video_scale=1.0)
    lab = labels.to(torch.uint8).to(cuda)
    out = {}
    for mode in ("all", "off"):
        store = variables.VariableStore(cuda, seed=7)
        eng = NetVladEngine(NetVladConfig(model="NetVladV1", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V,
                                          fused_gating=mode, remove_diag=remove_diag), store)
        perturb(store, seed=3)
        pred, ctx = eng.forward(x.to(cuda), nf.to(cuda), training, save_for_backward=training, return_intermediates=True)
        res = {"pred": pred.clone(), "hidden": ctx["inter"]["hidden"].clone(), "gated": ctx["inter"]["gated"].clone(),
               "mm": store.vars["gating_bn/moving_mean"].clone(), "mv": store.vars["gating_bn/moving_variance"].clone()}
        if training:
            loss, _ = ops.xent_fwd(pred, lab)
            grads = eng.backward(ctx, ops.xent_bwd(pred, lab, 1.0 / B))
            res["grads"] = {k: v.clone() for k, v in grads.items()}
        torch.cuda.synchronize()
        out[mode] = res
    a, b = out["all"], out["off"]
    assert rel(a["hidden"], b["hidden"]) < 1e-6 and rel(a["gated"], b["gated"]) < 2e-5
    assert float((a["pred"] - b["pred"]).abs().max()) < 2e-5
    assert rel(a["mm"], b["mm"]) < 1e-5 and rel(a["mv"], b["mv"]) < 1e-5
    if training:
        assert set(a["grads"]) == set(b["grads"])
        for k in a["grads"]:
            assert rel(a["grads"][k], b["grads"][k]) < 1e-2, k      # fp16 activation gradients downstream of a 1e-5 change through a 12-row gating_bn
