"""Gradients of the CUDA NetVladV1 path vs autograd through the CPU oracle (identical weights / inputs).
fp16 activation gradients with loss scaling: per-parameter rel-L2 <= 3e-2 (most are ~1e-3..1e-2)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("gating,tol", [(False, 2e-2), (True, 3e-1)])
@pytest.mark.parametrize("B,K,Hd,V,T", [(4, 64, 64, 100, 256), (3, 128, 64, 200, 128), (3, 512, 128, 100, 256)])
def test_netvlad_v1_gradients(cuda, B, K, Hd, V, T, gating, tol):
    """gating=False isolates the kernels (tight bound).  With context gating the batch-statistics BN over a
    batch of 3-4 videos has |mean|/std ~ 10 on this data and amplifies the fp16 forward error ~10x into
    dLoss/dpred, so that case only guards against gross errors (numerics)."""
    from learnablepoolingmethods_b200 import ops, variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from oracle import netvlad_oracle as O
    from tests.helpers import oracle_params as _oracle_params, perturb as _perturb
    store = variables.VariableStore(cuda, seed=7)
    cfg = NetVladConfig(model="NetVladV1", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V, gating=gating)
    eng = NetVladEngine(cfg, store)
    _perturb(store, seed=3)
    x, nf, labels = O.synthetic_batch(B, seed=20181001, vocab=V)
    P, S = _oracle_params(store)
    for p in P.values():
        p.requires_grad_(True)
    pred_ref = O.netvlad_v1(x, nf, P, S, vocab_size=V, iterations=T, cluster_size=K, is_training=True, gating=gating)
    loss_ref = O.cross_entropy_loss(pred_ref, labels)
    loss_ref.backward()

    pred, ctx = eng.forward(x.to(cuda), nf.to(cuda), True, save_for_backward=True)
    lab = labels.to(torch.uint8).to(cuda)
    loss, _ = ops.xent_fwd(pred, lab)
    dpred = ops.xent_bwd(pred, lab, 1.0 / B)
    grads = eng.backward(ctx, dpred)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_ref)) / float(loss_ref) < 2e-2
    bad = []
    print()
    gmax = max(float(p.grad.norm()) for p in P.values() if p.grad is not None)
    if K > 256:
        tol *= 2      # wide config: 1024 gates / 512-way attention at random init (conditioning)
    for name in sorted(P):
        if not gating and name.startswith("gating"):
            continue
        assert name in grads, f"missing gradient for {name}"
        e = rel(grads[name].reshape(P[name].shape), P[name].grad)
        gn = float(P[name].grad.norm())
        print(f"  {name:60s} rel-L2 {e:.2e}  |g| {gn:.2e}")
        # vanishing gradients (the 512-way near-uniform attention leaves |dWq|, |dWk| ~ 1e-9 of the largest tensor)
        # sit at the fp16 activation-gradient noise floor: guard against gross errors only
        if not (e < (tol if gn > 1e-7 * gmax else 10 * tol)):
            bad.append((name, e))
    assert not bad, bad


@pytest.mark.parametrize("B,K,Hd,V,T", [(4, 64, 64, 100, 256), (3, 32, 64, 120, 128)])
def test_netvlad_v2_gradients(cuda, B, K, Hd, V, T):
    """NetVladV2 backward (BN-on-logits attention, four batch norms, dropout 0.9 with injected masks) vs oracle
    autograd; gating off isolates the kernels from the batch-of-4 gating-BN conditioning."""
    from learnablepoolingmethods_b200 import ops, variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from oracle import netvlad_oracle as O
    from tests.helpers import oracle_params as _oracle_params, perturb as _perturb
    store = variables.VariableStore(cuda, seed=7)
    cfg = NetVladConfig(model="NetVladV2", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V, gating=False)
    eng = NetVladEngine(cfg, store)
    _perturb(store, seed=3)
    x, nf, labels = O.synthetic_batch(B, seed=20181001, vocab=V)
    P, S = _oracle_params(store)
    for p in P.values():
        p.requires_grad_(True)
    g = torch.Generator().manual_seed(9)
    masks = {"video": (torch.rand(B, T, 1024, generator=g) >= 0.9).float(), "audio": (torch.rand(B, T, 128, generator=g) >= 0.9).float()}
    pred_ref = O.netvlad_v2(x, nf, P, S, vocab_size=V, iterations=T, cluster_size=K, is_training=True, dropout_masks=masks,
                            gating=False)
    # head without gating for the oracle as well
    loss_ref = O.cross_entropy_loss(pred_ref, labels)
    loss_ref.backward()
    dm = {k: mk.reshape(B * T, -1).half().to(cuda) for k, mk in masks.items()}
    pred, ctx = eng.forward(x.to(cuda), nf.to(cuda), True, save_for_backward=True, dropout_masks=dm)
    lab = labels.to(torch.uint8).to(cuda)
    loss, _ = ops.xent_fwd(pred, lab)
    grads = eng.backward(ctx, ops.xent_bwd(pred, lab, 1.0 / B))
    torch.cuda.synchronize()
    bad = []
    print()
    for name in sorted(P):
        if name.startswith("gating"):
            continue
        assert name in grads, f"missing gradient for {name}"
        e = rel(grads[name].reshape(P[name].shape), P[name].grad)
        print(f"  {name:66s} rel-L2 {e:.2e}  |g| {float(P[name].grad.norm()):.2e}")
        if not (e < 1e-1):
            bad.append((name, e))
    # The two new backward kernels are exact in isolation (tests/test_kernels_gpu.py: BN-logits attention backward
    # 3e-4, BN-over-rows backward < 2e-3).  At model level every batch norm removes the common mode of its incoming
    # gradient, which amplifies the ~1e-3 upstream fp16 error to 2-5 % on the encoder parameters of this random-init
    # configuration (head / cluster_centers stay at 1e-3).
    assert not bad, bad


def test_netvlad_v1_d5_raw_reshape(cuda):
    """SURVEY defect D5 switch: the literal reading of frame_level_models.py:2290-2292 feeds the d-major flattened
    descriptor [B, D*K], re-interpreted as [B, K, D], to the attention block.  Forward and gradients vs the oracle's
    `d5_raw_reshape=True`."""
    from learnablepoolingmethods_b200 import ops, variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from oracle import netvlad_oracle as O
    from tests.helpers import oracle_params as _oracle_params, perturb as _perturb
    B, K, Hd, V, T = 3, 64, 64, 100, 128
    store = variables.VariableStore(cuda, seed=11)
    cfg = NetVladConfig(model="NetVladV1", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V, gating=False,
                        d5_raw_reshape=True)
    eng = NetVladEngine(cfg, store)
    _perturb(store, seed=5)
    x, nf, labels = O.synthetic_batch(B, seed=20181002, vocab=V)
    P, S = _oracle_params(store)
    # inference with context gating on (the conditioning of the other parity configurations); same variables
    eng_g = NetVladEngine(NetVladConfig(model="NetVladV1", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V,
                                        d5_raw_reshape=True), store)
    with torch.no_grad():
        ref_inf, inter = O.netvlad_v1(x, nf, P, {k: v.clone() for k, v in S.items()}, vocab_size=V, iterations=T, cluster_size=K,
                                      is_training=False, d5_raw_reshape=True, return_intermediates=True)
        other = O.netvlad_v1(x, nf, P, {k: v.clone() for k, v in S.items()}, vocab_size=V, iterations=T, cluster_size=K,
                             is_training=False)
    pred, ctx = eng_g.forward(x.to(cuda), nf.to(cuda), False, return_intermediates=True)
    e_att = rel(ctx["inter"]["att_video"], inter["att_video"])
    e_h = rel(ctx["inter"]["hidden"], inter["hidden"])
    e_p = float((pred.cpu() - ref_inf).abs().max())
    e_med = float((pred.cpu() - ref_inf).abs().median())
    print(f"\n[d5 raw reshape] att rel-L2 {e_att:.2e}, hidden {e_h:.2e}, pred max-abs {e_p:.2e} median {e_med:.2e} "
          f"(transpose reading differs by {float((other - ref_inf).abs().max()):.2e})")
    # random-init weights with this layout put single predictions on the steep part of the sigmoid gates (DESIGN.md,
    # numerics): bound the activations and the bulk of the predictions, guard the maximum against gross errors
    assert e_att < 3e-3 and e_h < 1e-3 and e_med < 1e-3 and e_p < 1e-1
    assert float((other - ref_inf).abs().max()) > 5 * e_p         # the switch really selects a different computation
    for p in P.values():
        p.requires_grad_(True)
    pred_ref = O.netvlad_v1(x, nf, P, S, vocab_size=V, iterations=T, cluster_size=K, is_training=True, gating=False,
                            d5_raw_reshape=True)
    O.cross_entropy_loss(pred_ref, labels).backward()
    pred, ctx = eng.forward(x.to(cuda), nf.to(cuda), True, save_for_backward=True)
    lab = labels.to(torch.uint8).to(cuda)
    grads = eng.backward(ctx, ops.xent_bwd(pred, lab, 1.0 / B))
    torch.cuda.synchronize()
    gmax = max(float(p.grad.norm()) for p in P.values() if p.grad is not None)
    bad = []
    for name in sorted(P):
        if name.startswith("gating"):
            continue
        e = rel(grads[name].reshape(P[name].shape), P[name].grad)
        gn = float(P[name].grad.norm())
        # 5e-2: the raw reinterpretation hands the audio block 16 rows that mix clusters and features, and its small
        # FFN gradients sit at 3e-2 of fp16 activation-gradient noise (a layout bug would show as O(1) errors)
        if not (e < (5e-2 if gn > 1e-7 * gmax else 2e-1)):
            bad.append((name, e, gn))
    assert not bad, bad
