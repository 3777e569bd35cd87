"""SURVEY 8f row 4 on the GPU: WillowModelReg (baseline NetVLAD + orthogonal regulariser, random frame sampling),
the LightVLAD / NetVladOrthoReg / NetVladAttenCluster / TransformerEncoder(Mod) module boundaries -- CUDA path vs the
CPU oracle on identical weights and inputs.  Tolerances as in test_parity_gpu / test_backward_gpu."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.helpers import oracle_params, perturb, rel


def _draws(B, T, seed=5):
    return np.random.RandomState(seed).rand(B, T).astype(np.float32)


@pytest.mark.parametrize("mode", [0, 1])
def test_random_frame_index_matches_reference_rule(cuda, mode):
    """int32(u * float32(nf)) / SampleRandomSequence arithmetic, bit-exact (model_utils.py:26-73), incl. nf = 1, 300."""
    from learnablepoolingmethods_b200 import ops
    from oracle import netvlad_oracle as O
    B, T = 64, 256
    nf = np.random.RandomState(1).randint(1, 301, size=B).astype(np.int32)
    nf[:4] = (1, 300, 2, 255)
    u = _draws(B, T)
    u[0, 0], u[1, 1] = 0.0, np.float32(1.0) - np.float32(2.0 ** -24)
    if mode == 0:
        want = O.sample_random_frame_indices(nf, u)
        got = ops.random_frame_index(torch.from_numpy(nf).to(cuda), T, 300, mode=0, uniform=torch.from_numpy(u))
    else:
        want = O.sample_random_sequence_indices(nf, T, u[:, 0])
        got = ops.random_frame_index(torch.from_numpy(nf).to(cuda), T, 300, mode=1, uniform=torch.from_numpy(u[:, 0].copy()))
    assert np.array_equal(got.cpu().numpy(), np.clip(want, 0, 299))
    # the built-in generator: indices in range, both halves of the range used, different seeds differ
    a = ops.random_frame_index(torch.from_numpy(nf).to(cuda), T, 300, mode=mode, seed=1).cpu().numpy()
    b = ops.random_frame_index(torch.from_numpy(nf).to(cuda), T, 300, mode=mode, seed=2).cpu().numpy()
    assert (a >= 0).all() and (a < nf[:, None]).all() and not np.array_equal(a, b)
    if mode == 0:
        frac = a[nf >= 100] / nf[nf >= 100, None]
        assert 0.45 < frac.mean() < 0.55 and frac.max() > 0.95 and frac.min() < 0.05


@pytest.mark.parametrize("D,K,scale", [(1024, 256, 1e-4), (128, 64, 1e-4), (96, 24, 0.5)])
def test_ortho_regulariser_value_and_gradient(cuda, D, K, scale):
    from learnablepoolingmethods_b200 import ops
    from oracle import netvlad_oracle as O
    g = torch.Generator().manual_seed(3)
    w = (torch.randn(D, K, generator=g) / D ** 0.5).requires_grad_(True)
    ref = O.orthogonal_regularizer(w.double(), scale)
    ref.backward()
    base = torch.randn(D, K, generator=g)
    dw = base.clone().to(cuda)
    val = ops.ortho_reg(w.detach().to(cuda), scale, dw=dw, grad_scale=0.5, accumulate=True)
    torch.cuda.synchronize()
    assert abs(float(val) - float(ref)) <= 2e-5 * abs(float(ref))
    # |.| is non-differentiable at 0: an fp32-vs-fp64 sign flip of a ~1e-8 Gram entry moves a few rows by 2*scale*N
    e = rel(dw.cpu() - base, 0.5 * w.grad)
    assert e < 2e-3, e
    dw2 = torch.empty(D, K, device=cuda)
    ops.ortho_reg(w.detach().to(cuda), scale, dw=dw2, accumulate=False, want_value=False)
    assert rel(dw2, w.grad) < 2e-3


@pytest.mark.parametrize("is_training", [False, True])
@pytest.mark.parametrize("codes", [False, True])
@pytest.mark.parametrize("B,K,Hd,V,T", [(4, 64, 64, 100, 256), (3, 256, 128, 3862, 256)])
def test_willow_forward_parity(cuda, B, K, Hd, V, T, codes, is_training):
    from learnablepoolingmethods_b200 import variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from oracle import netvlad_oracle as O
    store = variables.VariableStore(cuda, seed=1810)
    eng = NetVladEngine(NetVladConfig(model="WillowModelReg", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V), store)
    perturb(store)
    assert "video_VLAD/cluster_weightsnetvlad_rgb_scope" in store.vars and "audio_VLAD/cluster_weightsnetvlad_audio_scope" in store.vars
    assert tuple(store.vars["video_VLAD/cluster_weights2"].shape) == (1024, K)
    x, nf, _, q = O.synthetic_batch(B, seed=20181000, vocab=V, return_codes=True)
    if codes:   # the reader's codes: frames are dequantised + normalised inside the gather; padded frames are never drawn
        x = O.l2_normalize(O.dequantize(q.float()), 2)
    idx = O.sample_random_frame_indices(nf.numpy(), _draws(B, T))
    P, S = oracle_params(store)
    with torch.no_grad():
        ref, inter = O.willow_model_reg(x, nf, P, S, vocab_size=V, iterations=T, cluster_size=K, is_training=is_training,
                                        frame_index=idx, return_intermediates=True)
    xin = q.to(cuda) if codes else x.to(cuda)
    pred, ctx = eng.forward(xin, nf.to(cuda), is_training, return_intermediates=True, frame_index=torch.from_numpy(idx))
    torch.cuda.synchronize()
    gi = ctx["inter"]
    e_v, e_a = rel(gi["vlad_video"], inter["vlad_video"]), rel(gi["vlad_audio"], inter["vlad_audio"])
    e_h, e_p = rel(gi["hidden"], inter["hidden"]), float((pred.cpu() - ref).abs().max())
    print(f"\n[Willow B={B} K={K} codes={codes} train={is_training}] vlad rgb {e_v:.2e} audio {e_a:.2e} | hidden {e_h:.2e} | pred {e_p:.2e}")
    assert e_v < 1e-3 and e_a < 1e-3 and e_h < 2e-3
    assert e_p < (5e-3 if not is_training else 1e-1)
    want_reg = float(O.willow_regularization(P))
    assert abs(float(eng.regularization_loss()) - want_reg) < 1e-4 * want_reg
    if is_training:
        for k in ("input_bn/moving_variance", "video_VLAD/cluster_bn/moving_mean", "gating_bn/moving_variance"):
            assert rel(store.vars[k], S[k]) < 2e-3, k


@pytest.mark.parametrize("B,K,Hd,V,T", [(4, 64, 64, 100, 256), (3, 128, 64, 200, 128)])
def test_willow_gradients(cuda, B, K, Hd, V, T):
    """Backward incl. the regulariser's gradient on cluster_weights2 (scaled up so that it matters next to the label
    loss); gating off isolates the kernels (see test_backward_gpu)."""
    from learnablepoolingmethods_b200 import ops, variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from oracle import netvlad_oracle as O
    store = variables.VariableStore(cuda, seed=7)
    cfg = NetVladConfig(model="WillowModelReg", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V, gating=False,
                        rgb_det_reg=3e-3, audio_det_reg=2e-2)
    eng = NetVladEngine(cfg, store)
    perturb(store, seed=3)
    x, nf, labels = O.synthetic_batch(B, seed=20181001, vocab=V)
    idx = O.sample_random_frame_indices(nf.numpy(), _draws(B, T, 8))
    P, S = oracle_params(store)
    for p in P.values():
        p.requires_grad_(True)
    pred_ref = O.willow_model_reg(x, nf, P, S, vocab_size=V, iterations=T, cluster_size=K, is_training=True,
                                  frame_index=idx, gating=False)
    (O.cross_entropy_loss(pred_ref, labels) + 1.0 * O.willow_regularization(P, 3e-3, 2e-2)).backward()
    pred, ctx = eng.forward(x.to(cuda), nf.to(cuda), True, save_for_backward=True, frame_index=torch.from_numpy(idx))
    lab = labels.to(torch.uint8).to(cuda)
    grads = eng.backward(ctx, ops.xent_bwd(pred, lab, 1.0 / B))
    torch.cuda.synchronize()
    bad = []
    print()
    for name in sorted(P):
        if name.startswith("gating"):
            continue
        assert name in grads, f"missing gradient for {name}"
        e = rel(grads[name].reshape(P[name].shape), P[name].grad)
        print(f"  {name:60s} rel-L2 {e:.2e}  |g| {float(P[name].grad.norm()):.2e}")
        if not e < 2e-2:
            bad.append((name, e))
    assert not bad, bad


def test_willow_training_steps_match_oracle(cuda):
    from learnablepoolingmethods_b200 import variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from learnablepoolingmethods_b200.trainer import Trainer
    from oracle import netvlad_oracle as O
    B, K, Hd, V, T = 4, 64, 64, 100, 128
    store = variables.VariableStore(cuda, seed=11)
    eng = NetVladEngine(NetVladConfig(model="WillowModelReg", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V,
                                      gating=False), store)
    perturb(store, seed=5)
    P, S = oracle_params(store)
    P0 = {k: v.clone() for k, v in P.items()}
    for p in P.values():
        p.requires_grad_(True)
    tr = Trainer(eng, base_learning_rate=2e-4, batch_size=B)
    opt_state = {}
    for step in range(3):
        x, nf, labels = O.synthetic_batch(B, seed=100 + step, vocab=V)
        idx = O.sample_random_frame_indices(nf.numpy(), _draws(B, T, 20 + step))
        fn = lambda xx, Pp, Ss: O.willow_model_reg(xx[0], xx[1], Pp, Ss, vocab_size=V, iterations=T, cluster_size=K,
                                                   is_training=True, frame_index=idx, gating=False)
        losses, _ = O.train_step(fn, P, S, opt_state, [(x, nf)], [labels], step=step + 1, lr=2e-4,
                                 extra_reg=O.willow_regularization)
        loss = tr.train_step(x.to(cuda), nf.to(cuda), labels.to(torch.uint8).to(cuda), frame_index=torch.from_numpy(idx))
        print(f"[willow train step {step}] loss {float(loss):.4f} vs oracle {losses[0]:.4f}")
        assert abs(float(loss) - losses[0]) / losses[0] < (1e-3, 4e-3, 3e-2)[step]
    assert not tr.overflowed()
    for name, p in P.items():
        ours = store.vars[name].detach().cpu().double()
        du, dr = (ours - P0[name].double()).flatten(), (p.detach().double() - P0[name].double()).flatten()
        if float(du.norm()) == 0.0 and float(dr.norm()) == 0.0:
            continue
        cos = float((du @ dr) / (du.norm() * dr.norm()).clamp_min(1e-30))
        assert rel(ours, p.detach()) < 5e-3 and cos > 0.97, (name, cos)
    # without injected indices the draw changes from call to call (tf.random_uniform) and stays within num_frames
    x, nf, labels = O.synthetic_batch(B, seed=1, vocab=V)
    p1, c1 = eng.forward(x.to(cuda), nf.to(cuda), True, save_for_backward=True)
    p2, c2 = eng.forward(x.to(cuda), nf.to(cuda), True, save_for_backward=True)
    i1, i2 = c1["frame_index"].cpu(), c2["frame_index"].cpu()
    assert not torch.equal(i1, i2) and bool((i1 < nf[:, None]).all()) and bool((i1 >= 0).all())


def test_create_model_registry_willow(cuda):
    """find_class_by_name('WillowModelReg') -> create_model(...) as train.py:187-190,280-284 call it."""
    from learnablepoolingmethods_b200 import frame_level_models, utils, variables, video_level_models
    variables.reset_default_store(cuda, seed=3)
    model = utils.find_class_by_name("WillowModelReg", [frame_level_models, video_level_models])()
    from oracle import netvlad_oracle as O
    x, nf, labels = O.synthetic_batch(2, seed=4, vocab=40)
    res = model.create_model(x.to(cuda), vocab_size=40, num_frames=nf.to(cuda), iterations=64, cluster_size=32,
                             hidden_size=64, labels=labels, is_training=False)
    p = res["predictions"]
    assert tuple(p.shape) == (2, 40) and bool(((p > 0) & (p < 1)).all())
    assert float(res["regularization_loss"]) > 0


# ------------------------------------------------------------------------------------------------------------------
# module boundaries: same constructors / forward contracts as the reference classes
# ------------------------------------------------------------------------------------------------------------------
def _load(store, P, S, prefix=""):
    store.load_state_dict({**{k: v.detach() for k, v in P.items()}, **S})


@pytest.mark.parametrize("is_training", [False, True])
def test_light_vlad_and_ortho_reg_modules(cuda, is_training):
    from learnablepoolingmethods_b200 import frame_level_models, variables, video_pooling_modules
    from oracle import netvlad_oracle as O
    B, T, D, K = 3, 128, 128, 32
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B * T, D, generator=g) * 0.5
    # LightVLAD
    store = variables.VariableStore(cuda, seed=5)
    with store.variable_scope("video_VLAD"):
        mod = frame_level_models.LightVLAD(D, T, K, True, is_training)
        out = mod.forward(x.to(cuda), store=store)           # creates the variables, as TF does
        perturb(store, seed=1)
        P, S = oracle_params(store)
        out = mod.forward(x.to(cuda), store=store)
    with torch.no_grad():
        ref = O.light_vlad_forward(x, P, S, "video_VLAD", T, True, is_training)
    assert tuple(out.shape) == (B, K * D) and rel(out, ref) < 1e-3
    # NetVladOrthoReg
    store = variables.VariableStore(cuda, seed=6)
    with store.variable_scope("audio_VLAD"):
        mod = video_pooling_modules.NetVladOrthoReg(D, T, K, True, is_training, 1e-3, "netvlad_audio_scope")
        mod.forward(x.to(cuda), store=store)
        perturb(store, seed=2)
        P, S = oracle_params(store)
        out = mod.forward(x.to(cuda), store=store)
        reg = mod.regularization_loss()
    assert "audio_VLAD/cluster_weightsnetvlad_audio_scope" in P
    with torch.no_grad():
        ref = O.netvlad_ortho_reg_forward(x, P, S, "audio_VLAD", T, True, is_training, scope_id="netvlad_audio_scope")
        want = float(O.orthogonal_regularizer(P["audio_VLAD/cluster_weights2"], 1e-3))
    assert rel(out, ref) < 1e-3 and abs(float(reg) - want) < 1e-4 * want
    with pytest.raises(ValueError):
        m2 = video_pooling_modules.NetVladOrthoReg(D, T, K, True, False, 1, None)
        with store.variable_scope("x"):
            m2.forward(x.to(cuda), store=store)
            m2.regularization_loss()


def test_transformer_encoder_module(cuda):
    """TransformerEncoder.forward([B, L, hidden]) as NetVladV1 builds it (frame_level_models.py:2282-2290)."""
    from learnablepoolingmethods_b200 import transformer_utils, variables
    from oracle import netvlad_oracle as O
    B, L, D, H = 3, 64, 128, 16
    store = variables.VariableStore(cuda, seed=5)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, L, D, generator=g) * 0.1
    enc = transformer_utils.TransformerEncoder(feature_size=D, hidden_size=D, num_heads=H, attention_dropout=0.1,
                                               ff_filter_size=4 * D, ff_relu_dropout=0.1, is_train=True, scope_id="encode2")
    with store.variable_scope("audio_attention"):
        enc.forward(x.to(cuda), store=store)
        perturb(store, seed=1)
        P, S = oracle_params(store)
        out = enc.forward(x.to(cuda), store=store)
    with torch.no_grad():
        ref = O.transformer_encoder(x, P, "audio_attention", H, "encode2")
    assert tuple(out.shape) == (B, L, D)
    assert rel(out, ref) < 3e-3


@pytest.mark.parametrize("is_training", [False, True])
def test_netvlad_atten_cluster_module(cuda, is_training):
    """NetVladAttenCluster.forward([(B*T), D]) -> [B, K*D] (video_pooling_modules.py:1617-1663), dropout mask injected."""
    from learnablepoolingmethods_b200 import variables, video_pooling_modules
    from oracle import netvlad_oracle as O
    B, T, D, K = 3, 128, 128, 32
    store = variables.VariableStore(cuda, seed=5)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B * T, D, generator=g) * 0.5
    mask = (torch.rand(B, T, D, generator=g) >= 0.9).float()
    mod = video_pooling_modules.NetVladAttenCluster(D, T, K, True, is_training)
    m16 = mask.reshape(B * T, D).half().to(cuda)
    with store.variable_scope("audio_VLAD"):
        mod.forward(x.to(cuda), store=store, dropout_mask=m16)
        perturb(store, seed=1)
        P, S = oracle_params(store)
        out = mod.forward(x.to(cuda), store=store, dropout_mask=m16)
    with torch.no_grad():
        ref = O.netvlad_atten_cluster_forward(x, P, S, "audio_VLAD", T, is_training, dropout_mask=mask)
    assert tuple(out.shape) == (B, K * D)
    assert rel(out, ref) < 8e-3


def test_l2_normalize_prelude(cuda):
    """train.py:262-264: per-frame l2_normalize over rgb|audio jointly; zero-padded frames stay zero; in place allowed."""
    from learnablepoolingmethods_b200 import model_utils, ops
    from oracle import netvlad_oracle as O
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 300, 1152, generator=g) * 3
    x[1, 100:] = 0
    x[2] = 0
    want = O.l2_normalize(x.double(), 2)
    got = model_utils.l2_normalize_frames(x.to(cuda))
    assert float((got.cpu().double() - want).abs().max()) < 1e-6 and float(got[2].abs().max()) == 0.0
    xi = x.to(cuda).clone()
    ops.l2_normalize_frames(xi, out=xi)
    assert torch.equal(xi, got)


def test_sample_random_mirrors(cuda):
    """model_utils.SampleRandomFrames / SampleRandomSequence with injected draws vs the oracle's index rules; utils.Dequantize."""
    from learnablepoolingmethods_b200 import model_utils, utils
    from oracle import netvlad_oracle as O
    B, S = 6, 40
    x, nf, _, q = O.synthetic_batch(B, seed=3, vocab=20, return_codes=True)
    u = _draws(B, S, 4)
    want = O.gather_frames(x, O.sample_random_frame_indices(nf.numpy(), u))
    got = model_utils.SampleRandomFrames(x.to(cuda), nf.to(cuda), S, uniform=torch.from_numpy(u))
    assert tuple(got.shape) == (B, S, 1152) and float((got.float().cpu() - want).abs().max()) < 1e-3
    want = O.gather_frames(x, O.sample_random_sequence_indices(nf.numpy(), S, u[:, 0]))
    got = model_utils.SampleRandomSequence(x.to(cuda), nf.to(cuda), S, uniform=torch.from_numpy(u[:, 0].copy()))
    assert float((got.float().cpu() - want).abs().max()) < 1e-3
    got = model_utils.SampleRandomFrames(q.to(cuda), nf.to(cuda), S, uniform=torch.from_numpy(u))       # uint8 codes
    full = O.l2_normalize(O.dequantize(q.float()), 2)
    want = O.gather_frames(full, O.sample_random_frame_indices(nf.numpy(), u))
    assert float((got.float().cpu() - want).abs().max()) < 1e-3
    assert torch.equal(utils.Dequantize(q.float()), O.dequantize(q.float()))
