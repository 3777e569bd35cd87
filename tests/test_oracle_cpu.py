"""CPU tests of the oracle: golden self-pins, closed-form and brute-force cross-checks of the restated
reference lines, and (when /root/reference is present) the reference's own numpy metric code."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import netvlad_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _run(model, is_training, dtype):
    sys.path.insert(0, os.path.dirname(os.path.dirname(__file__)))
    from oracle import make_golden
    return make_golden.run(model, is_training, dtype)


@pytest.mark.parametrize("model", ["NetVladV1", "NetVladV2", "WillowModelReg"])
@pytest.mark.parametrize("is_training", [False, True])
def test_oracle_matches_golden(model, is_training):
    gold = np.load(os.path.join(GOLD, f"{model}_{'train' if is_training else 'infer'}_tiny_f64.npz"))
    out = _run(model, is_training, torch.float64)
    for k in gold.files:
        np.testing.assert_allclose(out[k], gold[k], rtol=1e-9, atol=1e-12, err_msg=k)
    out32 = _run(model, is_training, torch.float32)
    assert np.abs(out32["pred"] - gold["pred"]).max() < 1e-4
    assert abs(float(out32["loss"]) - float(gold["loss"])) / float(gold["loss"]) < 1e-4


def test_sample_indices_closed_form_and_golden():
    """model_utils.py:101-122: int32(fl32(i/S) * nf) == floor(i*nf/S); always < nf; nf=0 -> frame 0."""
    gold = np.load(os.path.join(GOLD, "sample_indices.npz"))
    nf = np.arange(0, 301)
    for T in (30, 64, 256, 300):
        idx = O.sample_uniform_indices(nf, T)
        np.testing.assert_array_equal(idx, gold[f"T{T}"])
        exact = (np.arange(T)[None, :].astype(np.int64) * nf[:, None]) // T
        if T in (64, 256):                      # i/T exact in fp32 => bit-exact floor rule
            np.testing.assert_array_equal(idx, exact)
        assert np.abs(idx - exact).max() <= 1   # non power-of-two T: fp32 rounding may move a boundary index
        assert (idx[1:] < nf[1:, None]).all() and (idx[0] == 0).all()
        assert (np.diff(idx, axis=1) >= 0).all()


def test_netvlad_forward_against_bruteforce_loops():
    """frame_level_models.py:2775-2822 restated with explicit loops (independent of the matmul form)."""
    g = torch.Generator().manual_seed(0)
    B, T, D, K = 2, 5, 6, 3
    x = torch.randn(B * T, D, generator=g, dtype=torch.float64)
    P = {"s/cluster_weights": torch.randn(D, K, generator=g, dtype=torch.float64),
         "s/cluster_biases": torch.randn(K, generator=g, dtype=torch.float64),
         "s/cluster_weights2": torch.randn(1, D, K, generator=g, dtype=torch.float64)}
    out = O.netvlad_forward(x, P, None, "s", T, False, False)
    xr = x.reshape(B, T, D)
    ref = torch.zeros(B, D, K, dtype=torch.float64)
    for b in range(B):
        for t in range(T):
            s = xr[b, t] @ P["s/cluster_weights"] + P["s/cluster_biases"]
            a = torch.exp(s - s.max()); a = a / a.sum()
            for k in range(K):
                ref[b, :, k] += a[k] * (xr[b, t] - P["s/cluster_weights2"][0, :, k])
    ref = ref / ref.norm(dim=1, keepdim=True)
    ref = ref.reshape(B, -1)
    ref = ref / ref.norm(dim=1, keepdim=True)
    torch.testing.assert_close(out, ref, rtol=1e-10, atol=1e-12)


def test_v2_aggregation_matches_d6_broadcast_form():
    """video_pooling_modules.py:1646-1652 with decision D6: sum_n a[b,n,c] (x[b,n,f] - c[f,c])."""
    g = torch.Generator().manual_seed(1)
    B, T, D, K = 2, 4, 5, 3
    xs = torch.randn(B, T, D, generator=g, dtype=torch.float64)
    A = torch.randn(B, T, K, generator=g, dtype=torch.float64)
    C = torch.randn(D, K, generator=g, dtype=torch.float64)
    residuals = xs.unsqueeze(3) - C                     # B x N x F x C
    ref = (residuals * A.unsqueeze(2)).sum(dim=1)       # expand_dims(axis=2) -> B x N x 1 x C
    got = torch.matmul(xs.transpose(1, 2), A) - A.sum(dim=1, keepdim=True) * C
    torch.testing.assert_close(got, ref, rtol=1e-12, atol=1e-12)


def test_tf_library_semantics():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(7, 5, generator=g, dtype=torch.float64)
    # slim.batch_norm: eps 1e-3, biased batch variance, Bessel-corrected moving variance (fused rank-2 path)
    P = {"bn/gamma": torch.full((5,), 2.0, dtype=torch.float64), "bn/beta": torch.full((5,), 0.5, dtype=torch.float64)}
    S = {"bn/moving_mean": torch.zeros(5, dtype=torch.float64), "bn/moving_variance": torch.ones(5, dtype=torch.float64)}
    y = O.batch_norm(x, P, S, "bn", True)
    ref = (x - x.mean(0)) / torch.sqrt(x.var(0, unbiased=False) + 1e-3) * 2.0 + 0.5
    torch.testing.assert_close(y, ref)
    torch.testing.assert_close(S["bn/moving_variance"], 0.999 * torch.ones(5, dtype=torch.float64) + 0.001 * x.var(0, unbiased=True))
    torch.testing.assert_close(S["bn/moving_mean"], 0.001 * x.mean(0))
    # rank-3 input: non-fused path feeds the biased variance
    x3 = torch.randn(2, 3, 5, generator=g, dtype=torch.float64)
    S3 = {"bn/moving_mean": torch.zeros(5, dtype=torch.float64), "bn/moving_variance": torch.ones(5, dtype=torch.float64)}
    O.batch_norm(x3, P, S3, "bn", True)
    torch.testing.assert_close(S3["bn/moving_variance"], 0.999 + 0.001 * x3.reshape(-1, 5).var(0, unbiased=False))
    # layer_norm: joint over all non-batch axes, eps 1e-12, params on the last axis
    z = torch.randn(2, 3, 5, generator=g, dtype=torch.float64)
    PL = {"ln/gamma": torch.ones(5, dtype=torch.float64), "ln/beta": torch.zeros(5, dtype=torch.float64)}
    out = O.layer_norm_joint(z, PL, "ln")
    flat = out.reshape(2, -1)
    assert flat.mean(1).abs().max() < 1e-12 and (flat.var(1, unbiased=False) - 1).abs().max() < 1e-9
    # l2_normalize of a zero vector stays zero (eps inside the max)
    assert O.l2_normalize(torch.zeros(1, 4), 1).abs().sum() == 0
    # clip_by_norm / exponential_decay staircase
    gvec = torch.tensor([3.0, 4.0])
    torch.testing.assert_close(O.clip_by_norm(gvec, 1.0), gvec / 5.0)
    torch.testing.assert_close(O.clip_by_norm(gvec, 10.0), gvec)
    assert O.learning_rate(2e-4, 0.85, 4e6, 49999, 80, 1) == 2e-4
    assert abs(O.learning_rate(2e-4, 0.85, 4e6, 50000, 80, 1) - 2e-4 * 0.85) < 1e-15


def test_adam_matches_tf_formula():
    p, gr = torch.tensor([1.0, -2.0]), torch.tensor([0.1, -0.3])
    m, v = torch.zeros(2), torch.zeros(2)
    O.adam_step(p, gr, m, v, step=1, lr=1e-3)
    # first TF Adam step moves every weight by ~lr * sign(g)
    torch.testing.assert_close(p, torch.tensor([1.0 - 1e-3, -2.0 + 1e-3]), rtol=0, atol=2e-8)


@pytest.mark.skipif(not os.path.exists("/root/reference/eval_util.py"), reason="reference tree not mounted")
def test_topk_and_gap_with_reference_eval_util():
    """The reference's own numpy metric code (eval_util.py:92-135) is importable: use it unmodified to pin the
    top-20 label-set comparison used by the parity report."""
    sys.path.insert(0, "/root/reference")
    import eval_util
    rng = np.random.RandomState(0)
    pred = rng.rand(6, 100).astype(np.float32)
    labels = rng.rand(6, 100) < 0.05
    for b in range(6):
        trip = eval_util.top_k_triplets(pred[b], labels[b], 20)
        ref_set = {int(t[0]) for t in trip}
        mine = set(np.argsort(-pred[b])[:20].tolist())
        assert ref_set == mine
    gap = eval_util.calculate_gap(pred, labels)
    assert 0.0 <= gap <= 1.0


def test_eval_oracle_matches_reference_golden():
    """tests/golden/eval_golden.npz holds the outputs of the REFERENCE's eval_util.py / average_precision_calculator.py
    (oracle/make_eval_golden.py, run where /root/reference is mounted): hit@1, PERR, GAP and the top-20 class sets
    of seeded batches.  The numpy restatement in oracle/eval_oracle.py must reproduce them."""
    import numpy as np
    from oracle import eval_oracle as E
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_golden.npz"))
    seeds = sorted(int(k[4:]) for k in G.files if k.startswith("case"))
    assert len(seeds) >= 4
    for seed in seeds:
        meta = G[f"case{seed}"]
        pred, labels = E.synthetic_eval_batch(int(meta[0]), int(meta[1]), int(meta[2]), tuple(int(z) for z in meta[3:]))
        assert abs(E.hit_at_one(pred, labels) - float(G[f"hit{seed}"])) < 1e-12
        assert abs(E.perr(pred, labels) - float(G[f"perr{seed}"])) < 1e-12
        assert abs(E.gap(pred, labels) - float(G[f"gap{seed}"])) < 1e-9
        assert E.top_k_sets(pred, 20) == [sorted(r.tolist()) for r in G[f"topk{seed}"]]


@pytest.mark.skipif(not os.path.exists("/root/reference/eval_util.py"), reason="reference tree not mounted")
def test_eval_golden_is_current():
    """The committed golden file equals what the reference's module returns today (guards the generator script)."""
    import numpy as np
    sys.path.insert(0, "/root/reference")
    import eval_util
    from oracle import eval_oracle as E
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_golden.npz"))
    pred, labels = E.synthetic_eval_batch(102, 7, 50, (3,))
    assert abs(eval_util.calculate_gap(pred, labels) - float(G["gap102"])) < 1e-12
    assert abs(eval_util.calculate_precision_at_equal_recall_rate(pred, labels) - float(G["perr102"])) < 1e-12


# ------------------------------------------------------------------------------------------------------------------
# baseline NetVLAD (WillowModelReg / NetVladOrthoReg / LightVLAD), random sampling, regulariser
# ------------------------------------------------------------------------------------------------------------------
def test_random_sampling_rules():
    """model_utils.py:54-73: int32(u*nf) in [0, nf); :26-51: a contiguous window clipped to nf-1."""
    rng = np.random.RandomState(0)
    nf = np.concatenate([[1, 2, 300, 299], rng.randint(1, 301, size=60)]).astype(np.int32)
    u = rng.rand(len(nf), 256).astype(np.float32)
    u[0, 0], u[2, 0] = 0.0, np.float32(1.0) - np.float32(2.0 ** -24)
    idx = O.sample_random_frame_indices(nf, u)
    assert idx.dtype == np.int32 and (idx >= 0).all() and (idx < nf[:, None]).all()
    np.testing.assert_array_equal(idx, np.floor(u.astype(np.float64) * nf[:, None]).clip(max=nf[:, None] - 1).astype(np.int32))
    seq = O.sample_random_sequence_indices(nf, 30, u[:, 0])
    assert (seq >= 0).all() and (seq < nf[:, None]).all()
    d = np.diff(seq, axis=1)
    assert ((d == 1) | (d == 0)).all()                      # consecutive frames, then stuck at the last frame
    long = nf >= 30
    assert (d[long] == 1).all() and (seq[long, -1] <= nf[long] - 1).all()
    assert (seq[~long, 0] == 0).all()                       # max_start = 0 -> start = int(u*1) = 0


def test_orthogonal_regularizer_closed_forms():
    """module_utils.py:55-90: rows of W are normalised (axis=1), the Gram matrix is K x K."""
    eye = torch.eye(6, dtype=torch.float64)
    assert float(O.orthogonal_regularizer(eye, 0.5)) == 0.0
    assert float(O.orthogonal_regularizer(3.0 * eye, 0.5)) == 0.0               # scale invariance of the rows
    D, K = 10, 4
    ones = torch.ones(D, K, dtype=torch.float64)                               # N = 1/sqrt(K): N^T N = D/K everywhere
    want = 0.25 * (K * abs(D / K - 1.0) + K * (K - 1) * D / K)
    assert abs(float(O.orthogonal_regularizer(ones, 0.25)) - want) < 1e-12
    # brute force against explicit loops
    g = torch.Generator().manual_seed(0)
    w = torch.randn(5, 3, generator=g, dtype=torch.float64)
    n = w / w.norm(dim=1, keepdim=True)
    tot = 0.0
    for i in range(3):
        for j in range(3):
            tot += abs(float((n[:, i] * n[:, j]).sum()) - (1.0 if i == j else 0.0))
    assert abs(float(O.orthogonal_regularizer(w, 2.0)) - 2.0 * tot) < 1e-12


def test_ortho_reg_pooling_equals_netvlad_and_light_drops_centres():
    """NetVladOrthoReg.forward (video_pooling_modules.py:1522-1586) is NetVLAD.forward with renamed variables;
    LightVLAD (frame_level_models.py:2827-2877) is the same with zero centres."""
    g = torch.Generator().manual_seed(3)
    B, T, D, K = 2, 6, 8, 4
    x = torch.randn(B * T, D, generator=g, dtype=torch.float64)
    Wc, C, b = (torch.randn(D, K, generator=g, dtype=torch.float64), torch.randn(D, K, generator=g, dtype=torch.float64),
                torch.randn(K, generator=g, dtype=torch.float64))
    ref = O.netvlad_forward(x, {"s/cluster_weights": Wc, "s/cluster_biases": b, "s/cluster_weights2": C[None]}, None, "s", T, False, False)
    got = O.netvlad_ortho_reg_forward(x, {"s/cluster_weightsnetvlad_rgb_scope": Wc, "s/cluster_biasesnetvlad_rgb_scope": b,
                                          "s/cluster_weights2": C}, None, "s", T, False, False, scope_id="netvlad_rgb_scope")
    torch.testing.assert_close(got, ref, rtol=1e-12, atol=1e-14)
    light = O.light_vlad_forward(x, {"s/cluster_weights": Wc, "s/cluster_biases": b}, None, "s", T, False, False)
    zero = O.netvlad_forward(x, {"s/cluster_weights": Wc, "s/cluster_biases": b, "s/cluster_weights2": torch.zeros(1, D, K, dtype=torch.float64)},
                             None, "s", T, False, False)
    torch.testing.assert_close(light, zero, rtol=1e-12, atol=1e-14)


def test_willow_train_step_adds_the_regulariser_gradient():
    sp = O.param_specs("WillowModelReg", iterations=8, cluster_size=8, hidden_size=16, vocab_size=10, rgb_dim=16, audio_dim=8)
    assert sp["video_VLAD/cluster_weights2"][0] == (16, 8) and "audio_VLAD/cluster_weightsnetvlad_audio_scope" in sp
    x, nf, labels = O.synthetic_batch(3, seed=1, max_frames=12, feat=24, vocab=10)
    idx = O.sample_random_frame_indices(nf.numpy(), np.random.RandomState(0).rand(3, 8).astype(np.float32))
    fn = lambda b, P, S: O.willow_model_reg(b[0], b[1], P, S, vocab_size=10, iterations=8, cluster_size=8, is_training=True,
                                            frame_index=idx, rgb_dim=16)
    grads = []
    for extra in (None, lambda P: O.willow_regularization(P, 0.5, 0.5)):
        P, S = O.init_params(sp, seed=3)
        _, g = O.train_step(fn, P, S, {}, [(x, nf)], [labels], step=1, lr=1e-3, extra_reg=extra)
        grads.append(g)
    P, _ = O.init_params(sp, seed=3)
    O.willow_regularization(P, 0.5, 0.5).backward()
    for k in ("video_VLAD/cluster_weights2", "audio_VLAD/cluster_weights2"):
        torch.testing.assert_close(grads[1][k] - grads[0][k], P[k].grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(grads[1]["hidden1_weights"], grads[0]["hidden1_weights"])
