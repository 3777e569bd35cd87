"""Per-kernel GPU parity: each C-ABI kernel against the CPU oracle's restatement of the same
reference lines, on the same seeded inputs.  Tolerances are stated per test."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_sample_and_input_bn(cuda):
    """model_utils.py:101-122 + frame_level_models.py:2265-2271 (train and inference statistics)."""
    from learnablepoolingmethods_b200 import ops
    from oracle import netvlad_oracle as O
    B, Fmax, F, T = 6, 300, 1152, 256
    x, nf, _ = O.synthetic_batch(B, seed=5, max_frames=Fmax, feat=F, vocab=10)
    nf[0] = 1
    nf[1] = 300
    g = torch.Generator().manual_seed(0)
    P = {"input_bn/gamma": torch.rand(F, generator=g) + 0.5, "input_bn/beta": torch.randn(F, generator=g) * 0.1}
    S = {"input_bn/moving_mean": torch.randn(F, generator=g) * 0.01, "input_bn/moving_variance": torch.rand(F, generator=g) * 0.01 + 0.001}
    S0 = {k: v.clone() for k, v in S.items()}
    xs = O.sample_uniform_frames(x, nf, T).reshape(-1, F)
    xd, nfd = x.to(cuda), nf.to(cuda)
    for training in (True, False):
        Sref = {k: v.clone() for k, v in S0.items()}
        ref = O.batch_norm(xs, P, Sref, "input_bn", training)
        mm, mv = S0["input_bn/moving_mean"].clone().to(cuda), S0["input_bn/moving_variance"].clone().to(cuda)
        part = ops.sample_bn_stats(xd, nfd, T) if training else None
        scale, shift = ops.bn_finalize(None if part is None else part[:, 0].contiguous(), None if part is None else part[:, 1].contiguous(),
                                       B * T, P["input_bn/gamma"].to(cuda), P["input_bn/beta"].to(cuda), mm, mv,
                                       training=training, bessel=True)
        y = ops.sample_bn_apply(xd, nfd, T, scale, shift)
        assert rel(y.float(), ref) < 6e-4          # fp16 output rounding (2^-11) only
        if training:
            assert rel(mm, Sref["input_bn/moving_mean"]) < 1e-5
            assert rel(mv, Sref["input_bn/moving_variance"]) < 1e-5


def test_sample_from_uint8_codes(cuda):
    """Ingest side: gather kernels fed with the reader's uint8 codes (readers.py:185-193) dequantise
    (utils.py:28-43) and L2-normalise (train.py:264) on the fly; same statistics / frames as the fp32 entry points fed
    with the oracle's dequantised + normalised tensor, and the model output follows."""
    from learnablepoolingmethods_b200 import ops, variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from oracle import netvlad_oracle as O
    B, T, F = 5, 256, 1152
    x, nf, _, codes = O.synthetic_batch(B, seed=11, return_codes=True)
    nf[0] = 1          # (only shrink: the fp32 tensor is zero beyond the generated num_frames, the codes are not)
    xd, cd, nfd = x.to(cuda), codes.to(cuda), nf.to(cuda)
    p32, p8 = ops.sample_bn_stats(xd, nfd, T), ops.sample_bn_stats(cd, nfd, T)
    assert rel(p8.sum(0), p32.sum(0)) < 1e-6
    g = torch.Generator().manual_seed(0)
    scale, shift = (torch.rand(F, generator=g) + 0.5).to(cuda), torch.randn(F, generator=g).to(cuda)
    y32, y8 = ops.sample_bn_apply(xd, nfd, T, scale, shift), ops.sample_bn_apply(cd, nfd, T, scale, shift)
    assert float((y32.float() - y8.float()).abs().max()) <= 2e-3 and rel(y8.float(), y32.float()) < 2e-4   # fp16 ulp flips only
    ya, yb = ops.sample_bn_apply(cd, nfd, T, scale, shift, split_col=1024)
    assert torch.equal(ya, y8[:, :1024]) and torch.equal(yb, y8[:, 1024:])
    store = variables.VariableStore(cuda, seed=1810)
    eng = NetVladEngine(NetVladConfig(iterations=T, cluster_size=64, hidden_size=64, vocab_size=100), store)
    for training in (False, True):
        pa, _ = eng.forward(xd, nfd, training)
        pb, _ = eng.forward(cd, nfd, training)
        # the two inputs differ by fp16 ulp flips of a few frames; random-init sigmoid gates amplify that (DESIGN.md numerics)
        d = (pa - pb).abs()
        assert float(d.median()) < 1e-4 and float(d.max()) < 5e-2


@pytest.mark.parametrize("B,T,D,K", [(3, 256, 1024, 256), (2, 256, 128, 64), (2, 200, 256, 128), (2, 96, 128, 32), (1, 30, 64, 8),
                                     (3, 256, 1024, 512), (2, 100, 256, 384), (2, 256, 128, 264), (5, 256, 1024, 192)])
def test_netvlad_pool_fwd(cuda, B, T, D, K):
    """NetVLAD.forward (frame_level_models.py:2775-2822) fused kernel vs oracle: rel-L2 <= 1e-3 on the descriptor."""
    from learnablepoolingmethods_b200 import ops
    from oracle import netvlad_oracle as O
    g = torch.Generator().manual_seed(B * 1000 + T + D + K)
    x = torch.randn(B * T, D, generator=g)
    P = {"v/cluster_weights": torch.randn(D, K, generator=g) / D ** 0.5,
         "v/cluster_weights2": torch.randn(1, D, K, generator=g) / D ** 0.5,
         "v/cluster_bn/gamma": torch.rand(K, generator=g) + 0.5, "v/cluster_bn/beta": torch.randn(K, generator=g) * 0.2}
    S = {"v/cluster_bn/moving_mean": torch.randn(K, generator=g) * 0.1, "v/cluster_bn/moving_variance": torch.rand(K, generator=g) + 0.5}
    x16 = x.half()
    ref, A_ref = O.netvlad_forward(x16.float(), P, S, "v", T, True, False, return_assign=True)
    dev = cuda
    scale, shift = ops.bn_finalize(None, None, 1, P["v/cluster_bn/gamma"].to(dev), P["v/cluster_bn/beta"].to(dev),
                                   S["v/cluster_bn/moving_mean"].to(dev), S["v/cluster_bn/moving_variance"].to(dev),
                                   training=False, bessel=True)
    wc16 = ops.cast_f16(P["v/cluster_weights"].to(dev))
    z, rs, a_sum, assign = ops.netvlad_pool_fwd(x16.to(dev), B, T, wc16, scale, shift,
                                                P["v/cluster_weights2"][0].contiguous().to(dev), save_assign=True)
    torch.cuda.synchronize()
    assert rel(assign.float(), A_ref) < 2e-3
    assert rel(a_sum, A_ref.sum(dim=1)) < 1e-3
    vlad = ops.netvlad_finalize(z, rs, d_major=True)
    err = rel(vlad, ref)
    print(f"\n[pool B={B} T={T} D={D} K={K}] vlad rel-L2 {err:.2e}")
    assert err < 1e-3
    vk = ops.netvlad_finalize(z, rs, d_major=False)
    assert rel(vk, ref.reshape(B, D, K).transpose(1, 2)) < 1e-3


def test_netvlad_pool_large_logits(cuda):
    """One-pass softmax with a running max: logits that span +-60 across the 64-cluster chunks (the maximum
    arrives late, early, and in the second CTA of the K=512 cluster pair) stay within the descriptor bound."""
    from learnablepoolingmethods_b200 import ops
    from oracle import netvlad_oracle as O
    for K, hot, spread, boost in ((256, 250, 20.0, 60.0), (256, 3, 20.0, 60.0), (512, 300, 20.0, 60.0), (512, 10, 20.0, 60.0),
                                  (256, 200, 1.5, 4.0), (512, 400, 1.5, 4.0)):
        B, T, D = 2, 256, 128
        g = torch.Generator().manual_seed(K + hot)
        x = torch.randn(B * T, D, generator=g).half()
        P = {"v/cluster_weights": torch.randn(D, K, generator=g) / D ** 0.5, "v/cluster_weights2": torch.randn(1, D, K, generator=g) / D ** 0.5,
             "v/cluster_biases": torch.randn(K, generator=g) * spread}
        P["v/cluster_biases"][hot] += boost
        dev = cuda
        z, rs, a_sum, assign = ops.netvlad_pool_fwd(x.to(dev), B, T, ops.cast_f16(P["v/cluster_weights"].to(dev)), torch.ones(K, device=dev),
                                                    P["v/cluster_biases"].to(dev), P["v/cluster_weights2"][0].contiguous().to(dev),
                                                    save_assign=True)
        ref, A_ref = O.netvlad_forward(x.float(), P, None, "v", T, False, False, return_assign=True)
        assert rel(assign.float(), A_ref) < 2e-3
        assert rel(a_sum, A_ref.sum(dim=1)) < 2e-3
        if spread < 5:
            # (with +-60 logits most assignments underflow fp16 and the intra-normalised rows of those clusters
            # are not comparable; the descriptor bound is checked where every cluster keeps mass)
            assert rel(ops.netvlad_finalize(z, rs), ref) < 1e-3


def test_netvlad_pool_masked_frames(cuda):
    """Masked mode (extension): frames t >= valid_frames[b] contribute nothing."""
    from learnablepoolingmethods_b200 import ops
    from oracle import netvlad_oracle as O
    B, T, D, K = 3, 256, 128, 64
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B * T, D, generator=g).half()
    P = {"v/cluster_weights": torch.randn(D, K, generator=g) / D ** 0.5, "v/cluster_weights2": torch.randn(1, D, K, generator=g) / D ** 0.5,
         "v/cluster_biases": torch.randn(K, generator=g) * 0.1}
    valid = torch.tensor([256, 100, 1], dtype=torch.int32)
    dev = cuda
    one = torch.ones(K, device=dev)
    z, rs, a_sum, _ = ops.netvlad_pool_fwd(x.to(dev), B, T, ops.cast_f16(P["v/cluster_weights"].to(dev)), one,
                                           P["v/cluster_biases"].to(dev), P["v/cluster_weights2"][0].contiguous().to(dev),
                                           valid_frames=valid.to(dev))
    out = ops.netvlad_finalize(z, rs)
    for b in range(B):
        n = int(valid[b])
        ref = O.netvlad_forward(x.float().reshape(B, T, D)[b, :n], P, None, "v", n, False, False)
        assert rel(out[b], ref[0]) < 1e-3


@pytest.mark.parametrize("B,L,Dm,H", [(3, 256, 1024, 64), (2, 64, 128, 16), (2, 16, 128, 16), (2, 256, 128, 8), (1, 80, 64, 4),
                                      (2, 512, 128, 8), (1, 128, 128, 16)])
def test_mha_core_fwd(cuda, B, L, Dm, H):
    """transformer_utils.py:563-581: softmax(q*depth^-0.5 k^T) v per head."""
    from learnablepoolingmethods_b200 import ops
    g = torch.Generator().manual_seed(L + Dm + H)
    qkv = (torch.randn(B * L, 3 * Dm, generator=g)).half()
    dh = Dm // H
    q, k, v = [t.float().reshape(B, L, H, dh).permute(0, 2, 1, 3) for t in qkv.split(Dm, dim=1)]
    logits = (q * dh ** -0.5) @ k.transpose(-1, -2)
    ref = (torch.softmax(logits, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, Dm)
    out, lse = ops.mha_core_fwd(qkv.to(cuda), B, L, Dm, H, scale=dh ** -0.5, want_lse=True)
    assert rel(out.float(), ref) < 2e-3
    assert rel(lse, torch.logsumexp(logits, -1)) < 1e-4
    # per-key affine (V2: batch norm on the logits)
    ks, kb = torch.rand(L, generator=g) + 0.5, torch.randn(L, generator=g)
    ref2 = (torch.softmax((q @ k.transpose(-1, -2)) * ks + kb, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, Dm)
    out2 = ops.mha_core_fwd(qkv.to(cuda), B, L, Dm, H, scale=1.0, key_scale=ks.to(cuda), key_shift=kb.to(cuda))
    assert rel(out2.float(), ref2) < 2e-3


@pytest.mark.parametrize("B,L,Dm,H", [(2, 256, 1024, 64), (2, 64, 128, 16), (2, 16, 128, 16), (1, 80, 64, 4), (1, 240, 64, 4),
                                      (2, 512, 128, 8), (1, 512, 64, 8), (1, 400, 64, 4)])
def test_mha_core_bwd(cuda, B, L, Dm, H):
    """Backward of transformer_utils.py:563-581 against torch autograd (fp64) on the same fp16 inputs."""
    from learnablepoolingmethods_b200 import ops
    g = torch.Generator().manual_seed(7 * L + Dm + H)
    qkv = torch.randn(B * L, 3 * Dm, generator=g).half()
    dout = (torch.randn(B * L, Dm, generator=g) * 0.5).half()
    dh = Dm // H
    ref_in = qkv.double().requires_grad_(True)
    q, k, v = [t.reshape(B, L, H, dh).permute(0, 2, 1, 3) for t in ref_in.split(Dm, dim=1)]
    out_ref = (torch.softmax((q * dh ** -0.5) @ k.transpose(-1, -2), -1) @ v).permute(0, 2, 1, 3).reshape(B * L, Dm)
    out_ref.backward(dout.double())
    qg = qkv.to(cuda)
    o, lse = ops.mha_core_fwd(qg, B, L, Dm, H, scale=dh ** -0.5, want_lse=True)
    dqkv = ops.mha_core_bwd(qg, o, dout.to(cuda), lse, B, L, Dm, H, scale=dh ** -0.5)
    for i, name in enumerate("qkv"):
        e = rel(dqkv[:, i * Dm:(i + 1) * Dm].float(), ref_in.grad[:, i * Dm:(i + 1) * Dm])
        assert e < 4e-3, (name, e)       # fp16 P / dS operands, fp32 accumulation


def test_layernorm_joint(cuda):
    """tf.contrib.layers.layer_norm (begin_norm_axis=1) + residual (transformer_utils.py:406-407)."""
    from learnablepoolingmethods_b200 import ops
    from oracle import netvlad_oracle as O
    B, R, D = 3, 64, 128
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(B, R, D, generator=g).half(), torch.randn(B, R, D, generator=g).half()
    rs = torch.rand(B, R, generator=g) + 0.5
    P = {"ln/gamma": torch.rand(D, generator=g) + 0.5, "ln/beta": torch.randn(D, generator=g)}
    ref = O.layer_norm_joint(a.float() + b.float() * rs[:, :, None], P, "ln")
    ad = a.to(cuda).clone()
    y, sm = ops.layernorm_joint_fwd(ad, b.to(cuda), rs.to(cuda), B, R, D, P["ln/gamma"].to(cuda), P["ln/beta"].to(cuda), save=True)
    assert rel(y.float(), ref) < 1e-3
    u = a.float() + b.float() * rs[:, :, None]
    assert rel(ad.float(), u) < 1e-3
    assert rel(sm[:, 0], u.reshape(B, -1).mean(1)) < 1e-2
    # no residual, strided output
    out = torch.zeros(B, 2 * R * D, dtype=torch.float16, device=cuda)
    ops.layernorm_joint_fwd(a.to(cuda).clone(), None, None, B, R, D, P["ln/gamma"].to(cuda), P["ln/beta"].to(cuda),
                            out=out[:, R * D:], out_stride=2 * R * D)
    assert rel(out[:, R * D:].float().reshape(B, R, D), O.layer_norm_joint(a.float(), P, "ln")) < 1e-3
    assert float(out[:, :R * D].abs().max()) == 0.0


@pytest.mark.parametrize("B,R,D,save", [(3, 256, 1024, True), (2, 256, 1024, False), (3, 64, 512, True), (2, 64, 128, True),
                                        (2, 8, 1024, False), (2, 2, 128, True), (1, 512, 1024, True)])
def test_layernorm_chain(cuda, B, R, D, save):
    """One-pass cluster kernel for LN(LN(f + h1) + h1) (transformer_utils.py:712-713 + :410-411), every cluster size
    (1..8 CTAs per sample), against the oracle's layer norm applied twice."""
    from learnablepoolingmethods_b200 import ops
    from oracle import netvlad_oracle as O
    assert ops.layernorm_chain_supported(R, D)
    g = torch.Generator().manual_seed(R + D)
    a, b = torch.randn(B, R, D, generator=g).half(), torch.randn(B, R, D, generator=g).half()
    P = {"l1/gamma": torch.rand(D, generator=g) + 0.5, "l1/beta": torch.randn(D, generator=g) * 0.3,
         "l2/gamma": torch.rand(D, generator=g) + 0.5, "l2/beta": torch.randn(D, generator=g) * 0.3}
    u1 = a.float() + b.float()
    y1 = O.layer_norm_joint(u1, P, "l1")
    u2 = y1 + b.float()
    ref = O.layer_norm_joint(u2, P, "l2")
    dev = cuda
    out = torch.zeros(B, 2 * R * D, dtype=torch.float16, device=dev)
    ad = a.to(dev)
    r = ops.layernorm_chain_fwd(ad, b.to(dev), B, R, D, P["l1/gamma"].to(dev), P["l1/beta"].to(dev), P["l2/gamma"].to(dev),
                                P["l2/beta"].to(dev), out=out[:, R * D:], out_stride=2 * R * D, save=save)
    y = out[:, R * D:].float().reshape(B, R, D)
    assert rel(y, ref) < 2e-3
    assert float(out[:, :R * D].abs().max()) == 0.0
    assert torch.equal(ad.cpu(), a)                     # the first operand is preserved (it is the ReLU mask of the backward)
    if save:
        _, g1, s1, g2, s2 = r
        assert rel(g1.float(), u1) < 1e-3 and rel(g2.float(), u2) < 2e-3
        m1, v1 = u1.reshape(B, -1).mean(1), u1.reshape(B, -1).var(1, unbiased=False)
        m2, v2 = u2.reshape(B, -1).mean(1), u2.reshape(B, -1).var(1, unbiased=False)
        assert float((s1[:, 0].cpu() - m1).abs().max()) < 2e-3 and rel(s1[:, 1], (v1 + 1e-12).rsqrt()) < 2e-3
        assert float((s2[:, 0].cpu() - m2).abs().max()) < 2e-3 and rel(s2[:, 1], (v2 + 1e-12).rsqrt()) < 2e-3


def test_head_kernels(cuda):
    """Context gating (frame_level_models.py:2342-2368), MoE mix (video_level_models.py:116-126), xent (losses.py:44-51)."""
    from learnablepoolingmethods_b200 import ops
    from oracle import netvlad_oracle as O
    B, H, V, M = 10, 64, 37, 2
    g = torch.Generator().manual_seed(11)
    act, gt = torch.randn(B, H, generator=g), torch.randn(B, H, generator=g)
    P = {"gating_bn/gamma": torch.rand(H, generator=g) + 0.5, "gating_bn/beta": torch.randn(H, generator=g)}
    for training in (True, False):
        S = {"gating_bn/moving_mean": torch.randn(H, generator=g) * 0.1, "gating_bn/moving_variance": torch.rand(H, generator=g) + 0.5}
        mm, mv = S["gating_bn/moving_mean"].clone().to(cuda), S["gating_bn/moving_variance"].clone().to(cuda)
        ref = act * torch.sigmoid(O.batch_norm(gt, P, S, "gating_bn", training))
        o32, o16 = ops.gating_fwd(act.to(cuda), gt.to(cuda), P["gating_bn/gamma"].to(cuda), P["gating_bn/beta"].to(cuda), mm, mv, training=training)
        assert rel(o32, ref) < 1e-5
        assert rel(mv, S["gating_bn/moving_variance"]) < 1e-5
    logits = torch.randn(B, V * (2 * M + 1) + 3, generator=g)
    Pm = {"gates/weights": torch.eye(1), "experts/weights": torch.eye(1)}
    gate = logits[:, :V * (M + 1)].reshape(-1, M + 1)
    ex = logits[:, V * (M + 1):V * (2 * M + 1)].reshape(-1, M)
    ref = (torch.softmax(gate, -1)[:, :M] * torch.sigmoid(ex)).sum(1).reshape(B, V)
    pred = ops.moe_mix_fwd(logits.to(cuda), V, M)
    assert rel(pred, ref) < 1e-5
    labels = torch.rand(B, V, generator=g) < 0.1
    loss, _ = ops.xent_fwd(pred, labels.to(torch.uint8).to(cuda))
    assert abs(float(loss) - float(O.cross_entropy_loss(ref, labels))) < 1e-4 * float(O.cross_entropy_loss(ref, labels))


def test_batch_norm_cols_backward(cuda):
    """slim.batch_norm (training statistics) backward over rows, with the ReLU mask of the preceding activation,
    vs torch autograd in fp64 on the same fp16-rounded inputs."""
    from learnablepoolingmethods_b200 import ops
    g = torch.Generator().manual_seed(21)
    rows, C, T = 512, 64, 128
    x = torch.relu(torch.randn(rows, C, generator=g) + 0.3).half()
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    dy = torch.randn(rows, C, generator=g)
    q = torch.randn(rows // T, C, generator=g)
    pre = (x.double() - 0.0).requires_grad_(True)       # treat the post-ReLU value as the BN input
    xd = pre
    mean, var = xd.mean(0), xd.var(0, unbiased=False)
    y = (xd - mean) / torch.sqrt(var + 1e-3) * gamma.double() + beta.double()
    dy_eff = dy.double() - q.double().repeat_interleave(T, dim=0)
    (y * dy_eff).sum().backward()
    ref_dx = pre.grad * (x.double() > 0)
    xg = x.to(cuda)
    mm, mv = torch.zeros(C, device=cuda), torch.ones(C, device=cuda)
    out = torch.empty_like(xg)
    r = ops.batch_norm_cols_f16(xg, gamma.to(cuda), beta.to(cuda), mm, mv, training=True, bessel=False, save=True, out=out)
    assert rel(out.float(), y.detach()) < 1e-3
    dx, dgam, dbet = ops.batch_norm_cols_bwd(dy.to(cuda), xg, r[2], gamma.to(cuda), inv_scale=1.0, relu=True, q=q.to(cuda), T=T)
    assert rel(dx.float(), ref_dx) < 2e-3, rel(dx.float(), ref_dx)
    xhat = (xd.detach() - mean.detach()) / torch.sqrt(var.detach() + 1e-3)
    assert rel(dbet, dy_eff.sum(0)) < 1e-4 and rel(dgam, (dy_eff * xhat).sum(0)) < 1e-3
    # fp16 dy path, no ReLU, no q
    dx2, _, _ = ops.batch_norm_cols_bwd(dy.half().to(cuda), xg, r[2], gamma.to(cuda), inv_scale=1.0, relu=False)
    pre.grad = None
    (y2 := (xd - xd.mean(0)) / torch.sqrt(xd.var(0, unbiased=False) + 1e-3) * gamma.double() + beta.double())
    (y2 * dy.half().double()).sum().backward()
    assert rel(dx2.float(), pre.grad) < 2e-3


@pytest.mark.parametrize("B,L,Dm,H", [(2, 256, 128, 8), (3, 64, 64, 4)])
def test_mha_bn_logits_backward(cuda, B, L, Dm, H):
    """Attention with batch-normed logits (transformer_utils.py:634-664): two-launch backward vs fp64 autograd."""
    from learnablepoolingmethods_b200 import ops
    g = torch.Generator().manual_seed(5 + L)
    qkv = (torch.randn(B * L, 3 * Dm, generator=g) * 0.7).half()
    gamma, beta = torch.rand(L, generator=g) + 0.5, torch.randn(L, generator=g) * 0.2
    dout = torch.randn(B * L, Dm, generator=g).half()
    t = qkv.double().requires_grad_(True)
    q, k, v = [u.reshape(B, L, H, 16).permute(0, 2, 1, 3) for u in t.split(Dm, dim=1)]
    logits = q @ k.transpose(-1, -2)                                  # [B,H,L,L], channel = last axis
    flat = logits.reshape(-1, L)
    mean, var = flat.mean(0), flat.var(0, unbiased=False)
    lb = (logits - mean) / torch.sqrt(var + 1e-3) * gamma.double() + beta.double()
    out = (torch.softmax(lb, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, Dm)
    (out * dout.double()).sum().backward()
    dgam_ref = torch.autograd.grad((torch.softmax(((logits.detach() - mean.detach()) / torch.sqrt(var.detach() + 1e-3)) * (gm := gamma.double().requires_grad_(True)) + beta.double(), -1) @ v.detach()).permute(0, 2, 1, 3).reshape(B * L, Dm).mul(dout.double()).sum(), gm)[0]
    dev = cuda
    qg = qkv.to(dev)
    part = ops.mha_logit_stats(qg, B, L, Dm, H)
    mm, mv = torch.zeros(L, device=dev), torch.ones(L, device=dev)
    ks, kb, st = ops.bn_finalize(part[:, 0], part[:, 1], B * H * L, gamma.to(dev), beta.to(dev), mm, mv, training=True,
                                 bessel=True, psum_stride=2 * L, save=True)
    assert rel(st[0], mean.detach()) < 1e-3 and rel(st[1], 1 / torch.sqrt(var.detach() + 1e-3)) < 1e-3
    o, lse = ops.mha_core_fwd(qg, B, L, Dm, H, scale=1.0, key_scale=ks, key_shift=kb, want_lse=True)
    assert rel(o.float(), out.detach()) < 2e-3
    p1 = ops.mha_core_bwd_bn(1, qg, o, dout.to(dev), lse, B, L, Dm, H, ks, kb, st[0], st[1])
    n = float(B * H * L)
    m12 = torch.empty(2, L, device=dev)
    ops.colsum_final(p1, B * H, 2 * L, 2 * L, alpha=1.0 / n, out=m12.view(-1))
    dgam = ops.colsum_final(p1[:, 1], B * H, 2 * L, L)
    dqkv = ops.mha_core_bwd_bn(2, qg, o, dout.to(dev), lse, B, L, Dm, H, ks, kb, st[0], st[1], m1=m12[0], m2=m12[1])
    e = rel(dqkv.float(), t.grad)
    print(f"\n[mha bn bwd L={L}] dqkv rel-L2 {e:.2e}, dgamma {rel(dgam, dgam_ref):.2e}")
    assert e < 1e-2
    assert rel(dgam, dgam_ref) < 1e-2


@pytest.mark.parametrize("R,Kd,N", [(80, 2048 + 128, 512), (4, 384, 64), (20, 136, 96)])
def test_rank_adam_step(cuda, R, Kd, N):
    """Factored hidden-projection update (train.py:321-336 + utils.py:181-188 on hidden1_weights): clip factor from
    the Gram matrices and Adam from the factors must equal the dense path dW = alpha*A^T G -> clip_by_norm -> Adam."""
    from learnablepoolingmethods_b200 import ops
    g = torch.Generator().manual_seed(R + Kd + N)
    A = torch.randn(R, Kd, generator=g).half()
    G = (torch.randn(R, N, generator=g) * 3).half()
    w = torch.randn(Kd, N, generator=g) * 0.05
    m, v = torch.randn(Kd, N, generator=g) * 1e-3, torch.rand(Kd, N, generator=g) * 1e-5
    alpha, clip, lr_t, b1, b2, eps = 1.0 / 64, 1.0, 3e-4, 0.9, 0.999, 1e-8
    dW = alpha * (A.double().t() @ G.double())
    norm = float(dW.norm())
    gr = dW * (clip / max(norm, clip))
    m_ref = b1 * m.double() + (1 - b1) * gr
    v_ref = b2 * v.double() + (1 - b2) * gr * gr
    w_ref = w.double() - lr_t * m_ref / (v_ref.sqrt() + eps)
    dev = cuda
    Ad, Gd = A.to(dev), G.to(dev)
    Rp = (R + 7) // 8 * 8
    Ap, Gp = torch.zeros(Rp, Kd, dtype=torch.float16, device=dev), torch.zeros(Rp, N, dtype=torch.float16, device=dev)
    Ap[:R], Gp[:R] = Ad, Gd
    ga = ops.gemm(Ap, Ap, b_mn=False, out_dtype=torch.float32)
    gg = ops.gemm(Gp, Gp, b_mn=False, out_dtype=torch.float32)
    sc = torch.zeros(2, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.rank_grad_clip(ga, gg, alpha, clip, sc[0:1], sc[1:2], flag)
    assert abs(float(sc[1]) - norm) / norm < 1e-4 and int(flag) == 0
    wd, md, vd = w.to(dev), m.to(dev), v.to(dev)
    w16 = torch.zeros(Kd, N, dtype=torch.float16, device=dev)
    ops.rank_adam_step(Ad, Gd, alpha, sc[0:1], flag, wd, md, vd, w16, lr_t=lr_t, b1=b1, b2=b2, eps=eps)
    assert rel(md, m_ref) < 1e-5 and rel(vd, v_ref) < 1e-5
    assert float((wd.double().cpu() - w_ref).abs().max()) < 1e-6
    assert torch.equal(w16.cpu(), wd.cpu().half())
    # a raised flag (loss-scale overflow elsewhere) skips the update
    flag.fill_(1)
    before = wd.clone()
    ops.rank_adam_step(Ad, Gd, alpha, sc[0:1], flag, wd, md, vd, w16, lr_t=lr_t)
    assert torch.equal(before, wd)


@pytest.mark.parametrize("R,Kd,N", [(80, 4096 + 128, 512), (4, 384, 64), (20, 136, 96), (96, 256, 512)])
def test_rank_adam_tiled_equals_persistent(cuda, R, Kd, N):
    """The small-CTA kernel that runs under the backward (lpm_rank_adam_step_ex, tiled=1, step size from device memory)
    must produce bit-identical w / m / v / fp16 shadow to the persistent kernel."""
    from learnablepoolingmethods_b200 import ops
    g = torch.Generator().manual_seed(R * 7 + Kd + N)
    A = torch.randn(R, Kd, generator=g).half().to(cuda)
    G = (torch.randn(R, N, generator=g) * 3).half().to(cuda)
    w0 = (torch.randn(Kd, N, generator=g) * 0.05).to(cuda)
    m0, v0 = (torch.randn(Kd, N, generator=g) * 1e-3).to(cuda), (torch.rand(Kd, N, generator=g) * 1e-5).to(cuda)
    factor = torch.full((1,), 0.37, device=cuda)
    flag = torch.zeros(1, dtype=torch.int32, device=cuda)
    lr = 3e-4
    outs = []
    for tiled in (False, True):
        w, m, v = w0.clone(), m0.clone(), v0.clone()
        w16 = torch.zeros(Kd, N, dtype=torch.float16, device=cuda)
        kw = dict(lr_dev=torch.full((1,), lr, device=cuda), tiled=True,
                  workspace=torch.empty(ops.rank_adam_workspace_bytes(R, N), dtype=torch.uint8, device=cuda)) if tiled else dict(lr_t=lr)
        for _ in range(2):
            ops.rank_adam_step(A, G, 1.0 / 64, factor, flag, w, m, v, w16, **kw)
        outs.append((w, m, v, w16))
    torch.cuda.synchronize()
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert not torch.equal(outs[0][0], w0)


def test_step_begin_latches_overflow_flag(cuda):
    """lpm_step_begin: a raised skip flag is counted once and cleared, so one overflow skips exactly one update."""
    from learnablepoolingmethods_b200 import ops
    flag = torch.ones(1, dtype=torch.int32, device=cuda)
    skipped = torch.zeros(1, dtype=torch.int32, device=cuda)
    ops.step_begin(flag, skipped)
    ops.step_begin(flag, skipped)
    assert int(flag) == 0 and int(skipped) == 1


def test_eval_metrics_against_reference_golden(cuda):
    """lpm_eval_topk / lpm_eval_metrics against the frozen outputs of the reference's own
    eval_util.py (tests/golden/eval_golden.npz) and the numpy oracle: top-20 class sets identical, hit@1 / PERR exact to
    fp32, GAP to 1e-6; videos without labels, V < k and a full 80 x 3862 batch included."""
    import os
    from learnablepoolingmethods_b200 import eval_util as GE
    from oracle import eval_oracle as E
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_golden.npz"))
    for seed in sorted(int(k[4:]) for k in G.files if k.startswith("case")):
        meta = G[f"case{seed}"]
        pred, labels = E.synthetic_eval_batch(int(meta[0]), int(meta[1]), int(meta[2]), tuple(int(z) for z in meta[3:]))
        p, a = torch.from_numpy(pred).to(cuda), torch.from_numpy(labels).to(cuda)
        idx, val, lab = GE.top_k_triplets(p, a, 20)
        k = min(20, pred.shape[1])
        got = [sorted(r[:k]) for r in idx.cpu().tolist()]
        assert got == [sorted(r.tolist()) for r in G[f"topk{seed}"]]
        v = val.cpu().numpy()[:, :k]
        assert (np.diff(v, axis=1) <= 0).all()                                     # ranked
        assert np.array_equal(v, np.take_along_axis(pred, idx.cpu().numpy()[:, :k].astype(np.int64), 1))
        assert np.array_equal(lab.cpu().numpy()[:, :k], np.take_along_axis(labels, idx.cpu().numpy()[:, :k].astype(np.int64), 1))
        m = GE.batch_metrics(p, a).cpu().double().numpy()
        assert abs(m[0] - float(G[f"hit{seed}"])) < 1e-6 and abs(m[1] - float(G[f"perr{seed}"])) < 1e-6
        assert abs(m[2] - float(G[f"gap{seed}"])) < 1e-6
        assert abs(m[2] - E.gap(pred, labels)) < 1e-6
    assert abs(GE.calculate_gap(p, a) - float(G[f"gap{seed}"])) < 1e-6
    # the reference's default --batch_size (1024) at k = 20 exceeds the single-CTA ranking: large-batch path vs the oracle
    pred, labels = E.synthetic_eval_batch(77, 1024, 500)
    p, a = torch.from_numpy(pred).to(cuda), torch.from_numpy(labels).to(cuda)
    m = GE.batch_metrics(p, a).cpu().double().numpy()
    assert abs(m[0] - E.hit_at_one(pred, labels)) < 1e-6 and abs(m[1] - E.perr(pred, labels)) < 1e-6
    assert abs(m[2] - E.gap(pred, labels)) < 1e-6


@pytest.mark.parametrize("splits,splits2", [(5, 0), (40, 0), (74, 37), (3, 2)])
def test_splitk_reduce_variants_and_split_operands(cuda, splits, splits2):
    """lpm_splitk_reduce_ex (narrow and wide kernels, second partial set, split-precision output) and
    lpm_split_hi_lo_f16 (all three layouts; 8-wide, pair and scalar kernels): hi + lo reproduces fp32 to ~2^-22."""
    from learnablepoolingmethods_b200 import ops
    g = torch.Generator().manual_seed(splits * 100 + splits2)
    B, H = 80, 512
    parts = torch.randn(splits, B, H, generator=g).to(cuda)
    parts2 = torch.randn(splits2, B, H, generator=g).to(cuda) if splits2 else None
    bias = torch.randn(H, generator=g).to(cuda)
    want = parts.double().sum(0) + (parts2.double().sum(0) if splits2 else 0) + bias.double()
    out32 = torch.empty(B, H, device=cuda)
    a3 = torch.empty(B, 3 * H, dtype=torch.float16, device=cuda)
    ops.splitk_reduce(parts, bias=bias, out32=out32, out16=a3, parts2=parts2, split3=True)
    assert float((out32.double() - want).abs().max()) < 1e-4
    assert torch.equal(a3[:, :H], a3[:, 2 * H:]) and torch.equal(a3[:, :H], out32.half())
    rec = a3[:, :H].double() + a3[:, H:2 * H].double()
    # hi + lo carries ~22 bits; the low-order half bottoms out at fp16's subnormal spacing (6e-8 absolute)
    assert bool(((rec - out32.double()).abs() <= 3e-7 * out32.double().abs() + 6.1e-8).all())
    if not splits2:       # the plain entry point (no second set, plain fp16 output) agrees with the extended one
        out16 = torch.empty(B, H, dtype=torch.float16, device=cuda)
        ops.splitk_reduce(parts, bias=bias, out32=out32, out16=out16)
        assert torch.equal(out16, a3[:, :H])
    for cols in (64, 3 * 3862 % 1000 + 2, 37):              # 8-wide, pair and scalar kernels
        w = (torch.randn(48, cols, generator=g) * 0.1).to(cuda)
        act = ops.split_hi_lo(w)
        assert torch.equal(act[:, :cols], w.half()) and torch.equal(act[:, :cols], act[:, 2 * cols:])
        assert float((act[:, :cols].double() + act[:, cols:2 * cols].double() - w.double()).abs().max()) < 1e-7
        pad = (cols + 7) // 8 * 8
        dst = torch.zeros(3 * 48, pad, dtype=torch.float16, device=cuda)
        ops.split_hi_lo(w, dst, along_rows=True)
        assert torch.equal(dst[:48, :cols], w.half()) and torch.equal(dst[48:96, :cols], w.half())
        assert torch.equal(dst[96:, :cols], act[:, cols:2 * cols])
        lo = torch.zeros(48, pad, dtype=torch.float16, device=cuda)
        ops.split_hi_lo(w, lo, along_rows=2)
        assert torch.equal(lo[:, :cols], act[:, cols:2 * cols])


@pytest.mark.parametrize("B", [1, 3])
def test_mha_tcgen05_paths_agree(cuda, B):
    """lpm_attn_tc.cu: the tcgen05 / TMEM attention kernels (four heads per CTA; backward with the operands in shared memory
    or in TMEM; opt-in forward) against fp64 autograd of transformer_utils.py:563-581 and against the warp-level kernels."""
    from learnablepoolingmethods_b200 import ops
    from learnablepoolingmethods_b200._lib import load
    lib = load()
    L, Dm, H = 256, 1024, 64
    g = torch.Generator().manual_seed(11 + B)
    qkv = torch.randn(B * L, 3 * Dm, generator=g).half()
    dout = (torch.randn(B * L, Dm, generator=g) * 0.5).half()
    ref_in = qkv.double().requires_grad_(True)
    q, k, v = [t.reshape(B, L, H, 16).permute(0, 2, 1, 3) for t in ref_in.split(Dm, dim=1)]
    logits = (q * 0.25) @ k.transpose(-1, -2)
    out_ref = (torch.softmax(logits, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, Dm)
    out_ref.backward(dout.double())
    qg, dg = qkv.to(cuda), dout.to(cuda)
    res = {}
    try:
        for mode in (0, 1, 2, 4, 8):
            lib.lpm_debug_set_mha_tc_mode(mode)
            o, lse = ops.mha_core_fwd(qg, B, L, Dm, H, scale=0.25, want_lse=True)
            dqkv = ops.mha_core_bwd(qg, o, dg, lse, B, L, Dm, H, scale=0.25)
            torch.cuda.synchronize()
            assert rel(o.float(), out_ref.detach()) < 2e-3, mode
            assert rel(lse.reshape(-1), torch.logsumexp(logits.detach(), -1).reshape(-1)) < 1e-4, mode
            for i in range(3):
                e = rel(dqkv[:, i * Dm:(i + 1) * Dm].float(), ref_in.grad[:, i * Dm:(i + 1) * Dm])
                assert e < 4e-3, (mode, "qkv"[i], e)
            res[mode] = (o, dqkv)
    finally:
        lib.lpm_debug_set_mha_tc_mode(2)
    # same fp16 P / dS operands, fp32 accumulation in a different order
    assert rel(res[1][1].float(), res[0][1].float()) < 1e-4 and rel(res[2][1].float(), res[0][1].float()) < 1e-4
    assert rel(res[4][0].float(), res[0][0].float()) < 1e-3 and rel(res[8][0].float(), res[0][0].float()) < 1e-3


def test_mha_tcgen05_strided_views_and_fallbacks(cuda):
    """The tcgen05 attention path takes q/k/v as a column view of a wider buffer (row stride > 3*Dm, as the engine's fused
    projections produce) and writes dq/dk/dv likewise; shapes outside its envelope (length != 256, depth 8, heads not a
    multiple of 4) must quietly take the warp-level kernels with the same results as lpm_debug_set_mha_tc_mode(0)."""
    from learnablepoolingmethods_b200 import ops
    from learnablepoolingmethods_b200._lib import load
    lib = load()
    g = torch.Generator().manual_seed(5)
    B, L, Dm, H = 2, 256, 256, 16
    wide = torch.randn(B * L, 3 * Dm + 64, generator=g).half().to(cuda)
    qkv = wide[:, 32:32 + 3 * Dm]                       # 64-byte column offset, row stride 3*Dm + 64
    dout = (torch.randn(B * L, Dm, generator=g) * 0.5).half().to(cuda)
    res = {}
    try:
        for mode in (0, 2):
            lib.lpm_debug_set_mha_tc_mode(mode)
            o, lse = ops.mha_core_fwd(qkv, B, L, Dm, H, scale=0.25, want_lse=True)
            res[mode] = ops.mha_core_bwd(qkv, o, dout, lse, B, L, Dm, H, scale=0.25)
        torch.cuda.synchronize()
        assert rel(res[2].float(), res[0].float()) < 1e-4
        # outside the envelope: identical bits with the switch on or off (same kernel runs)
        for (L2, Dm2, H2) in ((128, 256, 16), (256, 128, 16), (256, 96, 6)):
            q2 = torch.randn(B * L2, 3 * Dm2, generator=g).half().to(cuda)
            d2 = torch.randn(B * L2, Dm2, generator=g).half().to(cuda)
            outs = []
            for mode in (0, 2):
                lib.lpm_debug_set_mha_tc_mode(mode)
                o, lse = ops.mha_core_fwd(q2, B, L2, Dm2, H2, scale=(Dm2 // H2) ** -0.5, want_lse=True)
                outs.append(ops.mha_core_bwd(q2, o, d2, lse, B, L2, Dm2, H2, scale=(Dm2 // H2) ** -0.5))
            torch.cuda.synchronize()
            assert torch.equal(outs[0], outs[1]), (L2, Dm2, H2)
    finally:
        lib.lpm_debug_set_mha_tc_mode(2)


def test_mha_tcgen05_benchmark_batch(cuda):
    """Config-1 shape [80, 64 heads, 256, 16]: every CTA of the 1280-CTA launch against the warp-level kernel, twice (the
    second launch exercises the L2-prefetched tiles of the first)."""
    from learnablepoolingmethods_b200 import ops
    from learnablepoolingmethods_b200._lib import load
    lib = load()
    B, L, Dm, H = 80, 256, 1024, 64
    g = torch.Generator(device=cuda).manual_seed(3)
    qkv = (torch.randn(B * L, 3 * Dm, device=cuda, generator=g) * 0.5).half()
    dout = (torch.randn(B * L, Dm, device=cuda, generator=g) * 0.1).half()
    try:
        lib.lpm_debug_set_mha_tc_mode(0)
        o, lse = ops.mha_core_fwd(qkv, B, L, Dm, H, scale=0.25, want_lse=True)
        ref = ops.mha_core_bwd(qkv, o, dout, lse, B, L, Dm, H, scale=0.25)
        lib.lpm_debug_set_mha_tc_mode(2)
        a = ops.mha_core_bwd(qkv, o, dout, lse, B, L, Dm, H, scale=0.25)
        b = ops.mha_core_bwd(qkv, o, dout, lse, B, L, Dm, H, scale=0.25)
        torch.cuda.synchronize()
    finally:
        lib.lpm_debug_set_mha_tc_mode(2)
    assert torch.equal(a, b)                                   # deterministic
    per_sample = (a.float() - ref.float()).reshape(B, -1).norm(dim=1) / ref.float().reshape(B, -1).norm(dim=1)
    assert float(per_sample.max()) < 1e-4, per_sample.max()


@pytest.mark.parametrize("B,Kd,H,splits,training,diag,second", [(80, 270336 // 16, 512, 74, True, False, False),
                                                               (80, 8192, 512, 74, False, False, True),
                                                               (5, 1024, 128, 8, True, True, False),
                                                               (128, 4096, 1024, 37, False, True, True)])
def test_gemm_splitk_gated_k3(cuda, B, Kd, H, splits, training, diag, second):
    """K3 (frame_level_models.py:2314-2368) in one launch: hidden projection (+ the partials of a second pass), bias, gate
    product, gating_bn (batch or moving statistics), sigmoid, product -- against fp64 torch on the same fp16 operands."""
    from learnablepoolingmethods_b200 import ops
    g = torch.Generator().manual_seed(B + H + splits)
    a = (torch.randn(B, Kd, generator=g) * 0.05).half()
    w = (torch.randn(Kd, H, generator=g) * 0.05).half()
    bias = torch.randn(H, generator=g) * 0.1
    wg = torch.randn(H, H, generator=g) / H ** 0.5
    gamma, beta = torch.rand(H, generator=g) + 0.5, torch.randn(H, generator=g) * 0.2
    mm, mv = torch.randn(H, generator=g) * 0.1, torch.rand(H, generator=g) + 0.5
    hidden = a.double() @ w.double() + bias.double()
    parts2 = None
    if second:
        a2 = (torch.randn(B, Kd, generator=g) * 0.01).half()
        parts2 = ops.gemm(a2.to(cuda), w.to(cuda), splits=4)
        hidden = hidden + a2.double() @ w.double()
    gp = hidden @ wg.double()
    v = gp - (torch.diagonal(wg).double() * hidden if diag else 0)
    if training:
        mean, var = v.mean(0), v.var(0, unbiased=False)
    else:
        mean, var = mm.double(), mv.double()
    gates = (v - mean) / torch.sqrt(var + 1e-3) * gamma.double() + beta.double()
    ref = hidden * torch.sigmoid(gates)
    mmd, mvd = mm.to(cuda).clone(), mv.to(cuda).clone()
    act32, a3, out32, g3, st, g_sum, _ = ops.gemm_splitk_gated(
        a.to(cuda), w.to(cuda), splits=splits, bias=bias.to(cuda), wg=wg.to(cuda), gamma=gamma.to(cuda), beta=beta.to(cuda),
        moving_mean=mmd, moving_var=mvd, training=training, wg_diag=torch.diagonal(wg).contiguous().to(cuda) if diag else None,
        save=True, parts2=parts2)
    torch.cuda.synchronize()
    assert rel(act32, hidden) < 1e-5
    assert rel(g_sum, gp) < 1e-5
    assert rel(out32, ref) < 2e-5
    assert rel(st[0], mean) < 1e-4 and rel(st[1], 1 / torch.sqrt(var + 1e-3)) < 1e-4
    # split-precision operands: hi + lo reproduces the fp32 value to 2^-22
    assert rel(a3[:, :H].float() + a3[:, H:2 * H].float(), act32) < 1e-6 and torch.equal(a3[:, :H], a3[:, 2 * H:])
    assert rel(g3[:, :H].float() + g3[:, H:2 * H].float(), out32) < 1e-6
    if training:
        corr = B / (B - 1)
        assert rel(mmd, mm.double() * 0.999 + mean * 0.001) < 1e-5 and rel(mvd, mv.double() * 0.999 + var * corr * 0.001) < 1e-5
    # second launch: the barrier counters were left at zero
    o2 = ops.gemm_splitk_gated(a.to(cuda), w.to(cuda), splits=splits, bias=bias.to(cuda), wg=wg.to(cuda), gamma=gamma.to(cuda),
                               beta=beta.to(cuda), moving_mean=mm.to(cuda).clone(), moving_var=mv.to(cuda).clone(),
                               training=training, wg_diag=torch.diagonal(wg).contiguous().to(cuda) if diag else None, parts2=parts2)[2]
    torch.cuda.synchronize()
    assert torch.equal(o2, out32)
