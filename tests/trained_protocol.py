"""Trained-weights parity protocol at the config-1 shape (BASELINE.json north_star acceptance).

The reference ships no checkpoint, so "trained weights" are produced here: the CUDA trainer runs `steps` training
steps at the benchmarked shape (B=80, 256 of 300 frames, K=256/64, hidden 512, vocab 3862) on STRUCTURED synthetic
videos -- every class owns a random direction in feature space and a video's frames are noise plus the directions
of its labels, so that the loss really falls and the hidden activations differ from video to video.  The trained
variables (weights + batch-norm moving statistics) are then copied into the CPU oracle and both sides run inference
on the same held-out videos.  Reported: the three north-star numbers (VLAD rel-L2, prediction max-abs, identical
top-20 label sets) plus what explains them.

Test infrastructure: the data generator uses torch ops on the GPU only to make inputs; every model operation on the
product side goes through the C-ABI kernels, every operation on the checker side through `oracle/`.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SHAPE = dict(B=80, T=256, K=256, Hd=512, V=3862, F=1152, max_frames=300)


def class_directions(V, F, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(V, F, generator=g)


def structured_batch(B, seed, protos, device, *, V=SHAPE["V"], max_frames=SHAPE["max_frames"], strength=1.0,
                     min_frames=30):
    """(x fp32 [B, max_frames, F] L2-normalised per frame (train.py:264), nf int32 [B], labels uint8 [B, V]) on `device`.
    Labels: 1 + Poisson(2) positives, Zipf(1.0) over classes; frames: clipped N(0,1) codes of
    noise + strength * (sum of the label directions) / sqrt(#labels), dequantised (utils.py:28-43), zero-padded."""
    g = torch.Generator().manual_seed(seed)
    w = 1.0 / torch.arange(1, V + 1, dtype=torch.float64)
    labels = torch.zeros(B, V, dtype=torch.uint8)
    npos = 1 + torch.poisson(torch.full((B,), 2.0), generator=g).long()
    for b in range(B):
        labels[b, torch.multinomial(w, int(npos[b]), replacement=False, generator=g)] = 1
    nf = torch.randint(min_frames, max_frames + 1, (B,), generator=g, dtype=torch.int32)
    off = (labels.float() @ protos) / npos.float().sqrt()[:, None]                 # [B, F]
    gd = torch.Generator(device=device).manual_seed(seed)
    z = torch.randn(B, max_frames, protos.shape[1], generator=gd, device=device)
    z = (z + strength * off.to(device)[:, None, :]) / math.sqrt(1.0 + strength * strength)
    q = torch.clamp(torch.round((z + 2) * 255 / 4), 0, 255)
    x = q * (4.0 / 255.0) + (4.0 / 512.0 - 2.0)
    mask = (torch.arange(max_frames, device=device)[None, :] < nf.to(device)[:, None]).to(x.dtype)
    x = x * mask[:, :, None]
    x = x * torch.rsqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=1e-12))
    return x, nf.to(device), labels.to(device)


def train_model(device, *, steps, model="NetVladV1", lr=2e-4, seed=1810, log_every=100, strength=1.0, shape=SHAPE,
                resume=None):
    """Returns (engine, trainer, class directions, [(step, loss)]).  resume=(engine, trainer, directions): continue that
    run up to `steps` total steps (batch i is always seeded 5000 + i)."""
    from learnablepoolingmethods_b200 import variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from learnablepoolingmethods_b200.trainer import Trainer
    s = shape
    if resume is None:
        store = variables.VariableStore(device, seed=seed)
        cfg = NetVladConfig(model=model, iterations=s["T"], cluster_size=s["K"], hidden_size=s["Hd"], vocab_size=s["V"])
        eng = NetVladEngine(cfg, store)
        tr = Trainer(eng, base_learning_rate=lr, learning_rate_decay=0.85, batch_size=s["B"])
        protos = class_directions(s["V"], s["F"])
    else:
        eng, tr, protos = resume
    losses = []
    for i in range(tr.global_step, steps):
        x, nf, lab = structured_batch(s["B"], 5000 + i, protos, device, V=s["V"], max_frames=s["max_frames"], strength=strength)
        loss = tr.train_step(x, nf, lab)
        if i % log_every == 0 or i == steps - 1:
            losses.append((i, float(loss)))
    tr.sync_parameters()
    torch.cuda.synchronize()
    return eng, tr, protos, losses


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def top20_compare(ref, other, rel_tie=0.02, abs_floor=1e-37):
    """ref / other: float [N, V] numpy.  (identical fraction, tie-aware fraction, mean overlap): a video is tie-aware
    identical when every label in the symmetric difference is a near-tie with the 20th REFERENCE score: within 2 % of it
    (a logit difference of 0.02) or, for scores in the denormal range, within 1e-37."""
    from oracle import eval_oracle as E
    a_sets, g_sets = E.top_k_sets(ref, 20), E.top_k_sets(other, 20)
    ident = tie = 0
    overlap = 0
    diffs = []
    for r, (a, g) in enumerate(zip(a_sets, g_sets)):
        a, g = set(a), set(g)
        overlap += len(a & g)
        if a == g:
            ident += 1
            tie += 1
            continue
        kth = np.sort(ref[r])[-20]
        diffs.append({"video": r, "kth_ref_score": float(kth),
                      "labels": [(int(lab), float(ref[r, lab]), float(other[r, lab])) for lab in sorted(a ^ g)]})
        if all(abs(float(ref[r, lab]) - kth) <= max(rel_tie * float(kth), abs_floor) for lab in (a ^ g)):
            tie += 1
    n = ref.shape[0]
    top20_compare.last_diffs = diffs
    return ident / n, tie / n, overlap / n


def evaluate(eng, protos, device, *, n_videos=1024, chunk=128, emulate=(), strength=1.0, shape=SHAPE, batch=None,
             fp64=False):
    """Held-out inference on both sides.  Returns a dict of the parity numbers; `emulate` adds the same comparison for
    the ORACLE with its matmul operands rounded to the given formats (precision study); fp64=True also runs the oracle
    in double precision: `fp32_noise` = how far the fp32 oracle (the reference's own arithmetic) is from it, and
    `pred_max_abs_vs_fp64` = how far the product is."""
    from learnablepoolingmethods_b200 import ops
    from oracle import eval_oracle as E
    from oracle import netvlad_oracle as O
    from tests.helpers import oracle_params
    s = shape
    B = batch or s["B"]
    P, S = oracle_params(eng.store)
    P64, S64 = oracle_params(eng.store, dtype=torch.float64) if fp64 else (None, None)
    kw = dict(vocab_size=s["V"], iterations=s["T"], cluster_size=s["K"], is_training=False)
    ref_pred, gpu_pred, gpu_top, labels, ref64 = [], [], [], [], []
    emu_pred = {m: [] for m in emulate}
    vl = {"vlad_video": [0.0, 0.0], "vlad_audio": [0.0, 0.0], "att_video": [0.0, 0.0], "hidden": [0.0, 0.0], "gated": [0.0, 0.0]}
    worst_vlad = 0.0
    # Denormals are kept on both sides.  (Flushing them -- TensorFlow's CPU runtime does -- turns the tail of a saturated
    # model's ranking into a mass of exact zeros ordered by class index, and any score that crosses FLT_MIN on one side only
    # jumps that whole queue: measured 15 % identical top-20 sets with flushing on both sides, 99.8-100 % without.)
    for c0 in range(0, n_videos, chunk):
        n = min(chunk, n_videos - c0)
        x, nf, lab = structured_batch(n, 900000 + c0, protos, device, V=s["V"], max_frames=s["max_frames"], strength=strength)
        xc, nfc = x.cpu(), nf.cpu()
        with torch.no_grad():
            ref, inter = O.netvlad_v1(xc, nfc, P, S, return_intermediates=True, **kw)
            for m in emulate:
                # "fp16-body": operands rounded everywhere except the head's three products (what a split-precision head
                # on top of the fp16 body could reach at best)
                O.OPERAND_ROUND, O.OPERAND_ROUND_HEAD = m.split("-")[0], not m.endswith("-body")
                try:
                    emu_pred[m].append(O.netvlad_v1(xc, nfc, P, S, **kw).numpy())
                finally:
                    O.OPERAND_ROUND, O.OPERAND_ROUND_HEAD = None, True
            if fp64:
                ref64.append(O.netvlad_v1(xc.double(), nfc, P64, S64, **kw).numpy())
        ref_pred.append(ref.numpy())
        labels.append(lab.cpu().numpy())
        # product side: batches of B (the benchmarked tower batch; the tail batch is smaller)
        for b0 in range(0, n, B):
            b1 = min(n, b0 + B)
            pred, ctx = eng.forward(x[b0:b1], nf[b0:b1], False, return_intermediates=True)
            _, ti, _, _ = ops.eval_topk(pred, lab[b0:b1], 20)
            gpu_pred.append(pred.cpu().numpy())
            gpu_top.append(ti.cpu().numpy())
            gi = ctx["inter"]
            for k in vl:
                a, r = gi[k].double().cpu().reshape(b1 - b0, -1), inter[k][b0:b1].double().reshape(b1 - b0, -1)
                vl[k][0] += float(((a - r) ** 2).sum())
                vl[k][1] += float((r ** 2).sum())
                if k == "vlad_video":
                    worst_vlad = max(worst_vlad, float(((a - r).norm(dim=1) / r.norm(dim=1)).max()))
    ref_pred, gpu_pred = np.concatenate(ref_pred), np.concatenate(gpu_pred)
    gpu_top, labels = np.concatenate(gpu_top), np.concatenate(labels)
    err = np.abs(gpu_pred - ref_pred)
    out = {k + "_rel_l2": math.sqrt(v[0] / max(v[1], 1e-300)) for k, v in vl.items()}
    out["vlad_video_rel_l2_worst_video"] = worst_vlad
    out["pred_max_abs"] = float(err.max())
    out["pred_median_abs"] = float(np.median(err))
    out["pred_p999_abs"] = float(np.quantile(err, 0.999))
    ident, tie, ov = top20_compare(ref_pred, gpu_pred)
    out.update(top20_identical=ident, top20_identical_up_to_ties=tie, top20_mean_overlap=ov,
               top20_differences=top20_compare.last_diffs[:8])
    # the product's own on-GPU top-k (lpm_eval_topk) must select exactly the labels numpy selects from its predictions
    own = sum(sorted(gpu_top[r].tolist()) == sorted(np.argsort(-gpu_pred[r], kind="stable")[:20].tolist())
              for r in range(gpu_pred.shape[0]))
    out["gpu_topk_kernel_consistent"] = own / gpu_pred.shape[0]
    out["gap_oracle"] = E.gap(ref_pred, labels, 20)
    out["gap_gpu"] = E.gap(gpu_pred, labels, 20)
    out["hit1_oracle"] = E.hit_at_one(ref_pred, labels)
    out["n_videos"] = int(ref_pred.shape[0])
    # margin statistics that decide whether "identical top-20 sets" is attainable: gap between the 20th and 21st score
    srt = -np.sort(-ref_pred, axis=1)
    out["rank20_21_gap_median"] = float(np.median(srt[:, 19] - srt[:, 20]))
    out["rank20_score_median"] = float(np.median(srt[:, 19]))
    if fp64:
        r64 = np.concatenate(ref64)
        e32 = np.abs(ref_pred.astype(np.float64) - r64)
        eg = np.abs(gpu_pred.astype(np.float64) - r64)
        i3, t3, _ = top20_compare(r64, ref_pred.astype(np.float64))
        i4, t4, _ = top20_compare(r64, gpu_pred.astype(np.float64))
        out["fp32_noise"] = dict(pred_max_abs=float(e32.max()), top20_identical=i3, top20_identical_up_to_ties=t3,
                                 what="fp32 oracle vs fp64 oracle (the reference's own arithmetic against exact)")
        out["pred_max_abs_vs_fp64"] = float(eg.max())
        out["top20_identical_vs_fp64"] = i4
    for m in emulate:
        ep = np.concatenate(emu_pred[m])
        e = np.abs(ep - ref_pred)
        i2, t2, o2 = top20_compare(ref_pred, ep)
        out["oracle_" + m] = dict(pred_max_abs=float(e.max()), pred_median_abs=float(np.median(e)), top20_identical=i2,
                                  top20_identical_up_to_ties=t2, top20_mean_overlap=o2)
    return out
