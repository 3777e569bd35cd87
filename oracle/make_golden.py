"""Generate tests/golden/*.npz from the CPU oracle (self-pins: they freeze the oracle's behaviour so a later
edit cannot silently change it; they do NOT validate it against TensorFlow -- parity is unpinned, see the
oracle header).  Run from the repo root:  python oracle/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import netvlad_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TINY = dict(iterations=16, cluster_size=8, hidden_size=32, vocab_size=50, rgb_dim=64, audio_dim=16)


def run(model, is_training, dtype):
    sp = O.param_specs(model, **TINY)
    P, S = O.init_params(sp, seed=1810, dtype=dtype, perturb=0.1)
    x, nf, labels = O.synthetic_batch(3, seed=20181000, max_frames=20, feat=80, vocab=50, dtype=dtype)
    kw = dict(vocab_size=50, iterations=16, cluster_size=8, is_training=is_training, rgb_dim=64, return_intermediates=True)
    reg = 0.0
    if model == "NetVladV1":
        pred, inter = O.netvlad_v1(x, nf, P, S, rgb_heads=4, audio_heads=2, **kw)
    elif model == "WillowModelReg":
        u = np.random.RandomState(7).rand(3, 16).astype(np.float32)          # the tf.random_uniform draws (:66-67)
        idx = O.sample_random_frame_indices(nf.numpy(), u)
        pred, inter = O.willow_model_reg(x, nf, P, S, frame_index=idx, **kw)
        inter["frame_index"] = torch.from_numpy(idx)
        reg = O.willow_regularization(P, 1e-2, 1e-2)
    else:
        g = torch.Generator().manual_seed(5)
        masks = {"video": (torch.rand(3, 16, 64, generator=g) >= 0.9).to(dtype),
                 "audio": (torch.rand(3, 16, 16, generator=g) >= 0.9).to(dtype)}
        pred, inter = O.netvlad_v2(x, nf, P, S, dropout_masks=masks, **kw)
    loss = O.cross_entropy_loss(pred, labels)
    (loss + reg).backward()
    out = {"pred": pred.detach().numpy(), "loss": np.array(float(loss.detach())), "num_frames": nf.numpy()}
    if model == "WillowModelReg":
        out["reg"] = np.array(float(reg.detach()))
    for k, v in inter.items():
        out["inter/" + k] = v.detach().numpy()
    for k in ("hidden1_weights", {"NetVladV1": "video_VLAD/cluster_weights", "NetVladV2": "video_VLAD/cluster_centers",
                                  "WillowModelReg": "video_VLAD/cluster_weights2"}[model], "input_bn/gamma", "gates/weights"):
        out["grad/" + k] = P[k].grad.numpy()
    for k in ("input_bn/moving_variance", "gating_bn/moving_mean"):
        out["state/" + k] = S[k].numpy()
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for model in ("NetVladV1", "NetVladV2", "WillowModelReg"):
        for is_training in (False, True):
            d = run(model, is_training, torch.float64)
            name = f"{model}_{'train' if is_training else 'infer'}_tiny_f64.npz"
            np.savez_compressed(os.path.join(OUT, name), **d)
            print(name, {k: v.shape for k, v in list(d.items())[:3]})
    idx = {f"T{T}": O.sample_uniform_indices(np.arange(0, 301), T) for T in (30, 64, 256, 300)}
    np.savez_compressed(os.path.join(OUT, "sample_indices.npz"), **idx)
    print("sample_indices.npz")
