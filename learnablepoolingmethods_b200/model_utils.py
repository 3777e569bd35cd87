"""model_utils.SampleUniformFrames (model_utils.py:101-122) and the caller prelude (train.py:262-264) on the GPU."""
from __future__ import annotations

import torch

from . import ops


def SampleUniformFrames(model_input, num_frames, num_samples):
    """Deterministic uniform sampling: frame_index[b, i] = int32(fl32(i / num_samples) * fl32(num_frames[b])).
    model_input: fp32 [B, max_frames, F] on the GPU; returns fp16 [B, num_samples, F] (the hot path consumes
    the sampled frames as fp16 operands; NetVladV1/V2 fuse this gather with input_bn)."""
    B, _, F = model_input.shape
    nf = num_frames.reshape(-1).to(device=model_input.device, dtype=torch.int32).contiguous()
    one = torch.ones(F, dtype=torch.float32, device=model_input.device)
    zero = torch.zeros(F, dtype=torch.float32, device=model_input.device)
    y = ops.sample_bn_apply(model_input.contiguous().float(), nf, int(num_samples), one, zero)
    return y.view(B, int(num_samples), F)


def l2_normalize_frames(model_input):
    """tf.nn.l2_normalize(model_input, 2) as train.py:262-264 / eval.py:140-143 / export_model.py:91-92 apply it before
    `create_model`: fp32 [B, max_frames, F] on the GPU -> same shape (zero-padded frames stay zero)."""
    if not model_input.is_cuda:
        raise RuntimeError("model_input must live on the GPU (there is no CPU path)")
    return ops.l2_normalize_frames(model_input.contiguous().float())
