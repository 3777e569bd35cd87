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


def _gather(model_input, frame_index):
    B, _, F = model_input.shape
    S = frame_index.shape[1]
    x = model_input.contiguous() if model_input.dtype == torch.uint8 else model_input.contiguous().float()
    one = torch.ones(F, dtype=torch.float32, device=x.device)
    zero = torch.zeros(F, dtype=torch.float32, device=x.device)
    return ops.gather_bn_apply(x, frame_index, S, one, zero).view(B, S, F)


def SampleRandomFrames(model_input, num_frames, num_samples, uniform=None, seed=0):
    """model_utils.py:54-73: `num_samples` frames drawn independently and uniformly from the first num_frames[b] frames.
    Returns fp16 [B, num_samples, F] (uint8 codes are dequantised + L2-normalised on the way).  `uniform` (fp32
    [B, num_samples] in [0,1)) replaces the generator (tf.random_uniform's values in a parity test)."""
    nf = num_frames.reshape(-1).to(device=model_input.device, dtype=torch.int32).contiguous()
    idx = ops.random_frame_index(nf, int(num_samples), model_input.shape[1], mode=0, uniform=uniform, seed=seed)
    return _gather(model_input, idx)


def SampleRandomSequence(model_input, num_frames, num_samples, uniform=None, seed=0):
    """model_utils.py:26-51: a random window of `num_samples` consecutive frames (clipped to the last frame)."""
    nf = num_frames.reshape(-1).to(device=model_input.device, dtype=torch.int32).contiguous()
    idx = ops.random_frame_index(nf, int(num_samples), model_input.shape[1], mode=1, uniform=uniform, seed=seed)
    return _gather(model_input, idx)
