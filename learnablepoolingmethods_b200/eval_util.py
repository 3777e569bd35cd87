"""Batch evaluation metrics with the reference's names (eval_util.py:27-135), computed on the GPU.

The reference calls these with numpy copies of `predictions` / `labels` on every logged training step
(train.py:448-449) and in the eval loop (eval.py); here the inputs are CUDA tensors, the work runs in
`lpm_eval_topk` / `lpm_eval_metrics`, and only the scalar(s) asked for are read back."""
from __future__ import annotations

import torch

from . import ops


def _prep(predictions, actuals):
    if not predictions.is_cuda:
        raise RuntimeError("predictions must live on the GPU (there is no CPU path)")
    p = predictions.detach().float().contiguous()
    a = actuals.to(device=p.device)
    a = (a != 0).to(torch.uint8).contiguous() if a.dtype != torch.uint8 else a.contiguous()
    return p, a


def batch_metrics(predictions, actuals, top_k=20):
    """Device tensor [3] = (hit@1, PERR, GAP) of the batch: one fused evaluation, no host transfer."""
    p, a = _prep(predictions, actuals)
    tv, _, tl, rs = ops.eval_topk(p, a, top_k)
    return ops.eval_metrics(tv, tl, rs)


def calculate_hit_at_one(predictions, actuals):
    """eval_util.py:27-42."""
    return float(batch_metrics(predictions, actuals)[0])


def calculate_precision_at_equal_recall_rate(predictions, actuals):
    """eval_util.py:45-70."""
    return float(batch_metrics(predictions, actuals)[1])


def calculate_gap(predictions, actuals, top_k=20):
    """eval_util.py:73-91."""
    return float(batch_metrics(predictions, actuals, top_k)[2])


def top_k_triplets(predictions, labels, k=20):
    """eval_util.py:128-135 for a whole batch: (class index [B,k] int32, prediction [B,k], label [B,k] uint8),
    ranked by descending prediction (the reference's argpartition leaves the k entries unordered)."""
    p, a = _prep(predictions, labels)
    tv, ti, tl, _ = ops.eval_topk(p, a, k)
    return ti, tv, tl
