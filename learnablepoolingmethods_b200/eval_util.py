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


GAP_KERNEL_LIMIT = 16384     # lpm_eval_metrics ranks the B*k triplets of a batch in one CTA's shared memory


def batch_metrics(predictions, actuals, top_k=20):
    """Device tensor [3] = (hit@1, PERR, GAP) of the batch: one fused evaluation, no host transfer.
    Batches with B*k > 16384 (the reference's default --batch_size 1024 at k = 20) exceed the single-CTA ranking of
    `lpm_eval_metrics`: hit@1 / PERR are then the means of the per-video statistics `lpm_eval_topk` already produced, and
    the global ranking behind GAP (average_precision_calculator.py:203-262) is one stable device sort of the B*k scores."""
    p, a = _prep(predictions, actuals)
    tv, _, tl, rs = ops.eval_topk(p, a, top_k)
    if tv.numel() <= GAP_KERNEL_LIMIT:
        return ops.eval_metrics(tv, tl, rs)
    hit, perr, numpos = rs[:, 0].mean(), rs[:, 1].mean(), rs[:, 2].sum()
    order = torch.sort(tv.reshape(-1), descending=True, stable=True).indices        # ties keep (video, rank) order
    lab = tl.reshape(-1)[order].to(torch.float64)
    ranks = torch.arange(1, lab.numel() + 1, device=lab.device, dtype=torch.float64)
    gap = (torch.cumsum(lab, 0) / ranks * lab).sum() / numpos.to(torch.float64).clamp_min(1.0)
    gap = torch.where(numpos > 0, gap, torch.zeros_like(gap))
    return torch.stack([hit, perr, gap.to(torch.float32)])


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
