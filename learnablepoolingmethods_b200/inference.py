"""Prediction CSV in the reference's `inference.py` format (inference.py:88-96, 182-196), top-k on the GPU.

    VideoId,LabelConfidencePairs
    <video id>,<label> <score> <label> <score> ...        (top_k pairs, descending score, "%i %g")

The reference picks the top_k classes with numpy.argpartition on a host copy of the [B, vocab] predictions and
sorts them in Python; here `lpm_eval_topk` ranks them on the device and only the [B, top_k] winners are copied
back (ties: lower class index first; the reference's order among exactly equal scores is unspecified)."""
from __future__ import annotations

import torch

from . import ops

HEADER = "VideoId,LabelConfidencePairs\n"


def format_lines(video_ids, predictions, top_k):
    """Generator of CSV lines for one batch.  video_ids: sequence of bytes or str; predictions: CUDA fp32 [B, vocab]."""
    if not predictions.is_cuda:
        raise RuntimeError("predictions must live on the GPU (there is no CPU path)")
    p = predictions.detach().float().contiguous()
    B, V = p.shape
    if len(video_ids) != B:
        raise ValueError(f"{len(video_ids)} video ids for {B} prediction rows")
    k = min(int(top_k), V)
    labels = torch.zeros((B, V), dtype=torch.uint8, device=p.device)
    tv, ti, _, _ = ops.eval_topk(p, labels, k)
    tv, ti = tv.cpu().numpy(), ti.cpu().numpy()
    for b in range(B):
        vid = video_ids[b]
        vid = vid.decode("utf-8") if isinstance(vid, (bytes, bytearray)) else str(vid)
        yield vid + "," + " ".join("%i %g" % (int(l), float(s)) for l, s in zip(ti[b], tv[b])) + "\n"


def write_predictions(out_file, batches, top_k=20):
    """out_file: path or text file object; batches: iterable of (video_ids, predictions).  Returns #examples."""
    own = isinstance(out_file, str)
    f = open(out_file, "w+") if own else out_file
    n = 0
    try:
        f.write(HEADER)
        for ids, pred in batches:
            for line in format_lines(ids, pred, top_k):
                f.write(line)
            f.flush()
            n += len(ids)
    finally:
        if own:
            f.close()
    return n
