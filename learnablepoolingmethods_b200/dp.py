"""Data-parallel gradient exchange: all-reduce SUM (utils.py:192-213 `combine_gradients` sums the towers)
over a flat gradient buffer, in fixed-size buckets that are launched as soon as the backward has produced
every gradient they cover, so the exchange overlaps the rest of the backward (NCCL over NVLink on the GPU
box; the same code runs on gloo/CPU tensors in the tests)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class BucketedAllReduce:
    def __init__(self, flat_grad: torch.Tensor, bucket_elems: int, group=None, skip=()):
        """`skip`: [(begin, end)] flat ranges that are NOT all-reduced (their cross-rank sum is formed another way,
        e.g. FactorGather for hidden1_weights)."""
        self.g, self.bucket, self.group = flat_grad, int(bucket_elems), group
        self.total = flat_grad.numel()
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.buckets: List[tuple] = []
        pos = 0
        for s0, s1 in sorted(skip) + [(self.total, self.total)]:
            while pos < s0:
                end = min(s0, pos + self.bucket)
                self.buckets.append((pos, end))
                pos = end
            pos = max(pos, s1)
        self.handles: List = []
        self.launched: List[tuple] = []      # (begin, end) of every bucket launched this step, in order
        self.reset()

    def reset(self):
        self.done_upto, self.next = 0, 0     # self.next indexes self.buckets
        self.handles, self.launched = [], []

    def mark_done(self, end_offset: int):
        """The gradients of flat range [0, end_offset) are final (the buffer is laid out in the order the
        backward produces them)."""
        self.done_upto = max(self.done_upto, int(end_offset))
        self._launch(final=False)

    def flush(self):
        self.done_upto = self.total
        self._launch(final=True)

    def _launch(self, final: bool):
        if self.world == 1:
            return
        while self.next < len(self.buckets):
            b0, b1 = self.buckets[self.next]
            if not final and self.done_upto < b1:
                break
            self.handles.append(dist.all_reduce(self.g[b0:b1], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.launched.append((b0, b1))
            self.next += 1

    def wait(self):
        for h in self.handles:
            h.wait()
        self.handles = []


class FactorGather:
    """Cross-rank SUM of a rank-B weight gradient dW = alpha * A^T G without moving dW:
    sum_r A_r^T G_r = [A_0; ...; A_{W-1}]^T [G_0; ...; G_{W-1}], so the ranks all-gather the factors
    (hidden1_weights: A = the fp16 descriptor [B, 270336], 43 MB per rank, available right after the forward;
    G = dLoss/dhidden [B, 512]) instead of all-reducing the 554 MB fp32 gradient, and every rank forms the summed
    gradient with one GEMM over W*B rows.  utils.py:205-211 (sum over towers) is unchanged in value."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.bufs = {}
        self.handles = {}

    def start(self, key: str, local: torch.Tensor) -> torch.Tensor:
        local = local.contiguous()
        shape = (self.world * local.shape[0],) + tuple(local.shape[1:])
        buf = self.bufs.get(key)
        if buf is None or tuple(buf.shape) != shape or buf.dtype != local.dtype:
            buf = torch.empty(shape, dtype=local.dtype, device=local.device)
            self.bufs[key] = buf
        self.handles[key] = dist.all_gather_into_tensor(buf, local, group=self.group, async_op=True)
        return buf

    def wait(self, key: str) -> torch.Tensor:
        h = self.handles.pop(key, None)
        if h is not None:
            h.wait()
        return self.bufs[key]
