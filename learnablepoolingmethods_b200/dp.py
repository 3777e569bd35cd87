"""Data-parallel gradient exchange: all-reduce SUM (utils.py:192-213 `combine_gradients` sums the towers)
over a flat gradient buffer, in fixed-size buckets that are launched as soon as the backward has produced
every gradient they cover, so the exchange overlaps the rest of the backward (NCCL over NVLink on the GPU
box; the same code runs on gloo/CPU tensors in the tests)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class BucketedAllReduce:
    def __init__(self, flat_grad: torch.Tensor, bucket_elems: int, group=None):
        self.g, self.bucket, self.group = flat_grad, int(bucket_elems), group
        self.total = flat_grad.numel()
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.handles: List = []
        self.launched: List[tuple] = []      # (begin, end) of every bucket launched this step, in order
        self.reset()

    def reset(self):
        self.done_upto, self.next = 0, 0
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
        while self.next < self.total:
            b0, b1 = self.next, min(self.total, self.next + self.bucket)
            if not final and self.done_upto < b1:
                break
            self.handles.append(dist.all_reduce(self.g[b0:b1], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.launched.append((b0, b1))
            self.next = b1

    def wait(self):
        for h in self.handles:
            h.wait()
        self.handles = []
