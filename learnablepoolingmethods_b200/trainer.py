"""Training step of the reference trainer for the hot path (train.py:193-345, utils.py:170-213).

One process per GPU.  Reference semantics kept: `--batch_size` is per tower (= per rank), batch-norm
statistics are per tower, gradients are SUMMED over towers (utils.py:205-211 -> NCCL all-reduce SUM,
bucketed and overlapped with the backward), per-tensor clip_by_norm after the sum (utils.py:170-189),
Adam, learning rate decayed on global examples (train.py:244-249).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch

from . import ops
from .dp import BucketedAllReduce, FactorGather
from .engine import NetVladConfig, NetVladEngine, no_gc_during_capture, nvtx_range
from .variables import VariableStore

CHUNK = 32768           # elements per optimiser chunk
ALIGN = 32              # segment alignment inside the flat buffers (elements)


class FlatState:
    """Parameters, gradients and Adam moments as views of four flat fp32 buffers (fixed addresses)."""

    def __init__(self, store: VariableStore, order: List[str], wd: Dict[str, float], factored=()):
        """`factored`: variables updated by ops.rank_adam_step straight from their gradient factors.  They live at
        the end of the parameter / moment buffers, have no gradient storage and no optimiser chunks."""
        tr = store.trainable()
        self.factored = [n for n in factored if n in tr]
        self.store, self.order = store, [n for n in order if n not in self.factored]
        missing = [n for n in tr if n not in self.order and n not in self.factored]
        self.order += missing + self.factored
        dev = store.device
        offs, off, g_total = {}, 0, 0
        for n in self.order:
            offs[n] = off
            off += (tr[n].numel() + ALIGN - 1) // ALIGN * ALIGN
            if n not in self.factored:
                g_total = off
        self.total, self.offsets, self.g_total = off, offs, g_total
        self.p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.g = torch.zeros(g_total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.v = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad_views: Dict[str, torch.Tensor] = {}
        self.moment_views: Dict[str, tuple] = {}
        table, chunk_begin = [], [0]
        for t, n in enumerate(self.order):
            numel, o = tr[n].numel(), offs[n]
            view = self.p[o:o + numel].view(tr[n].shape)
            view.copy_(tr[n])
            store.vars[n] = view                       # re-home the variable into the flat buffer
            if n in self.factored:
                self.moment_views[n] = (self.m[o:o + numel].view(tr[n].shape), self.v[o:o + numel].view(tr[n].shape))
                chunk_begin.append(len(table))
                continue
            self.grad_views[n] = self.g[o:o + numel].view(tr[n].shape)
            for c0 in range(0, numel, CHUNK):
                table.append((t, (o + c0) // ALIGN, min(CHUNK, numel - c0), c0 // ALIGN))
            chunk_begin.append(len(table))
        self.table = torch.tensor(table, dtype=torch.int32, device=dev)
        self.chunk_begin = torch.tensor(chunk_begin, dtype=torch.int32, device=dev)
        self.chunk_begin_host = chunk_begin
        self.wd = torch.tensor([wd.get(n, 0.0) for n in self.order], dtype=torch.float32, device=dev)
        nt = len(self.order)
        # partial sums | clip factors | norms | overflow flag (cleared at the start of every step) | skipped-step counter
        self.scratch = (torch.zeros(len(table), dtype=torch.float32, device=dev), torch.zeros(nt, dtype=torch.float32, device=dev),
                        torch.zeros(nt, dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev),
                        torch.zeros(1, dtype=torch.int32, device=dev))
        self.shadow = None
        store.mark_dirty()
        store.layout_version += 1                      # every trainable variable now lives at a new address

    def bind_shadows(self, engine):
        """Per-tensor fp16 shadow destinations so the Adam kernel refreshes the GEMM operands in the same pass."""
        engine.refresh_shadows(force=True)
        dev = self.store.device
        ptrs, cols, lds = [0] * len(self.order), [1] * len(self.order), [0] * len(self.order)
        idx = {n: i for i, n in enumerate(self.order)}
        for var, key, shape, col0 in engine.shadow_specs():
            dst = self.store.shadows[key]
            i = idx[var]
            ptrs[i] = dst.data_ptr() + col0 * 2
            cols[i] = self.store.vars[var].shape[-1]
            lds[i] = dst.stride(0)
        self.shadow = (torch.tensor(ptrs, dtype=torch.int64, device=dev).view(torch.uint64) if hasattr(torch, "uint64")
                       else torch.tensor(ptrs, dtype=torch.int64, device=dev),
                       torch.tensor(cols, dtype=torch.int32, device=dev), torch.tensor(lds, dtype=torch.int64, device=dev))

    def end_offset(self, name: str) -> int:
        return self.offsets[name] + self.store.vars[name].numel()


def head_gradient_span(flat: "FlatState"):
    """(end offset, ok): the flat-gradient prefix [0, end) that holds every gradient produced by the head of the backward
    (MoE, gating, hidden bias / batch norm) -- they are final before the modalities' backward starts, so their all-reduce
    can travel underneath it.  ok is False when the layout does not have them in front (then one all-reduce at the end)."""
    body = [n for n in flat.order if n not in flat.factored and n.startswith(("video_", "audio_", "input_bn"))]
    head = [n for n in flat.order if n not in flat.factored and n not in body]
    if not head or not body:
        return 0, False
    end = max(flat.end_offset(n) for n in head)
    return end, all(flat.offsets[n] >= end for n in body)


def head_tensor_range(flat: "FlatState"):
    """(n_tensors, n_chunks) of the leading tensors of the flat layout whose gradients the head of the backward produces
    (MoE, gating, hidden bias / batch norm), or None when the layout does not start with exactly those.  Clipping is per
    tensor (utils.py:181-188), so their clip + Adam may run as soon as the head of the backward is done."""
    end, ok = head_gradient_span(flat)
    if not ok:
        return None
    n = 0
    while n < len(flat.order) and flat.order[n] not in flat.factored and flat.end_offset(flat.order[n]) <= end:
        n += 1
    if n == 0 or any(flat.order[i] not in flat.factored and flat.offsets[flat.order[i]] < end for i in range(n, len(flat.order))):
        return None
    return n, flat.chunk_begin_host[n]


def shard_row_range(n_rows: int, world: int, rank: int):
    """Rows [r0, r1) of hidden1_weights [n_rows, H] owned by `rank` (equal contiguous shards)."""
    rows = n_rows // world
    return rank * rows, (rank + 1) * rows


def pack_descriptor_slices(vlad: torch.Tensor, world: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """vlad [B, Kd] of this tower -> send buffer [world, B, Kd/world]: slice w holds the descriptor COLUMNS that are the
    ROWS of hidden1_weights owned by rank w.  After all_to_all_single, recv.view(world*B, Kd/world) on rank r is
    A_all[:, rows_r] for the towers in rank order -- the left factor of dW[rows_r] = A_all[:, rows_r]^T G_all."""
    B, Kd = vlad.shape
    rows = Kd // world
    src = vlad.view(B, world, rows).transpose(0, 1)
    if out is None:
        return src.contiguous()
    out.copy_(src)
    return out


def three_stage_order(order: List[str]) -> List[str]:
    """Flat-gradient order for the three-stage NetVladV1 backward: head | rgb attention block | audio (all) | rgb pooling |
    input_bn -- the rgb pooling's gradients (video_VLAD/*) are the only ones still open after stage "body1"."""
    pool = [n for n in order if n.startswith("video_VLAD/")]
    inbn = [n for n in order if n.startswith("input_bn")]
    rest = [n for n in order if n not in pool and n not in inbn]
    return rest + pool + inbn


def body_gradient_spans(flat: "FlatState"):
    """(head_end, mid_end, ok) for the three-stage backward: [0, head_end) = head gradients, [head_end, mid_end) = final after
    "body1" (rgb attention block, audio modality), [mid_end, g_total) = rgb pooling + input_bn (final after "body2")."""
    head_end, ok = head_gradient_span(flat)
    tail = [n for n in flat.order if n not in flat.factored and (n.startswith("video_VLAD/") or n.startswith("input_bn"))]
    mid = [n for n in flat.order if n not in flat.factored and n.startswith(("video_attention", "audio_"))]
    if not ok or not tail or not mid:
        return head_end, flat.g_total, False
    mid_end = min(flat.offsets[n] for n in tail)
    good = all(flat.end_offset(n) <= mid_end for n in mid) and all(flat.offsets[n] >= head_end for n in mid)
    return head_end, mid_end, good


class ShardedHiddenUpdate:
    """Data-parallel update of hidden1_weights (85 % of the parameters) without moving its gradient or repeating its
    optimiser step on every rank.  Rank r owns rows [r*Kd/W, (r+1)*Kd/W) of W_h [Kd, H]:

      * after the forward the ranks exchange descriptor COLUMN slices (all-to-all, 43 MB/W received per peer): rank r
        ends up with A_all[:, rows_r] for every video of the global batch;  dLoss/dhidden (80 KB) is all-gathered;
      * dW[rows_r] = inv * A_all[:, rows_r]^T G_all  -- the tower-summed gradient (utils.py:205-211) of the shard, one
        tcgen05 GEMM;  tf.clip_by_norm (utils.py:181-188) needs the norm of the whole tensor: a scalar all-reduce;
      * clip + Adam run on the shard only (1/W of the 3.6 GB optimiser stream) and refresh the fp16 operand shard;
      * the fp16 shards are all-gathered in place into the GEMM operand; the gather rides under the next forward
        (the engine waits for it right before the hidden projection).

    fp32 master rows of other ranks go stale on purpose; `sync_master()` all-gathers them (checkpointing)."""

    def __init__(self, trainer, flat):
        import torch.distributed as dist
        self.dist, self.pg = dist, trainer.pg
        self.world, self.rank = dist.get_world_size(trainer.pg), dist.get_rank(trainer.pg)
        store, c = trainer.store, trainer.cfg
        Kd, H = c.vlad_dim, c.hidden_size
        self.rows, self.H = Kd // self.world, H
        r0, _ = shard_row_range(Kd, self.world, self.rank)
        dev = store.device
        w = store.vars["hidden1_weights"]
        m, v = flat.moment_views["hidden1_weights"]
        self.w, self.m_full, self.v_full = w, m, v
        self.w_s, self.m_s, self.v_s = w[r0:r0 + self.rows], m[r0:r0 + self.rows], v[r0:r0 + self.rows]
        self.wh16 = store.shadows["wh16"]
        assert self.wh16.is_contiguous() and tuple(self.wh16.shape) == (Kd, H)
        self.wh16_s = self.wh16[r0:r0 + self.rows]
        n = self.rows * H
        table = [(0, c0 // ALIGN, min(CHUNK, n - c0), c0 // ALIGN) for c0 in range(0, n, CHUNK)]
        self.table = torch.tensor(table, dtype=torch.int32, device=dev)
        self.wd1 = torch.zeros(1, dtype=torch.float32, device=dev)
        self.partial = torch.zeros(len(table), dtype=torch.float32, device=dev)
        self.sumsq, self.factor, self.norm = (torch.zeros(1, dtype=torch.float32, device=dev) for _ in range(3))
        self.flag = flat.scratch[3]
        ptrs = torch.tensor([self.wh16_s.data_ptr()], dtype=torch.int64, device=dev)
        self.shadow = (ptrs.view(torch.uint64) if hasattr(torch, "uint64") else ptrs,
                       torch.tensor([H], dtype=torch.int32, device=dev),
                       torch.tensor([self.wh16.stride(0)], dtype=torch.int64, device=dev))
        self.dw = torch.empty((self.rows, H), dtype=torch.float32, device=dev)
        self.send = self.recv = self.g_all = None
        self.h_a2a = self.h_gather = None

    @staticmethod
    def supported(cfg, world) -> bool:
        return world > 1 and cfg.vlad_dim % world == 0 and (cfg.vlad_dim // world) % 64 == 0 and cfg.hidden_size % 8 == 0

    def start_exchange(self, vlad):
        """vlad fp16 [B, Kd] of this tower -> every rank receives its row slice of every tower's descriptor."""
        B = vlad.shape[0]
        if self.send is None or self.send.shape[1] != B:
            self.send = torch.empty((self.world, B, self.rows), dtype=vlad.dtype, device=vlad.device)
            self.recv = torch.empty_like(self.send)
            self.g_all = torch.empty((self.world * B, self.H), dtype=torch.float16, device=vlad.device)
        pack_descriptor_slices(vlad, self.world, out=self.send)                    # pack (plumbing for the collective)
        self.h_a2a = self.dist.all_to_all_single(self.recv, self.send, group=self.pg, async_op=True)

    def wait_weights(self):
        if self.h_gather is not None:
            self.h_gather.wait()
            self.h_gather = None

    def step(self, dact16, inv, clip, lr_t):
        d = self.dist
        d.all_gather_into_tensor(self.g_all, dact16.contiguous(), group=self.pg)
        self.h_a2a.wait()
        a_s = self.recv.view(-1, self.rows)                                        # [W*B, rows] = A_all[:, my rows]
        ops.gemm(a_s, self.g_all, a_mn=True, b_mn=True, out_dtype=torch.float32, alpha=inv, out=self.dw)
        ops.shard_sqnorm(self.w_s, self.dw, self.table, self.wd1, self.partial, self.sumsq)
        d.all_reduce(self.sumsq, op=d.ReduceOp.SUM, group=self.pg)
        return lambda: self._apply(clip, lr_t)

    def _apply(self, clip, lr_t):
        ops.shard_adam(self.w_s, self.dw, self.m_s, self.v_s, self.table, self.wd1, self.sumsq, clip=clip, lr_t=lr_t,
                       factor=self.factor, norm=self.norm, flag=self.flag, shadow=self.shadow)
        self.h_gather = self.dist.all_gather_into_tensor(self.wh16, self.wh16_s, group=self.pg, async_op=True)

    def sync_master(self):
        """All-gather the fp32 master rows (and Adam moments) so that every rank holds the full tensors again."""
        self.wait_weights()
        r0, _ = shard_row_range(self.rows * self.world, self.world, self.rank)
        for full in (self.w, self.m_full, self.v_full):
            self.dist.all_gather_into_tensor(full, full[r0:r0 + self.rows], group=self.pg)


class Trainer:
    def __init__(self, engine: NetVladEngine, *, base_learning_rate=0.01, learning_rate_decay=0.95,
                 learning_rate_decay_examples=4000000, clip_gradient_norm=1.0, regularization_penalty=1.0,
                 batch_size: int = 80, process_group=None, bucket_elems: int = 48 * 1024 * 1024):
        """Keyword defaults are the reference's flag defaults (train.py:74-86: base_learning_rate 0.01,
        learning_rate_decay 0.95, learning_rate_decay_examples 4000000, clip_gradient_norm 1.0,
        regularization_penalty 1.0); the YT8M NetVLAD runs pass --base_learning_rate=0.0002 --learning_rate_decay=0.8."""
        self.engine, self.store, self.cfg = engine, engine.store, engine.cfg
        self.base_lr, self.decay, self.decay_examples = base_learning_rate, learning_rate_decay, learning_rate_decay_examples
        self.clip, self.reg_penalty, self.batch_size = clip_gradient_norm, regularization_penalty, batch_size
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.global_step = 0
        self.flat: Optional[FlatState] = None
        self.bucket_elems = bucket_elems
        self.reducer: Optional[BucketedAllReduce] = None
        self.rank_scratch = None
        # data parallel: hidden1_weights' gradient (85 % of the bytes) is summed from all-gathered factors
        self.gather = FactorGather(process_group) if self.world > 1 and self.gather_hidden_factors else None
        self.shard: Optional[ShardedHiddenUpdate] = None      # created with the flat state (first step)
        self.use_shard = (self.world > 1 and self.shard_hidden_update and self.gather_hidden_factors
                          and ShardedHiddenUpdate.supported(self.cfg, self.world))
        # single tower: forward + loss + backward (~130 launches) are captured in a CUDA graph after the first steps and
        # replayed with one launch; the optimiser (learning rate changes every step) stays outside the graph, and so do the per-step random draws
        # (WillowModelReg frame indices: a static buffer filled before the replay; NetVladV2 dropout: a device-side seed)
        self.graph = None
        self.graph_after = 2           # eager steps before the capture (flat optimiser state, workspaces, attributes)
        # single tower: the optimiser is captured too.  The step size lives in device memory (`lr_dev`, written before each
        # replay); the hidden1_weights update (85 % of the parameters, HBM-bound) forks onto a second stream right
        # after the head of the backward -- the last reader of its fp16 operand -- and runs underneath the tensor-bound
        # backward of the modalities as small CTAs that share the SMs with it (ops.rank_adam_step(tiled=True)).
        self.lr_dev = None
        self._prio = None              # (capture stream, optimiser stream, fork event, join event)
        self.rank_ws = None

    def _factored_hidden(self, batch: int) -> bool:
        """hidden1_weights (85 % of the parameters) is updated from its gradient factors on a single tower: with
        several towers the summed gradient is no longer rank-`batch` on this rank, so the dense all-reduce path
        is kept (optimiser)."""
        c = self.cfg
        return (self.world == 1 and c.vlad_dim % 8 == 0 and not self.disable_factored_hidden
                and ops.rank_adam_supported(batch, c.hidden_size))

    on_flat_created = None
    # capture the optimiser inside the single-tower step graph (see __init__).  Measured on B200 at config 1
    # (profiles/r2_adam_overlap.md): eager optimiser after the graph 3.86 ms; forked at the SAME priority as the backward
    # 3.76 ms; forked at a lower priority 4.07 ms (its CTAs are starved while multi-wave kernels have blocks pending and
    # the update ends up exposed at the join); shorter-lived CTAs (LPM_ADAM_SPLIT=2/4) 3.86 / 3.95 ms.  Hence -1 / 1.
    fuse_optimizer = os.environ.get("LPM_FUSE_OPT", "1") != "0"
    opt_priority = int(os.environ.get("LPM_OPT_PRIORITY", "-1"))
    # where the hidden1 update forks: "head" (after the head of the backward) | "attn" (right before the rgb attention-core
    # backward, which leaves a quarter of the register file free).  Measured: head 3.74 ms, attn 4.01 ms -- next to the
    # update's CTAs the issue-bound attention kernel loses more than the update gains.
    adam_fork = os.environ.get("LPM_ADAM_FORK", "head")
    three_stage = os.environ.get("LPM_DP_THREE_STAGE", "1") != "0"  # data parallel NetVladV1: body split around the rgb pooling
    adam_col_splits = int(os.environ.get("LPM_ADAM_SPLIT", "1"))   # column splits of the tiled hidden1 update (CTA lifetime)
    # clip + Adam of the head's own variables (MoE, gating: 43 % of the non-factored parameters) on the forked branch as well
    # -- legal, clipping is per tensor (utils.py:181-188), so there is no global norm to wait for -- measured on B200 at
    # config 1 (gpurun r2ad): 3.73 ms with it against 3.50 ms without: like the hidden1 update, an HBM-streaming kernel
    # next to the backward's GEMMs costs them more than the 85 us it removes from the tail.  Off; kept as a switch.
    early_head_adam = os.environ.get("LPM_EARLY_HEAD_ADAM", "0") != "0"
    disable_factored_hidden = False
    gather_hidden_factors = True
    shard_hidden_update = True

    def _hidden_dw(self, dact16, inv, out):
        """sum over ranks of inv * vlad_r^T dact_r from the gathered factors (one GEMM over world*B rows)."""
        self.gather.start("dact", dact16)
        a_all = self.gather.wait("vlad")
        g_all = self.gather.wait("dact")
        return ops.gemm(a_all, g_all, a_mn=True, b_mn=True, out_dtype=torch.float32, alpha=inv, out=out)

    def _dp_side_stream(self):
        if self._dp_side is None:
            self._dp_side = torch.cuda.Stream(device=self.store.device)
        return self._dp_side

    _dp_side = None

    def _priority_streams(self):
        if self._prio is None:
            dev = self.store.device
            # torch: lower number = higher priority; 0 is the lowest the device offers
            self._prio = (torch.cuda.Stream(device=dev, priority=-1), torch.cuda.Stream(device=dev, priority=self.opt_priority),
                          torch.cuda.Event(), torch.cuda.Event())
        return self._prio

    def _fork_hidden_update(self, ctx):
        """engine.backward's after-head hook (captured): clip norm + Adam of hidden1_weights on the low-priority stream."""
        _, opt, fork, join = self._priority_streams()
        fork.record(torch.cuda.current_stream())
        opt.wait_event(fork)
        with torch.cuda.stream(opt):
            f, hr = self.flat, self._head_range()
            if hr is not None:
                # the head's own variables (MoE, gating: 43 % of the non-factored parameters) are final too: their clip +
                # Adam leaves the tail of the step and runs under the modalities' backward
                ops.adam_clip_step(f.p, f.g, f.m, f.v, f.table, f.chunk_begin, f.wd, clip=self.clip, lr_t=0.0, scratch=f.scratch,
                                   shadow=f.shadow, lr_dev=self.lr_dev, tensor_range=(0, hr[0]), chunk_range=(0, hr[1]))
                ctx["_head_adam_done"] = hr
            self._factored_hidden_step(ctx, 0.0, lr_dev=self.lr_dev, tiled=self.adam_col_splits)()
            join.record(opt)
        ctx["_opt_join"] = join

    def _head_range(self):
        if not self.early_head_adam:
            return None
        if self._head_rng is None:
            self._head_rng = (head_tensor_range(self.flat),)
        return self._head_rng[0]

    _head_rng = None

    def _factored_hidden_step(self, ctx, lr_t, lr_dev=None, tiled=False):
        f = self.flat
        vlad, dact16, inv = ctx["hidden_factors"]
        R = vlad.shape[0]
        a, g = vlad, dact16
        if R % 8:                                    # tiny test batches: the Gram GEMMs want 16-byte rows
            Rp = (R + 7) // 8 * 8
            a = torch.zeros((Rp, vlad.shape[1]), dtype=vlad.dtype, device=vlad.device); a[:R] = vlad
            g = torch.zeros((Rp, dact16.shape[1]), dtype=dact16.dtype, device=vlad.device); g[:R] = dact16
        gram_a = ops.gemm(a, a, b_mn=False, out_dtype=torch.float32)
        gram_g = ops.gemm(g, g, b_mn=False, out_dtype=torch.float32)
        if self.rank_scratch is None:
            self.rank_scratch = torch.zeros(2, dtype=torch.float32, device=vlad.device)
        ops.rank_grad_clip(gram_a, gram_g, inv, self.clip, self.rank_scratch[0:1], self.rank_scratch[1:2], f.scratch[3])
        if self.rank_ws is None:      # private scratch: the tiled update runs concurrently with users of the shared one
            self.rank_ws = torch.empty(max(16, ops.rank_adam_workspace_bytes(R, dact16.shape[1])), dtype=torch.uint8,
                                       device=vlad.device)
        return lambda: ops.rank_adam_step(vlad, dact16, inv, self.rank_scratch[0:1], f.scratch[3],
                                          self.store.vars["hidden1_weights"], *f.moment_views["hidden1_weights"],
                                          self.store.shadows["wh16"], lr_t=lr_t, lr_dev=lr_dev, tiled=tiled,
                                          workspace=self.rank_ws)

    # -- learning rate (train.py:244-249, tf.train.exponential_decay staircase) -------------------
    def learning_rate(self) -> float:
        ex = self.global_step * self.batch_size * self.world
        return self.base_lr * self.decay ** math.floor(ex / self.decay_examples)

    def _wd(self) -> Dict[str, float]:
        # slim.l2_regularizer(moe_l2) * regularization_penalty is part of EVERY tower's final_loss (train.py:299-311) and
        # combine_gradients sums the towers (utils.py:205-211): the regulariser's gradient is num_towers * l2 * w
        l2 = self.cfg.moe_l2 * self.reg_penalty * self.world
        return {"gates/weights": l2, "experts/weights": l2}

    # -- gradient all-reduce (SUM), bucketed over the flat buffer, overlapped with the backward -----
    def _hook(self, name, g):
        """Eager data-parallel steps: a bucket may go out once every gradient in front of it in the flat LAYOUT is final.  The
        layout need not be the production order (three_stage_order puts the rgb pooling's gradients behind the audio
        modality's), so the final prefix is tracked by name."""
        if self.reducer is None:
            return
        self._produced.add(name)
        order = self._layout_names
        while self._hook_pos < len(order) and order[self._hook_pos] in self._produced:
            self._hook_pos += 1
        if self._hook_pos > 0:
            self.reducer.mark_done(self.flat.end_offset(order[self._hook_pos - 1]))

    def _hook_reset(self):
        self._produced, self._hook_pos = set(), 0
        self._layout_names = [n for n in self.flat.order if n not in self.flat.factored]

    use_graph = True

    def _graph_ok(self, model_input) -> bool:
        if not (self.use_graph and self.flat is not None and self.global_step >= self.graph_after and model_input.is_cuda):
            return False
        # several towers: only the sharded hidden update has all of its collectives outside forward / backward
        return self.world == 1 or (self.use_shard and self.shard is not None)

    def _graph_step(self, model_input, num_frames, labels_u8, frame_index):
        """Replay (capture on first use) forward + cross-entropy + backward on static input buffers; then the eager
        optimiser.  Same arithmetic as the eager step: the kernels and their order are identical.
        Single tower: one graph.  Data parallel: three graphs -- (A) frames -> descriptor, (B) head + loss, (C) backward --
        with the collectives in between, exactly where the eager step has them: the wait for the all-gathered fp16 weight
        shards before the hidden projection, the all-to-all of descriptor slices after the forward, the gradient all-reduce
        (the backward is itself split after the head so that the MoE / gating gradients travel under the modalities'
        backward)."""
        eng, f = self.engine, self.flat
        g = self.graph
        key = (tuple(model_input.shape), model_input.dtype)
        eng.refresh_shadows()                                       # no-op unless values changed outside train_step
        if g is not None and (g["key"] != key or g["layout"] != self.store.layout_version):
            g = self.graph = None                                   # batch shape or buffer addresses changed: capture again
        willow, dp = self.cfg.model == "WillowModelReg", self.world > 1
        if g is None:
            dev = model_input.device
            g = {"key": key, "x": torch.empty_like(model_input), "nf": torch.empty(num_frames.shape, dtype=torch.int32, device=dev),
                 "lab": torch.empty_like(labels_u8),
                 "idx": torch.zeros((model_input.shape[0], self.cfg.iterations), dtype=torch.int32, device=dev) if willow else None}
        g["x"].copy_(model_input, non_blocking=True)
        g["nf"].copy_(num_frames.to(torch.int32), non_blocking=True)
        g["lab"].copy_(labels_u8, non_blocking=True)
        if willow:
            if frame_index is None:      # tf.random_uniform: a fresh draw per step, made outside the graph
                frame_index = ops.random_frame_index(g["nf"], self.cfg.iterations, model_input.shape[1],
                                                     mode=0 if self.cfg.random_frames else 1, seed=0x5EED0000 + eng.draws)
                eng.draws += 1
            g["idx"].copy_(frame_index.to(torch.int32), non_blocking=True)
        from . import _lib
        if "graphs" not in g:
            def seg_a():
                _, ctx = eng.forward(g["x"], g["nf"], True, save_for_backward=True, frame_index=g["idx"],
                                     device_seed=self.cfg.model == "NetVladV2", head=False)
                ctx["reg_penalty"] = self.reg_penalty
                return ctx

            def seg_b(ctx):
                pred = eng.forward_head(ctx)
                loss, _ = ops.xent_fwd(pred, g["lab"])
                return loss, ops.xent_bwd(pred, g["lab"], 1.0 / pred.shape[0])

            def seg_c(ctx, dpred, stage=None, fused=False):
                if stage != "body":
                    ctx["factored_hidden"] = bool(f.factored)
                    ctx["grad_views"] = f.grad_views
                    if fused and f.factored:
                        # NetVladV1: fork right before the rgb attention-core backward (see engine._v1_modality_bwd);
                        # other models (or LPM_ADAM_FORK=head): right after the head of the backward
                        at_attn = self.cfg.model == "NetVladV1" and self.adam_fork == "attn"
                        ctx["before_attention_bwd_hook" if at_attn else "after_head_hook"] = self._fork_hidden_update
                eng.backward(ctx, dpred, stage=stage)

            # data parallel: the head's gradients (MoE, gating: the first ~40 % of the flat buffer) are all-reduced while
            # the modalities' backward runs; valid when the flat layout really has them in front
            g["head_end"], ok = head_gradient_span(f)
            g["split"] = dp and ok
            # NetVladV1: split the body once more -- everything but the rgb pooling's gradients is final after "body1", so
            # that all-reduce (13 M parameters) travels under the rgb pooling backward instead of after the step
            _, g["mid_end"], ok3 = body_gradient_spans(f)
            g["split3"] = g["split"] and ok3 and self.cfg.model == "NetVladV1" and self.three_stage

            eng.pre_head_hook = None                                 # the wait for the weight shards happens between graphs
            if self.shard is not None:
                self.shard.wait_weights()
            side = torch.cuda.Stream(device=model_input.device)
            side.wait_stream(torch.cuda.current_stream())
            snap = {k: v.clone() for k, v in self.store.vars.items() if k.endswith(("moving_mean", "moving_variance"))}
            draws0 = eng.draws                                       # warm-up / capture passes do not consume random draws
            with torch.cuda.stream(side):                            # warm-up on a side stream (allocator, attributes)
                c0 = seg_a()
                seg_c(c0, seg_b(c0)[1])
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for k, v in snap.items():                                # the warm-up pass must not count as a training step
                self.store.vars[k].copy_(v)
            n0 = _lib.launch_count
            graphs = [torch.cuda.CUDAGraph()]
            g["fused"] = fuse = (not dp) and self.fuse_optimizer
            if not dp:
                if fuse and self.lr_dev is None:
                    self.lr_dev = torch.zeros(1, dtype=torch.float32, device=model_input.device)
                # fused: captured on a high-priority stream so that the optimiser branch (priority 0) only takes what
                # the forward / backward kernels leave free
                cap = self._priority_streams()[0] if fuse else None
                with no_gc_during_capture(), torch.cuda.graph(graphs[0], stream=cap, capture_error_mode="thread_local"):
                    if fuse:
                        ops.step_begin(f.scratch[3], f.scratch[4])
                    g["ctx"] = seg_a()
                    g["loss"], dpred = seg_b(g["ctx"])
                    seg_c(g["ctx"], dpred, fused=fuse)
                    if fuse:
                        self._captured_optimizer(g["ctx"])
            else:
                with no_gc_during_capture(), torch.cuda.graph(graphs[0], capture_error_mode="thread_local"):
                    g["ctx"] = seg_a()
                graphs.append(torch.cuda.CUDAGraph())
                with no_gc_during_capture(), torch.cuda.graph(graphs[1], pool=graphs[0].pool(), capture_error_mode="thread_local"):
                    g["loss"], g["dpred"] = seg_b(g["ctx"])
                for stage in (("head", "body1", "body2") if g["split3"] else ("head", "body") if g["split"] else (None,)):
                    graphs.append(torch.cuda.CUDAGraph())
                    with no_gc_during_capture(), torch.cuda.graph(graphs[-1], pool=graphs[0].pool(),
                                                                  capture_error_mode="thread_local"):
                        seg_c(g["ctx"], g["dpred"], stage)
            g["graphs"], g["launches"] = graphs, _lib.launch_count - n0     # kernels one replay launches
            g["layout"] = self.store.layout_version
            _lib.launch_count = n0
            self.graph = g
            eng.draws = draws0
        if self.cfg.model == "NetVladV2":
            eng.seed_dev.fill_(2 * eng.draws)        # NetVladV2's dropout: a fresh mask per replay (engine.forward, device_seed)
            eng.draws += 1
        graphs = g["graphs"]
        if g["fused"]:
            self.lr_dev.fill_(self._lr_t())          # read by the captured Adam kernels
        elif dp:
            ops.step_begin(f.scratch[3], f.scratch[4])    # before anything of this step can raise the skip flag
            g["ctx"]["_latched"] = True
        graphs[0].replay()
        if dp:
            self.shard.wait_weights()                # the fp16 weight shards gathered under the forward
            graphs[1].replay()
            self.shard.start_exchange(g["ctx"]["head"]["vlad"])      # rides under the backward
            graphs[2].replay()
            d = torch.distributed
            if g["split"]:
                # the head of the backward is done: dLoss/dhidden exists and hidden1's fp16 operand has had its last reader.
                # The whole sharded update of hidden1_weights (all-gather of dLoss/dhidden, the shard's gradient GEMM, norm
                # all-reduce, clip + Adam on the shard, start of the fp16 all-gather) runs on a second stream underneath the
                # modalities' backward, like the single-tower step's forked update.
                cur = torch.cuda.current_stream()
                side = self._dp_side_stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    _, dact16, inv = g["ctx"]["hidden_factors"]
                    self.shard.step(dact16, inv, self.clip, self._lr_t())()
                g["ctx"]["_hidden_done"] = True
                h0 = d.all_reduce(f.g[:g["head_end"]], op=d.ReduceOp.SUM, group=self.pg, async_op=True)
                graphs[3].replay()
                if g["split3"]:
                    h1 = d.all_reduce(f.g[g["head_end"]:g["mid_end"]], op=d.ReduceOp.SUM, group=self.pg, async_op=True)
                    graphs[4].replay()
                    h2 = d.all_reduce(f.g[g["mid_end"]:], op=d.ReduceOp.SUM, group=self.pg, async_op=True)
                    h0.wait()
                    h1.wait()
                    h2.wait()
                else:
                    h1 = d.all_reduce(f.g[g["head_end"]:], op=d.ReduceOp.SUM, group=self.pg, async_op=True)
                    h0.wait()
                    h1.wait()
                cur.wait_stream(side)
            else:
                d.all_reduce(f.g, op=d.ReduceOp.SUM, group=self.pg)
        _lib.launch_count += g["launches"]
        return g["loss"], g["ctx"]

    @nvtx_range("lpm.train_step")
    def train_step(self, model_input, num_frames, labels_u8, frame_index=None):
        """One step on this rank's tower batch.  Returns the label loss (device scalar, fp32).
        frame_index: optional int32 [B, iterations] (WillowModelReg: replaces the random frame draw)."""
        eng = self.engine
        if self._graph_ok(model_input):
            captured = self.graph is not None and "graphs" in self.graph
            try:
                loss, ctx = self._graph_step(model_input, num_frames, labels_u8, frame_index)
                if self.graph["fused"]:                    # the optimiser ran inside the graph
                    self.store.version += 1
                    self.store.shadow_version = self.store.version
                    self.global_step += 1
                    return loss
                return self._optimizer_step(ctx, bool(self.flat.factored), loss)
            except Exception as e:                         # noqa: BLE001
                if captured or self.world > 1:
                    raise                                  # a failing replay is a real error; under data parallelism the
                                                           # ranks must issue the same collectives, so no silent switch
                # the CAPTURE failed (driver / allocator state): keep training with the same kernels issued eagerly
                import sys
                import traceback
                print(f"lpm-b200: CUDA-graph capture of the training step failed ({e!r}); continuing eagerly", file=sys.stderr)
                if os.environ.get("LPM_DEBUG"):
                    traceback.print_exc()
                self.use_graph, self.graph = False, None
                torch.cuda.synchronize()
                if self.shard is not None:
                    eng.pre_head_hook = self.shard.wait_weights
        pred, ctx = eng.forward(model_input, num_frames, True, save_for_backward=True, frame_index=frame_index)
        ctx["reg_penalty"] = self.reg_penalty
        B = pred.shape[0]
        if self.use_shard and self.shard is not None:
            self.shard.start_exchange(ctx["head"]["vlad"])       # rides under the whole backward
        elif self.gather is not None and not self.use_shard:
            self.gather.start("vlad", ctx["head"]["vlad"])
            ctx["hidden_dw"] = self._hidden_dw
        loss, _ = ops.xent_fwd(pred, labels_u8)
        dpred = ops.xent_bwd(pred, labels_u8, 1.0 / B)
        order: List[str] = []
        factored = self._factored_hidden(B) or self.use_shard
        ctx["factored_hidden"] = factored
        if self.flat is None:
            ctx["grad_hook"] = lambda n, g: order.append(n)
            grads = eng.backward(ctx, dpred)
            if self.world > 1 and self.cfg.model == "NetVladV1":
                order = three_stage_order(order)        # lets the body's all-reduce start before the rgb pooling backward
            self.flat = FlatState(self.store, order, self._wd(), factored=("hidden1_weights",) if factored else ())
            self.flat.bind_shadows(eng)
            if self.on_flat_created is not None:        # checkpoint.load_into_store: restored Adam moments
                self.on_flat_created(self.flat)
                self.on_flat_created = None
            for n, g in grads.items():
                self.flat.grad_views[n].copy_(g.reshape(self.flat.grad_views[n].shape))
            if self.use_shard:
                self.shard = ShardedHiddenUpdate(self, self.flat)
                eng.pre_head_hook = self.shard.wait_weights
                self.shard.start_exchange(ctx["head"]["vlad"])
            if self.world > 1:
                skip = ()
                if self.gather is not None and not self.use_shard:
                    o = self.flat.offsets["hidden1_weights"]
                    skip = ((o, self.flat.end_offset("hidden1_weights")),)
                self.reducer = BucketedAllReduce(self.flat.g, self.bucket_elems, self.pg, skip=skip)
                self.reducer.flush()
        else:
            ctx["grad_views"] = self.flat.grad_views
            ctx["grad_hook"] = self._hook
            if self.reducer is not None:
                self.reducer.reset()
                self._hook_reset()
            eng.backward(ctx, dpred)
            if self.reducer is not None:
                self.reducer.flush()
        if self.reducer is not None:
            self.reducer.wait()
        return self._optimizer_step(ctx, factored, loss)

    def _lr_t(self) -> float:
        """TF Adam's bias-corrected step size for the step about to be applied (train.py:321-336)."""
        t = self.global_step + 1
        return self.learning_rate() * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)

    def _captured_optimizer(self, ctx):
        """Optimiser tail of the fused single-tower graph: clip + Adam of everything except hidden1_weights (whose update
        forked after the head of the backward), then the join with that branch and the two odd operand layouts."""
        f = self.flat
        hr = ctx.pop("_head_adam_done", None)
        if hr is None:
            ops.adam_clip_step(f.p, f.g, f.m, f.v, f.table, f.chunk_begin, f.wd, clip=self.clip, lr_t=0.0, scratch=f.scratch,
                               shadow=f.shadow, lr_dev=self.lr_dev)
        else:
            nt, nc = len(f.chunk_begin_host) - 1, f.chunk_begin_host[-1]
            ops.adam_clip_step(f.p, f.g, f.m, f.v, f.table, f.chunk_begin, f.wd, clip=self.clip, lr_t=0.0, scratch=f.scratch,
                               shadow=f.shadow, lr_dev=self.lr_dev, tensor_range=(hr[0], nt - hr[0]),
                               chunk_range=(hr[1], nc - hr[1]))
        join = ctx.pop("_opt_join", None)
        if join is not None:
            torch.cuda.current_stream().wait_event(join)
        self.engine.refresh_small_shadows()

    @nvtx_range("lpm.optimizer_step")
    def _optimizer_step(self, ctx, factored, loss):
        eng = self.engine
        f = self.flat
        lr_t = self._lr_t()
        if factored != bool(f.factored):
            raise RuntimeError("the tower batch size changed across the factored-update limit after the first step")
        if not ctx.pop("_latched", False):
            ops.step_begin(f.scratch[3], f.scratch[4])   # a skip flag raised by the previous step is counted and cleared
        # norm first: it may raise the skip flag
        if ctx.pop("_hidden_done", False):
            hidden_update = None                         # already applied underneath the backward (_graph_step)
        elif self.use_shard:
            _, dact16, inv = ctx["hidden_factors"]
            hidden_update = self.shard.step(dact16, inv, self.clip, lr_t)
        else:
            hidden_update = self._factored_hidden_step(ctx, lr_t) if factored else None
        ops.adam_clip_step(f.p, f.g, f.m, f.v, f.table, f.chunk_begin, f.wd, clip=self.clip, lr_t=lr_t, scratch=f.scratch,
                           shadow=f.shadow)
        if hidden_update is not None:
            hidden_update()
        # the fp16 GEMM operands were refreshed by the Adam kernel; only the two odd layouts remain
        self.store.version += 1
        eng.refresh_small_shadows()
        self.store.shadow_version = self.store.version
        self.global_step += 1
        return loss

    def apply_gradients(self, grads: Optional[Dict[str, torch.Tensor]] = None):
        """Optimiser half of the reference loop (train.py:321-336: clip_gradient_norms + apply_gradients) for callers that
        own the autograd edge themselves -- `create_model(..., is_training=True)` + `CrossEntropyLoss` + `loss.backward()`
        (autograd.NetVladFunction).  grads: {variable name: fp32 gradient}; default: the variables' `.grad` (consumed and
        reset).  Dense path: every gradient, hidden1_weights included, goes through the flat multi-tensor clip + Adam."""
        if self.world > 1:
            raise NotImplementedError("apply_gradients is the single-tower registry path; data parallel runs use train_step")
        tr = self.store.trainable()
        if self.flat is None:
            if grads is None:
                grads = {n: tr[n].grad for n in tr if tr[n].grad is not None}
            with torch.no_grad():     # the variables may carry requires_grad from the autograd edge
                self.flat = FlatState(self.store, list(tr.keys()), self._wd(), factored=())
                self.flat.bind_shadows(self.engine)
        elif self.flat.factored:
            raise RuntimeError("this Trainer already runs the factored train_step path; use a separate Trainer")
        f = self.flat
        for n, view in f.grad_views.items():
            g = grads[n] if grads is not None else tr[n].grad
            if g is None:
                raise KeyError(f"no gradient for {n}")
            view.copy_(g.reshape(view.shape))
            if grads is None:
                tr[n].grad = None
        ops.step_begin(f.scratch[3], f.scratch[4])
        ops.adam_clip_step(f.p, f.g, f.m, f.v, f.table, f.chunk_begin, f.wd, clip=self.clip, lr_t=self._lr_t(), scratch=f.scratch,
                           shadow=f.shadow)
        self.store.version += 1
        self.engine.refresh_small_shadows()
        self.store.shadow_version = self.store.version
        self.global_step += 1

    def skipped_steps(self) -> int:
        """Optimiser steps dropped so far because a gradient norm was not finite (fp16 activation-gradient overflow).
        The flag is cleared on the device at the start of every step, so one overflow skips exactly one update of the
        tensors that follow it; hidden1_weights (its own clip norm, its own stream) and the rest skip independently.
        global_step counts attempted steps, as a TF run would (the reference has no skip logic).  Host sync."""
        if self.flat is None:
            return 0
        return int(self.flat.scratch[4].item()) + int(self.flat.scratch[3].item())

    _seen_skips = 0

    def overflowed(self) -> bool:
        """True if any step since the last call saw a non-finite gradient norm (host sync)."""
        n = self.skipped_steps()
        r = n > self._seen_skips
        self._seen_skips = n
        return r

    def sync_parameters(self):
        """Data parallel: make every rank's fp32 copy of hidden1_weights (and its Adam moments) complete again
        (rows owned by other ranks are only refreshed as fp16 GEMM operands during training).  Call before saving."""
        if self.shard is not None:
            self.shard.sync_master()
