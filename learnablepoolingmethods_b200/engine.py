"""Execution engine for the NetVladV1 / NetVladV2 hot path.

Sequences the C-ABI kernels (liblpm_b200.so) for the forward pass and -- in training -- records what
the hand-written backward needs.  torch owns memory and the autograd edge at the model boundary
(`NetVladFunction`); all arithmetic runs in the CUDA kernels.  There is no fallback path.

Reference lines mirrored: frame_level_models.py:2222-2377 (V1), :2383-2513 (V2), :2765-2824 (NetVLAD),
transformer_utils.py:374-457,507-767, video_level_models.py:48-159, model_utils.py:101-122; baseline NetVLAD
(SURVEY 8f row 4): frame_level_models.py:2516-2635 (WillowModelReg), video_pooling_modules.py:1499-1586
(NetVladOrthoReg), module_utils.py:55-90 (orthogonal regulariser), model_utils.py:26-73 (random frame sampling).
"""
from __future__ import annotations

import os

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

from . import ops
from .variables import VariableStore


class no_gc_during_capture:
    """Cyclic garbage collection stays off while a CUDA graph is being captured.  A collection that happens to run in the
    middle of a capture can destroy an OLD torch.cuda.CUDAGraph (e.g. a previous Trainer caught in a reference cycle);
    its destructor resets the graph -- an operation that is not permitted while a stream is capturing and that
    invalidates the capture in progress.  Collect first, then capture with the collector disabled."""

    def __enter__(self):
        import gc
        self._was = gc.isenabled()
        gc.collect()
        gc.collect()
        gc.disable()
        return self

    def __exit__(self, *exc):
        import gc
        if self._was:
            gc.enable()
        return False


import functools
import os

_NVTX = os.environ.get("LPM_NVTX", "0") == "1"


def nvtx_range(name):
    """LPM_NVTX=1: wrap a stage of the engine in an NVTX range (nsys / ncu --nvtx timelines; SURVEY section 5 tracing).
    Off by default: the decorator then returns the function unchanged (no per-call cost)."""
    def deco(fn):
        if not _NVTX:
            return fn

        @functools.wraps(fn)
        def wrapped(*a, **kw):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **kw)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapped
    return deco


def _ceil8(n: int) -> int:
    return (n + 7) // 8 * 8


@dataclass
class NetVladConfig:
    model: str = "NetVladV1"
    iterations: int = 256
    cluster_size: int = 256
    hidden_size: int = 512
    vocab_size: int = 3862
    num_mixtures: int = 2
    rgb_dim: int = 1024
    audio_dim: int = 128
    rgb_heads: int = 64        # frame_level_models.py:2285
    audio_heads: int = 16      # frame_level_models.py:2297
    add_batch_norm: bool = True
    gating: bool = True
    remove_diag: bool = False
    netvlad_relu: bool = False     # --netvlad_relu: hidden1_bn instead of hidden1_biases, then relu6 (:2321-2340)
    moe_l2: float = 1e-8
    d5_raw_reshape: bool = False   # SURVEY.md defect D5 switch (raw reinterpret instead of transpose)
    dropout_rate: float = 0.9      # D7: tf.layers.dropout(rate=1-0.1) in TransformerEncoderMod
    loss_scale: float = 0.0        # fp16 activation-gradient scale inside the backward; 0 = auto (8 x batch)
    hidden_splits: int = 74        # split-K factor of the hidden projection (2 N-tiles x 74 = 148 CTAs)
    # K3: reduce the split-K partials and run the context gating in the tail of the hidden-projection GEMM kernel (one launch
    # instead of GEMM + reduce + gate GEMM + gating; the gate product in exact fp32).  "infer" (default): inference / eval
    # forward only; "all" (LPM_FUSED_GATING=1): training too; "off" (=0): the four-launch path.  Training keeps the
    # four-launch head by default: the gain there is 10 us of 3.5 ms, and a different summation order moves the training
    # trajectory (same accuracy per step, another trained state: DESIGN.md section 2 lists both).
    fused_gating: str = {"1": "all", "0": "off"}.get(os.environ.get("LPM_FUSED_GATING", "infer"), "infer")
    split_hidden_infer: bool = True  # NetVladV1 inference: hidden projection with split-precision operands (see _head)
    overlap_audio: bool = True     # run the audio modality (0.4 % of the FLOPs, ~1/3 of the launches) on a second stream
    overlap_wgrad: bool = True     # rgb weight-gradient GEMMs on a third stream: they have no consumer before the optimiser
                                   # and fill the SMs the data-gradient GEMMs leave idle (partial last waves, 128-tile grids)
    # WillowModelReg only (frame_level_models.py:2209-2216, 2535-2544)
    rgb_det_reg: float = 1e-4      # --rgb_det_reg: orthogonal-regulariser scale of the rgb cluster centres
    audio_det_reg: float = 1e-4    # --audio_det_reg
    random_frames: bool = True     # SampleRandomFrames (True) or SampleRandomSequence (False)

    def modalities(self):
        ka = self.cluster_size // 4          # D7: integer division (frame_level_models.py:2263)
        return (("video", 0, self.rgb_dim, self.cluster_size, self.rgb_heads, "encode1"),
                ("audio", self.rgb_dim, self.audio_dim, ka, self.audio_heads, "encode2"))

    @property
    def feature_size(self):
        return self.rgb_dim + self.audio_dim

    @property
    def vlad_dim(self):
        return sum(D * K for _, _, D, K, _, _ in self.modalities())


class NetVladEngine:
    def __init__(self, cfg: NetVladConfig, store: VariableStore):
        self.cfg = cfg
        self.store = store
        if cfg.model not in ("NetVladV1", "NetVladV2", "WillowModelReg"):
            raise ValueError(f"unknown model {cfg.model}")
        if not cfg.add_batch_norm:
            # frame_level_models.py:2236 `add_batch_norm or FLAGS...` can never be False and the no-BN
            # gating branch raises in TF.
            raise NotImplementedError("netvlad_add_batch_norm=False is unreachable in the reference (D7)")
        self.build_variables()
        self._side = None          # second CUDA stream + fork / join events (created on first use)
        self._wside = None         # third stream for the rgb weight-gradient GEMMs of the backward
        self.draws = 0             # training/eval forwards so far: keys the dropout / frame-sampling generators
        self.seed_dev = None       # device copy of 2 * draws for graph-captured steps (NetVladV2 dropout)
        self.pre_head_hook = None  # callable run right before the hidden projection reads its fp16 weights

    def _wgrad_stream(self):
        if self._wside is None:
            self._wside = (torch.cuda.Stream(device=self.store.device, priority=-1), torch.cuda.Event(), torch.cuda.Event())
        return self._wside

    def _side_stream(self):
        if self._side is None:
            # priority -1 (high) like the trainer's capture stream: only the optimiser branch runs at the low priority 0
            self._side = (torch.cuda.Stream(device=self.store.device, priority=-1), torch.cuda.Event(), torch.cuda.Event())
        return self._side

    # ------------------------------------------------------------------------------------------
    # variables (names = TF variable names)
    # ------------------------------------------------------------------------------------------
    def build_variables(self):
        c, s = self.cfg, self.store
        s.batch_norm_vars("input_bn", c.feature_size)
        for name, _, D, K, H, sid in c.modalities():
            with s.variable_scope(name + "_VLAD"):
                if c.model == "WillowModelReg":
                    # NetVladOrthoReg: the scope id is glued to the variable name (video_pooling_modules.py:1527-1531)
                    s.get_variable("cluster_weights" + self.wc_suffix(name), (D, K), "normal", 1 / math.sqrt(D))
                    s.batch_norm_vars("cluster_bn", K)
                    s.get_variable("cluster_weights2", (D, K), "normal", 1 / math.sqrt(D))
                elif c.model == "NetVladV1":
                    s.get_variable("cluster_weights", (D, K), "normal", 1 / math.sqrt(D))
                    s.batch_norm_vars("cluster_bn", K)
                    s.get_variable("cluster_weights2", (1, D, K), "normal", 1 / math.sqrt(D))
                else:
                    with s.variable_scope("cluster_attention"):
                        self._dense_vars("q", D, D, False); self._dense_vars("k", D, D, False); self._dense_vars("v", D, D, False)
                        s.batch_norm_vars("logits_bn", c.iterations)
                        s.batch_norm_vars("attention_bn", D)
                        self._dense_vars("output_transform", D, D, True)
                        self._ln_vars("LayerNorm", D)
                        self._dense_vars("filter_outputencode", D, 4 * D, True)
                        s.batch_norm_vars("filter_bn", 4 * D)
                        self._dense_vars("ff_outputencode", 4 * D, K, True)
                        s.batch_norm_vars("feed_output_bn", K)
                    s.get_variable("cluster_centers", (D, K), "normal", 1 / math.sqrt(D))
            if c.model == "NetVladV1":
                with s.variable_scope(name + "_attention"):
                    self._dense_vars("q", D, D, False); self._dense_vars("k", D, D, False); self._dense_vars("v", D, D, False)
                    self._dense_vars("output_transform", D, D, True)
                    for ln in ("LayerNorm", "LayerNorm_1", "LayerNorm_2"):
                        self._ln_vars(ln, D)
                    self._dense_vars("filter_output" + sid, D, 4 * D, True)
                    self._dense_vars("ff_output" + sid, 4 * D, D, True)
        s.get_variable("hidden1_weights", (c.vlad_dim, c.hidden_size), "normal", 1 / math.sqrt(c.cluster_size))
        if c.netvlad_relu:
            s.batch_norm_vars("hidden1_bn", c.hidden_size)        # `add_batch_norm and relu` (:2321-2327)
        else:
            s.get_variable("hidden1_biases", (c.hidden_size,), "normal", 0.01)
        s.get_variable("gating_weights_2", (c.hidden_size, c.hidden_size), "normal", 1 / math.sqrt(c.hidden_size))
        s.batch_norm_vars("gating_bn", c.hidden_size)
        V, M = c.vocab_size, c.num_mixtures
        with s.variable_scope("gates"):
            s.get_variable("weights", (c.hidden_size, V * (M + 1)), "glorot")
        with s.variable_scope("experts"):
            s.get_variable("weights", (c.hidden_size, V * M), "glorot")
            s.get_variable("biases", (V * M,), "zeros")

    def wc_suffix(self, name: str) -> str:
        """frame_level_models.py:2552-2557: scope ids "netvlad_rgb_scope" / "netvlad_audio_scope"."""
        if self.cfg.model != "WillowModelReg":
            return ""
        return "netvlad_rgb_scope" if name == "video" else "netvlad_audio_scope"

    def _dense_vars(self, scope, i, o, bias):
        s = self.store
        with s.variable_scope(scope):
            s.get_variable("kernel", (i, o), "glorot")
            if bias:
                s.get_variable("bias", (o,), "zeros")

    def _ln_vars(self, scope, c):
        s = self.store
        with s.variable_scope(scope):
            s.get_variable("beta", (c,), "zeros")
            s.get_variable("gamma", (c,), "ones")

    # ------------------------------------------------------------------------------------------
    # fp16 operand shadows
    # ------------------------------------------------------------------------------------------
    def shadow_specs(self):
        """[(variable name, shadow key, shadow shape, column offset)]: fp16 row-major copies (possibly one
        column block of a wider, zero-padded shadow) consumed as tcgen05 operands."""
        c = self.cfg
        specs = []
        for name, _, D, K, H, sid in c.modalities():
            vs = name + "_VLAD"
            if c.model == "WillowModelReg":
                specs.append((vs + "/cluster_weights" + self.wc_suffix(name), vs + "/wc16", (D, K), 0))
                continue
            if c.model == "NetVladV1":
                specs.append((vs + "/cluster_weights", vs + "/wc16", (D, K), 0))
                a = name + "_attention"
                w1, w2, n2 = a + "/filter_output" + sid, a + "/ff_output" + sid, D
            else:
                a = vs + "/cluster_attention"
                w1, w2, n2 = a + "/filter_outputencode", a + "/ff_outputencode", K
            for i, n in enumerate(("q", "k", "v")):
                specs.append((f"{a}/{n}/kernel", a + "/wqkv16", (D, 3 * D), i * D))
            specs.append((a + "/output_transform/kernel", a + "/wo16", (D, D), 0))
            specs.append((w1 + "/kernel", a + "/w1_16", (D, 4 * D), 0))
            specs.append((w2 + "/kernel", a + "/w2_16", (4 * D, _ceil8(n2)), 0))
        specs.append(("hidden1_weights", "wh16", (c.vlad_dim, c.hidden_size), 0))
        specs.append(("gating_weights_2", "wg16", (c.hidden_size, c.hidden_size), 0))
        V, M = c.vocab_size, c.num_mixtures
        g8, e8 = _ceil8(V * (M + 1)), _ceil8(V * M)
        specs.append(("gates/weights", "wmoe16", (c.hidden_size, g8 + e8), 0))
        specs.append(("experts/weights", "wmoe16", (c.hidden_size, g8 + e8), g8))
        return specs

    # split-precision operands of the head (gate + MoE products): key -> [3*rows, cols] = [hi ; hi ; lo]; the plain fp16
    # shadow (used by the backward, refreshed by the optimiser) is the view of the first `rows` rows
    _SPLIT_SHADOWS = ("wg16", "wmoe16")

    def _shadow_buf(self, key, shape, dtype=torch.float16):
        sh = self.store.shadows
        if key in self._SPLIT_SHADOWS:
            big = self._shadow_buf(key + "x3", (3 * shape[0], shape[1]), dtype)
            t = sh.get(key)
            if t is None or t.data_ptr() != big.data_ptr() or tuple(t.shape) != tuple(shape):
                sh[key] = big[:shape[0]]
            return sh[key]
        t = sh.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.zeros(shape, dtype=dtype, device=self.store.device)
            sh[key] = t
            self.store.layout_version += 1
        return t

    def refresh_small_shadows(self):
        """Layouts that are not plain fp16 copies: transposed fp32 cluster centres, padded MoE bias."""
        s, c = self.store, self.cfg
        v, sh = s.vars, s.shadows
        for name, _, D, K, H, sid in c.modalities():
            vs = name + "_VLAD"
            src = (v[vs + "/cluster_weights2"][0] if c.model == "NetVladV1" else
                   v[vs + "/cluster_weights2"] if c.model == "WillowModelReg" else v[vs + "/cluster_centers"])
            o32, o16 = sh.get(vs + "/centers_t"), sh.get(vs + "/centers_t16")
            if o32 is None or tuple(o32.shape) != (K, D):
                o32 = o16 = None
                s.layout_version += 1
            # refreshed IN PLACE once they exist: a captured training / inference graph holds their addresses
            sh[vs + "/centers_t"], sh[vs + "/centers_t16"] = ops.transpose_f32_dual(src, out32=o32, out16=o16)
        V, M = c.vocab_size, c.num_mixtures
        g8, e8 = _ceil8(V * (M + 1)), _ceil8(V * M)
        bm = self._shadow_buf("bmoe", (g8 + e8,), torch.float32)
        bm[g8:g8 + V * M].copy_(v["experts/biases"])
        sh["moe_g8"] = g8
        # split-precision weight operands of the gate and MoE products: [hi ; hi ; lo] along the reduction
        if "wg16x3" in sh and "wmoe16x3" in sh:
            ops.split_hi_lo(v["gating_weights_2"], sh["wg16x3"], along_rows=True)
            ops.split_hi_lo(v["gates/weights"], sh["wmoe16x3"], along_rows=True)
            ops.split_hi_lo(v["experts/weights"], sh["wmoe16x3"][:, g8:], along_rows=True)

    def refresh_hidden_lo(self):
        """Low-order fp16 half of hidden1_weights (W = fp16(W) + wh16lo) for the split-precision inference projection;
        recomputed lazily when the variables changed since the last inference call (one 0.8 GB pass)."""
        s = self.store
        if s.__dict__.get("_hidden_lo_version") == s.version and "wh16lo" in s.shadows:
            return
        w = s.vars["hidden1_weights"]
        lo = self._shadow_buf("wh16lo", tuple(w.shape))
        ops.split_hi_lo(w, lo, along_rows=2)
        s._hidden_lo_version = s.version

    def refresh_shadows(self, force=False):
        s = self.store
        if not force and s.shadow_version == s.version:
            return s.shadows
        for var, key, shape, col0 in self.shadow_specs():
            src = s.vars[var]
            dst = self._shadow_buf(key, shape)
            ops.cast_f16(src, dst[:, col0:col0 + min(shape[1] - col0, _ceil8(src.shape[1]))],
                         cols_dst=min(shape[1] - col0, _ceil8(src.shape[1])))
        self.refresh_small_shadows()
        s.shadow_version = s.version
        return s.shadows

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    @nvtx_range("lpm.forward")
    def forward(self, model_input: torch.Tensor, num_frames: torch.Tensor, is_training: bool,
                save_for_backward: bool = False, dropout_masks=None, return_intermediates: bool = False,
                frame_index=None, device_seed: bool = False, head: bool = True):
        """model_input fp32 [B, max_frames, rgb+audio] (L2-normalised by the caller, train.py:264), or the uint8
        codes [B, max_frames, rgb+audio] as decoded by the reader (readers.py:185-193): those are dequantised
        (utils.py:28-43) and L2-normalised inside the gather kernels.
        num_frames int [B].  frame_index (WillowModelReg only): int32 [B, iterations] gather indices replacing the
        random draw of model_utils.py:26-73.  Returns (predictions fp32 [B, vocab], ctx).
        head=False stops after the concatenated descriptor (returns (None, ctx)); `forward_head(ctx)` finishes the pass:
        the data-parallel trainer replays the two halves as separate CUDA graphs around the wait for the weight shards."""
        c, s = self.cfg, self.store
        v = s.vars
        sh = self.refresh_shadows()
        if model_input.dim() != 3 or model_input.shape[2] != c.feature_size:
            raise ValueError(f"model_input must be [B, frames, {c.feature_size}], got {tuple(model_input.shape)}")
        if not model_input.is_cuda:
            raise RuntimeError("model_input must live on the GPU (there is no CPU path)")
        x = model_input.contiguous() if model_input.dtype == torch.uint8 else model_input.contiguous().float()
        nf = num_frames.to(device=x.device, dtype=torch.int32).contiguous()
        B, T, F = x.shape[0], c.iterations, c.feature_size
        ctx: Dict[str, object] = {"B": B, "training": is_training, "inter": {}, "seed": self.draws, "_head": head,
                                  "_save": save_for_backward, "_want_inter": return_intermediates}
        self.draws += 1
        if device_seed:
            # captured in a CUDA graph: the per-call part of the dropout seed lives in device memory (2 * draw counter,
            # written by the caller before each replay), so that eager and replayed steps draw the same masks
            if self.seed_dev is None:
                self.seed_dev = torch.zeros(1, dtype=torch.int64, device=x.device)
            ctx["seed"], ctx["seed_dev"] = 0, self.seed_dev
        save = save_for_backward
        if c.model == "WillowModelReg":
            return self._willow_forward(x, nf, B, T, is_training, save, ctx, return_intermediates, frame_index)
        if frame_index is not None:
            raise ValueError("frame_index applies to WillowModelReg only (NetVladV1/V2 sample uniformly)")

        # ---- a2 + a3: uniform frame sampling + input_bn -> fp16 frames [B*T, F] -----------------
        if is_training:
            part = ops.sample_bn_stats(x, nf, T)
            r = ops.bn_finalize(part[:, 0], part[:, 1], B * T, v["input_bn/gamma"], v["input_bn/beta"],
                                v["input_bn/moving_mean"], v["input_bn/moving_variance"], training=True,
                                bessel=True, save=save, psum_stride=2 * F)
        else:
            r = ops.bn_finalize(None, None, 1, v["input_bn/gamma"], v["input_bn/beta"], v["input_bn/moving_mean"],
                                v["input_bn/moving_variance"], training=False, bessel=True, save=save)
        in_scale, in_shift = r[0], r[1]
        if c.model == "NetVladV1":
            xb = ops.sample_bn_apply(x, nf, T, in_scale, in_shift)
            xmods = [xb[:, col0:col0 + D] for _, col0, D, _, _, _ in c.modalities()]
        else:
            # V2 adds the frames as a residual inside the encoder: contiguous per-modality matrices
            xb = None
            xmods = list(ops.sample_bn_apply(x, nf, T, in_scale, in_shift, split_col=c.rgb_dim))
        if save:
            ctx["xb"] = xb
            ctx["input_bn_stats"] = r[2]

        # inference (NetVladV1): rows [B, 2B) hold the low-order halves of the descriptor, vlad = hi + lo (see _head)
        split = c.model == "NetVladV1" and c.split_hidden_infer and not is_training and not save
        vlad2 = torch.empty(((2 if split else 1) * B, c.vlad_dim), dtype=torch.float16, device=x.device)
        vlad, vlad_lo = vlad2[:B], (vlad2[B:] if split else None)
        ctx["_vlad2"] = vlad2 if split else None
        off = 0
        main = torch.cuda.current_stream()
        side, ev_fork, ev_join = self._side_stream() if c.overlap_audio else (None, None, None)
        if side is not None:
            ev_fork.record(main)
        for (name, col0, D, K, H, sid), X in zip(c.modalities(), xmods):
            # the two modalities are independent between the sampled frames and the concatenated descriptor
            # (frame_level_models.py:2273-2309): the audio one is launch- not throughput-bound, so it rides along
            # on a second stream underneath the rgb kernels
            on_side = side is not None and name == "audio"
            if on_side:
                side.wait_event(ev_fork)
            with torch.cuda.stream(side if on_side else main):
                if c.model == "NetVladV1":
                    m = self._v1_modality(name, X, B, T, D, K, H, sid, is_training, save, vlad[:, off:off + K * D], ctx,
                                          return_intermediates, out_lo=None if vlad_lo is None else vlad_lo[:, off:off + K * D])
                else:
                    mask = None if dropout_masks is None else dropout_masks.get(name)
                    m = self._v2_modality(name, X, B, T, D, K, is_training, save, vlad[:, off:off + K * D], ctx,
                                          return_intermediates, mask)
            if on_side:
                ev_join.record(side)
            if save:
                ctx[name] = m
            off += K * D
        if side is not None:
            main.wait_event(ev_join)

        return self._finish(vlad, ctx)

    def _finish(self, vlad, ctx):
        if not ctx["_head"]:
            ctx["_vlad"] = vlad
            return None, ctx
        return self._head(vlad, ctx["B"], ctx["training"], ctx["_save"], ctx, ctx["_want_inter"]), ctx

    def forward_head(self, ctx):
        """Second half of a forward started with head=False: hidden projection, gating, MoE -> predictions."""
        return self._head(ctx.pop("_vlad"), ctx["B"], ctx["training"], ctx["_save"], ctx, ctx["_want_inter"])

    @nvtx_range("lpm.v1_modality")
    def _v1_modality(self, name, X, B, T, D, K, H, sid, training, save, out_view, ctx, want_inter, out_lo=None):
        c, v, sh = self.cfg, self.store.vars, self.store.shadows
        vs, a = name + "_VLAD", name + "_attention"
        m: Dict[str, object] = {}
        wc16 = sh[vs + "/wc16"]
        # ---- cluster_bn statistics over all B*T rows (recompute pass, nothing stored) ------------
        bn = vs + "/cluster_bn"
        if training:
            _, st = ops.gemm(wc16, X, a_mn=True, b_mn=False, out="none", stats=True)    # S^T = Wc^T X^T
            r = ops.bn_finalize(st[0].reshape(-1, K), st[1].reshape(-1, K), B * T, v[bn + "/gamma"], v[bn + "/beta"],
                                v[bn + "/moving_mean"], v[bn + "/moving_variance"], training=True, bessel=True, save=save)
        else:
            r = ops.bn_finalize(None, None, 1, v[bn + "/gamma"], v[bn + "/beta"], v[bn + "/moving_mean"],
                                v[bn + "/moving_variance"], training=False, bessel=True, save=save)
        lscale, lshift = r[0], r[1]
        # ---- K1: fused soft-assignment + aggregation + norms -----------------------------------
        z, rscale, a_sum, assign = ops.netvlad_pool_fwd(X, B, T, wc16, lscale, lshift, sh[vs + "/centers_t16"],
                                                        save_assign=save)
        if want_inter:
            ctx["inter"]["vlad_" + name] = ops.netvlad_finalize(z, rscale, d_major=True)
        # ---- a9: attention block over the K cluster descriptors (rows = (b,k), cols = d) -------
        Z2 = z.view(B * K, D)
        rs = rscale.view(B * K)
        if c.d5_raw_reshape:
            # literal reading of frame_level_models.py:2290-2292 (SURVEY defect D5): the normalised, d-major flattened
            # descriptor [B, D*K] is re-interpreted as [B, K, D] without a transpose
            zn = torch.empty((B * K, D), dtype=torch.float16, device=z.device)
            ops.netvlad_finalize_f16(z, rscale, zn.view(B, K * D), K * D)
            qkv = ops.gemm(zn, sh[a + "/wqkv16"])
            resid, resid_rs = zn, None
        elif save:
            # training keeps the normalised descriptor as a GEMM operand for the weight gradients
            zn = ops.scale_rows_f16(Z2, rs)
            qkv = ops.gemm(zn, sh[a + "/wqkv16"])
            resid, resid_rs = zn, None
        else:
            qkv = ops.gemm(Z2, sh[a + "/wqkv16"], row_scale=rs)
            resid, resid_rs = Z2, rs
        dh = D // H
        o = ops.mha_core_fwd(qkv, B, K, D, H, scale=dh ** -0.5, want_lse=save)
        lse = None
        if save:
            o, lse = o
        att = ops.gemm(o, sh[a + "/wo16"], bias=v[a + "/output_transform/bias"])
        h1 = ops.layernorm_joint_fwd(att, resid, resid_rs, B, K, D, v[a + "/LayerNorm/gamma"], v[a + "/LayerNorm/beta"], save=save)
        st1 = None
        if save:
            h1, st1 = h1
        f1 = ops.gemm(h1.view(B * K, D), sh[a + "/w1_16"], bias=v[f"{a}/filter_output{sid}/bias"], relu=True)
        f2 = ops.gemm(f1, sh[a + "/w2_16"], bias=v[f"{a}/ff_output{sid}/bias"], relu=True)
        la, lb = a + "/LayerNorm_1", a + "/LayerNorm_2"
        if ops.layernorm_chain_supported(K, D):
            # LN(f2 + h1) (FeedForwardNetwork, :712-713) and LN(. + h1) (TransformerEncoder, :410-411) in one pass
            rc = ops.layernorm_chain_fwd(f2, h1, B, K, D, v[la + "/gamma"], v[la + "/beta"], v[lb + "/gamma"], v[lb + "/beta"],
                                         out=out_view, out_stride=out_view.stride(0), save=save, out_lo=out_lo)
            u2, st2, h2, st3 = (rc[1], rc[2], rc[3], rc[4]) if save else (None, None, None, None)
        else:
            u2 = torch.empty_like(f2) if save else f2
            h2 = ops.layernorm_joint_fwd(f2, h1, None, B, K, D, v[la + "/gamma"], v[la + "/beta"], save=save, u_out=u2)
            st2 = None
            if save:
                h2, st2 = h2
            r3 = ops.layernorm_joint_fwd(h2, h1, None, B, K, D, v[lb + "/gamma"], v[lb + "/beta"],
                                         out=out_view, out_stride=out_view.stride(0), save=save)
            st3 = r3[1] if save else None
            if out_lo is not None:
                out_lo.zero_()          # the two-kernel fallback emits no low-order half
        if want_inter:
            ctx["inter"]["att_" + name] = out_view.float().reshape(B, K, D)
        if save:
            m.update(dict(X=X, z=z, rscale=rscale, a_sum=a_sum, assign=assign, cluster_bn_stats=r[2], lscale=lscale,
                          zn=zn, qkv=qkv, o=o, lse=lse, u1=att, st1=st1, h1=h1, f1=f1, f2=f2, u2=u2, st2=st2,
                          u3=h2, st3=st3))
        return m

    def _v2_modality(self, name, X, B, T, D, K, training, save, out_view, ctx, want_inter, dropout_mask):
        """NetVladAttenCluster.forward (video_pooling_modules.py:1617-1663): cluster similarities from
        TransformerEncoderMod over the frames (transformer_utils.py:443-457, 634-677, 737-767), then the same
        residual aggregation / norms as V1 through the pooling kernel's external-assignment mode."""
        c, v, sh = self.cfg, self.store.vars, self.store.shadows
        vs = name + "_VLAD"
        a = vs + "/cluster_attention"
        H = D // 16                                                   # video_pooling_modules.py:1612
        if K % 8:
            raise NotImplementedError("cluster sizes that are not multiples of 8")
        qkv = ops.gemm(X, sh[a + "/wqkv16"])                          # [B*T, 3D]; no q scaling (batch-normed logits)

        def bnv(scope):
            return (v[scope + "/gamma"], v[scope + "/beta"], v[scope + "/moving_mean"], v[scope + "/moving_variance"])

        if training:
            part = ops.mha_logit_stats(qkv, B, T, D, H)
            r = ops.bn_finalize(part[:, 0], part[:, 1], B * H * T, *bnv(a + "/logits_bn"), training=True, bessel=True,
                                psum_stride=2 * T, save=save)
        else:
            r = ops.bn_finalize(None, None, 1, *bnv(a + "/logits_bn"), training=False, bessel=True, save=save)
        ks, kb = r[0], r[1]
        lbn_stats = r[2] if save else None
        o = ops.mha_core_fwd(qkv, B, T, D, H, scale=1.0, key_scale=ks, key_shift=kb, want_lse=save)
        lse = None
        if save:
            o, lse = o
        o_bn = torch.empty_like(o) if save else None                  # rank-3 input: biased moving variance
        r = ops.batch_norm_cols_f16(o, *bnv(a + "/attention_bn"), training=training, bessel=False, save=save, out=o_bn)
        abn_stats = r[2] if save else None
        o_in = o_bn if save else o
        att = ops.gemm(o_in, sh[a + "/wo16"], bias=v[a + "/output_transform/bias"])
        mask = None
        if training and c.dropout_rate > 0:                           # D7: drop probability 0.9
            mask = torch.empty_like(att) if save else None
            ops.dropout_f16(att, c.dropout_rate, mask_in=dropout_mask, mask_out=mask,
                            seed=int(ctx.get("seed", 0)) * 2 + (name == "audio"), seed_dev=ctx.get("seed_dev"))
        h1 = ops.layernorm_joint_fwd(att, X, None, B, T, D, v[a + "/LayerNorm/gamma"], v[a + "/LayerNorm/beta"], save=save)
        st1 = None
        if save:
            h1, st1 = h1
        f = ops.gemm(h1.view(B * T, D), sh[a + "/w1_16"], bias=v[a + "/filter_outputencode/bias"], relu=True)
        f_bn = torch.empty_like(f) if save else None
        r = ops.batch_norm_cols_f16(f, *bnv(a + "/filter_bn"), training=training, bessel=False, save=save, out=f_bn)
        fbn_stats = r[2] if save else None
        f_in = f_bn if save else f
        f2 = ops.gemm(f_in, sh[a + "/w2_16"], bias=v[a + "/ff_outputencode/bias"], relu=True)     # [B*T, K]
        A = torch.empty_like(f2) if save else None
        r = ops.batch_norm_cols_f16(f2, *bnv(a + "/feed_output_bn"), training=training, bessel=False, save=save, out=A)
        obn_stats = r[2] if save else None
        A_in = A if save else f2
        z, rscale, a_sum, _ = ops.netvlad_pool_fwd(X, B, T, None, None, None, sh[vs + "/centers_t16"], assign_in=A_in)
        ops.netvlad_finalize_f16(z, rscale, out_view, out_view.stride(0))    # d-major flatten, normalised, fp16
        if want_inter:
            ctx["inter"]["vlad_" + name] = ops.netvlad_finalize(z, rscale, d_major=True)
            ctx["inter"]["assign_" + name] = A_in.float().reshape(B, T, K)
        if not save:
            return {}
        return dict(X=X, qkv=qkv, ks=ks, kb=kb, lbn_stats=lbn_stats, o=o, lse=lse, o_bn=o_bn, abn_stats=abn_stats,
                    mask=mask, u1=att, st1=st1, h1=h1, f=f, fbn_stats=fbn_stats, f_bn=f_bn, f2=f2, obn_stats=obn_stats,
                    A=A, z=z, rscale=rscale, a_sum=a_sum)

    # ------------------------------------------------------------------------------------------
    # WillowModelReg: baseline NetVLAD + context gating + orthogonal regulariser
    # ------------------------------------------------------------------------------------------
    def _willow_forward(self, x, nf, B, T, training, save, ctx, want_inter, frame_index):
        """frame_level_models.py:2516-2635: random frames -> input_bn -> NetVladOrthoReg per modality (K1 + d-major
        flatten, no attention block) -> shared head."""
        c, v, sh = self.cfg, self.store.vars, self.store.shadows
        F = c.feature_size
        if frame_index is None:
            # tf.random_uniform is redrawn on every session.run (training and evaluation alike)
            frame_index = ops.random_frame_index(nf, T, x.shape[1], mode=0 if c.random_frames else 1,
                                                 seed=0x5EED0000 + int(ctx["seed"]))
        else:
            frame_index = frame_index.to(device=x.device, dtype=torch.int32).contiguous()
            if tuple(frame_index.shape) != (B, T):
                raise ValueError(f"frame_index must be [{B}, {T}], got {tuple(frame_index.shape)}")
        bnv = (v["input_bn/gamma"], v["input_bn/beta"], v["input_bn/moving_mean"], v["input_bn/moving_variance"])
        if training:
            part = ops.gather_bn_stats(x, frame_index, T)
            r = ops.bn_finalize(part[:, 0], part[:, 1], B * T, *bnv, training=True, bessel=True, save=save, psum_stride=2 * F)
        else:
            r = ops.bn_finalize(None, None, 1, *bnv, training=False, bessel=True, save=save)
        xb = ops.gather_bn_apply(x, frame_index, T, r[0], r[1])
        if save:
            ctx["xb"], ctx["input_bn_stats"], ctx["frame_index"] = xb, r[2], frame_index
        vlad = torch.empty((B, c.vlad_dim), dtype=torch.float16, device=x.device)
        off = 0
        for name, col0, D, K, _, _ in c.modalities():
            vs = name + "_VLAD"
            X = xb[:, col0:col0 + D]
            wc16, bn = sh[vs + "/wc16"], vs + "/cluster_bn"
            bv = (v[bn + "/gamma"], v[bn + "/beta"], v[bn + "/moving_mean"], v[bn + "/moving_variance"])
            if training:
                _, st = ops.gemm(wc16, X, a_mn=True, b_mn=False, out="none", stats=True)
                rb = ops.bn_finalize(st[0].reshape(-1, K), st[1].reshape(-1, K), B * T, *bv, training=True, bessel=True, save=save)
            else:
                rb = ops.bn_finalize(None, None, 1, *bv, training=False, bessel=True, save=save)
            z, rscale, a_sum, assign = ops.netvlad_pool_fwd(X, B, T, wc16, rb[0], rb[1], sh[vs + "/centers_t16"],
                                                            save_assign=save)
            out_view = vlad[:, off:off + K * D]
            ops.netvlad_finalize_f16(z, rscale, out_view, out_view.stride(0))      # d-major flatten (:1583)
            if want_inter:
                ctx["inter"]["vlad_" + name] = ops.netvlad_finalize(z, rscale, d_major=True)
            if save:
                ctx[name] = dict(X=X, z=z, rscale=rscale, a_sum=a_sum, assign=assign, cluster_bn_stats=rb[2])
            off += K * D
        return self._finish(vlad, ctx)

    def regularization_loss(self):
        """REGULARIZATION_LOSSES of the model beyond the MoE weight decay (device scalar, fp32): the orthogonal
        regulariser of the two NetVladOrthoReg modules (module_utils.py:55-90), zero for NetVladV1/V2."""
        c, v = self.cfg, self.store.vars
        tot = torch.zeros(1, dtype=torch.float32, device=self.store.device)
        if c.model == "WillowModelReg":
            for name, scale in (("video", c.rgb_det_reg), ("audio", c.audio_det_reg)):
                if scale > 0:
                    tot += ops.ortho_reg(v[name + "_VLAD/cluster_weights2"], scale)
        return tot[0]

    def _willow_modality_bwd(self, ctx, name, col0, D, K, dv, dgamma_in, dbeta_in, put):
        """Backward of NetVladOrthoReg (same residual aggregation / soft-assignment backward as NetVladV1, entered from
        the d-major flatten) + the regulariser's gradient on cluster_weights2."""
        c, v, sh = self.cfg, self.store.vars, self.store.shadows
        inv = 1.0 / ctx["loss_scale"]
        m, B, T = ctx[name], ctx["B"], c.iterations
        vs = name + "_VLAD"
        f32 = torch.float32
        gout = ctx["_gout"]
        ct = sh[vs + "/centers_t"]
        dvh = ops.dmajor_to_kmajor_f16(dv, B, K, D)
        dz, q = ops.netvlad_norm_bwd(m["z"], m["rscale"], dvh, ct)
        X, wc16 = m["X"], sh[vs + "/wc16"]
        S16 = ops.gemm(X, wc16)                                               # logits recomputed: [B*T, K]
        X3 = ctx["xb"].view(B, T, -1)[:, :, col0:col0 + D]
        G = ops.gemm(X3, dz, b_mn=False, out_dtype=f32)                       # [B, T, K] = X dV^T per video
        bn = vs + "/cluster_bn"
        dS, dg, db = ops.assign_bwd(G.view(B * T, K), m["assign"].view(B * T, K), q, S16, m["cluster_bn_stats"],
                                    v[bn + "/gamma"], T, inv_scale=inv)
        put(bn + "/gamma", dg); put(bn + "/beta", db)
        wname = vs + "/cluster_weights" + self.wc_suffix(name)
        m_tiles = (D + 127) // 128
        parts = ops.gemm(X, dS, a_mn=True, b_mn=True, splits=max(2, min(64, 148 // max(1, m_tiles))))
        dWc = gout(wname)
        if dWc is None:
            dWc = torch.empty((D, K), dtype=f32, device=X.device)
        ops.splitk_reduce(parts, alpha=inv, out32=dWc)
        put(wname, dWc)
        dCt, E = ops.center_bwd(dz, m["z"], m["a_sum"], ct, v["input_bn/beta"][col0:col0 + D], inv)
        ops.input_bn_grad(v[wname], dWc, dCt, E, v["input_bn/gamma"][col0:col0 + D],
                          dgamma_in[col0:col0 + D], dbeta_in[col0:col0 + D])
        dC = ops.transpose_f32(dCt)                                           # [D, K]
        scale = c.rgb_det_reg if name == "video" else c.audio_det_reg
        if scale > 0:
            # loss = label_loss + regularization_penalty * reg_loss (train.py:301-303, 323-324)
            ops.ortho_reg(v[vs + "/cluster_weights2"], scale, dw=dC, grad_scale=float(ctx.get("reg_penalty", 1.0)),
                          want_value=False)
        put(vs + "/cluster_weights2", dC)

    @nvtx_range("lpm.head")
    def _head(self, vlad, B, training, save, ctx, want_inter):
        """frame_level_models.py:2309-2377 + video_level_models.py:48-159."""
        c, v, sh = self.cfg, self.store.vars, self.store.shadows
        Hn = c.hidden_size
        if self.pre_head_hook is not None:
            self.pre_head_hook()    # data parallel: the all-gather of the updated fp16 weight shards lands here
        vlad2 = ctx.get("_vlad2")
        fg = {True: "all", False: "off"}.get(c.fused_gating, c.fused_gating)      # booleans accepted too
        fused = (fg == "all" or (fg == "infer" and not training)) and c.gating and not c.netvlad_relu and B <= 128 and Hn % 4 == 0
        if fused:
            # K3 as one launch (ops.gemm_splitk_gated): every CTA of the split-K product stays for the tail -- grid barrier,
            # fixed-order sum of a column slice of all partials (+ bias), grid barrier, the gate product of its columns in fp32,
            # gating_bn over the batch rows, sigmoid, product, split-precision operand of the MoE product.
            sp = max(2, min(c.hidden_splits, 148 // -(-Hn // 256)))
            p1 = None
            a_in, w_in = vlad, sh["wh16"]
            if vlad2 is not None:
                # inference: hidden = [v_hi ; v_lo] W_hi + v_hi W_lo (see below); the first pass leaves its partials, the second
                # pass carries the tail and sums both
                self.refresh_hidden_lo()
                p1 = ops.gemm(vlad2, sh["wh16"], splits=max(2, c.hidden_splits))
                p1 = p1.view(p1.shape[0] * 2, B, Hn)
                w_in = sh["wh16lo"]
            diag = torch.diagonal(v["gating_weights_2"]).contiguous() if c.remove_diag else None
            act32, a3, gated32, g3, gstats, gates, _ = ops.gemm_splitk_gated(
                a_in, w_in, splits=sp, bias=v["hidden1_biases"], wg=v["gating_weights_2"], gamma=v["gating_bn/gamma"],
                beta=v["gating_bn/beta"], moving_mean=v["gating_bn/moving_mean"], moving_var=v["gating_bn/moving_variance"],
                training=training, wg_diag=diag, save=save, parts2=p1)
            act16, gated16 = a3[:, :Hn], g3[:, :Hn]
            hpre, hstats = None, None
            r = (gated32, g3, gstats, gates)
        else:
            act32 = torch.empty((B, Hn), dtype=torch.float32, device=vlad.device)
            hpre, hstats = None, None
            parts_lo = None
            if vlad2 is not None:
                # Inference: hidden = (v_hi + v_lo)(W_hi + W_lo) ~= [v_hi ; v_lo] W_hi + v_hi W_lo.  The 270 336-term products with
                # 11-bit operands are what is left of the prediction error once the gate / MoE products are split (trained-weights
                # protocol, DESIGN.md numerics); two passes over the fp16 weight bytes instead of one (+277 MB of HBM reads).
                # M = 2B rows run as one 2-CTA (cta_group::2) tile per split, so W_hi is still streamed once.
                self.refresh_hidden_lo()
                parts = ops.gemm(vlad2, sh["wh16"], splits=max(2, c.hidden_splits))        # [S, 2B, H]
                parts = parts.view(parts.shape[0] * 2, B, Hn)                              # hi and lo rows summed by the reduce
                parts_lo = ops.gemm(vlad, sh["wh16lo"], splits=max(2, c.hidden_splits))    # [S, B, H]
            else:
                parts = ops.gemm(vlad, sh["wh16"], splits=max(2, c.hidden_splits))
            if c.netvlad_relu:
                hpre = act32
                ops.splitk_reduce(parts, out32=hpre, parts2=parts_lo)
                r = ops.hidden_bn_relu6_fwd(hpre, v["hidden1_bn/gamma"], v["hidden1_bn/beta"], v["hidden1_bn/moving_mean"],
                                            v["hidden1_bn/moving_variance"], training=training, save=save)
                act32 = r[0]
                hstats = r[2] if save else None
                a3 = ops.split_hi_lo(act32)                               # [B, 3H] = [hi | lo | hi]
            else:
                # one kernel: sum of the partials (both passes), bias, fp32 activation and its split-precision operand
                a3 = torch.empty((B, 3 * Hn), dtype=torch.float16, device=vlad.device)
                ops.splitk_reduce(parts, bias=v["hidden1_biases"], out32=act32, out16=a3, parts2=parts_lo, split3=True)
            # The gate and MoE products run with split-precision operands (x = hi + lo in two fp16 terms, one GEMM over a 3x
            # longer reduction, ops.split_hi_lo): `hidden` is O(30) at this model's initialisation scale and feeds sigmoids, so
            # 10-bit-mantissa operands in these two small products are what limits the predictions (numerics).
            act16 = a3[:, :Hn]                                            # plain fp16 view for the backward
            if c.gating:
                # split-K partials of the gate product go straight into the gating kernel (summed there in a fixed order)
                gparts = ops.gemm(a3, sh["wg16x3"], splits=6)
                diag = torch.diagonal(v["gating_weights_2"]).contiguous() if c.remove_diag else None
                r = ops.gating_fwd(act32, gparts, v["gating_bn/gamma"], v["gating_bn/beta"], v["gating_bn/moving_mean"],
                                   v["gating_bn/moving_variance"], training=training, wg_diag=diag, save=save, split3=True)
                gated32, g3, gates = r[0], r[1], r[3].view(B, Hn)
            else:
                gates, gated32, g3, r = None, act32, a3, (None, None, None)
            gated16 = g3[:, :Hn]
        logits = ops.gemm(g3, sh["wmoe16x3"], bias=sh["bmoe"], out_dtype=torch.float32)
        pred = ops.moe_mix_fwd(logits, c.vocab_size, c.num_mixtures, expert_off=sh["moe_g8"])
        if want_inter:
            ctx["inter"].update(hidden=act32, gated=gated32)
        if save:
            ctx["head"] = dict(vlad=vlad, act32=act32, act16=act16, gates=gates, gating_stats=r[2] if c.gating else None,
                               gated32=gated32, gated16=gated16, logits=logits, pred=pred, hpre=hpre, hstats=hstats)
        return pred

    # ------------------------------------------------------------------------------------------
    # backward (NetVladV1): hand-written autodiff of the forward above
    # ------------------------------------------------------------------------------------------
    @nvtx_range("lpm.backward")
    def backward(self, ctx, dpred: torch.Tensor, stage: Optional[str] = None) -> Dict[str, torch.Tensor]:
        """dpred: fp32 [B, vocab] = dLoss/dpredictions.  Returns {variable name: fp32 gradient}.
        Activation gradients travel as fp16 scaled by cfg.loss_scale; parameter gradients are unscaled.
        stage: None = the whole backward; "head" = MoE, gating and hidden projection only (stops once dLoss/dvlad exists);
        "body" = the rest (after a "head" call on the same ctx); NetVladV1 can split the rest once more: "body1" = the
        attention blocks of both modalities and the audio pooling (every gradient except the rgb pooling's and input_bn's is
        final afterwards), "body2" = the rgb pooling backward.  The data-parallel trainer replays the two stages as
        separate CUDA graphs and starts the all-reduce of the head's gradients in between."""
        if stage == "body":
            return self._backward_body(ctx, *ctx.pop("_bwd_state"))
        if stage == "body1":      # NetVladV1: attention blocks of both modalities + the audio pooling (see _backward_body)
            return self._backward_body(ctx, *ctx["_bwd_state"], part="body1")
        if stage == "body2":      # the rgb pooling backward + input_bn
            return self._backward_body(ctx, *ctx.pop("_bwd_state"), part="body2")
        c, v, sh = self.cfg, self.store.vars, self.store.shadows
        B, hd = ctx["B"], ctx["head"]
        # dLoss/dpred scales as 1/B; LayerNorm over the L2-normalised descriptor has rstd ~ 200, so the scale
        # is kept modest (see DESIGN.md "numerics"): auto = 8*B clamped to [8, 4096]
        S = float(c.loss_scale) if c.loss_scale else float(min(4096, max(8, 8 * B)))
        ctx["loss_scale"] = S
        inv = 1.0 / S
        V, M, Hn = c.vocab_size, c.num_mixtures, c.hidden_size
        g8 = sh["moe_g8"]
        f32 = torch.float32
        grads: Dict[str, torch.Tensor] = {}
        hook = ctx.get("grad_hook")
        deferred_hidden = None
        views = ctx.get("grad_views")       # optional {name: preallocated fp32 view} (flat gradient buffer)
        ctx["_gout"] = (lambda n: views.get(n) if views is not None else None)
        gout = ctx["_gout"]

        def put(name, g):
            if views is not None and name in views:
                dst = views[name]
                if g.data_ptr() != dst.data_ptr():
                    dst.copy_(g.reshape(dst.shape))
                g = dst
            grads[name] = g
            if hook is not None:
                hook(name, g)

        # ---- MoE (video_level_models.py:86-126) ------------------------------------------------
        dl16 = ops.moe_mix_bwd(hd["logits"], dpred, V, M, g8, S)
        gated16 = hd["gated16"]
        put("experts/biases", ops.colsum(dl16[:, g8:], alpha=inv, cols=V * M))
        put("gates/weights", ops.gemm(gated16, dl16[:, :g8], a_mn=True, b_mn=True, out_dtype=f32, alpha=inv, N=V * (M + 1),
                                      out=gout("gates/weights")))
        put("experts/weights", ops.gemm(gated16, dl16[:, g8:], a_mn=True, b_mn=True, out_dtype=f32, alpha=inv, N=V * M,
                                        out=gout("experts/weights")))
        dgated = ops.gemm(dl16, sh["wmoe16"], b_mn=False, out_dtype=f32)
        # ---- context gating (frame_level_models.py:2342-2368) -----------------------------------
        if c.gating:
            diag = torch.diagonal(v["gating_weights_2"]).contiguous() if c.remove_diag else None
            gb = ops.gating_bwd(hd["act32"], hd["gates"], v["gating_bn/gamma"], v["gating_bn/beta"],
                                hd["gating_stats"], dgated, inv, wg_diag=diag)
            dact, dg16, dgam, dbet = gb[:4]
            put("gating_bn/gamma", dgam)
            put("gating_bn/beta", dbet)
            dwg = ops.gemm(hd["act16"], dg16, a_mn=True, b_mn=True, out_dtype=f32, alpha=inv, out=gout("gating_weights_2"))
            if c.remove_diag:
                ops.add_diag(dwg, gb[4])              # -sum_b dv*act on the diagonal (:2349-2352)
            put("gating_weights_2", dwg)
            ops.gemm(dg16, sh["wg16"], b_mn=False, out=dact, accumulate=True)
        else:
            dact = dgated
        # ---- hidden projection (frame_level_models.py:2314-2334) --------------------------------
        if c.netvlad_relu:
            dact, dgam, dbet = ops.hidden_bn_relu6_bwd(hd["hpre"], hd["act32"], dact, v["hidden1_bn/gamma"], hd["hstats"],
                                                       inv_scale=inv)
            put("hidden1_bn/gamma", dgam)
            put("hidden1_bn/beta", dbet)
        else:
            put("hidden1_biases", ops.colsum(dact, alpha=inv))
        dact16 = ops.cast_scaled_f16(dact)
        if ctx.get("factored_hidden"):
            # the trainer applies clip + Adam straight from the factors (dW = inv * vlad^T dact16 is never written)
            ctx["hidden_factors"] = (hd["vlad"], dact16, inv)
        elif ctx.get("hidden_dw") is not None:
            # data parallel: the trainer sums this gradient over the ranks from all-gathered factors (dp.FactorGather).
            # The product is formed at the END of the backward so that the gather of the 43 MB-per-rank descriptors
            # (started right after the forward) has the whole backward to hide under.
            deferred_hidden = (dact16, inv)
        else:
            put("hidden1_weights", ops.gemm(hd["vlad"], dact16, a_mn=True, b_mn=True, out_dtype=f32, alpha=inv,
                                            out=gout("hidden1_weights")))
        dvlad = ops.gemm(dact16, sh["wh16"], b_mn=False)                      # [B, vlad_dim] fp16
        after_head = ctx.get("after_head_hook")
        if after_head is not None:
            # single tower: that product was the last reader of the fp16 hidden1 operand in this step, so the trainer forks
            # the factored update of hidden1_weights from here onto its low-priority stream (trainer._fork_hidden_update)
            after_head(ctx)
        if stage == "head":
            ctx["_bwd_state"] = (dvlad, grads, put, deferred_hidden)
            return grads
        return self._backward_body(ctx, dvlad, grads, put, deferred_hidden)

    @nvtx_range("lpm.backward_body")
    def _backward_body(self, ctx, dvlad, grads, put, deferred_hidden, part=None):
        c, v = self.cfg, self.store.vars
        f32 = torch.float32
        gout = ctx["_gout"]
        if part == "body2":
            # second half of the split NetVladV1 body: the rgb pooling backward (main stream only), then input_bn
            dgamma_in, dbeta_in = ctx.pop("_dbn_in")
            name, col0, D, K, H, sid = list(c.modalities())[0]
            self._v1_modality_bwd(ctx, name, col0, D, K, H, sid, dvlad[:, 0:K * D], dgamma_in, dbeta_in, put, part="pool")
            put("input_bn/gamma", dgamma_in)
            put("input_bn/beta", dbeta_in)
            if deferred_hidden is not None:
                put("hidden1_weights", ctx["hidden_dw"](deferred_hidden[0], deferred_hidden[1], gout("hidden1_weights")))
            return grads
        if part == "body1" and c.model != "NetVladV1":
            raise ValueError("the three-stage backward exists for NetVladV1 only")
        dgamma_in = torch.zeros(c.feature_size, dtype=f32, device=dvlad.device)
        dbeta_in = torch.zeros(c.feature_size, dtype=f32, device=dvlad.device)
        off = 0
        mods = list(c.modalities())
        offs = []
        for name, col0, D, K, H, sid in mods:
            offs.append(off)
            off += K * D
        main = torch.cuda.current_stream()
        side, ev_fork, ev_join = self._side_stream() if c.overlap_audio else (None, None, None)
        if side is not None:
            ev_fork.record(main)
        deferred = []                       # audio gradients are announced after the join (hooks may start an all-reduce)
        # weight gradients of the rgb attention block on their own stream -- only when nobody listens for finished
        # gradients while the backward runs (the eager data-parallel path starts all-reduces from `put`)
        ctx["_wgrad"] = None
        if c.overlap_wgrad and c.model == "NetVladV1" and ctx.get("grad_hook") is None:
            ctx["_wgrad"] = self._wgrad_stream() + ([],)
        for (name, col0, D, K, H, sid), o0 in reversed(list(zip(mods, offs))):
            on_side = side is not None and name == "audio"
            put_m = (lambda n, g: deferred.append((n, g))) if on_side else put
            if on_side:
                side.wait_event(ev_fork)
            with torch.cuda.stream(side if on_side else main):
                if c.model == "NetVladV1":
                    self._v1_modality_bwd(ctx, name, col0, D, K, H, sid, dvlad[:, o0:o0 + K * D], dgamma_in, dbeta_in, put_m,
                                          part="attn" if (part == "body1" and name == "video") else None)
                elif c.model == "WillowModelReg":
                    self._willow_modality_bwd(ctx, name, col0, D, K, dvlad[:, o0:o0 + K * D], dgamma_in, dbeta_in, put_m)
                else:
                    self._v2_modality_bwd(ctx, name, col0, D, K, dvlad[:, o0:o0 + K * D], dgamma_in, dbeta_in, put_m)
            if on_side:
                ev_join.record(side)
        if side is not None:
            main.wait_event(ev_join)
        if ctx["_wgrad"] is not None:
            wstream, _, wjoin, keep = ctx["_wgrad"]
            wjoin.record(wstream)
            main.wait_event(wjoin)
            keep.clear()                    # operands of the side-stream products may be recycled from here on
            ctx["_wgrad"] = None
        for n, g in deferred:
            put(n, g)
        if part == "body1":
            ctx["_dbn_in"] = (dgamma_in, dbeta_in)      # the rgb pooling's share is still to come (body2)
            return grads
        put("input_bn/gamma", dgamma_in)
        put("input_bn/beta", dbeta_in)
        if deferred_hidden is not None:
            put("hidden1_weights", ctx["hidden_dw"](deferred_hidden[0], deferred_hidden[1], gout("hidden1_weights")))
        return grads

    def _wgrad_run(self, ctx, name, operands, fn):
        """Parameter-gradient work without a consumer before the optimiser.  For the rgb modality it runs on the wgrad
        stream as a parallel branch (forked after its operands exist on the main stream, joined at the end of the
        backward); the operands stay referenced until the join so that their memory cannot be recycled underneath it."""
        w = ctx.get("_wgrad")
        if w is None or name != "video":
            return fn()
        wstream, wfork, _, keep = w
        wfork.record(torch.cuda.current_stream())
        wstream.wait_event(wfork)
        keep.extend(operands)
        with torch.cuda.stream(wstream):
            return fn()

    def _wgrad_gemm(self, ctx, name, a, b, **kw):
        """Weight-gradient product dW = a^T b (see _wgrad_run)."""
        return self._wgrad_run(ctx, name, (a, b), lambda: ops.gemm(a, b, a_mn=True, b_mn=True, **kw))

    @nvtx_range("lpm.v1_modality_bwd")
    def _v1_modality_bwd(self, ctx, name, col0, D, K, H, sid, dv, dgamma_in, dbeta_in, put, part=None):
        """part: None = the whole modality; "attn" = the attention block only (stops once the gradient of the normalised
        descriptor exists), "pool" = the NetVLAD normalisation / aggregation / soft-assignment backward after an "attn" call."""
        c, v, sh = self.cfg, self.store.vars, self.store.shadows
        S = ctx["loss_scale"]
        inv = 1.0 / S
        m, B, T = ctx[name], ctx["B"], c.iterations
        a, vs = name + "_attention", name + "_VLAD"
        f32 = torch.float32
        gout = ctx["_gout"]
        if part != "pool":
            rows = B * K
            # ---- LN3: out = LN(u3), u3 = h2 + h1 ----------------------------------------------------
            du3, dg, db = ops.layernorm_joint_bwd(m["u3"], dv, dv.stride(0), B, K, D, m["st3"], v[a + "/LayerNorm_2/gamma"],
                                                  inv_scale=inv)
            put(a + "/LayerNorm_2/gamma", dg); put(a + "/LayerNorm_2/beta", db)
            # ---- LN2: h2 = LN(u2), u2 = relu(f2pre) + h1 -------------------------------------------
            (du2, dpre2), dg, db, db2 = ops.layernorm_joint_bwd(m["u2"], du3, K * D, B, K, D, m["st2"],
                                                                v[a + "/LayerNorm_1/gamma"], inv_scale=inv, mask=m["f2"],
                                                                want_du_colsum=True)
            put(a + "/LayerNorm_1/gamma", dg); put(a + "/LayerNorm_1/beta", db)
            put(f"{a}/ff_output{sid}/bias", db2)
            h1 = m["h1"].view(rows, D)
            f1 = m["f1"]
            dpre2 = dpre2.view(rows, D)
            put(f"{a}/ff_output{sid}/kernel", self._wgrad_gemm(ctx, name, f1, dpre2, out_dtype=f32, alpha=inv,
                                                               out=gout(f"{a}/ff_output{sid}/kernel")))
            dpre1 = ops.gemm(dpre2, sh[a + "/w2_16"], b_mn=False, mask=f1, N=4 * D, K=D)      # [rows, 4D], ReLU mask fused
            # (written straight into the gradient view: `put` would otherwise copy on the main stream, racing the side stream)
            bname = f"{a}/filter_output{sid}/bias"
            put(bname, self._wgrad_run(ctx, name, (dpre1,), lambda: ops.colsum(dpre1, alpha=inv, out=gout(bname))))
            put(f"{a}/filter_output{sid}/kernel", self._wgrad_gemm(ctx, name, h1, dpre1, out_dtype=f32, alpha=inv,
                                                                   out=gout(f"{a}/filter_output{sid}/kernel")))
            dh1 = ops.gemm(dpre1, sh[a + "/w1_16"], b_mn=False, add1=du3.view(rows, D), add2=du2.view(rows, D))
            # ---- LN1: h1 = LN(u1), u1 = att + vlad --------------------------------------------------
            du1, dg, db, dbo = ops.layernorm_joint_bwd(m["u1"], dh1, K * D, B, K, D, m["st1"], v[a + "/LayerNorm/gamma"],
                                                       inv_scale=inv, want_du_colsum=True)
            put(a + "/LayerNorm/gamma", dg); put(a + "/LayerNorm/beta", db)
            put(a + "/output_transform/bias", dbo)
            du1 = du1.view(rows, D)
            put(a + "/output_transform/kernel", self._wgrad_gemm(ctx, name, m["o"], du1, out_dtype=f32, alpha=inv,
                                                                 out=gout(a + "/output_transform/kernel")))
            do = ops.gemm(du1, sh[a + "/wo16"], b_mn=False)
            pre_attn = ctx.get("before_attention_bwd_hook") if name == "video" else None
            if pre_attn is not None:
                # single tower: the factored update of hidden1_weights forks HERE -- the attention-core backward that follows is
                # issue-bound, moves < 1 TB/s and leaves a quarter of the register file free, so the HBM-bound update's small
                # CTAs run next to it (trainer._fork_hidden_update)
                pre_attn(ctx)
            dqkv = ops.mha_core_bwd(m["qkv"], m["o"], do, m["lse"], B, K, D, H, scale=(D // H) ** -0.5)
            zn = m["zn"]
            for i, n in enumerate(("q", "k", "v")):
                put(f"{a}/{n}/kernel", self._wgrad_gemm(ctx, name, zn, dqkv[:, i * D:(i + 1) * D], out_dtype=f32, alpha=inv,
                                                        out=gout(f"{a}/{n}/kernel")))
            dzn = ops.gemm(dqkv, sh[a + "/wqkv16"], b_mn=False, add1=du1)        # + residual branch
            if part == "attn":
                ctx["_dzn_" + name] = dzn
                return
        else:
            dzn = ctx.pop("_dzn_" + name)
        # ---- NetVLAD normalisation + aggregation + soft-assignment ------------------------------
        ct = sh[vs + "/centers_t"]
        if c.d5_raw_reshape:     # the block's input was the d-major flattened descriptor: its gradient is d-major too
            dzn = ops.dmajor_to_kmajor_f16(dzn.view(B, K * D), B, K, D)
        dz, q = ops.netvlad_norm_bwd(m["z"], m["rscale"], dzn.view(B, K, D), ct)
        X = m["X"]                                                            # [B*T, D] view, row stride = feature size
        wc16 = sh[vs + "/wc16"]
        S16 = ops.gemm(X, wc16)                                               # logits recomputed: [B*T, K]
        X3 = ctx["xb"].view(B, T, -1)[:, :, col0:col0 + D]
        G = ops.gemm(X3, dz, b_mn=False, out_dtype=f32)                       # [B, T, K] = X dV^T per video
        bn = vs + "/cluster_bn"
        dS, dg, db = ops.assign_bwd(G.view(B * T, K), m["assign"].view(B * T, K), q, S16, m["cluster_bn_stats"],
                                    v[bn + "/gamma"], T, inv_scale=inv)
        put(bn + "/gamma", dg); put(bn + "/beta", db)
        m_tiles = (D + 127) // 128
        splits = max(2, min(64, 148 // max(1, m_tiles)))
        parts = ops.gemm(X, dS, a_mn=True, b_mn=True, splits=splits)           # [s, D, K] fp32
        dWc = gout(vs + "/cluster_weights")
        if dWc is None:
            dWc = torch.empty((D, K), dtype=f32, device=X.device)
        ops.splitk_reduce(parts, alpha=inv, out32=dWc)
        put(vs + "/cluster_weights", dWc)
        dCt, E = ops.center_bwd(dz, m["z"], m["a_sum"], ct, v["input_bn/beta"][col0:col0 + D], inv)
        put(vs + "/cluster_weights2", ops.transpose_f32(dCt).view(1, D, K))
        ops.input_bn_grad(v[vs + "/cluster_weights"], dWc, dCt, E, v["input_bn/gamma"][col0:col0 + D],
                          dgamma_in[col0:col0 + D], dbeta_in[col0:col0 + D])

    def _v2_modality_bwd(self, ctx, name, col0, D, K, dv, dgamma_in, dbeta_in, put):
        """Backward of NetVladAttenCluster + TransformerEncoderMod (video_pooling_modules.py:1617-1663,
        transformer_utils.py:443-457,634-677,737-767).  Unlike V1 the frames feed the encoder, so dX is formed
        (three contributions fused into one GEMM epilogue) and reduced into input_bn's gamma / beta."""
        c, v, sh = self.cfg, self.store.vars, self.store.shadows
        S = ctx["loss_scale"]
        inv = 1.0 / S
        m, B, T = ctx[name], ctx["B"], c.iterations
        vs = name + "_VLAD"
        a = vs + "/cluster_attention"
        H = D // 16
        f32 = torch.float32
        gout = ctx["_gout"]
        rows = B * T
        X = m["X"]
        ct = sh[vs + "/centers_t"]
        # ---- d-major flatten + norms + residual aggregation ---------------------------------------
        dvh = ops.dmajor_to_kmajor_f16(dv, B, K, D)
        dz, q = ops.netvlad_norm_bwd(m["z"], m["rscale"], dvh, ct)
        X3 = X.view(B, T, D)
        G = ops.gemm(X3, dz, b_mn=False, out_dtype=f32)                       # [B,T,K] = X dV^T
        dXpool = ops.gemm(m["A"].view(B, T, K), dz, b_mn=True)                # [B,T,D] = A dV
        dCt, _ = ops.center_bwd(dz, m["z"], m["a_sum"], ct, v["input_bn/beta"][col0:col0 + D], inv)
        put(vs + "/cluster_centers", ops.transpose_f32(dCt))
        # ---- FeedForwardNetworkMod: BN(relu(BN(relu(h1 W1 + b1)) W2 + b2)) -------------------------
        bn = a + "/feed_output_bn"
        # grad wrt the BN'd similarities = G - q (kept in fp32: the batch-norm backward cancels its common mode)
        dpre2, dg, db = ops.batch_norm_cols_bwd(G.view(rows, K), m["f2"], m["obn_stats"], v[bn + "/gamma"], inv_scale=inv,
                                                relu=True, q=q, T=T)
        put(bn + "/gamma", dg); put(bn + "/beta", db)
        put(a + "/ff_outputencode/bias", ops.colsum(dpre2, alpha=inv))
        put(a + "/ff_outputencode/kernel", ops.gemm(m["f_bn"], dpre2, a_mn=True, b_mn=True, out_dtype=f32, alpha=inv,
                                                    out=gout(a + "/ff_outputencode/kernel")))
        df_bn = ops.gemm(dpre2, sh[a + "/w2_16"], b_mn=False, N=4 * D, K=K, out_dtype=f32)
        bn = a + "/filter_bn"
        dpre1, dg, db = ops.batch_norm_cols_bwd(df_bn, m["f"], m["fbn_stats"], v[bn + "/gamma"], inv_scale=inv, relu=True)
        put(bn + "/gamma", dg); put(bn + "/beta", db)
        put(a + "/filter_outputencode/bias", ops.colsum(dpre1, alpha=inv))
        h1 = m["h1"].view(rows, D)
        put(a + "/filter_outputencode/kernel", ops.gemm(h1, dpre1, a_mn=True, b_mn=True, out_dtype=f32, alpha=inv,
                                                        out=gout(a + "/filter_outputencode/kernel")))
        dh1 = ops.gemm(dpre1, sh[a + "/w1_16"], b_mn=False)
        # ---- LayerNorm(dropout(att) + X) ---------------------------------------------------------
        du1, dg, db = ops.layernorm_joint_bwd(m["u1"], dh1, T * D, B, T, D, m["st1"], v[a + "/LayerNorm/gamma"], inv_scale=inv)
        put(a + "/LayerNorm/gamma", dg); put(a + "/LayerNorm/beta", db)
        du1 = du1.view(rows, D)
        datt = du1
        if m["mask"] is not None:
            datt = torch.empty_like(du1)
            ops.dropout_f16(du1, c.dropout_rate, mask_in=m["mask"], out=datt)
        # ---- MultiHeadAttentionBN ------------------------------------------------------------------
        put(a + "/output_transform/bias", ops.colsum(datt, alpha=inv))
        put(a + "/output_transform/kernel", ops.gemm(m["o_bn"], datt, a_mn=True, b_mn=True, out_dtype=f32, alpha=inv,
                                                     out=gout(a + "/output_transform/kernel")))
        do_bn = ops.gemm(datt, sh[a + "/wo16"], b_mn=False, out_dtype=f32)
        bn = a + "/attention_bn"
        do, dg, db = ops.batch_norm_cols_bwd(do_bn, m["o"], m["abn_stats"], v[bn + "/gamma"], inv_scale=inv, relu=False)
        put(bn + "/gamma", dg); put(bn + "/beta", db)
        lmean, lrstd = m["lbn_stats"][0], m["lbn_stats"][1]
        part = ops.mha_core_bwd_bn(1, m["qkv"], m["o"], do, m["lse"], B, T, D, H, m["ks"], m["kb"], lmean, lrstd)
        n_l = float(B * H * T)
        m12 = torch.empty((2, T), dtype=f32, device=X.device)
        ops.colsum_final(part, B * H, 2 * T, 2 * T, alpha=1.0 / n_l, out=m12.view(-1))
        bn = a + "/logits_bn"
        put(bn + "/beta", ops.colsum_final(part, B * H, 2 * T, T, alpha=inv))
        put(bn + "/gamma", ops.colsum_final(part[:, 1], B * H, 2 * T, T, alpha=inv))
        dqkv = ops.mha_core_bwd_bn(2, m["qkv"], m["o"], do, m["lse"], B, T, D, H, m["ks"], m["kb"], lmean, lrstd,
                                   m1=m12[0], m2=m12[1])
        for i, n in enumerate(("q", "k", "v")):
            put(f"{a}/{n}/kernel", ops.gemm(X, dqkv[:, i * D:(i + 1) * D], a_mn=True, b_mn=True, out_dtype=f32, alpha=inv,
                                            out=gout(f"{a}/{n}/kernel")))
        # ---- dX = dqkv Wqkv^T + du1 (residual) + A dV (aggregation); input_bn gamma / beta --------
        dX = ops.gemm(dqkv, sh[a + "/wqkv16"], b_mn=False, add1=du1, add2=dXpool.view(rows, D))
        dg, db = ops.bn_output_param_grads(dX, X, v["input_bn/beta"][col0:col0 + D], v["input_bn/gamma"][col0:col0 + D],
                                           inv_scale=inv)
        dgamma_in[col0:col0 + D].copy_(dg)
        dbeta_in[col0:col0 + D].copy_(db)


class InferenceGraph:
    """The `is_training=False` forward of one engine captured in a CUDA graph (fixed batch / frame count / input dtype).

    The forward is 28 launches at config 1 (15 for WillowModelReg) issued from Python; at small batches the ~10 us of host
    work per launch, not the GPU, sets the latency.  Capturing the whole sequence -- both streams: the audio modality
    forks onto the side stream and joins before the head -- replays it with one launch.  Inputs are copied into static
    buffers; the returned predictions tensor is static too (clone it to keep it past the next call).  Parameter VALUES may
    change between calls (training steps refresh the fp16 operand shadows in place); the capture is redone only when a
    variable or shadow buffer moves (`VariableStore.layout_version`: first trainer step, load of new variables)."""

    def __init__(self, engine: NetVladEngine, batch: int, max_frames: int, input_dtype=torch.float32):
        self.engine = engine
        c, dev = engine.cfg, engine.store.device
        self.x = torch.zeros((batch, max_frames, c.feature_size), dtype=input_dtype, device=dev)
        self.nf = torch.full((batch,), max_frames, dtype=torch.int32, device=dev)
        # WillowModelReg draws random frames on every call: the draw stays outside the graph (a seed baked into a
        # captured launch would repeat it), its indices enter through a static buffer
        self.idx = (torch.zeros((batch, c.iterations), dtype=torch.int32, device=dev) if c.model == "WillowModelReg" else None)
        self.graph, self.pred, self.version, self.calls = None, None, None, 0

    def _capture(self):
        eng = self.engine
        eng.refresh_shadows()
        side = torch.cuda.Stream(device=self.x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():          # warm-up off the default stream (attribute setup, allocator)
            for _ in range(2):
                eng.forward(self.x, self.nf, False, frame_index=self.idx)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        from . import _lib
        n0 = _lib.launch_count
        with no_gc_during_capture(), torch.cuda.graph(self.graph), torch.no_grad():
            self.pred, _ = eng.forward(self.x, self.nf, False, frame_index=self.idx)
        self.launches = _lib.launch_count - n0       # kernels one replay launches (bench.py's gpu_launches)
        _lib.launch_count = n0
        self.version = eng.store.layout_version

    def __call__(self, model_input, num_frames, frame_index=None):
        eng = self.engine
        if tuple(model_input.shape) != tuple(self.x.shape) or model_input.dtype != self.x.dtype:
            raise ValueError(f"graph captured for {tuple(self.x.shape)} {self.x.dtype}, got {tuple(model_input.shape)} {model_input.dtype}")
        if model_input.data_ptr() != self.x.data_ptr():     # callers may fill `self.x` directly and skip this copy
            self.x.copy_(model_input, non_blocking=True)
        self.nf.copy_(num_frames.to(torch.int32), non_blocking=True)
        if self.idx is not None:
            if frame_index is None:
                frame_index = ops.random_frame_index(self.nf, eng.cfg.iterations, self.x.shape[1],
                                                     mode=0 if eng.cfg.random_frames else 1, seed=0x5EED0000 + eng.draws)
                eng.draws += 1
            self.idx.copy_(frame_index.to(torch.int32), non_blocking=True)
        eng.refresh_shadows()                          # in place; a no-op unless values changed outside the trainer
        if eng.cfg.model == "NetVladV1" and eng.cfg.split_hidden_infer:
            eng.refresh_hidden_lo()                    # low-order half of hidden1_weights follows the variables (outside the graph)
        if self.graph is None or self.version != eng.store.layout_version:
            self._capture()
        self.graph.replay()
        from . import _lib
        _lib.launch_count += self.launches
        self.calls += 1
        return self.pred
