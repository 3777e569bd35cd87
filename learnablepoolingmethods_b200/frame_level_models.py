"""Frame-level models of the hot path with the reference's registry names and call signatures.

Drop-in for `frame_level_models.NetVladV1` / `NetVladV2` (frame_level_models.py:2222-2377, 2383-2513),
the `NetVLAD` pooling module (:2765-2824) and -- SURVEY 8f row 4 -- the baseline `WillowModelReg` (:2516-2635)
and the `LightVLAD` module (:2827-2877).  `create_model` executes eagerly on the GPU (there is no
graph): it returns `{"predictions": tensor [B, vocab]}`; when called with `is_training=True` under
torch autograd the tensor carries the backward of the whole path (hand-written CUDA).
"""
from __future__ import annotations

import torch

from . import models, ops, variables
from .engine import NetVladConfig, NetVladEngine
from .flags import FLAGS, ensure_parsed


def _resolve(cls_name, model_input, vocab_size, iterations, cluster_size, hidden_size, unused, sample_random_frames=None):
    ensure_parsed()
    # kwargs resolve as `kw or FLAGS.<name>` (frame_level_models.py:2235-2242)
    iterations = iterations or FLAGS.iterations
    cluster_size = cluster_size or FLAGS.netvlad_cluster_size
    hidden_size = hidden_size or FLAGS.netvlad_hidden_size
    feat = int(model_input.shape[2])
    rgb_dim = int(unused.get("rgb_dim", 1024))      # the reference hard-codes 1024 (:2261,2274)
    return NetVladConfig(model=cls_name, iterations=int(iterations), cluster_size=int(cluster_size),
                         hidden_size=int(hidden_size), vocab_size=int(vocab_size),
                         num_mixtures=int(unused.get("num_mixtures") or FLAGS.moe_num_mixtures),
                         rgb_dim=rgb_dim, audio_dim=feat - rgb_dim,
                         rgb_heads=int(unused.get("rgb_heads", 64)), audio_heads=int(unused.get("audio_heads", 16)),
                         add_batch_norm=True, gating=FLAGS.gating, remove_diag=FLAGS.gating_remove_diag,
                         netvlad_relu=FLAGS.netvlad_relu,
                         moe_l2=FLAGS.moe_l2, rgb_det_reg=float(FLAGS.rgb_det_reg), audio_det_reg=float(FLAGS.audio_det_reg),
                         # `sample_random_frames or FLAGS.sample_random_frames` (frame_level_models.py:2531)
                         random_frames=bool(sample_random_frames or FLAGS.sample_random_frames))


def get_engine(cfg: NetVladConfig, store=None) -> NetVladEngine:
    """One engine per (store, config): calling create_model again reuses the variables (tower reuse)."""
    store = store or variables.default_store()
    cache = store.__dict__.setdefault("_engines", {})
    key = tuple(sorted(cfg.__dict__.items()))
    if key not in cache:
        cache[key] = NetVladEngine(cfg, store)
    return cache[key]


class _NetVladBase(models.BaseModel):
    _NAME = None

    def create_model(self, model_input, vocab_size, num_frames, iterations=None, add_batch_norm=None,
                     sample_random_frames=None, cluster_size=None, hidden_size=None, is_training=True,
                     **unused_params):
        cfg = _resolve(self._NAME, model_input, vocab_size, iterations, cluster_size, hidden_size, unused_params,
                       sample_random_frames)
        engine = get_engine(cfg, unused_params.get("store"))
        from .autograd import netvlad_apply
        if unused_params.get("cuda_graph") and not is_training:
            # serving loops (eval.py / inference.py call the same graph once per batch): replay the captured forward;
            # one engine.InferenceGraph per input shape, cached on the engine.  The returned tensor is static.
            from .engine import InferenceGraph
            graphs = engine.__dict__.setdefault("_inference_graphs", {})
            key = (tuple(model_input.shape), model_input.dtype)
            if key not in graphs:
                graphs[key] = InferenceGraph(engine, model_input.shape[0], model_input.shape[1], model_input.dtype)
            pred = graphs[key](model_input, num_frames, frame_index=unused_params.get("frame_index"))
        else:
            ensure_parsed()
            penalty = unused_params.get("regularization_penalty")
            pred = netvlad_apply(engine, model_input, num_frames, is_training,
                                 dropout_masks=unused_params.get("dropout_masks"), frame_index=unused_params.get("frame_index"),
                                 reg_penalty=FLAGS.regularization_penalty if penalty is None else penalty)
        result = {"predictions": pred}
        if self._NAME == "WillowModelReg":
            # TF collects the orthogonal regulariser through REGULARIZATION_LOSSES (train.py:301-303); the eager
            # mirror hands its VALUE back under the key train.py:296-297 already honours.  It is detached: its gradient is
            # part of this model's hand-written backward, scaled by `regularization_penalty=` (default: the flag), so a
            # caller must NOT add `penalty * regularization_loss` to the loss it differentiates.  The MoE weight decay
            # (slim.l2_regularizer(moe_l2)) is applied by Trainer.apply_gradients / Trainer.train_step.
            result["regularization_loss"] = engine.regularization_loss()
        return result


class NetVladV1(_NetVladBase):
    """ NetVlad with Context Gating + attention over the cluster descriptors (paper 3.1). """
    _NAME = "NetVladV1"


class NetVladV2(_NetVladBase):
    """ NetVlad with attention-based cluster similarities (paper 3.2). """
    _NAME = "NetVladV2"


class WillowModelReg(_NetVladBase):
    """ WILLOW model with orthogonal regularization for robust features (frame_level_models.py:2516-2635):
    random frame sampling, NetVladOrthoReg pooling per modality, context gating, MoE.  `frame_index=` (int32
    [B, iterations]) in **unused_params replaces the random draw (deterministic evaluation / parity tests). """
    _NAME = "WillowModelReg"


class NetVLAD():
    """frame_level_models.py:2765-2824; the 6th ctor argument of the call sites (:2261-2264) is a scope label."""

    def __init__(self, feature_size, max_frames, cluster_size, add_batch_norm, is_training, scope_id=None):
        self.feature_size = feature_size
        self.max_frames = max_frames
        self.is_training = is_training
        self.add_batch_norm = add_batch_norm
        self.cluster_size = int(cluster_size)
        self.scope_id = scope_id

    def forward(self, reshaped_input, store=None):
        """reshaped_input: [(B*max_frames), feature_size] (fp32 or fp16, GPU) -> [B, cluster_size*feature_size]
        fp32, d-major flatten (index d*K + k) as in the reference."""
        s = store or variables.default_store()
        D, K, T = self.feature_size, self.cluster_size, self.max_frames
        import math
        wc = s.get_variable("cluster_weights", (D, K), "normal", 1 / math.sqrt(D))
        if self.add_batch_norm:
            beta, gamma, mm, mv = s.batch_norm_vars("cluster_bn", K)
        else:
            bias = s.get_variable("cluster_biases", (K,), "normal", 1 / math.sqrt(D))
        c2 = s.get_variable("cluster_weights2", (1, D, K), "normal", 1 / math.sqrt(D))
        x16 = reshaped_input if reshaped_input.dtype == torch.float16 else ops.cast_f16(reshaped_input.contiguous())
        B = x16.shape[0] // T
        wc16 = ops.cast_f16(wc)
        if self.add_batch_norm:
            if self.is_training:
                _, st = ops.gemm(wc16, x16, a_mn=True, b_mn=False, out="none", stats=True)
                scale, shift = ops.bn_finalize(st[0].reshape(-1, K), st[1].reshape(-1, K), B * T, gamma, beta, mm, mv,
                                               training=True, bessel=True)
            else:
                scale, shift = ops.bn_finalize(None, None, 1, gamma, beta, mm, mv, training=False, bessel=True)
        else:
            scale, shift = torch.ones(K, device=x16.device), bias
        z, rs, _, _ = ops.netvlad_pool_fwd(x16, B, T, wc16, scale, shift, c2[0])
        return ops.netvlad_finalize(z, rs, d_major=True)


class LightVLAD():
    """frame_level_models.py:2827-2877: NetVLAD without the cluster-centre residual.  Runs the same fused pooling
    kernel with zero centres."""

    def __init__(self, feature_size, max_frames, cluster_size, add_batch_norm, is_training):
        self.feature_size = feature_size
        self.max_frames = max_frames
        self.is_training = is_training
        self.add_batch_norm = add_batch_norm
        self.cluster_size = int(cluster_size)

    def forward(self, reshaped_input, store=None):
        """reshaped_input: [(B*max_frames), feature_size] -> [B, cluster_size*feature_size] fp32 (d-major flatten)."""
        import math
        s = store or variables.default_store()
        D, K, T = self.feature_size, self.cluster_size, self.max_frames
        wc = s.get_variable("cluster_weights", (D, K), "normal", 1 / math.sqrt(D))
        x16 = reshaped_input if reshaped_input.dtype == torch.float16 else ops.cast_f16(reshaped_input.contiguous())
        B = x16.shape[0] // T
        wc16 = ops.cast_f16(wc)
        if self.add_batch_norm:
            beta, gamma, mm, mv = s.batch_norm_vars("cluster_bn", K)
            if self.is_training:
                _, st = ops.gemm(wc16, x16, a_mn=True, b_mn=False, out="none", stats=True)
                scale, shift = ops.bn_finalize(st[0].reshape(-1, K), st[1].reshape(-1, K), B * T, gamma, beta, mm, mv,
                                               training=True, bessel=True)
            else:
                scale, shift = ops.bn_finalize(None, None, 1, gamma, beta, mm, mv, training=False, bessel=True)
        else:
            scale = torch.ones(K, device=x16.device)
            shift = s.get_variable("cluster_biases", (K,), "normal", 1 / math.sqrt(D))
        zero_centers = torch.zeros((K, D), dtype=torch.float16, device=x16.device)
        z, rs, _, _ = ops.netvlad_pool_fwd(x16, B, T, wc16, scale, shift, zero_centers)
        return ops.netvlad_finalize(z, rs, d_major=True)
