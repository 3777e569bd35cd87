"""CPU oracle for the NetVLAD learnable-pooling hot path (TEST INFRASTRUCTURE ONLY).

This file is a CPU restatement (torch-CPU ops, fp32 or fp64) of the *intended*
computation of pomonam/LearnablePoolingMethods' `NetVladV1` / `NetVladV2`
path, with TensorFlow-1.x library semantics encoded by hand.  It is the checker
for the CUDA product path -- only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  The product
package (`learnablepoolingmethods_b200`) never imports it.

PARITY UNPINNED: the reference ships no tests, golden vectors or checkpoints
for this path, TensorFlow is not installable here, and the reference does not
import as shipped (defects D1-D7).  The pins that do exist:
  * `tests/golden/*.npz` produced by `oracle/make_golden.py` from this file
    (self-pins: they freeze the oracle, they do not validate it against TF);
  * the frame-sampling index rule is checked against its closed form;
  * top-k / GAP are checked with the reference's own importable numpy code
    (`eval_util.py`) when `/root/reference` is present.

Reference lines followed (all relative to /root/reference):
  train.py:262-264                  per-frame L2 normalise (caller prelude)
  model_utils.py:101-122            SampleUniformFrames
  frame_level_models.py:2222-2377   NetVladV1.create_model
  frame_level_models.py:2383-2513   NetVladV2.create_model
  frame_level_models.py:2765-2824   NetVLAD.forward
  video_pooling_modules.py:1592-1663 NetVladAttenCluster.forward
  transformer_utils.py:374-457,507-767  encoder blocks
  video_level_models.py:48-159      MoeModel
  losses.py:41-51                   CrossEntropyLoss
  utils.py:170-213, train.py:244-252,321-336  combine / clip / Adam / LR decay
  model_utils.py:26-73              SampleRandomSequence / SampleRandomFrames      
  frame_level_models.py:2516-2635   WillowModelReg.create_model
  video_pooling_modules.py:1499-1586 NetVladOrthoReg.forward
  frame_level_models.py:2827-2877   LightVLAD.forward
  module_utils.py:55-90             orthogonal_regularizer
Decisions D1-D7 of SURVEY.md section 0 are applied (see the `D5`/`D6`/`D7` notes inline).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

BN_EPS = 1e-3        # slim.batch_norm default epsilon
BN_DECAY = 0.999     # slim.batch_norm default decay
LN_EPS = 1e-12       # tf.contrib.layers.layer_norm variance_epsilon
L2N_EPS = 1e-12      # tf.nn.l2_normalize epsilon
XENT_EPS = 10e-6     # losses.py:46

# Precision study only (tests/test_trained_parity_gpu.py, DESIGN.md "Numerics"): when set to "tf32" / "fp16" / "bf16"
# every matrix product of the NetVladV1 path rounds its two operands to that format first (fp32 accumulation), which
# is what a tensor-core run of the reference does (TF32 is TensorFlow's default matmul mode on Ampere and later).
# None (the default, and the only mode parity is asserted against) = plain fp32 / fp64 products.
OPERAND_ROUND = None
OPERAND_ROUND_HEAD = True     # False: the head's products (hidden projection, gate, MoE) stay exact while the body rounds


def _round_operand(x: torch.Tensor) -> torch.Tensor:
    if OPERAND_ROUND is None:
        return x
    if OPERAND_ROUND == "fp16":
        return x.half().to(x.dtype)
    if OPERAND_ROUND == "bf16":
        return x.bfloat16().to(x.dtype)
    if OPERAND_ROUND == "tf32":      # 10 explicit mantissa bits, round to nearest (ties away: cvt.rna.tf32.f32)
        bits = x.detach().float().contiguous().view(torch.int32)
        return ((bits + 0x1000) & ~0x1FFF).view(torch.float32).to(x.dtype)
    raise ValueError(OPERAND_ROUND)


def mm(a: torch.Tensor, b: torch.Tensor, head: bool = False) -> torch.Tensor:
    """tf.matmul (fp32 on the reference's CPU path); see OPERAND_ROUND."""
    if OPERAND_ROUND is None or (head and not OPERAND_ROUND_HEAD):
        return a @ b
    return _round_operand(a) @ _round_operand(b)


# --------------------------------------------------------------------------- #
# TF library semantics
# --------------------------------------------------------------------------- #
def l2_normalize(x: torch.Tensor, axis) -> torch.Tensor:
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2, axis), eps))."""
    ss = (x * x).sum(dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=L2N_EPS))


def batch_norm(x, P, S, scope, is_training, update_moving=True):
    """slim.batch_norm(center=True, scale=True): channel = last axis.

    Training: batch mean / biased variance over all other axes.  Moving
    statistics: decay 0.999; the fused kernel (used by slim for rank 2 and 4
    inputs) feeds the Bessel-corrected variance into the moving average, the
    non-fused path (rank 3) feeds the biased one.
    """
    gamma, beta = P[scope + "/gamma"], P[scope + "/beta"]
    C = x.shape[-1]
    if is_training:
        flat = x.reshape(-1, C)
        n = flat.shape[0]
        mean = flat.mean(dim=0)
        var = ((flat - mean) ** 2).mean(dim=0)
        if update_moving and S is not None:
            with torch.no_grad():
                corr = n / max(n - 1, 1) if x.dim() in (2, 4) else 1.0
                S[scope + "/moving_mean"].mul_(BN_DECAY).add_(mean.detach() * (1 - BN_DECAY))
                S[scope + "/moving_variance"].mul_(BN_DECAY).add_(var.detach() * corr * (1 - BN_DECAY))
    else:
        mean, var = S[scope + "/moving_mean"], S[scope + "/moving_variance"]
    return (x - mean) * torch.rsqrt(var + BN_EPS) * gamma + beta


def layer_norm_joint(x, P, scope):
    """tf.contrib.layers.layer_norm defaults: begin_norm_axis=1 (moments over ALL
    non-batch axes jointly), begin_params_axis=-1 (gamma/beta on last axis)."""
    B = x.shape[0]
    flat = x.reshape(B, -1)
    mean = flat.mean(dim=1).reshape([B] + [1] * (x.dim() - 1))
    var = ((flat - flat.mean(dim=1, keepdim=True)) ** 2).mean(dim=1).reshape([B] + [1] * (x.dim() - 1))
    return (x - mean) * torch.rsqrt(var + LN_EPS) * P[scope + "/gamma"] + P[scope + "/beta"]


def dense(x, P, scope, bias=True, relu=False):
    """tf.layers.dense: kernel [in, out] applied to the last axis."""
    y = mm(x, P[scope + "/kernel"])
    if bias:
        y = y + P[scope + "/bias"]
    return torch.relu(y) if relu else y


# --------------------------------------------------------------------------- #
# model_utils.py:101-122  SampleUniformFrames
# --------------------------------------------------------------------------- #
def sample_uniform_indices(num_frames: np.ndarray, num_samples: int) -> np.ndarray:
    """frame_index = int32( linspace(0,1,S+1)[:S] * float32(num_frames) ), fp32 math.

    TF's LinSpace kernel computes `start + step * i` in float32 with
    step = (stop - start) / (num - 1); the cast to int32 truncates toward zero.
    """
    step = np.float32(1.0) / np.float32(num_samples)
    grid = (np.arange(num_samples, dtype=np.float32) * step).astype(np.float32)
    prod = (grid[None, :] * np.asarray(num_frames).astype(np.float32)[:, None]).astype(np.float32)
    return prod.astype(np.int32)


def sample_uniform_frames(model_input: torch.Tensor, num_frames, num_samples: int) -> torch.Tensor:
    idx = torch.from_numpy(sample_uniform_indices(np.asarray(num_frames), num_samples)).long()
    B = model_input.shape[0]
    return model_input[torch.arange(B)[:, None], idx]            # [B, S, F]


# --------------------------------------------------------------------------- #
# model_utils.py:26-73  SampleRandomSequence / SampleRandomFrames (uniform draws injected)
# --------------------------------------------------------------------------- #
def sample_random_frame_indices(num_frames: np.ndarray, uniform: np.ndarray) -> np.ndarray:
    """frame_index = int32( uniform[B,S] * tile(float32(num_frames)) ), fp32 product, truncation (:66-69)."""
    nf = np.asarray(num_frames).astype(np.float32)[:, None]
    return (np.asarray(uniform, dtype=np.float32) * nf).astype(np.float32).astype(np.int32)


def sample_random_sequence_indices(num_frames: np.ndarray, num_samples: int, uniform: np.ndarray) -> np.ndarray:
    """start = int32(uniform[B,1] * float32(max(nf - S, 0) + 1)); index = min(start + range(S), int32(nf - 1)) (:38-47)."""
    nf = np.asarray(num_frames).astype(np.float32)[:, None]
    max_start = np.maximum(nf - np.float32(num_samples), np.float32(0))
    start = (np.asarray(uniform, dtype=np.float32).reshape(-1, 1) * (max_start + np.float32(1))).astype(np.float32).astype(np.int32)
    return np.minimum(start + np.arange(num_samples, dtype=np.int32)[None, :], (nf - 1).astype(np.int32))


def gather_frames(model_input: torch.Tensor, frame_index: np.ndarray) -> torch.Tensor:
    """tf.gather_nd(model_input, stack([batch_index, frame_index], 2)) (:48-51, 70-73).  Indices are clamped to the
    padded frame range: TF's CPU kernel raises on an out-of-range index (only reachable with num_frames = 0)."""
    idx = torch.from_numpy(np.clip(np.asarray(frame_index), 0, model_input.shape[1] - 1)).long()
    return model_input[torch.arange(model_input.shape[0])[:, None], idx]


# --------------------------------------------------------------------------- #
# frame_level_models.py:2765-2824  NetVLAD.forward
# --------------------------------------------------------------------------- #
def netvlad_forward(x, P, S, scope, max_frames, add_batch_norm, is_training, return_assign=False):
    """x: [(B*T), D] -> [B, D*K] (d-major flatten)."""
    Wc = P[scope + "/cluster_weights"]                              # [D, K]
    D, K = Wc.shape
    act = mm(x, Wc)                                                 # :2781
    if add_batch_norm:
        act = batch_norm(act, P, S, scope + "/cluster_bn", is_training)   # :2783-2789
    else:
        act = act + P[scope + "/cluster_biases"]                    # :2790-2796
    act = torch.softmax(act, dim=-1)                                # :2798
    act = act.reshape(-1, max_frames, K)                            # :2801
    a_sum = act.sum(dim=-2, keepdim=True)                           # :2803  [B,1,K]
    a = a_sum * P[scope + "/cluster_weights2"]                      # :2805-2810  [B,D,K]
    xr = x.reshape(-1, max_frames, D)
    vlad = mm(act.transpose(1, 2), xr)                              # :2812-2815  [B,K,D]
    vlad = vlad.transpose(1, 2) - a                                 # :2816-2817  [B,D,K]
    vlad = l2_normalize(vlad, 1)                                    # :2819 intra-norm over D
    vlad = vlad.reshape(-1, K * D)                                  # :2821 d-major flatten
    vlad = l2_normalize(vlad, 1)                                    # :2822
    return (vlad, act) if return_assign else vlad


# --------------------------------------------------------------------------- #
# module_utils.py:55-90  orthogonal_regularizer
# --------------------------------------------------------------------------- #
def orthogonal_regularizer(weights: torch.Tensor, scale: float) -> torch.Tensor:
    """scale * sum | l2n(W, axis=1)^T l2n(W, axis=1) - I |   (W is [feature_size, cluster_size]: ROWS are normalised)."""
    n = l2_normalize(weights, 1)                                    # :78
    det = n.transpose(0, 1) @ n                                     # :79-80
    det = det - torch.eye(det.shape[0], dtype=det.dtype)            # :81-82
    return scale * det.abs().sum()                                  # :83, 88


# --------------------------------------------------------------------------- #
# video_pooling_modules.py:1499-1586  NetVladOrthoReg.forward
# --------------------------------------------------------------------------- #
def netvlad_ortho_reg_forward(x, P, S, scope, max_frames, batch_norm_, is_training, scope_id=None):
    """Same arithmetic as NetVLAD.forward; `cluster_weights<scope_id>` is the variable name (:1527-1531) and
    `cluster_weights2` is [D, K] (no leading 1, :1558-1568).  The regulariser is collected separately
    (`willow_regularization`): TF attaches it to the variable, train.py:301-303 adds it to the loss."""
    sid = "" if scope_id is None else str(scope_id)
    Wc = P[scope + "/cluster_weights" + sid]
    D, K = Wc.shape
    act = x @ Wc                                                    # :1536
    if batch_norm_:
        act = batch_norm(act, P, S, scope + "/cluster_bn", is_training)   # :1538-1544
    else:
        act = act + P[scope + "/cluster_biases" + sid]              # :1545-1552
    act = torch.softmax(act, dim=-1)                                # :1554
    act = act.reshape(-1, max_frames, K)                            # :1557
    a_sum = act.sum(dim=-2, keepdim=True)                           # :1559
    a = a_sum * P[scope + "/cluster_weights2"].unsqueeze(0)         # :1573-1574
    xr = x.reshape(-1, max_frames, D)
    vlad = torch.matmul(act.transpose(1, 2), xr).transpose(1, 2) - a      # :1576-1581
    vlad = l2_normalize(vlad, 1)                                    # :1582
    vlad = vlad.reshape(-1, K * D)                                  # :1583
    return l2_normalize(vlad, 1)                                    # :1584


# --------------------------------------------------------------------------- #
# frame_level_models.py:2827-2877  LightVLAD.forward (no cluster centres)
# --------------------------------------------------------------------------- #
def light_vlad_forward(x, P, S, scope, max_frames, add_batch_norm, is_training):
    Wc = P[scope + "/cluster_weights"]
    D, K = Wc.shape
    act = x @ Wc                                                    # :2841
    if add_batch_norm:
        act = batch_norm(act, P, S, scope + "/cluster_bn", is_training)   # :2843-2849
    else:
        act = act + P[scope + "/cluster_biases"]
    act = torch.softmax(act, dim=-1).reshape(-1, max_frames, K)     # :2858-2860
    vlad = torch.matmul(act.transpose(1, 2), x.reshape(-1, max_frames, D)).transpose(1, 2)   # :2862-2869
    vlad = l2_normalize(vlad, 1)                                    # :2871
    return l2_normalize(vlad.reshape(-1, K * D), 1)                 # :2873-2874


# --------------------------------------------------------------------------- #
# transformer_utils.py  (V1 attention block over cluster descriptors)
# --------------------------------------------------------------------------- #
def _split_heads(x, H):
    B, L, D = x.shape
    return x.reshape(B, L, H, D // H).permute(0, 2, 1, 3)          # [B,H,L,depth]


def _combine_heads(x):
    B, H, L, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, L, H * d)


def multi_head_attention(x, P, scope, num_heads):
    """transformer_utils.py:552-586 (self-attention, q scaled by depth^-0.5, no dropout)."""
    hidden = x.shape[-1]
    q = dense(x, P, scope + "/q", bias=False)
    k = dense(x, P, scope + "/k", bias=False)
    v = dense(x, P, scope + "/v", bias=False)
    q, k, v = _split_heads(q, num_heads), _split_heads(k, num_heads), _split_heads(v, num_heads)
    q = q * (hidden // num_heads) ** -0.5
    w = torch.softmax(mm(q, k.transpose(-1, -2)), dim=-1)
    out = _combine_heads(mm(w, v))
    return dense(out, P, scope + "/output_transform", bias=True)


def transformer_encoder(x, P, scope, num_heads, scope_id):
    """transformer_utils.py:399-413 + FeedForwardNetwork.forward :696-715.

    Variable names: the three layer_norm calls in one variable scope are
    auto-uniquified by TF as LayerNorm, LayerNorm_1, LayerNorm_2 in call order
    (encoder :407, FFN :713, encoder :411)."""
    att = multi_head_attention(x, P, scope, num_heads) + x
    h1 = layer_norm_joint(att, P, scope + "/LayerNorm")
    f = dense(h1, P, scope + "/filter_output" + scope_id, relu=True)
    f = dense(f, P, scope + "/ff_output" + scope_id, relu=True)     # ReLU on the output too (:708-711)
    h2 = layer_norm_joint(f + h1, P, scope + "/LayerNorm_1")
    out = layer_norm_joint(h2 + h1, P, scope + "/LayerNorm_2")
    return out


# --------------------------------------------------------------------------- #
# transformer_utils.py  (V2 assignment network over frames)
# --------------------------------------------------------------------------- #
def multi_head_attention_bn(x, P, S, scope, num_heads, is_training):
    """transformer_utils.py:634-677: BN on logits (channel = key axis), BN on combined heads."""
    q = dense(x, P, scope + "/q", bias=False)
    k = dense(x, P, scope + "/k", bias=False)
    v = dense(x, P, scope + "/v", bias=False)
    q, k, v = _split_heads(q, num_heads), _split_heads(k, num_heads), _split_heads(v, num_heads)
    logits = q @ k.transpose(-1, -2)                                # [B,H,T,T]
    logits = batch_norm(logits, P, S, scope + "/logits_bn", is_training)
    w = torch.softmax(logits, dim=-1)
    out = _combine_heads(w @ v)                                     # [B,T,D]
    out = batch_norm(out, P, S, scope + "/attention_bn", is_training)
    return dense(out, P, scope + "/output_transform", bias=True)


def transformer_encoder_mod(x, P, S, scope, num_heads, is_training, dropout_rate=0.9, dropout_mask=None):
    """transformer_utils.py:443-457 + FeedForwardNetworkMod :737-767.

    D7: tf.layers.dropout(rate=1-0.1) => drop probability 0.9 in training.  The
    keep-mask (1 = keep) may be injected for deterministic parity."""
    att = multi_head_attention_bn(x, P, S, scope, num_heads, is_training)
    if is_training and dropout_rate > 0:
        if dropout_mask is None:
            dropout_mask = (torch.rand(att.shape) >= dropout_rate).to(att.dtype)
        att = att * dropout_mask.to(att.dtype) / (1.0 - dropout_rate)
    att = att + x
    h1 = layer_norm_joint(att, P, scope + "/LayerNorm")
    f = dense(h1, P, scope + "/filter_outputencode", relu=True)
    f = batch_norm(f, P, S, scope + "/filter_bn", is_training)
    f = dense(f, P, scope + "/ff_outputencode", relu=True)
    f = batch_norm(f, P, S, scope + "/feed_output_bn", is_training)
    return f                                                        # [B,T,K]; no softmax over K


def netvlad_atten_cluster_forward(x, P, S, scope, max_frames, is_training, dropout_mask=None,
                                  dropout_rate=0.9):
    """video_pooling_modules.py:1617-1663 with D6 (expand_dims(axis=2))."""
    D = x.shape[-1]
    xs = x.reshape(-1, max_frames, D)
    A = transformer_encoder_mod(xs, P, S, scope + "/cluster_attention", D // 16, is_training,
                                dropout_rate=dropout_rate, dropout_mask=dropout_mask)   # [B,T,K]
    C = P[scope + "/cluster_centers"]                               # [D,K]
    vlad = torch.matmul(xs.transpose(1, 2), A) - A.sum(dim=1, keepdim=True) * C         # [B,D,K]
    vlad = l2_normalize(vlad, 1)
    vlad = vlad.reshape(vlad.shape[0], -1)
    return l2_normalize(vlad, 1)


# --------------------------------------------------------------------------- #
# video_level_models.py:48-159  MoeModel (low_rank_gating=-1, prob gating off)
# --------------------------------------------------------------------------- #
def moe_forward(act, P, vocab_size, num_mixtures):
    gate = mm(act, P["gates/weights"], head=True)                              # no bias (:86-92)
    expert = mm(act, P["experts/weights"], head=True) + P["experts/biases"]     # :109-114
    gating = torch.softmax(gate.reshape(-1, num_mixtures + 1), dim=-1)
    experts = torch.sigmoid(expert.reshape(-1, num_mixtures))
    prob = (gating[:, :num_mixtures] * experts).sum(dim=1)
    return prob.reshape(-1, vocab_size)


# --------------------------------------------------------------------------- #
# frame_level_models.py:2309-2377  shared head
# --------------------------------------------------------------------------- #
def head_forward(vlad, P, S, vocab_size, is_training, num_mixtures=2, gating=True,
                 remove_diag=False, return_intermediates=False, relu=False):
    act = mm(vlad, P["hidden1_weights"], head=True)                            # :2319
    if relu:                                                        # `add_batch_norm and relu` (:2321-2327)
        act = batch_norm(act, P, S, "hidden1_bn", is_training)
        act = torch.clamp(act, 0.0, 6.0)                            # tf.nn.relu6 (:2339-2340)
    else:
        act = act + P["hidden1_biases"]                             # :2329-2334 (netvlad_relu False)
    hidden = act
    if gating:
        Wg = P["gating_weights_2"]
        gates = mm(act, Wg, head=True)                                         # :2347
        if remove_diag:
            gates = gates - torch.diagonal(Wg) * act                # :2349-2352
        gates = batch_norm(gates, P, S, "gating_bn", is_training)   # :2354-2360
        act = act * torch.sigmoid(gates)                            # :2367-2368
    pred = moe_forward(act, P, vocab_size, num_mixtures)
    if return_intermediates:
        return pred, {"hidden": hidden, "gated": act}
    return pred


# --------------------------------------------------------------------------- #
# Models
# --------------------------------------------------------------------------- #
def _shell(model_input, num_frames, iterations, P, S, is_training):
    x = sample_uniform_frames(model_input, num_frames, iterations)  # :2255
    B, T, F = x.shape
    x = x.reshape(-1, F)
    x = batch_norm(x, P, S, "input_bn", is_training)                # :2265-2271
    return x, B, T


def netvlad_v1(model_input, num_frames, P, S, *, vocab_size, iterations, cluster_size,
               is_training, num_mixtures=2, rgb_dim=1024, rgb_heads=64, audio_heads=16,
               d5_raw_reshape=False, remove_diag=False, gating=True, return_intermediates=False, relu=False):
    """NetVladV1.create_model forward.  model_input [B, max_frames, rgb+audio] is
    already L2-normalised by the caller (train.py:264)."""
    x, B, T = _shell(model_input, num_frames, iterations, P, S, is_training)
    audio_dim = x.shape[1] - rgb_dim
    Kr, Ka = cluster_size, cluster_size // 4                        # D7: integer division
    inter = {}
    outs = []
    for name, sl, D, K, H, sid in (("video", slice(0, rgb_dim), rgb_dim, Kr, rgb_heads, "encode1"),
                                   ("audio", slice(rgb_dim, None), audio_dim, Ka, audio_heads, "encode2")):
        v = netvlad_forward(x[:, sl], P, S, name + "_VLAD", T, True, is_training)      # [B, D*K]
        inter["vlad_" + name] = v
        if d5_raw_reshape:
            z = v.reshape(B, K, D)
        else:
            z = v.reshape(B, D, K).transpose(1, 2)                  # D5: [B,K,D] cluster-major
        z = transformer_encoder(z, P, name + "_attention", H, sid)
        inter["att_" + name] = z
        outs.append(z.reshape(B, K * D))                            # :2292, :2304 k-major flatten
    vlad = torch.cat(outs, dim=1)                                   # :2309
    pred, hi = head_forward(vlad, P, S, vocab_size, is_training, num_mixtures, gating=gating,
                            remove_diag=remove_diag, return_intermediates=True, relu=relu)
    inter.update(hi)
    return (pred, inter) if return_intermediates else pred


def netvlad_v2(model_input, num_frames, P, S, *, vocab_size, iterations, cluster_size,
               is_training, num_mixtures=2, rgb_dim=1024, dropout_masks=None, dropout_rate=0.9,
               remove_diag=False, gating=True, return_intermediates=False, relu=False):
    x, B, T = _shell(model_input, num_frames, iterations, P, S, is_training)
    inter = {}
    outs = []
    for name, sl in (("video", slice(0, rgb_dim)), ("audio", slice(rgb_dim, None))):
        mask = None if dropout_masks is None else dropout_masks[name]
        v = netvlad_atten_cluster_forward(x[:, sl], P, S, name + "_VLAD", T, is_training,
                                          dropout_mask=mask, dropout_rate=dropout_rate)
        inter["vlad_" + name] = v
        outs.append(v)
    vlad = torch.cat(outs, dim=1)
    pred, hi = head_forward(vlad, P, S, vocab_size, is_training, num_mixtures, gating=gating,
                            remove_diag=remove_diag, return_intermediates=True, relu=relu)
    inter.update(hi)
    return (pred, inter) if return_intermediates else pred


WILLOW_SCOPES = (("video", "netvlad_rgb_scope"), ("audio", "netvlad_audio_scope"))   # frame_level_models.py:2552-2557


def willow_model_reg(model_input, num_frames, P, S, *, vocab_size, iterations, cluster_size, is_training,
                     frame_index, num_mixtures=2, rgb_dim=1024, remove_diag=False, gating=True,
                     return_intermediates=False, relu=False):
    """WillowModelReg.create_model forward (frame_level_models.py:2516-2635; netvlad_relu False).
    `frame_index` int [B, iterations]: the indices SampleRandomFrames / SampleRandomSequence drew (:2539-2544)."""
    x = gather_frames(model_input, frame_index)
    B, T, F = x.shape
    x = batch_norm(x.reshape(-1, F), P, S, "input_bn", is_training)  # :2558-2564
    inter, outs = {}, []
    for (name, sid), sl in zip(WILLOW_SCOPES, (slice(0, rgb_dim), slice(rgb_dim, None))):
        v = netvlad_ortho_reg_forward(x[:, sl], P, S, name + "_VLAD", T, True, is_training, scope_id=sid)
        inter["vlad_" + name] = v
        outs.append(v)
    vlad = torch.cat(outs, dim=1)                                   # :2573
    pred, hi = head_forward(vlad, P, S, vocab_size, is_training, num_mixtures, gating=gating,
                            remove_diag=remove_diag, return_intermediates=True, relu=relu)
    inter.update(hi)
    return (pred, inter) if return_intermediates else pred


def willow_regularization(P, rgb_det_reg=1e-4, audio_det_reg=1e-4):
    """REGULARIZATION_LOSSES contributed by the two NetVladOrthoReg modules (flags :2209-2216)."""
    return (orthogonal_regularizer(P["video_VLAD/cluster_weights2"], rgb_det_reg)
            + orthogonal_regularizer(P["audio_VLAD/cluster_weights2"], audio_det_reg))


# --------------------------------------------------------------------------- #
# losses.py:41-51, regulariser, utils.py:170-213, Adam
# --------------------------------------------------------------------------- #
def cross_entropy_loss(pred, labels):
    y = labels.to(pred.dtype)
    ce = -(y * torch.log(pred + XENT_EPS) + (1 - y) * torch.log(1 - pred + XENT_EPS))
    return ce.sum(dim=1).mean()


def moe_regularization(P, l2_penalty=1e-8):
    """slim.l2_regularizer(l2) = l2 * sum(w^2)/2 on gates/weights and experts/weights."""
    return l2_penalty * 0.5 * ((P["gates/weights"] ** 2).sum() + (P["experts/weights"] ** 2).sum())


def clip_by_norm(g, max_norm):
    """tf.clip_by_norm: g * max_norm / max(||g||, max_norm)."""
    n = torch.sqrt((g * g).sum())
    return g * (max_norm / torch.clamp(n, min=max_norm))


def learning_rate(base_lr, decay, decay_examples, global_step, batch_size, num_towers):
    """tf.train.exponential_decay(staircase=True) on global_step*batch*towers (train.py:244-249)."""
    return base_lr * decay ** math.floor(global_step * batch_size * num_towers / decay_examples)


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update (t = step, 1-based), in place."""
    lr_t = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    m.mul_(beta1).add_(g * (1 - beta1))
    v.mul_(beta2).add_(g * g * (1 - beta2))
    p.sub_(lr_t * m / (torch.sqrt(v) + eps))


TRAINABLE_EXCLUDE = ("moving_mean", "moving_variance")


def train_step(model_fn, P, S, opt_state, batches, labels_list, *, step, lr, clip_norm=1.0,
               reg_penalty=1.0, l2_penalty=1e-8, extra_reg=None):
    """One reference training step over `len(batches)` towers: per-tower loss/grads
    (per-tower BN statistics), SUM over towers (utils.py:205-211), per-tensor clip
    (utils.py:170-189), Adam.  `model_fn(batch, P, S) -> pred`.  Returns (losses, grads)."""
    names = [n for n in P if P[n].requires_grad]
    total = {n: torch.zeros_like(P[n]) for n in names}
    losses = []
    for batch, labels in zip(batches, labels_list):
        for n in names:
            P[n].grad = None
        pred = model_fn(batch, P, S)
        label_loss = cross_entropy_loss(pred, labels)
        reg = moe_regularization(P, l2_penalty)
        if extra_reg is not None:                                   # e.g. willow_regularization (train.py:301-303)
            reg = reg + extra_reg(P)
        loss = label_loss + reg_penalty * reg
        loss.backward()
        losses.append(float(label_loss))
        for n in names:
            if P[n].grad is not None:
                total[n] += P[n].grad
    grads = {n: (clip_by_norm(g, clip_norm) if clip_norm > 0 else g) for n, g in total.items()}
    with torch.no_grad():
        for n in names:
            st = opt_state.setdefault(n, {"m": torch.zeros_like(P[n]), "v": torch.zeros_like(P[n])})
            adam_step(P[n], grads[n], st["m"], st["v"], step, lr)
    return losses, total


# --------------------------------------------------------------------------- #
# Parameter shapes / initialisers (SURVEY 8a "Parameter initialisers") and synthetic data (8d)
# --------------------------------------------------------------------------- #
def param_specs(model: str, *, iterations, cluster_size, hidden_size, vocab_size, num_mixtures=2,
                rgb_dim=1024, audio_dim=128):
    """name -> (shape, init kind, init arg).  kinds: normal(std) | glorot | zeros | ones."""
    sp = {}

    def bn(scope, c):
        sp[scope + "/beta"] = ((c,), "zeros", None)
        sp[scope + "/gamma"] = ((c,), "ones", None)
        sp[scope + "/moving_mean"] = ((c,), "zeros", None)
        sp[scope + "/moving_variance"] = ((c,), "ones", None)

    def ln(scope, c):
        sp[scope + "/beta"] = ((c,), "zeros", None)
        sp[scope + "/gamma"] = ((c,), "ones", None)

    def dn(scope, i, o, bias=True):
        sp[scope + "/kernel"] = ((i, o), "glorot", None)
        if bias:
            sp[scope + "/bias"] = ((o,), "zeros", None)

    bn("input_bn", rgb_dim + audio_dim)
    vdim = 0
    for name, D, K, sid in (("video", rgb_dim, cluster_size, "encode1"),
                            ("audio", audio_dim, cluster_size // 4, "encode2")):
        vs = name + "_VLAD"
        if model == "WillowModelReg":
            sid = dict(WILLOW_SCOPES)[name]
            sp[vs + "/cluster_weights" + sid] = ((D, K), "normal", 1 / math.sqrt(D))
            bn(vs + "/cluster_bn", K)
            sp[vs + "/cluster_weights2"] = ((D, K), "normal", 1 / math.sqrt(D))
        elif model == "NetVladV1":
            sp[vs + "/cluster_weights"] = ((D, K), "normal", 1 / math.sqrt(D))
            bn(vs + "/cluster_bn", K)
            sp[vs + "/cluster_weights2"] = ((1, D, K), "normal", 1 / math.sqrt(D))
            a = name + "_attention"
            for n in ("q", "k", "v"):
                dn(a + "/" + n, D, D, bias=False)
            dn(a + "/output_transform", D, D)
            for s in ("/LayerNorm", "/LayerNorm_1", "/LayerNorm_2"):
                ln(a + s, D)
            dn(a + "/filter_output" + sid, D, 4 * D)
            dn(a + "/ff_output" + sid, 4 * D, D)
        else:
            a = vs + "/cluster_attention"
            for n in ("q", "k", "v"):
                dn(a + "/" + n, D, D, bias=False)
            bn(a + "/logits_bn", iterations)
            bn(a + "/attention_bn", D)
            dn(a + "/output_transform", D, D)
            ln(a + "/LayerNorm", D)
            dn(a + "/filter_outputencode", D, 4 * D)
            bn(a + "/filter_bn", 4 * D)
            dn(a + "/ff_outputencode", 4 * D, K)
            bn(a + "/feed_output_bn", K)
            sp[vs + "/cluster_centers"] = ((D, K), "normal", 1 / math.sqrt(D))
        vdim += D * K
    sp["hidden1_weights"] = ((vdim, hidden_size), "normal", 1 / math.sqrt(cluster_size))
    sp["hidden1_biases"] = ((hidden_size,), "normal", 0.01)
    sp["gating_weights_2"] = ((hidden_size, hidden_size), "normal", 1 / math.sqrt(hidden_size))
    bn("gating_bn", hidden_size)
    sp["gates/weights"] = ((hidden_size, vocab_size * (num_mixtures + 1)), "glorot", None)
    sp["experts/weights"] = ((hidden_size, vocab_size * num_mixtures), "glorot", None)
    sp["experts/biases"] = ((vocab_size * num_mixtures,), "zeros", None)
    return sp


def init_params(specs, seed=1810, dtype=torch.float32, perturb=0.0):
    """Seeded initial values.  `perturb` > 0 jitters the ones/zeros parameters
    (BN/LN affine, biases, moving stats) so parity tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    P, S = {}, {}
    for name in sorted(specs):
        shape, kind, arg = specs[name]
        if kind == "normal":
            t = torch.randn(shape, generator=g, dtype=torch.float32) * arg
        elif kind == "glorot":
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * lim
        elif kind == "zeros":
            t = torch.zeros(shape)
            if perturb:
                t = t + perturb * torch.randn(shape, generator=g)
        else:
            t = torch.ones(shape)
            if perturb:
                t = t + perturb * torch.rand(shape, generator=g)
        t = t.to(dtype)
        if name.endswith(TRAINABLE_EXCLUDE):
            S[name] = t
        else:
            P[name] = t.requires_grad_(True)
    return P, S


def dequantize(feat_vector, max_quantized_value=2, min_quantized_value=-2):
    """utils.py:28-43 Dequantize: byte codes -> floats."""
    assert max_quantized_value > min_quantized_value
    quantized_range = max_quantized_value - min_quantized_value
    scalar = quantized_range / 255.0
    bias = (quantized_range / 512.0) + min_quantized_value
    return feat_vector * scalar + bias


def synthetic_batch(batch, *, seed, max_frames=300, feat=1152, vocab=3862, fixed_num_frames=None,
                    dtype=torch.float32, return_codes=False, video_scale=0.0):
    """uint8 codes from clipped N(0,1), dequantised (utils.py:28-43),
    zero-padded to max_frames (readers.py:193), per-frame L2 normalised (train.py:264)."""
    g = torch.Generator().manual_seed(seed)
    if fixed_num_frames is None:
        nf = torch.randint(1, max_frames + 1, (batch,), generator=g, dtype=torch.int32)
    else:
        nf = torch.full((batch,), fixed_num_frames, dtype=torch.int32)
    z = torch.randn(batch, max_frames, feat, generator=g)
    if video_scale > 0:
        # videos that differ from each other (a per-video offset in feature space): i.i.d. noise videos are nearly
        # identical after pooling, which makes every batch statistic downstream degenerate
        z = (z + video_scale * torch.randn(batch, 1, feat, generator=g)) / math.sqrt(1.0 + video_scale ** 2)
    q = torch.clamp(torch.round((z + 2) * 255 / 4), 0, 255)
    x = dequantize(q)
    mask = (torch.arange(max_frames)[None, :] < nf[:, None]).to(x.dtype)
    x = x * mask[:, :, None]
    x = l2_normalize(x, 2)
    # labels: 1 + Poisson(2) positives, Zipf(1.0) over classes
    w = 1.0 / torch.arange(1, vocab + 1, dtype=torch.float64)
    labels = torch.zeros(batch, vocab, dtype=torch.bool)
    npos = 1 + torch.poisson(torch.full((batch,), 2.0), generator=g).long()
    for b in range(batch):
        cls = torch.multinomial(w, int(npos[b]), replacement=False, generator=g)
        labels[b, cls] = True
    if return_codes:
        # what the reader holds before Dequantize (readers.py:185-193); padded frames carry arbitrary codes here
        # (they are never sampled: model_utils.py:101-122 draws indices < num_frames)
        return x.to(dtype), nf, labels, q.to(torch.uint8)
    return x.to(dtype), nf, labels
