"""Attention modules of the hot path with the reference's class names and constructor signatures
(transformer_utils.py:374-457, 507-767), forward only, on the C-ABI kernels.

`TransformerEncoder` is the block NetVladV1 runs over the K cluster descriptors (frame_level_models.py:2282-2304);
`TransformerEncoderMod` is the assignment network of NetVladV2 (video_pooling_modules.py:1605-1615, 1628-1638).
Variables are created in the store's current `variable_scope` under the TF names.  Training with the
hand-written backward runs through `frame_level_models.NetVladV1/NetVladV2.create_model`; these classes are the
module-level boundary (`modules.BaseModule.forward(inputs, **unused_params)`).
"""
from __future__ import annotations

import torch

from . import modules, ops, variables


def _f16_rows(x: torch.Tensor, D: int) -> torch.Tensor:
    x = x.reshape(-1, D)
    return x.contiguous() if x.dtype == torch.float16 else ops.cast_f16(x.contiguous().float())


def _dense(s, scope, i, o, bias=True):
    with s.variable_scope(scope):
        k = s.get_variable("kernel", (i, o), "glorot")
        b = s.get_variable("bias", (o,), "zeros") if bias else None
    return k, b


def _ln(s, scope, c):
    with s.variable_scope(scope):
        return s.get_variable("gamma", (c,), "ones"), s.get_variable("beta", (c,), "zeros")


def _qkv16(s, D):
    w = torch.empty((D, 3 * D), dtype=torch.float16, device=s.device)
    for i, n in enumerate(("q", "k", "v")):
        k, _ = _dense(s, n, D, D, bias=False)                     # tf.layers.dense(use_bias=False) (:559-561, 641-643)
        ops.cast_f16(k, w[:, i * D:(i + 1) * D], cols_dst=D)
    return w


class TransformerEncoder(modules.BaseModule):
    """transformer_utils.py:374-413 (+ MultiHeadAttention :507-586, FeedForwardNetwork :679-715).
    out = LN(LN(ffn(h1) + h1) + h1), h1 = LN(MHA(x, x) + x); joint-axis layer norms, no dropout (commented out in the
    reference, :576-577, 705-706)."""

    def __init__(self, feature_size, hidden_size, num_heads, attention_dropout, ff_filter_size, ff_relu_dropout,
                 is_train, scope_id):
        self.feature_size, self.hidden_size, self.num_heads = feature_size, hidden_size, num_heads
        self.attention_dropout, self.ff_filter_size, self.ff_relu_dropout = attention_dropout, ff_filter_size, ff_relu_dropout
        self.is_train, self.scope_id = is_train, scope_id

    def forward(self, inputs, store=None, **unused_params):
        """inputs: [batch_size, input_length, hidden_size] (fp32 or fp16, GPU) -> fp32 of the same shape."""
        s = store or variables.default_store()
        B, L, D = inputs.shape
        H, Fs, sid = self.num_heads, self.ff_filter_size, str(self.scope_id)
        if D != self.hidden_size:
            raise ValueError(f"inputs must be [B, L, {self.hidden_size}], got {tuple(inputs.shape)}")
        x = _f16_rows(inputs, D)
        wqkv = _qkv16(s, D)
        wo, bo = _dense(s, "output_transform", D, D)
        g1, b1 = _ln(s, "LayerNorm", D)                             # auto-uniquified names in call order (:407, 713, 411)
        w1, c1 = _dense(s, "filter_output" + sid, D, Fs)
        w2, c2 = _dense(s, "ff_output" + sid, Fs, D)
        g2, b2 = _ln(s, "LayerNorm_1", D)
        g3, b3 = _ln(s, "LayerNorm_2", D)
        qkv = ops.gemm(x, wqkv)
        o = ops.mha_core_fwd(qkv, B, L, D, H, scale=(D // H) ** -0.5)
        att = ops.gemm(o, ops.cast_f16(wo), bias=bo)
        h1 = ops.layernorm_joint_fwd(att, x, None, B, L, D, g1, b1)
        f1 = ops.gemm(h1.view(B * L, D), ops.cast_f16(w1), bias=c1, relu=True)
        f2 = ops.gemm(f1, ops.cast_f16(w2), bias=c2, relu=True)   # ReLU on the output layer too (:708-711)
        h2 = ops.layernorm_joint_fwd(f2, h1, None, B, L, D, g2, b2)
        out = ops.layernorm_joint_fwd(h2, h1, None, B, L, D, g3, b3)
        return out.float().view(B, L, D)


class TransformerEncoderMod(modules.BaseModule):
    """transformer_utils.py:415-457 (+ MultiHeadAttentionBN :589-677, FeedForwardNetworkMod :718-767):
    BN(relu(BN(relu(h1 W1 + b1)) W2 + b2)), h1 = LN(dropout(MHA_BN(x, x)) + x) -> [B, L, final_size]; no softmax."""

    def __init__(self, feature_size, hidden_size, num_heads, attention_dropout, ff_filter_size, ff_relu_dropout,
                 is_train, scope_id, final_size):
        self.feature_size, self.hidden_size, self.num_heads = feature_size, hidden_size, num_heads
        self.attention_dropout, self.ff_filter_size, self.ff_relu_dropout = attention_dropout, ff_filter_size, ff_relu_dropout
        self.is_train, self.scope_id, self.final_size = is_train, scope_id, final_size

    def forward(self, inputs, store=None, dropout_mask=None, seed=0, as_f16=False, **unused_params):
        """inputs: [B, L, hidden] -> fp32 [B, L, final_size].  dropout_mask (fp16 0/1, [B*L, hidden]) replaces the
        generated keep-mask of tf.layers.dropout(rate = 1 - attention_dropout) (:450)."""
        s = store or variables.default_store()
        B, L, D = inputs.shape
        H, Fs, K, train = self.num_heads, self.ff_filter_size, int(self.final_size), bool(self.is_train)
        if D // H != 16 or K % 8:
            raise NotImplementedError("this path implements head depth 16 (feature_size // 16 heads) and final sizes "
                                      "that are multiples of 8")
        x = _f16_rows(inputs, D)
        wqkv = _qkv16(s, D)
        lbn = s.batch_norm_vars("logits_bn", L)                      # channel = key axis: tied to L (:653)
        abn = s.batch_norm_vars("attention_bn", D)
        wo, bo = _dense(s, "output_transform", D, D)
        g1, b1 = _ln(s, "LayerNorm", D)
        w1, c1 = _dense(s, "filter_outputencode", D, Fs)
        fbn = s.batch_norm_vars("filter_bn", Fs)
        w2, c2 = _dense(s, "ff_outputencode", Fs, K)
        obn = s.batch_norm_vars("feed_output_bn", K)

        def bnv(t):                                                  # (beta, gamma, mm, mv) -> (gamma, beta, mm, mv)
            return (t[1], t[0], t[2], t[3])

        qkv = ops.gemm(x, wqkv)
        if train:
            part = ops.mha_logit_stats(qkv, B, L, D, H)
            r = ops.bn_finalize(part[:, 0], part[:, 1], B * H * L, *bnv(lbn), training=True, bessel=True, psum_stride=2 * L)
        else:
            r = ops.bn_finalize(None, None, 1, *bnv(lbn), training=False, bessel=True)
        o = ops.mha_core_fwd(qkv, B, L, D, H, scale=1.0, key_scale=r[0], key_shift=r[1])
        ops.batch_norm_cols_f16(o, *bnv(abn), training=train, bessel=False)
        att = ops.gemm(o, ops.cast_f16(wo), bias=bo)
        rate = 1.0 - float(self.attention_dropout)
        if train and rate > 0:
            ops.dropout_f16(att, rate, mask_in=dropout_mask, seed=seed)
        h1 = ops.layernorm_joint_fwd(att, x, None, B, L, D, g1, b1)
        f = ops.gemm(h1.view(B * L, D), ops.cast_f16(w1), bias=c1, relu=True)
        ops.batch_norm_cols_f16(f, *bnv(fbn), training=train, bessel=False)
        f2 = ops.gemm(f, ops.cast_f16(w2), bias=c2, relu=True)
        ops.batch_norm_cols_f16(f2, *bnv(obn), training=train, bessel=False)
        if as_f16:
            return f2
        return f2.float().view(B, L, K)
