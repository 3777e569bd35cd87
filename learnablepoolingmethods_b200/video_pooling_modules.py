"""Pooling modules of the hot path with the reference's class names and constructor signatures, forward only,
on the fused pooling kernel (K1):

  NetVladAttenCluster  video_pooling_modules.py:1592-1663   (NetVladV2's pooling: attention-based cluster similarities)
  NetVladOrthoReg      video_pooling_modules.py:1499-1586   (baseline NetVLAD + orthogonal regulariser)

`forward(inputs: [(B*max_frames), feature_size]) -> [B, cluster_size*feature_size]` fp32, d-major flatten (index
d*K + k) as in the reference.  Variables live in the store's current `variable_scope` under the TF names.
"""
from __future__ import annotations

import math

import torch

from . import modules, ops, variables
from .transformer_utils import TransformerEncoderMod, _f16_rows


class NetVladAttenCluster(modules.BaseModule):
    """ NetVLAD whose soft assignments come from a transformer encoder over the frames. """

    def __init__(self, feature_size, max_frames, cluster_size, batch_norm, is_training, scope_id=None):
        self.feature_size, self.max_frames, self.cluster_size = feature_size, max_frames, int(cluster_size)
        self.batch_norm, self.is_training, self.scope_id = batch_norm, is_training, scope_id
        # attention encoder parameters (video_pooling_modules.py:1610-1615)
        self.encoder_hidden_size = feature_size
        self.num_heads = feature_size // 16
        self.dropout_ratio = 0.1
        self.filter_size = 4 * self.encoder_hidden_size

    def forward(self, inputs, store=None, dropout_mask=None, seed=0, **unused_params):
        s = store or variables.default_store()
        D, K, T = self.feature_size, self.cluster_size, self.max_frames
        x16 = _f16_rows(inputs, D)
        B = x16.shape[0] // T
        enc = TransformerEncoderMod(feature_size=D, hidden_size=self.encoder_hidden_size, num_heads=self.num_heads,
                                    attention_dropout=self.dropout_ratio, ff_filter_size=self.filter_size,
                                    ff_relu_dropout=self.dropout_ratio, is_train=self.is_training, scope_id=self.scope_id,
                                    final_size=K)
        with s.variable_scope("cluster_attention"):                  # :1628
            A = enc.forward(x16.view(B, T, D), store=s, dropout_mask=dropout_mask, seed=seed, as_f16=True)   # [B*T, K]
        centers = s.get_variable("cluster_centers", (D, K), "normal", 1 / math.sqrt(D))    # :1640-1644
        z, rs, _, _ = ops.netvlad_pool_fwd(x16, B, T, None, None, None, centers, assign_in=A)
        return ops.netvlad_finalize(z, rs, d_major=True)


class NetVladOrthoReg(modules.BaseModule):
    """ NetVLAD from WILLOW's model with orthogonal regularization. """

    def __init__(self, feature_size, max_frames, cluster_size, batch_norm, is_training, det_reg=None, scope_id=None):
        self.feature_size, self.max_frames, self.cluster_size = feature_size, max_frames, int(cluster_size)
        self.batch_norm, self.is_training, self.det_reg, self.scope_id = batch_norm, is_training, det_reg, scope_id
        self._centers = None

    def forward(self, inputs, store=None, **unused_params):
        s = store or variables.default_store()
        D, K, T = self.feature_size, self.cluster_size, self.max_frames
        sid = "" if self.scope_id is None else str(self.scope_id)
        wc = s.get_variable("cluster_weights" + sid, (D, K), "normal", 1 / math.sqrt(D))      # :1527-1531
        x16 = _f16_rows(inputs, D)
        B = x16.shape[0] // T
        wc16 = ops.cast_f16(wc)
        if self.batch_norm:
            beta, gamma, mm, mv = s.batch_norm_vars("cluster_bn", K)
            if self.is_training:
                _, st = ops.gemm(wc16, x16, a_mn=True, b_mn=False, out="none", stats=True)
                scale, shift = ops.bn_finalize(st[0].reshape(-1, K), st[1].reshape(-1, K), B * T, gamma, beta, mm, mv,
                                               training=True, bessel=True)
            else:
                scale, shift = ops.bn_finalize(None, None, 1, gamma, beta, mm, mv, training=False, bessel=True)
        else:
            scale = torch.ones(K, device=x16.device)
            shift = s.get_variable("cluster_biases" + sid, (K,), "normal", 1 / math.sqrt(D))   # :1545-1552
        self._centers = s.get_variable("cluster_weights2", (D, K), "normal", 1 / math.sqrt(D))  # :1561-1571
        z, rs, _, _ = ops.netvlad_pool_fwd(x16, B, T, wc16, scale, shift, self._centers)
        return ops.netvlad_finalize(z, rs, d_major=True)

    def regularization_loss(self):
        """The regulariser TF attaches to cluster_weights2 when det_reg is given (:1566-1571; module_utils.py:55-90):
        a device scalar, or None (det_reg None / 0.0 disables it, module_utils.py:67-69)."""
        if self._centers is None:
            raise RuntimeError("call forward() first (the variable is created there, as in the reference)")
        if self.det_reg is None or float(self.det_reg) == 0.0:
            return None
        if isinstance(self.det_reg, int):
            raise ValueError("scale cannot be an integer: %s" % (self.det_reg,))          # module_utils.py:62-63
        if self.det_reg < 0.0:
            raise ValueError("Setting a scale less than 0 on a regularizer: %g." % self.det_reg)
        return ops.ortho_reg(self._centers, float(self.det_reg))[0]
