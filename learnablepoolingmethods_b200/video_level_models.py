"""video_level_models.MoeModel (video_level_models.py:48-159) on the hot-path kernels."""
from __future__ import annotations

import torch

from . import models, ops, variables
from .flags import FLAGS, ensure_parsed


class MoeModel(models.BaseModel):
    """A softmax over a mixture of logistic models (low_rank_gating=-1, prob gating off: the configuration
    NetVladV1/V2 use).  Forward only at this boundary; training runs through NetVladV1/V2.create_model."""

    def create_model(self, model_input, vocab_size, is_training=True, num_mixtures=None, l2_penalty=1e-8,
                     **unused_params):
        ensure_parsed()
        M = num_mixtures or FLAGS.moe_num_mixtures
        if FLAGS.moe_low_rank_gating != -1 or FLAGS.moe_prob_gating:
            raise NotImplementedError("moe_low_rank_gating / moe_prob_gating are outside the NetVlad hot path")
        s = unused_params.get("store") or variables.default_store()
        H = int(model_input.shape[1])
        V = int(vocab_size)
        with s.variable_scope("gates"):
            wg = s.get_variable("weights", (H, V * (M + 1)), "glorot")
        with s.variable_scope("experts"):
            we = s.get_variable("weights", (H, V * M), "glorot")
            be = s.get_variable("biases", (V * M,), "zeros")
        g8, e8 = (V * (M + 1) + 7) // 8 * 8, (V * M + 7) // 8 * 8
        w16 = torch.zeros((H, g8 + e8), dtype=torch.float16, device=model_input.device)
        ops.cast_f16(wg, w16[:, :g8], cols_dst=g8)
        ops.cast_f16(we, w16[:, g8:], cols_dst=e8)
        bias = torch.zeros(g8 + e8, dtype=torch.float32, device=model_input.device)
        bias[g8:g8 + V * M].copy_(be)
        x16 = model_input if model_input.dtype == torch.float16 else ops.cast_f16(model_input.contiguous().float())
        logits = ops.gemm(x16, w16, bias=bias, out_dtype=torch.float32)
        return {"predictions": ops.moe_mix_fwd(logits, V, M, expert_off=g8)}
