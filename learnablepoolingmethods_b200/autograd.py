"""torch.autograd bridge at the model boundary: one Function for the whole NetVlad path.

Forward and backward both run the hand-written CUDA kernels through the engine; torch only carries
the graph edge from `predictions` to the parameters (so `loss.backward()` fills `.grad`)."""
from __future__ import annotations

import torch


class NetVladFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, model_input, num_frames, dropout_masks, frame_index, reg_penalty, names, *params):
        pred, ectx = engine.forward(model_input, num_frames, True, save_for_backward=True,
                                    dropout_masks=dropout_masks, frame_index=frame_index)
        # WillowModelReg: the orthogonal regulariser's gradient is added inside the hand-written backward, scaled by the
        # caller's --regularization_penalty (train.py:301-303, 323-324); `result["regularization_loss"]` is its VALUE only
        ectx["reg_penalty"] = float(reg_penalty)
        ctx.engine, ctx.ectx, ctx.names = engine, ectx, names
        return pred

    @staticmethod
    def backward(ctx, dpred):
        grads = ctx.engine.backward(ctx.ectx, dpred.contiguous())
        out = tuple(grads.get(n) for n in ctx.names)
        ctx.ectx = None
        return (None, None, None, None, None, None, None) + out


def netvlad_apply(engine, model_input, num_frames, is_training, dropout_masks=None, frame_index=None, reg_penalty=1.0):
    train_graph = is_training and torch.is_grad_enabled()
    if not train_graph:
        pred, _ = engine.forward(model_input, num_frames, is_training, dropout_masks=dropout_masks, frame_index=frame_index)
        return pred
    tr = engine.store.trainable()
    names = tuple(tr.keys())
    params = tuple(tr[n] for n in names)
    for p in params:
        if not p.requires_grad:
            p.requires_grad_(True)
    return NetVladFunction.apply(engine, model_input, num_frames, dropout_masks, frame_index, reg_penalty, names, *params)
