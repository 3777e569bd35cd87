"""Registry / trainer helpers with the reference's names (train.py:187-190, utils.py:170-213)."""
from __future__ import annotations

import torch


def Dequantize(feat_vector, max_quantized_value=2, min_quantized_value=-2):
    """utils.py:28-43: byte codes -> floats, `q * range/255 + range/512 + min` (readers.py:185-193 calls it per feature).
    Host-side helper for callers that want fp32 frames; the model entry points take the uint8 codes directly and
    dequantise inside the gather kernels."""
    assert max_quantized_value > min_quantized_value
    quantized_range = max_quantized_value - min_quantized_value
    scalar = quantized_range / 255.0
    bias = (quantized_range / 512.0) + min_quantized_value
    return feat_vector * scalar + bias


def find_class_by_name(name, modules):
    """Searches the provided modules for the named class and returns it (train.py:187-190)."""
    modules = [getattr(module, name, None) for module in modules]
    return next(a for a in modules if a)


def clip_gradient_norms(gradients_to_variables, max_norm):
    """utils.py:170-189 on (grad, var) pairs of torch tensors: tf.clip_by_norm per tensor.  The training loop
    does this inside the fused optimiser kernel; this helper mirrors the reference API for callers that
    handle gradients themselves."""
    out = []
    for grad, var in gradients_to_variables:
        if grad is not None:
            n = torch.linalg.vector_norm(grad)
            grad = grad * (max_norm / torch.clamp(n, min=max_norm))
        out.append((grad, var))
    return out


def combine_gradients(tower_grads):
    """utils.py:192-213: sum each variable's gradient over the towers."""
    filtered = [[x for x in gl if x[0] is not None] for gl in tower_grads]
    final = []
    for i in range(len(filtered[0])):
        grads = [filtered[t][i] for t in range(len(filtered))]
        final.append((torch.stack([g[0] for g in grads], 0).sum(0), filtered[0][i][1]))
    return final
