"""Freeze the outputs of the REFERENCE's own metric code (eval_util.py / average_precision_calculator.py, imported
unmodified from /root/reference) on seeded batches into tests/golden/eval_golden.npz.
Run in the build container (the GPU box has no /root/reference):  python oracle/make_eval_golden.py"""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import eval_util  # noqa: E402  (the reference module)
from oracle import eval_oracle as E  # noqa: E402

CASES = [(101, 48, 3862, ()), (102, 7, 50, (3,)), (103, 5, 12, (0, 4)), (104, 80, 3862, (10,))]
out = {}
for seed, B, V, zero in CASES:
    pred, labels = E.synthetic_eval_batch(seed, B, V, zero)
    out[f"case{seed}"] = np.array([seed, B, V] + list(zero), dtype=np.int64)
    out[f"hit{seed}"] = np.float64(eval_util.calculate_hit_at_one(pred, labels))
    out[f"perr{seed}"] = np.float64(eval_util.calculate_precision_at_equal_recall_rate(pred, labels))
    out[f"gap{seed}"] = np.float64(eval_util.calculate_gap(pred, labels))
    k = min(20, V)
    trip = [sorted(int(t[0]) for t in eval_util.top_k_triplets(pred[b], labels[b], 20)) for b in range(B)]
    out[f"topk{seed}"] = np.array(trip, dtype=np.int32).reshape(B, k)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "eval_golden.npz"), **out)
print({k: (v if v.ndim == 0 else v.shape) for k, v in out.items()})
