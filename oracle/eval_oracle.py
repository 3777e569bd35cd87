"""TEST INFRASTRUCTURE ONLY (see oracle/netvlad_oracle.py): numpy restatement of the reference's batch metrics.

Pinned against the reference's own code: `oracle/make_eval_golden.py` imports /root/reference/eval_util.py in this
container and freezes its outputs into tests/golden/eval_golden.npz; tests/test_oracle_cpu.py checks this restatement
against that file (and against the live reference module when /root/reference is mounted)."""
import numpy as np


def hit_at_one(pred, actuals):
    """eval_util.py:27-42."""
    top = np.argmax(pred, 1)
    return float(np.average(actuals[np.arange(actuals.shape[0]), top]))


def perr(pred, actuals):
    """eval_util.py:45-70 (a video without labels contributes 0: numpy's [-0:] slice is the whole row)."""
    tot = 0.0
    for r in range(actuals.shape[0]):
        n = int(np.sum(actuals[r]))
        if n == 0:
            continue
        order = np.argsort(-pred[r], kind="stable")[:n]
        tot += float(sum(actuals[r][i] for i in order if pred[r][i] > 0)) / n
    return tot / actuals.shape[0]


def top_k_sets(pred, k=20):
    """eval_util.py:128-135: the k best classes of every video as sorted index lists."""
    k = min(k, pred.shape[1])
    return [sorted(np.argsort(-pred[r], kind="stable")[:k].tolist()) for r in range(pred.shape[0])]


def gap(pred, actuals, k=20):
    """eval_util.py:73-91 + average_precision_calculator.py:203-262 (ties ranked by video, then class)."""
    k = min(k, pred.shape[1])
    items = []
    for r in range(pred.shape[0]):
        for i in np.argsort(-pred[r], kind="stable")[:k]:
            items.append((pred[r][i], actuals[r][i]))
    numpos = float(np.sum(actuals))
    if numpos == 0 or not items:
        return 0.0
    order = sorted(range(len(items)), key=lambda j: -items[j][0])
    ap, pos = 0.0, 0.0
    for rank, j in enumerate(order):
        if items[j][1] > 0:
            pos += 1
            ap += pos / (rank + 1) / numpos
    return ap


def synthetic_eval_batch(seed, B, V, zero_label_rows=()):
    """Scores and multi-hot labels from numpy's legacy generator (bit-stable across machines)."""
    rs = np.random.RandomState(seed)
    pred = rs.rand(B, V).astype(np.float32) ** 4              # skewed towards 0 like sigmoid outputs
    labels = np.zeros((B, V), dtype=np.uint8)
    for b in range(B):
        if b in zero_label_rows:
            continue
        n = 1 + rs.poisson(2.0)
        cls = rs.choice(V, size=min(n, V), replace=False)
        labels[b, cls] = 1
        pred[b, cls[: max(1, len(cls) // 2)]] += 0.5           # some positives score high
    return np.clip(pred, 0.0, 1.0).astype(np.float32), labels


def format_lines(video_ids, predictions, top_k):
    """inference.py:88-96 restated (numpy only; the reference module imports tensorflow at the top and cannot be
    imported here): argpartition for the top_k classes, sorted by descending score, "%i %g" pairs."""
    batch_size = len(video_ids)
    for video_index in range(batch_size):
        top_indices = np.argpartition(predictions[video_index], -top_k)[-top_k:]
        line = [(class_index, predictions[video_index][class_index]) for class_index in top_indices]
        line = sorted(line, key=lambda p: -p[1])
        yield video_ids[video_index].decode('utf-8') + "," + " ".join("%i %g" % (label, score) for (label, score) in line) + "\n"
