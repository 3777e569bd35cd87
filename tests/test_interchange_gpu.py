"""SURVEY 8f row 3 on the GPU: inference.py-format CSV from device top-k, and a trainer -> TF-format checkpoint ->
trainer round trip (weights, Adam moments, global step) that continues bit-identically."""
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_inference_csv_matches_reference_format(cuda):
    from learnablepoolingmethods_b200 import inference
    from oracle import eval_oracle as E
    rng = np.random.RandomState(3)
    pred = rng.rand(37, 3862).astype(np.float32)          # distinct scores: the order is then fully determined
    ids = [("vid%04d" % i).encode() for i in range(37)]
    want = list(E.format_lines(ids, pred, 20))
    got = list(inference.format_lines(ids, torch.from_numpy(pred).to(cuda), 20))
    assert got == want
    buf = io.StringIO()
    n = inference.write_predictions(buf, [(ids[:20], torch.from_numpy(pred[:20]).to(cuda)),
                                          (ids[20:], torch.from_numpy(pred[20:]).to(cuda))], top_k=20)
    assert n == 37 and buf.getvalue() == "VideoId,LabelConfidencePairs\n" + "".join(want)


def test_checkpoint_round_trip_continues_training(cuda, tmp_path):
    from learnablepoolingmethods_b200 import checkpoint as ck, variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from learnablepoolingmethods_b200.trainer import Trainer
    from oracle import netvlad_oracle as O
    B, K, Hd, V, T = 4, 64, 64, 100, 128
    cfg = NetVladConfig(model="NetVladV1", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V)
    batches = [O.synthetic_batch(B, seed=50 + i, vocab=V) for i in range(3)]

    def step(tr, i):
        x, nf, lab = batches[i]
        return float(tr.train_step(x.to(cuda), nf.to(cuda), lab.to(torch.uint8).to(cuda)))

    a = variables.VariableStore(cuda, seed=4)
    tra = Trainer(NetVladEngine(cfg, a), base_learning_rate=2e-4, batch_size=B)
    step(tra, 0); step(tra, 1)
    prefix = str(tmp_path / "model.ckpt-2")
    ck.write_model_flags(str(tmp_path), ck.model_flags("NetVladV1"))
    ck.save_from_store(a, prefix, trainer=tra)
    names = ck.list_tf_checkpoint(prefix)
    assert names["global_step"]["dtype"] == 3 and "tower/hidden1_weights/Adam_1" in names and "beta2_power" in names
    assert ck.latest_checkpoint(str(tmp_path)) == prefix
    la = step(tra, 2)

    b = variables.VariableStore(cuda, seed=99)            # different initial values: everything must come from the file
    trb = Trainer(NetVladEngine(cfg, b), base_learning_rate=2e-4, batch_size=B)
    rep = ck.load_into_store(b, ck.latest_checkpoint(str(tmp_path)), trainer=trb)
    assert not rep["missing"] and not rep["unused"] and rep["global_step"] == 2 and trb.global_step == 2
    lb = step(trb, 2)
    assert la == lb
    for k in a.vars:
        assert torch.equal(a.vars[k], b.vars[k]), k
    # restoring INTO a trainer whose step is already replayed from a CUDA graph: values land in place (same addresses),
    # the fp16 operand shadows are refreshed before the next replay, and the step repeats exactly
    step(tra, 0)                                          # move away from the checkpointed state (graph replay)
    assert tra.graph is not None
    ck.load_into_store(a, prefix, trainer=tra)
    assert tra.global_step == 2 and step(tra, 2) == la
    for k in a.vars:
        assert torch.equal(a.vars[k], b.vars[k]), k
