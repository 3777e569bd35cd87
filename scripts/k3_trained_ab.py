"""Trained-weights acceptance statistics (tests/trained_protocol.py, 300 steps) for the fused K3 head and the four-launch head:
trained with either, evaluated with both, so that the effect of the evaluation path and of the trained state separate."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import trained_protocol as TP
dev = torch.device("cuda:0")
out = {}
for train_fused in (True, False):
    import dataclasses, learnablepoolingmethods_b200.engine as E
    E.NetVladConfig.__dataclass_fields__["fused_gating"].default = "all" if train_fused else "off"
    _init = E.NetVladConfig.__init__
    E.NetVladConfig.__init__ = lambda self, *a, _i=_init, _m=("all" if train_fused else "off"), **k: (_i(self, *a, **k), setattr(self, "fused_gating", k.get("fused_gating", _m)))[0]
    eng, tr, protos, losses = TP.train_model(dev, steps=300)
    E.NetVladConfig.__init__ = _init
    for eval_fused in (True, False):
        eng.cfg.fused_gating = "all" if eval_fused else "off"
        rep = TP.evaluate(eng, protos, dev, n_videos=1024)
        key = f"trained_{'fused' if train_fused else 'plain'}__eval_{'fused' if eval_fused else 'plain'}"
        out[key] = {k: rep[k] for k in ("pred_max_abs", "top20_identical", "top20_identical_up_to_ties", "hidden_rel_l2", "gated_rel_l2", "gap_gpu", "gap_oracle")}
        print(key, json.dumps(out[key]))
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/k3_trained_ab.json", "w"), indent=1)
