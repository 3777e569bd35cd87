import os, sys, torch
sys.path.insert(0, "/root/repo")
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from oracle import netvlad_oracle as O
dev = torch.device("cuda:0")
res = {}
for fused in (True, False):
    store = variables.VariableStore(dev, seed=7)
    eng = NetVladEngine(NetVladConfig(model="NetVladV1", iterations=256, cluster_size=256, hidden_size=512, vocab_size=3862, fused_gating="all" if fused else "off"), store)
    x, nf, _ = O.synthetic_batch(80, seed=3, vocab=3862, video_scale=1.0)
    for training in (True, False):
        pred, ctx = eng.forward(x.to(dev), nf.to(dev), training, save_for_backward=training, return_intermediates=True)
        torch.cuda.synchronize()
        res[(fused, training)] = (pred.clone(), ctx["inter"]["hidden"].clone(), ctx["inter"]["gated"].clone())
for training in (True, False):
    a, b = res[(True, training)], res[(False, training)]
    print("training" if training else "inference", "pred max-abs diff", float((a[0] - b[0]).abs().max()),
          "hidden rel", float((a[1] - b[1]).norm() / b[1].norm()), "gated rel", float((a[2] - b[2]).norm() / b[2].norm()),
          "gated max-abs", float((a[2] - b[2]).abs().max()))
