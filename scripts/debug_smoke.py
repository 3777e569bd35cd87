import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import _lib, ops, variables
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from oracle import netvlad_oracle as O
dev = torch.device("cuda:0")
B, K, Hd, V, T = 4, 64, 64, 100, 256
store = variables.VariableStore(dev, seed=1810)
eng = NetVladEngine(NetVladConfig(iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V), store)
x, nf, labels = O.synthetic_batch(B, seed=20181000, vocab=V)
pred, ctx = eng.forward(x.to(dev), nf.to(dev), True, save_for_backward=True)
lab = labels.to(torch.uint8).to(dev)
loss, _ = ops.xent_fwd(pred, lab)
order = []
ctx["grad_hook"] = lambda n, g: order.append((n, bool(torch.isfinite(g).all()), float(g.abs().max())))
grads = eng.backward(ctx, ops.xent_bwd(pred, lab, 1.0 / B))
for o in order: print(o)
for name in ("video", "audio"):
    m = ctx[name]
    for k, t in m.items():
        if torch.is_tensor(t) and t.is_floating_point():
            print(name, k, tuple(t.shape), "finite", bool(torch.isfinite(t.float()).all()), "absmax", float(t.float().abs().max()))
