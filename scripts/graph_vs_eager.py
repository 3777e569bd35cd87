"""Config-1-shaped check that the graph-replayed training step reproduces the eager step bit for bit (debug aid)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from learnablepoolingmethods_b200.trainer import Trainer
dev = torch.device("cuda:0")
C = bench.CFG
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
runs = []
for mode in sys.argv[2:] or ["eager", "graph"]:
    store = variables.VariableStore(dev, seed=1810)
    cfg = NetVladConfig(iterations=C["iterations"], cluster_size=C["cluster_size"], hidden_size=C["hidden_size"], vocab_size=C["vocab"])
    if "nowgrad" in mode:
        cfg.overlap_wgrad = False
    if "noaudio" in mode:
        cfg.overlap_audio = False
    eng = NetVladEngine(cfg, store)
    tr = Trainer(eng, base_learning_rate=2e-4, learning_rate_decay=0.85, batch_size=C["batch"])
    tr.use_graph = mode.startswith("graph")
    losses, grads = [], None
    for i in range(steps):
        x, nf, lab = bench.synthetic(C["batch"], 20181000 + i, device=dev, codes=True)
        losses.append(float(tr.train_step(x, nf, lab)))
        if i == steps - 1:
            grads = tr.flat.g.clone()
    runs.append((mode, losses, {k: v.clone() for k, v in store.vars.items()}, grads))
    del tr, eng, store
    torch.cuda.empty_cache()
base = runs[0]
for mode, losses, vars_, grads in runs[1:]:
    print(f"== {mode} vs {base[0]}: losses {losses} | {base[1]}")
    print("   last-step flat gradient: max abs diff", float((grads - base[3]).abs().max()), "of", float(base[3].abs().max()))
    bad = [(k, float((vars_[k] - base[2][k]).abs().max())) for k in vars_ if not torch.equal(vars_[k], base[2][k])]
    print(f"   {len(bad)} of {len(vars_)} variables differ", bad[:12])
