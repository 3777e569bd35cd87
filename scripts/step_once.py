"""A few train + infer steps at config 1 (for ncu launch lists): step_once.py [steps] [model]."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from learnablepoolingmethods_b200.trainer import Trainer
dev = torch.device("cuda:0")
C = bench.CFG
store = variables.VariableStore(dev, seed=1810)
model = sys.argv[2] if len(sys.argv) > 2 else "NetVladV1"
eng = NetVladEngine(NetVladConfig(model=model, iterations=C["iterations"], cluster_size=C["cluster_size"], hidden_size=C["hidden_size"], vocab_size=C["vocab"]), store)
tr = Trainer(eng, base_learning_rate=2e-4, learning_rate_decay=0.85, batch_size=C["batch"])
x, nf, lab = bench.synthetic(C["batch"], 20181000, device=dev, codes=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    tr.train_step(x, nf, lab)
torch.cuda.synchronize()
print("overflow", tr.overflowed())
with torch.no_grad():
    eng.forward(x, nf, False)      # first inference call after training: refreshes the low-order half of hidden1_weights
    torch.cuda.synchronize()
    eng.forward(x, nf, False)      # steady state (the one scripts/launch_summary.py reports)
torch.cuda.synchronize()
