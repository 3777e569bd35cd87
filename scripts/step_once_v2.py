"""A few NetVladV2 train + infer steps at config-1 shape (for ncu launch lists)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from learnablepoolingmethods_b200.trainer import Trainer
dev = torch.device("cuda:0")
C = bench.CFG
store = variables.VariableStore(dev, seed=1810)
eng = NetVladEngine(NetVladConfig(model="NetVladV2", iterations=C["iterations"], cluster_size=C["cluster_size"], hidden_size=C["hidden_size"], vocab_size=C["vocab"]), store)
tr = Trainer(eng, batch_size=C["batch"])
x, nf, lab = bench.synthetic(C["batch"], 20181000, device=dev, codes=True)
for _ in range(3):
    tr.train_step(x, nf, lab)
torch.cuda.synchronize()
with torch.no_grad():
    eng.forward(x, nf, False)
torch.cuda.synchronize()
