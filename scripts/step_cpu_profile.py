import sys, os, cProfile, pstats, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from learnablepoolingmethods_b200.trainer import Trainer
dev = torch.device("cuda:0")
C = bench.CFG
store = variables.VariableStore(dev, seed=1810)
eng = NetVladEngine(NetVladConfig(iterations=C["iterations"], cluster_size=64, hidden_size=64, vocab_size=100), store)
B = 2
tr = Trainer(eng, batch_size=B)
x, nf, lab = bench.synthetic(B, 1, device=dev, codes=True)
lab = lab[:, :100].contiguous()
for _ in range(5):
    tr.train_step(x, nf, lab)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    tr.train_step(x, nf, lab)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
