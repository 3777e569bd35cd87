import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from oracle import netvlad_oracle as O
dev = torch.device("cuda:0")
B, K, Hd, V, T = 2, 64, 64, 100, 256
store = variables.VariableStore(dev, seed=1810)
eng = NetVladEngine(NetVladConfig(model="NetVladV2", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V), store)
x, nf, _ = O.synthetic_batch(B, seed=20181000, vocab=V)
pred, ctx = eng.forward(x.to(dev), nf.to(dev), False)
torch.cuda.synchronize()
print("infer ok", float(pred.mean()))
pred, ctx = eng.forward(x.to(dev), nf.to(dev), True)
torch.cuda.synchronize()
print("train fwd ok", float(pred.mean()))
