"""Does NetVladV1 fit a batch of DISTINCT videos?  (bench.py's i.i.d.-noise videos are nearly identical after pooling, so
its loss plateaus at the batch prior.)  usage: overfit_check.py [steps] [model] [video_scale]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from learnablepoolingmethods_b200.trainer import Trainer
from oracle import netvlad_oracle as O          # synthetic data only
dev = torch.device("cuda:0")
C = bench.CFG
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
model = sys.argv[2] if len(sys.argv) > 2 else "NetVladV1"
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
loss_scale = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0      # 0 = the engine's automatic choice (8 x batch)
store = variables.VariableStore(dev, seed=1810)
eng = NetVladEngine(NetVladConfig(model=model, iterations=C["iterations"], cluster_size=C["cluster_size"], hidden_size=C["hidden_size"],
                                  vocab_size=C["vocab"], loss_scale=loss_scale), store)
tr = Trainer(eng, base_learning_rate=2e-4, batch_size=C["batch"])
x, nf, lab = O.synthetic_batch(C["batch"], seed=7, vocab=C["vocab"], video_scale=scale)
x, nf, lab = x.to(dev), nf.to(dev), lab.to(torch.uint8).to(dev)
p = lab.float().mean(0)
prior = float(-(p * torch.log(p + 1e-5) + (1 - p) * torch.log(1 - p + 1e-5)).sum())
print(f"{model}, video_scale {scale}: loss of predicting the batch prior = {prior:.3f}")
for i in range(steps):
    loss = tr.train_step(x, nf, lab)
    if i % max(1, steps // 8) == 0 or i == steps - 1 or i in (10, 20, 40, 80):
        print(f"  step {i:5d}  loss {float(loss):.4f}", flush=True)
print("overflow", tr.overflowed())
