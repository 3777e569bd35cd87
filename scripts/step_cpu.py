"""Host-side cost of issuing one training step: tiny batch (the GPU finishes early), wall clock per step."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from learnablepoolingmethods_b200 import variables, _lib
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
l0 = _lib.launch_count
t0 = time.perf_counter()
n = 100
for _ in range(n):
    tr.train_step(x, nf, lab)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time per train step: {1e3 * (t1 - t0) / n:.3f} ms ({(_lib.launch_count - l0) // n} launches); drained after {1e3 * (t2 - t1):.2f} ms")
with torch.no_grad():
    for _ in range(5):
        eng.forward(x, nf, False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        eng.forward(x, nf, False)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
print(f"host issue time per inference forward: {1e3 * (t1 - t0) / n:.3f} ms")
