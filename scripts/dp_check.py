"""torchrun --nproc-per-node 2 scripts/dp_check.py: (zero-initialised biases move by ~lr*sign(g) per Adam step, so fp32 summation-order differences show up as ~5e-3 of their tiny norm)
data-parallel training with the hidden1_weights gradient summed
from all-gathered factors (dp.FactorGather) against the plain bucketed all-reduce of the dense gradient: same weights
after 5 steps (the sharded mode replays its last three steps from CUDA graphs) (up to fp32 summation order), identical on every rank."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from learnablepoolingmethods_b200.trainer import Trainer
from oracle import netvlad_oracle as O   # synthetic batches only (test infrastructure)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
B, K, Hd, V, T = 8, 64, 64, 100, 128
res = {}
# True = hidden1_weights sharded over the ranks (all-to-all of descriptor slices, shard-local clip + Adam, fp16 all-gather);
# "gather" = factors all-gathered, dense update on every rank; False = plain bucketed all-reduce of the dense gradient
for mode in (True, "gather", False):
    store = variables.VariableStore(dev, seed=11)
    eng = NetVladEngine(NetVladConfig(iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V), store)
    tr = Trainer(eng, base_learning_rate=2e-4, learning_rate_decay=0.85, batch_size=B)
    if mode is not True:
        tr.use_shard = False
    if mode is False:
        tr.gather = None
    for step in range(5):
        x, nf, lab = O.synthetic_batch(B, seed=100 + step * world + rank, vocab=V)
        loss = tr.train_step(x.to(dev), nf.to(dev), lab.to(torch.uint8).to(dev))
    assert (tr.shard is not None) == (mode is True)
    tr.sync_parameters()
    torch.cuda.synchronize()
    if mode is True:   # the fp16 GEMM operand every rank holds is the rounding of the (synchronised) fp32 master
        assert torch.equal(store.shadows["wh16"], store.vars["hidden1_weights"].half())
    res[mode] = {k: v.detach().clone() for k, v in store.vars.items()}
    res[(mode, "loss")] = float(loss)
worst, per = 0.0, []
for mode in (True, "gather"):
    for k in res[mode]:
        if not torch.is_tensor(res[mode][k]):
            continue
        a, b = res[mode][k].double(), res[False][k].double()
        e = float((a - b).norm() / b.norm().clamp_min(1e-30))
        per.append((e, f"{mode}:{k}"))
        worst = max(worst, e)
if rank == 0:
    for e, k in sorted(per, reverse=True)[:6]:
        print(f"   {k:55s} {e:.2e}")
# every rank holds the same weights
w = res[True]["hidden1_weights"].clone()
w0 = w.clone()
dist.broadcast(w0, 0)
same = bool(torch.equal(w, w0))
flag = torch.tensor([worst, 0.0 if same else 1.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"dp_check world={world}: worst rel diff (sharded | gathered) vs dense all-reduce {float(flag[0]):.2e}; ranks identical: "
          f"{float(flag[1]) == 0.0}; loss {res[(True, 'loss')]:.4f} / {res[('gather', 'loss')]:.4f} / {res[(False, 'loss')]:.4f}")
    assert float(flag[0]) < 2e-2 and float(flag[1]) == 0.0
dist.destroy_process_group()
