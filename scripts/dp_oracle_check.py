"""torchrun --nproc-per-node 2 scripts/dp_oracle_check.py [out.json]
Data-parallel Trainer (the path the scaling bench runs: sharded hidden1_weights update, all-to-all of descriptor slices,
shard-local clip + Adam, fp16 all-gather, split gradient all-reduce; the last steps replayed from CUDA graphs) against
the ORACLE's multi-tower train_step: per-tower batch-norm statistics, gradients SUMMED over the towers
(utils.py:205-211), per-tensor clip (utils.py:170-189), Adam.  Context gating off, as in tests/test_train_gpu.py."""
import json, os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
from learnablepoolingmethods_b200.trainer import Trainer
from oracle import netvlad_oracle as O
from tests.helpers import oracle_params, perturb, rel
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
B, K, Hd, V, T, STEPS = 4, 64, 64, 100, 128, 4
store = variables.VariableStore(dev, seed=11)
cfg = NetVladConfig(model="NetVladV1", iterations=T, cluster_size=K, hidden_size=Hd, vocab_size=V, gating=False)
eng = NetVladEngine(cfg, store)
perturb(store, seed=5)
P, S = oracle_params(store)
P0 = {k: v.clone() for k, v in P.items()}
for p in P.values():
    p.requires_grad_(True)
tr = Trainer(eng, base_learning_rate=2e-4, learning_rate_decay=0.85, batch_size=B)
fn = lambda x, Pp, Ss: O.netvlad_v1(x[0], x[1], Pp, Ss, vocab_size=V, iterations=T, cluster_size=K, is_training=True, gating=False)
opt, rep = {}, {"world": world, "steps": STEPS, "loss": []}
for step in range(STEPS):
    towers = [O.synthetic_batch(B, seed=100 + step * world + r, vocab=V) for r in range(world)]
    x, nf, lab = towers[rank]
    loss = float(tr.train_step(x.to(dev), nf.to(dev), lab.to(torch.uint8).to(dev)))
    if rank == 0:       # the oracle plays all towers (moving statistics: every tower updates the shared variables in turn)
        ref_losses, _ = O.train_step(fn, P, S, opt, [(t[0], t[1]) for t in towers], [t[2] for t in towers], step=step + 1, lr=2e-4)
        rep["loss"].append({"step": step, "rank0": loss, "oracle_tower0": ref_losses[0], "rel": abs(loss - ref_losses[0]) / ref_losses[0]})
assert tr.shard is not None and tr.graph is not None, "the sharded, graph-replayed path must be the one under test"
tr.sync_parameters()
torch.cuda.synchronize()
w = store.vars["hidden1_weights"].clone()
w0 = w.clone()
dist.broadcast(w0, 0)
same = torch.tensor([0.0 if torch.equal(w, w0) else 1.0], device=dev)
dist.all_reduce(same, op=dist.ReduceOp.MAX)
if rank == 0:
    worst, worst_cos, per = 0.0, 1.0, []
    for name, p in P.items():
        ours = store.vars[name].detach().cpu().double()
        e = rel(ours, p.detach())
        du, dr = (ours - P0[name].double()).flatten(), (p.detach().double() - P0[name].double()).flatten()
        if float(du.norm()) == 0.0 and float(dr.norm()) == 0.0:
            continue
        cos = float((du @ dr) / (du.norm() * dr.norm()).clamp_min(1e-30))
        per.append((e, cos, name))
        worst, worst_cos = max(worst, e), min(worst_cos, cos)
    rep.update(worst_param_rel_l2=worst, worst_update_cosine=worst_cos, ranks_identical=float(same) == 0.0,
               hidden1_weights=[p for p in per if p[2] == "hidden1_weights"][0][:2], skipped_steps=tr.skipped_steps(),
               worst5=[(n, e, c) for e, c, n in sorted(per, reverse=True)[:5]])
    print(json.dumps(rep, indent=1))
    if len(sys.argv) > 1:
        json.dump(rep, open(sys.argv[1], "w"), indent=1)
    assert worst < 5e-3 and worst_cos > 0.97 and rep["ranks_identical"], rep
dist.destroy_process_group()
