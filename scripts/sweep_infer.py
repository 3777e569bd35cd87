"""BASELINE.json config 4: NetVladV1 inference sweep -- batch 80 -> 4096 per GPU, variable num_frames, K = 64/128/256.
One JSON line per point (CUDA events, inputs resident in HBM).  Under torchrun every rank is an independent replica on
its own GPU (inference has no exchange step): the line reports the aggregate rate, time = max over ranks.
usage: [torchrun --nproc-per-node N] scripts/sweep_infer.py [batches] [Ks]"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import InferenceGraph, NetVladConfig, NetVladEngine
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def agg(ms):
    if world == 1:
        return ms
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


C = bench.CFG
batches = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [80, 320, 1280, 4096]
Ks = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 128, 256]
for K in Ks:
    store = variables.VariableStore(dev, seed=1810)
    eng = NetVladEngine(NetVladConfig(iterations=C["iterations"], cluster_size=K, hidden_size=C["hidden_size"], vocab_size=C["vocab"]), store)
    for B in batches:
        g = torch.Generator().manual_seed(B + K + rank)
        x = torch.randn(B, C["max_frames"], C["feat"], generator=g).to(dev)
        x = x * torch.rsqrt((x * x).sum(-1, keepdim=True).clamp_min(1e-12))
        nf = torch.randint(1, C["max_frames"] + 1, (B,), generator=g, dtype=torch.int32).to(dev)
        it = 5 if B >= 1280 else 20
        with torch.no_grad():
            if world > 1:
                dist.barrier()
            ms = agg(bench.time_cuda(lambda: eng.forward(x, nf, False), it))
            ig = InferenceGraph(eng, B, C["max_frames"])       # the same forward replayed from a CUDA graph
            ig.x.copy_(x)                                      # a serving loop fills the graph's static input directly
            ig(ig.x, nf)
            if world > 1:
                dist.barrier()
            ms_g = agg(bench.time_cuda(lambda: ig(ig.x, nf), it))
        if rank == 0:
            print(json.dumps({"config": "infer sweep", "n_gpus": world, "K": K, "batch_per_gpu": B, "num_frames": "U{1..300}",
                              "ms": round(ms, 3), "videos_per_s": round(B * world / ms * 1e3, 1), "ms_graph": round(ms_g, 3),
                              "videos_per_s_graph": round(B * world / ms_g * 1e3, 1)}), file=bench._JSON_OUT, flush=True)
        del ig, x
    del eng, store
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
