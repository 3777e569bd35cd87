"""BASELINE.json config 4: NetVladV1 inference sweep -- batch 80 -> 4096, variable num_frames, K = 64/128/256.
Prints one JSON line per point (videos/s, CUDA events, inputs resident in HBM)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from learnablepoolingmethods_b200 import variables
from learnablepoolingmethods_b200.engine import InferenceGraph, NetVladConfig, NetVladEngine
dev = torch.device("cuda:0")
C = bench.CFG
batches = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [80, 320, 1280, 4096]
for K in (64, 128, 256):
    store = variables.VariableStore(dev, seed=1810)
    eng = NetVladEngine(NetVladConfig(iterations=C["iterations"], cluster_size=K, hidden_size=C["hidden_size"], vocab_size=C["vocab"]), store)
    for B in batches:
        g = torch.Generator().manual_seed(B + K)
        x = torch.randn(B, C["max_frames"], C["feat"], generator=g).to(dev)
        x = x * torch.rsqrt((x * x).sum(-1, keepdim=True).clamp_min(1e-12))
        nf = torch.randint(1, C["max_frames"] + 1, (B,), generator=g, dtype=torch.int32).to(dev)
        with torch.no_grad():
            ms = bench.time_cuda(lambda: eng.forward(x, nf, False), 5 if B >= 1280 else 20)
            ig = InferenceGraph(eng, B, C["max_frames"])       # the same forward replayed from a CUDA graph
            ig(x, nf)
            ms_g = bench.time_cuda(lambda: ig(x, nf), 5 if B >= 1280 else 20)
        print(json.dumps({"config": "infer sweep", "K": K, "batch": B, "num_frames": "U{1..300}", "ms": round(ms, 3),
                          "videos_per_s": round(B / ms * 1e3, 1), "ms_graph": round(ms_g, 3),
                          "videos_per_s_graph": round(B / ms_g * 1e3, 1)}), flush=True)
        del ig
        del x
    del eng, store
    torch.cuda.empty_cache()
