#!/usr/bin/env python
"""Benchmark of the NetVladV1 hot path (BASELINE.json metric: NetVladV1 videos/sec, train + infer).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the CPU oracle (port of the TF reference) on host cores

A "step" is one training step (forward + backward + gradient all-reduce + clip + Adam) over one
synthetic tower batch of 80 videos per GPU (config 1: 256 of 300 frames, rgb 1024 + audio 128,
K=256/64, hidden 512, vocab 3862).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: keep a private handle to it and send everything else that writes to fd 1
# (NCCL prints its version banner there from C) to stderr
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

import torch  # noqa: E402

CFG = dict(batch=80, max_frames=300, iterations=256, cluster_size=256, hidden_size=512, vocab=3862, feat=1152)
WORKLOAD = ("NetVladV1 train step (fwd+bwd+allreduce+clip+Adam), batch 80/GPU, 256 of 300 frames, "
            "rgb1024+audio128, K=256/64, hidden 512, vocab 3862, synthetic YT8M-shaped features")


def synthetic(batch, seed, device=None, pin=False, codes=False):
    """SURVEY 8(d) synthetic batch: uint8 codes from clipped N(0,1), dequantised, L2-normalised frames
    (codes=True: the uint8 codes themselves, as the reader decodes them; the kernels dequantise + normalise)."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(batch, CFG["max_frames"], CFG["feat"], generator=g)
    q = torch.clamp(torch.round((z + 2) * 255 / 4), 0, 255)
    x = q * (4.0 / 255.0) + (4.0 / 512.0 - 2.0)
    x = x * torch.rsqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=1e-12))
    if codes:
        x = q.to(torch.uint8)
    nf = torch.full((batch,), CFG["max_frames"], dtype=torch.int32)
    labels = torch.zeros(batch, CFG["vocab"], dtype=torch.uint8)
    w = 1.0 / torch.arange(1, CFG["vocab"] + 1, dtype=torch.float64)
    npos = 1 + torch.poisson(torch.full((batch,), 2.0), generator=g).long()
    for b in range(batch):
        labels[b, torch.multinomial(w, int(npos[b]), generator=g)] = 1
    if pin:
        x, nf, labels = x.pin_memory(), nf.pin_memory(), labels.pin_memory()
    if device is not None:
        x, nf, labels = x.to(device), nf.to(device), labels.to(device)
    return x, nf, labels


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = []
        for i, n in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(len(r) > i and r[i].lower().startswith("active") for r in self.rows):
                reasons.append(n)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def ncu_traffic(kernel_key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    `ncu --set full` capture of the same call (profiles/ncu_traffic.json, written by scripts/ncu_summary.py); None when
    no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p)).get(kernel_key)
    if not d:
        return None, None
    return float(d["dram_bytes_per_launch"]), d.get("source")


def time_cuda(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cpu_reference(steps, warmup, sample_batch=None, train=True, model="NetVladV1", budget_s=240.0):
    """The CPU oracle (torch-CPU port of the TF reference) on all host cores; one step = a train step on
    `sample_batch` videos of the config-1 shape (default: the whole tower batch).  The first warm-up step is timed and
    the number of timed steps is cut (never below 3) so that the whole leg stays inside `budget_s` seconds."""
    sample_batch = sample_batch or CFG["batch"]
    from oracle import netvlad_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sp = O.param_specs(model, iterations=CFG["iterations"], cluster_size=CFG["cluster_size"],
                       hidden_size=CFG["hidden_size"], vocab_size=CFG["vocab"])
    P, S = O.init_params(sp)
    x, nf, labels = synthetic(sample_batch, 20181000)
    opt = {}
    kw = dict(vocab_size=CFG["vocab"], iterations=CFG["iterations"], cluster_size=CFG["cluster_size"], is_training=True)
    extra = None
    if model == "WillowModelReg":
        import numpy as np
        rng = np.random.RandomState(0)
        fn = lambda xx, Pp, Ss: O.willow_model_reg(xx[0], xx[1], Pp, Ss, frame_index=O.sample_random_frame_indices(
            xx[1].numpy(), rng.rand(sample_batch, CFG["iterations"]).astype(np.float32)), **kw)
        extra = O.willow_regularization
    else:
        fn = lambda xx, Pp, Ss: getattr(O, {"NetVladV1": "netvlad_v1", "NetVladV2": "netvlad_v2"}[model])(xx[0], xx[1], Pp, Ss, **kw)
    times = []
    warmup = max(1, warmup)
    i = 0
    while i < warmup + steps:
        t0 = time.perf_counter()
        O.train_step(fn, P, S, opt, [(x, nf)], [labels.bool()], step=i + 1, lr=2e-4, extra_reg=extra)
        dt = time.perf_counter() - t0
        if i == 0:
            fit = int(budget_s / max(dt, 1e-3))
            if warmup + steps > fit:                              # bounded: drop warm-up first, then timed steps
                warmup = 1
                steps = max(3, min(steps, fit - 1))
        if i >= warmup:
            times.append(dt)
        i += 1
    sec = sorted(times)[len(times) // 2]                       # median step
    return (sample_batch / sec, sec, torch.get_num_threads(),
            f"train step on {sample_batch} of {CFG['batch']} videos, {len(times)} timed steps (median), {warmup} warm-up", len(times), warmup)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--cluster-size", type=int, default=CFG["cluster_size"], help="netvlad_cluster_size (512 = wide config)")
    ap.add_argument("--hidden-size", type=int, default=CFG["hidden_size"], help="netvlad_hidden_size (1024 = wide config)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=CFG["batch"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-registry-e2e", action="store_true", help="skip the create_model + CrossEntropyLoss + backward() timing")
    ap.add_argument("--cpu-sample-batch", type=int, default=0, help="videos per CPU-oracle step (0 = the whole tower batch)")
    ap.add_argument("--model", default="NetVladV1", choices=["NetVladV1", "NetVladV2", "WillowModelReg"])
    ap.add_argument("--input", default="u8", choices=["u8", "f32"],
                    help="u8: model_input = the reader's uint8 codes (dequantise + L2-normalise fused into the gather kernels); "
                         "f32: model_input = dequantised, L2-normalised fp32 frames (the reference's create_model contract)")
    args = ap.parse_args()
    CFG["cluster_size"], CFG["hidden_size"] = args.cluster_size, args.hidden_size
    global WORKLOAD
    WORKLOAD = WORKLOAD.replace("K=256/64, hidden 512", f"K={args.cluster_size}/{args.cluster_size // 4}, hidden {args.hidden_size}")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    base = {"metric": f"{args.model} videos/sec (train step; infer reported alongside)", "unit": "videos/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOAD.replace("NetVladV1", args.model), "per_gpu_batch": args.batch, "global_batch": args.batch * args.gpus,
                       "parallelism": f"dp{args.gpus}",
                       "model_input": "uint8 codes [B,300,1152] (readers.py:185-193); dequantise + l2_normalize fused" if args.input == "u8"
                       else "fp32 frames [B,300,1152], L2-normalised by the caller", "l2": "working set per step (>=3 GB of weights, moments and "
                       "activations) exceeds the 126 MB L2; no flush needed"}}

    if args.impl == "reference":
        if rank != 0:
            return
        CFG["batch"] = args.batch
        v, sec, cores, sample, steps, wu = cpu_reference(args.steps, args.warmup, sample_batch=args.cpu_sample_batch or None,
                                                        model=args.model)
        base["config"]["cpu_sample_batch"] = args.cpu_sample_batch or args.batch
        base["config"]["model_input"] = "fp32 frames [B,300,1152], dequantised and L2-normalised (the same synthetic videos)"
        base["config"]["parallelism"] = f"host cores x{cores} (rank 0 only)"
        out = dict(base, impl="reference", value=v, ms_per_step=sec * 1e3, steps=steps, warmup=wu,
                   dtype="f32", n_gpus=args.gpus,
                   cpu_baseline={"value": v, "unit": "videos/s", "cores": cores, "kind": "port", "sample": sample},
                   e2e={"value": v, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   gpu_launches=0)
        print(json.dumps(out), file=_JSON_OUT, flush=True)
        return

    import torch.distributed as dist
    from learnablepoolingmethods_b200 import _lib, ops, variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    from learnablepoolingmethods_b200.trainer import Trainer
    _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # one rank per GPU on one host: give every rank its own slice of the host cores (the launch thread of one rank must
        # not be descheduled behind another rank's) -- the topology reports one NUMA node for all eight GPUs
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]) or set(cores))
        except (AttributeError, OSError):
            pass
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    store = variables.VariableStore(dev, seed=1810)
    cfg = NetVladConfig(model=args.model, iterations=CFG["iterations"], cluster_size=CFG["cluster_size"],
                        hidden_size=CFG["hidden_size"], vocab_size=CFG["vocab"])
    eng = NetVladEngine(cfg, store)
    tr = Trainer(eng, base_learning_rate=2e-4, learning_rate_decay=0.85, batch_size=B)
    xh, nfh, lh = synthetic(B, 20181000 + rank, pin=True, codes=args.input == "u8")
    x, nf, lab = xh.to(dev), nfh.to(dev), lh.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---------------- training: device-resident inputs ----------------
    for _ in range(max(args.warmup, 3)):
        tr.train_step(x, nf, lab)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = tr.train_step(x, nf, lab)
    e1.record()
    barrier()
    train_ms = reduce_max(e0.elapsed_time(e1)) / args.steps
    loss = loss.clone()          # the graph-replayed step returns a static tensor that later steps overwrite
    launches = (_lib.launch_count - l0) // args.steps
    overflow = tr.overflowed()

    # ---------------- training end to end: pinned host inputs, H2D each step, loss read back ----------------
    # One packed pinned slab per step (frames | labels | num_frames) -> ONE cudaMemcpyAsync on a copy stream, double
    # buffered on the device; the model reads typed views of the device slab.
    copy_stream = torch.cuda.Stream()
    nx, nl = xh.numel() * xh.element_size(), lh.numel()
    o_nf = (nx + nl + 15) // 16 * 16                        # num_frames starts 16-byte aligned
    slab_h = torch.zeros(o_nf + nfh.numel() * 4, dtype=torch.uint8).pin_memory()
    slab_h[:nx].copy_(xh.view(-1).view(torch.uint8))
    slab_h[nx:nx + nl].copy_(lh.view(-1))
    slab_h[o_nf:].copy_(nfh.view(-1).view(torch.uint8))
    slabs = [torch.empty_like(slab_h, device=dev) for _ in range(2)]
    bufs = [(sl[:nx].view(xh.dtype).view(xh.shape), sl[o_nf:].view(torch.int32), sl[nx:nx + nl].view(lh.shape)) for sl in slabs]
    ready = [torch.cuda.Event() for _ in range(2)]

    def stage(i):
        with torch.cuda.stream(copy_stream):
            slabs[i % 2].copy_(slab_h, non_blocking=True)
            ready[i % 2].record(copy_stream)

    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n, step_fn=None):
        step_fn = step_fn or tr.train_step
        stage(0)
        for i in range(n):
            torch.cuda.current_stream().wait_event(ready[i % 2])
            if i + 1 < n:
                if i >= 1:
                    copy_stream.wait_event(consumed[(i + 1) % 2])
                stage(i + 1)
            bx, bn, bl = bufs[i % 2]
            ls = step_fn(bx, bn, bl)
            consumed[i % 2].record()
            loss_host.copy_(ls, non_blocking=True)
        torch.cuda.synchronize()

    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = reduce_max(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.stop() if rank == 0 else None      # sampled across both timed regions (device-resident and end-to-end)
    h2d = slab_h.numel()

    # ---------------- the same, through the reference-facing registry entry (the drop-in boundary) ----------------
    # find_class_by_name -> create_model(model_input, vocab_size, num_frames, ..., is_training=True)["predictions"]
    # -> CrossEntropyLoss().calculate_loss -> loss.backward() (autograd.NetVladFunction: the hand-written CUDA backward)
    # -> Trainer.apply_gradients (clip_gradient_norms + Adam, train.py:321-336).  Dense path: the 554 MB gradient of
    # hidden1_weights is materialised, so this is slower than Trainer.train_step; it is what a caller of the reference's
    # own loop structure gets.
    reg_ms, reg_err = None, None
    if world == 1 and not args.no_registry_e2e:
        try:
            from learnablepoolingmethods_b200 import frame_level_models, losses
            from learnablepoolingmethods_b200.utils import find_class_by_name
            store2 = variables.VariableStore(dev, seed=1810)
            model = find_class_by_name(args.model, [frame_level_models])()
            loss_fn = losses.CrossEntropyLoss()
            kw = dict(vocab_size=CFG["vocab"], iterations=CFG["iterations"], cluster_size=CFG["cluster_size"],
                      hidden_size=CFG["hidden_size"], is_training=True, store=store2)
            reg = {"tr": None}

            def registry_step(bx, bn, bl):
                res = model.create_model(bx, num_frames=bn, **kw)
                ls = loss_fn.calculate_loss(res["predictions"], bl)
                ls.backward()
                if reg["tr"] is None:
                    reg["tr"] = Trainer(next(iter(store2._engines.values())), base_learning_rate=2e-4, learning_rate_decay=0.85,
                                        batch_size=B)
                reg["tr"].apply_gradients()
                return ls.detach()

            e2e_loop(3, registry_step)
            barrier()
            e0.record()
            e2e_loop(args.steps, registry_step)
            e1.record()
            barrier()
            reg_ms = e0.elapsed_time(e1) / args.steps
            del store2, reg
            torch.cuda.empty_cache()
        except Exception as e:                                     # reported, never fatal for the headline numbers
            reg_err = repr(e)
            print(f"registry end-to-end path failed: {e!r}", file=sys.stderr)

    # ---------------- inference: forward only (is_training=False) ----------------
    with torch.no_grad():
        infer_ms = reduce_max(time_cuda(lambda: eng.forward(x, nf, False), max(5, args.steps)))

    # the same forward replayed from a CUDA graph (one launch instead of ~28 from Python)
    graph_ms = None
    try:
        from learnablepoolingmethods_b200.engine import InferenceGraph
        ig = InferenceGraph(eng, B, CFG["max_frames"], input_dtype=x.dtype)
        ig(x, nf)
        graph_ms = reduce_max(time_cuda(lambda: ig(x, nf), max(5, args.steps)))
    except Exception as e:                                     # reported, never fatal for the headline numbers
        print(f"inference graph unavailable: {e!r}", file=sys.stderr)

    out = None
    if rank == 0:
        burst, sustained, hbm, src = peaks()
        # dominant kernel: the tcgen05 GEMM (attention-block linears are 92% of the FLOPs); its largest call
        M, N, K = B * CFG["cluster_size"], 4 * 1024, 1024
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(K, N, device=dev) * 0.03).half()
        o = torch.empty(M, N, dtype=torch.float16, device=dev)
        g_ms = time_cuda(lambda: ops.gemm(a, w, out=o), 20)
        g_tf = 2.0 * M * N * K / g_ms / 1e9
        g_traffic, g_traffic_src = ncu_traffic("gemm_f16_ffn_20480x4096x1024")
        # yardstick only (never on the product path): cuBLAS through torch.matmul at the SAME shape, same timing loop -- the
        # burst peak in MEASURED_PEAKS.json is cuBLAS at a much larger product
        lib_tf = None
        try:
            lib_ms = time_cuda(lambda: torch.matmul(a, w, out=o), 20)
            lib_tf = 2.0 * M * N * K / lib_ms / 1e9
        except Exception as e:
            print(f"cuBLAS yardstick unavailable: {e!r}", file=sys.stderr)
        # north-star kernel: fused NetVLAD pooling (rgb), 4*T*D*K FLOPs per video
        T, D, Kc = CFG["iterations"], 1024, 256
        xb = torch.randn(B * T, D, device=dev).half()
        wc = (torch.randn(D, Kc, device=dev) / 32).half()
        ct = ops.transpose_f32_dual(torch.randn(D, Kc, device=dev) / 32, want32=False)[1]   # fp16 [K, D] shadow
        one, zero = torch.ones(Kc, device=dev), torch.zeros(Kc, device=dev)
        p_ms = time_cuda(lambda: ops.netvlad_pool_fwd(xb, B, T, wc, one, zero, ct), 20)
        p_tf = 4.0 * T * D * Kc * B / p_ms / 1e9
        Bl = 148 * 8
        xbl = torch.randn(Bl * T, D, device=dev).half()
        pl_ms = time_cuda(lambda: ops.netvlad_pool_fwd(xbl, Bl, T, wc, one, zero, ct), 5)
        pl_tf = 4.0 * T * D * Kc * Bl / pl_ms / 1e9
        del xbl
        out = dict(base, value=B * world / (train_ms / 1e3), ms_per_step=train_ms,
                   infer_value=B * world / (infer_ms / 1e3), infer_ms_per_step=infer_ms,
                   infer_graph_ms_per_step=graph_ms,
                   e2e={"value": B * world / (e2e_ms / 1e3), "unit": "videos/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms},
                   e2e_registry=({"value": B / (reg_ms / 1e3), "unit": "videos/s", "ms_per_step": reg_ms,
                                  "path": "create_model + CrossEntropyLoss + loss.backward() + Trainer.apply_gradients (dense gradients)",
                                  "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4} if reg_ms else {"unavailable": reg_err}),
                   gpu_launches=int(launches), clocks=clocks, loss=float(loss), loss_scale_overflow=bool(overflow),
                   skipped_steps=int(tr.skipped_steps()),
                   roofline={"kernel": "gemm_f16_kernel<256,6,0,1,TWO=1> (tcgen05.mma.cta_group::2, FFN 20480x4096x1024 call)", "bound": "tensor",
                             "achieved": g_tf, "peak": burst, "unit": "TFLOP/s", "frac": g_tf / burst,
                             "frac_of_sustained_peak": g_tf / sustained,
                             "traffic": g_traffic, "traffic_source": g_traffic_src,
                             "algorithmic_bytes": 2.0 * (M * K + K * N + M * N),
                             "cublas_same_shape": lib_tf, "frac_of_cublas_same_shape": (g_tf / lib_tf) if lib_tf else None,
                             "peak_source": f"{src} bf16 burst (the kernel is timed alone: 20 back-to-back launches, ~3 ms)"},
                   roofline_pool={"kernel": "netvlad_pool_fwd_kernel<256> (rgb, fused)", "bound": "tensor",
                                  "achieved": p_tf, "peak": burst, "unit": "TFLOP/s", "frac": p_tf / burst,
                                  "frac_of_occupied_sms": p_tf / burst * 148.0 / min(148, B),
                                  "achieved_full_waves": pl_tf, "frac_full_waves": pl_tf / burst,
                                  "note": "B=80 fills 80 of 148 SMs (one CTA per video); full-wave figure at B=1184",
                                  "peak_source": f"{src} bf16 burst (kernel timed alone)"})
        if args.gpus == 1 and not args.no_cpu_baseline:
            v, sec, cores, sample, _, _ = cpu_reference(3, 1, sample_batch=args.cpu_sample_batch or None, model=args.model, budget_s=40.0)
            out["cpu_baseline"] = {"value": v, "unit": "videos/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(out), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
