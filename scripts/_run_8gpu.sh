# BASELINE.json configs 3, 4, 5 on one 8 x B200 box (sequential runs; every bench line is kept under gpurun_out/).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
# config 5: wide NetVladV1 (K=512/128, hidden 1024), batch 80/GPU: 1 and 8 GPUs (global batch 640)
python bench.py --cluster-size 512 --hidden-size 1024 --steps 20 --warmup 5 --no-cpu-baseline --no-registry-e2e > gpurun_out/r2_bench_wide_n1.json 2> gpurun_out/r2_bench_wide_n1.err
$TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --cluster-size 512 --hidden-size 1024 --steps 20 --warmup 5 > gpurun_out/r2_bench_wide_n8.json 2> gpurun_out/r2_bench_wide_n8.err
# config 3: NetVladV2 data parallel at 2 / 4 / 8 (+ the single-GPU line for the efficiency)
python bench.py --model NetVladV2 --steps 20 --warmup 5 --no-cpu-baseline --no-registry-e2e > gpurun_out/r2_bench_v2_n1.json 2> gpurun_out/r2_bench_v2_n1.err
for n in 2 4 8; do
$TR --nproc-per-node $n --master-port 2971$n bench.py --gpus $n --model NetVladV2 --steps 20 --warmup 5 > gpurun_out/r2_bench_v2_n$n.json 2> gpurun_out/r2_bench_v2_n$n.err
done
# headline config at 1 / 8 (the driver runs 1/2/4/8 itself at round end; these are the builder's own lines)
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-registry-e2e > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
$TR --nproc-per-node 8 --master-port 29731 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
# config 4: inference sweep on 8 GPUs (replicas) and on 1
$TR --nproc-per-node 8 --master-port 29741 scripts/sweep_infer.py 80,320,1280,4096 64,128,256 > gpurun_out/r2_infer_sweep_n8.jsonl 2> gpurun_out/r2_infer_sweep_n8.err
python scripts/sweep_infer.py 80,320,1280,4096 64,128,256 > gpurun_out/r2_infer_sweep_n1.jsonl 2> gpurun_out/r2_infer_sweep_n1.err
for f in gpurun_out/r2_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], d["n_gpus"], round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"], 3), "infer", round(d["infer_ms_per_step"], 3), "skipped", d.get("skipped_steps"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -3 gpurun_out/r2_infer_sweep_n8.jsonl
tail -2 gpurun_out/*.err | tail -40
