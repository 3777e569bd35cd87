# Multi-GPU bench lines on one 8 x B200 box (sequential runs).  usage: run_8gpu_lines.sh <tag> [all]
set -x
TAG=${1:-r2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29818 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err
$TR --nproc-per-node 4 --master-port 29814 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_n4.json 2> gpurun_out/${TAG}_bench_n4.err
if [ "$2" = "all" ]; then
$TR --nproc-per-node 8 --master-port 29821 bench.py --gpus 8 --model NetVladV2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_v2_n8.json 2> gpurun_out/${TAG}_bench_v2_n8.err
$TR --nproc-per-node 8 --master-port 29831 bench.py --gpus 8 --cluster-size 512 --hidden-size 1024 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_wide_n8.json 2> gpurun_out/${TAG}_bench_wide_n8.err
fi
for f in gpurun_out/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], d["n_gpus"], round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"], 3), "infer", round(d["infer_ms_per_step"], 3), "skipped", d.get("skipped_steps"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
