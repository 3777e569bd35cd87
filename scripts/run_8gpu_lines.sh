# Final multi-GPU lines of round 2 (one 8 x B200 box, sequential runs).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err
for n in 2 4 8; do
$TR --nproc-per-node $n --master-port 2981$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/r2p_bench_n$n.json 2> gpurun_out/r2p_bench_n$n.err
done
$TR --nproc-per-node 8 --master-port 29821 bench.py --gpus 8 --model NetVladV2 --steps 20 --warmup 5 > gpurun_out/r2p_bench_v2_n8.json 2> gpurun_out/r2p_bench_v2_n8.err
python bench.py --model NetVladV2 --steps 20 --warmup 5 --no-cpu-baseline --no-registry-e2e > gpurun_out/r2p_bench_v2_n1.json 2> gpurun_out/r2p_bench_v2_n1.err
$TR --nproc-per-node 8 --master-port 29831 bench.py --gpus 8 --cluster-size 512 --hidden-size 1024 --steps 20 --warmup 5 > gpurun_out/r2p_bench_wide_n8.json 2> gpurun_out/r2p_bench_wide_n8.err
$TR --nproc-per-node 2 --master-port 29841 scripts/dp_oracle_check.py gpurun_out/r2p_dp_oracle_check_n2.json > gpurun_out/r2p_dp_oracle.log 2>&1; tail -3 gpurun_out/r2p_dp_oracle.log
for f in gpurun_out/r2p_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], d["n_gpus"], round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"], 3), "infer", round(d["infer_ms_per_step"], 3), "skipped", d.get("skipped_steps"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
