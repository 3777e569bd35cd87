set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
$TR --master-port 29911 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2ah_bench_n2.json 2> gpurun_out/r2ah_bench_n2.err
$TR --master-port 29912 bench.py --gpus 2 --model NetVladV2 --steps 20 --warmup 5 > gpurun_out/r2ah_bench_v2_n2.json 2> gpurun_out/r2ah_bench_v2_n2.err
$TR --master-port 29913 scripts/dp_oracle_check.py gpurun_out/r2ah_dp_oracle_check_n2.json > gpurun_out/r2ah_dp_oracle.log 2>&1; tail -2 gpurun_out/r2ah_dp_oracle.log
for f in gpurun_out/r2ah_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], d["n_gpus"], round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"], 3), "skipped", d.get("skipped_steps"), "loss", d.get("loss"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2ah_dp_oracle_check_n2.json"))
print({k: d[k] for k in ("worst_param_rel_l2", "worst_update_cosine", "ranks_identical")})
PY
