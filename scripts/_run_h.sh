set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "trained-weights|passed|failed|FAILED|Error|capture" | cut -c1-3000 > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/dp_check.py 2>&1 | tail -12
torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 scripts/dp_oracle_check.py gpurun_out/r2h_dp_oracle_check.json 2>&1 | tail -40
torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; tail -3 gpurun_out/r2h_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_n2.json')); print({k:d[k] for k in ['ms_per_step','value','infer_ms_per_step','gpu_launches','loss','skipped_steps']}, d['e2e'])"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e > gpurun_out/r2h_bench_n1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_n1.json')); print({k:d[k] for k in ['ms_per_step','value','infer_ms_per_step','gpu_launches']}, d['e2e'])"
