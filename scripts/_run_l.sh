set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "trained-weights|passed|failed|FAILED|Error|capture" | cut -c1-1500 > gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_pytest.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench.err; tail -2 gpurun_out/r2l_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_n1.json')); print({k:d[k] for k in ['ms_per_step','infer_ms_per_step','infer_graph_ms_per_step','gpu_launches','loss','skipped_steps']}, d['e2e'], d['e2e_registry'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2l_launches_train_step_b80.csv python scripts/step_once.py 4 > gpurun_out/r2l_ncu_launch.log 2>&1; tail -1 gpurun_out/r2l_ncu_launch.log
