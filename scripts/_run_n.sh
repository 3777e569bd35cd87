set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "trained-weights|passed|failed|FAILED|Error|capture" | cut -c1-1200 > gpurun_out/r2n_pytest.log; cat gpurun_out/r2n_pytest.log
for st in 1 0; do
LPM_LN_STAGE=$st python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e 2> gpurun_out/r2n_b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('LN_STAGE $st', d['ms_per_step'], d['e2e']['ms_per_step'], d['infer_ms_per_step'], d['infer_graph_ms_per_step'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2n_launches_train_step_b80.csv python scripts/step_once.py 4 > gpurun_out/r2n_ncu_launch.log 2>&1; tail -1 gpurun_out/r2n_ncu_launch.log
