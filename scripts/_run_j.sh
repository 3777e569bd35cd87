set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "trained-weights|passed|failed|FAILED|Error|capture" | cut -c1-1800 > gpurun_out/r2j_pytest.log; cat gpurun_out/r2j_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench.err; tail -2 gpurun_out/r2j_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_n1.json')); print({k:d[k] for k in ['ms_per_step','infer_ms_per_step','infer_graph_ms_per_step','gpu_launches','loss','skipped_steps','cpu_baseline']}, d['e2e'], d['e2e_registry'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2j_launches_train_step_b80.csv python scripts/step_once.py 4 > gpurun_out/r2j_ncu_launch.log 2>&1; tail -2 gpurun_out/r2j_ncu_launch.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mha_fwd|mha_bwd" -c 4 -o gpurun_out/r2j_mha python scripts/prof_attn.py 1 > gpurun_out/r2j_ncu_mha.log 2>&1; tail -1 gpurun_out/r2j_ncu_mha.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"netvlad_pool_fwd|gemm_f16" -c 9 -o gpurun_out/r2j_pool_gemm python scripts/prof_pool.py > gpurun_out/r2j_ncu_pool.log 2>&1; tail -1 gpurun_out/r2j_ncu_pool.log
for tool in memcheck racecheck; do for w in pool gemm misc; do
timeout 300 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_kernels.py $w > gpurun_out/r2j_sanitizer_${tool}_$w.log 2>&1; echo "$tool $w rc=$?"; tail -4 gpurun_out/r2j_sanitizer_${tool}_$w.log
done; done
