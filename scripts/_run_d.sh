set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2d_pytest.log; tail -12 gpurun_out/r2d_pytest.log
python scripts/prof_attn.py 20 > gpurun_out/r2d_attn.log 2>&1; cat gpurun_out/r2d_attn.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench.err; tail -3 gpurun_out/r2d_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_n1.json')); print({k:d[k] for k in ['ms_per_step','infer_ms_per_step','infer_graph_ms_per_step','gpu_launches','loss','skipped_steps']}, d['e2e'], d['e2e_registry'])"
python scripts/trained_parity.py 3000 1024 gpurun_out/r2d_trained_parity.json tf32,fp16-body 2e-4 1.0 > gpurun_out/r2d_trained.log 2>&1; grep -v "^  " gpurun_out/r2d_trained.log | tail -45
