set -x
mkdir -p gpurun_out
LPM_POOL_SPLIT=1 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -k "pool or parity or edge" 2>&1 | tail -4
LPM_POOL_SPLIT=1 python scripts/pool_timeline.py 2>&1 | tail -3
python scripts/pool_timeline.py 2>&1 | tail -3
LPM_POOL_SPLIT=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e 2> gpurun_out/r2t_b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH SPLIT', d['ms_per_step'], d['e2e']['ms_per_step'], d['infer_ms_per_step'], d['infer_graph_ms_per_step']); print(d['roofline_pool'])"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e 2> gpurun_out/r2t_b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH BASE', d['ms_per_step'], d['e2e']['ms_per_step'], d['infer_ms_per_step'], d['infer_graph_ms_per_step']); print(d['roofline_pool'])"
