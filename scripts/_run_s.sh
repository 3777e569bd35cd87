set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_willow_gpu.py tests/test_edge_cases_gpu.py -m gpu -q 2>&1 | tail -4
python scripts/pool_timeline.py 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e 2> gpurun_out/r2s_b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['ms_per_step'], d['e2e']['ms_per_step'], d['infer_ms_per_step'], d['infer_graph_ms_per_step']); print(d['roofline_pool'])"
