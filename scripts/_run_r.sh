set -x
mkdir -p gpurun_out
python scripts/prof_attn.py 20
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e 2> gpurun_out/r2r_b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['ms_per_step'], d['e2e']['ms_per_step'], d['infer_ms_per_step'], d['infer_graph_ms_per_step'])"
python __graft_entry__.py smoke 2>&1 | tail -2
