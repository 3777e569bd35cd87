set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "trained-weights acceptance,|passed|failed|FAILED|Error|capture" | cut -c1-400 > gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_pytest.log
python -m pytest tests/test_backward_gpu.py -m gpu -q -s -k "v1_gradients and True" 2>&1 | grep -E "rel-L2" | awk '{print $NF, $(NF-2), $1}' | sort -k2 -g -r | head -12
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e 2> gpurun_out/r2o_b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['ms_per_step'], d['e2e']['ms_per_step'], d['infer_ms_per_step'], d['infer_graph_ms_per_step'])"
