set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; tail -3 gpurun_out/r2q_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench_n2.json')); print({k:d[k] for k in ['ms_per_step','value','infer_ms_per_step','gpu_launches','loss','skipped_steps']}, d['e2e'])"
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29912 scripts/dp_check.py 2>&1 | tail -3
python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "eval_metrics" 2>&1 | tail -2
