set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
$TR --master-port 29921 scripts/dp_check.py 2>&1 | tail -3
$TR --master-port 29922 scripts/dp_oracle_check.py gpurun_out/r2v_dp_oracle_check_n2.json 2>&1 | grep -E "worst|ranks|Error|error" | head
$TR --master-port 29923 bench.py --gpus 2 --steps 30 --warmup 5 2> gpurun_out/r2v_bench_n2.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N2 three-stage', d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])"
LPM_DP_THREE_STAGE=0 $TR --master-port 29924 bench.py --gpus 2 --steps 30 --warmup 5 2> gpurun_out/r2v_bench_n2b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N2 two-stage', d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])"
tail -3 gpurun_out/r2v_bench_n2.err
