set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.log; tail -5 gpurun_out/r2a_pytest.log
python scripts/prof_attn.py 20 > gpurun_out/r2a_attn.log 2>&1; cat gpurun_out/r2a_attn.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench.err; tail -3 gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench_n1.json
LPM_FUSE_OPT=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e > gpurun_out/r2a_bench_nofuse_n1.json 2> gpurun_out/r2a_bench_nofuse.err; cat gpurun_out/r2a_bench_nofuse_n1.json
LPM_OPT_PRIORITY=-1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e > gpurun_out/r2a_bench_sameprio_n1.json 2> gpurun_out/r2a_bench_sameprio.err; cat gpurun_out/r2a_bench_sameprio_n1.json
python scripts/trained_parity.py 1500 1024 gpurun_out/r2a_trained_parity.json tf32,fp16 > gpurun_out/r2a_trained.log 2>&1; tail -80 gpurun_out/r2a_trained.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mha_ -c 4 -o gpurun_out/r2a_mha python scripts/prof_attn.py 1 > gpurun_out/r2a_ncu_mha.log 2>&1; tail -3 gpurun_out/r2a_ncu_mha.log
