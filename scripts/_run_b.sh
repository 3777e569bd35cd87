set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2b_pytest.log; tail -6 gpurun_out/r2b_pytest.log
for sp in 2 4; do
LPM_OPT_PRIORITY=-1 LPM_ADAM_SPLIT=$sp python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e 2> gpurun_out/r2b_b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('SAMEPRIO split $sp', d['ms_per_step'], d['e2e']['ms_per_step'])"
LPM_ADAM_SPLIT=$sp python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e 2> gpurun_out/r2b_b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('LOWPRIO split $sp', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
python scripts/trained_parity.py 3000 1024 gpurun_out/r2b_trained_parity_lr5e-4_s2.json tf32,fp16,fp16-body 5e-4 2.0 > gpurun_out/r2b_trained1.log 2>&1; grep -v "^  " gpurun_out/r2b_trained1.log | tail -60
python scripts/trained_parity.py 3000 1024 gpurun_out/r2b_trained_parity_lr2e-4_s1.json tf32,fp16,fp16-body 2e-4 1.0 > gpurun_out/r2b_trained2.log 2>&1; grep -v "^  " gpurun_out/r2b_trained2.log | tail -60
