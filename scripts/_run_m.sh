set -x
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -m gpu -q 2>&1 | tail -4
python scripts/prof_attn.py 20
for fk in attn head; do
LPM_ADAM_FORK=$fk python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-registry-e2e 2> gpurun_out/r2m_b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('FORK $fk', d['ms_per_step'], d['e2e']['ms_per_step'], d['infer_ms_per_step'], d['infer_graph_ms_per_step'])"
done
python scripts/trained_parity.py 2000 1024 gpurun_out/r2m_trained_parity_2000.json tf32 2e-4 1.0 > gpurun_out/r2m_trained.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_trained_parity_2000.json'))
print({k:d[k] for k in ['vlad_video_rel_l2','hidden_rel_l2','gated_rel_l2','pred_max_abs','pred_max_abs_vs_fp64','top20_identical','top20_identical_vs_fp64','top20_identical_up_to_ties','fp32_noise','oracle_tf32']})
print(d['top20_differences'][:4])
PY
