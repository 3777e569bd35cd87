set -x
mkdir -p gpurun_out
for st in 300 1000 2000; do
python scripts/trained_parity.py $st 1024 gpurun_out/r2k_trained_parity_$st.json none 2e-4 1.0 > gpurun_out/r2k_trained_$st.log 2>&1
python - <<PY
import json
d=json.load(open('gpurun_out/r2k_trained_parity_$st.json'))
print($st, {k:d[k] for k in ['vlad_video_rel_l2','pred_max_abs','pred_median_abs','pred_p999_abs','top20_identical','top20_identical_up_to_ties','gap_oracle','hit1_oracle','rank20_score_median','rank20_21_gap_median','gated_rel_l2']}, d['top20_differences'][:3], d['losses'][-1])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2k_launches_train_step_b80.csv python scripts/step_once.py 4 > gpurun_out/r2k_ncu_launch.log 2>&1; tail -2 gpurun_out/r2k_ncu_launch.log
python scripts/prof_attn.py 20
python -m pytest tests/test_kernels_gpu.py -q -k "mha" 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -3
