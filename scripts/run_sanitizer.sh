# compute-sanitizer on the kernels touched in the second half of round 2 (warp-collective MMA issue, tcgen05 attention)
set -x
mkdir -p gpurun_out
for w in attn_tc k3 gemm pool; do
  timeout 280 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_kernels.py $w > gpurun_out/r2ab_sanitizer_memcheck_$w.log 2>&1; tail -3 gpurun_out/r2ab_sanitizer_memcheck_$w.log
  timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_kernels.py $w > gpurun_out/r2ab_sanitizer_racecheck_$w.log 2>&1; tail -3 gpurun_out/r2ab_sanitizer_racecheck_$w.log
done
