set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mha_fwd|mha_bwd" -c 4 -o gpurun_out/r2u_mha python scripts/prof_attn.py 1 > gpurun_out/r2u_ncu_mha.log 2>&1; tail -1 gpurun_out/r2u_ncu_mha.log
