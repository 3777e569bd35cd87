set -x
mkdir -p gpurun_out
LPM_DEBUG=1 python -m pytest tests/test_train_gpu.py -x -q -s 2>&1 | tail -80 > gpurun_out/r2f_train_tests.log; grep -n "Traceback" -A40 gpurun_out/r2f_train_tests.log | head -80; tail -5 gpurun_out/r2f_train_tests.log
python -m pytest tests/test_backward_gpu.py -k d5 -x -q 2>&1 | tail -3
