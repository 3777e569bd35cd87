set -x
LPM_DEBUG=1 python -m pytest tests/test_train_gpu.py -k "graph_replayed and Willow" -x -q -s 2>&1 | tail -60
python -m pytest tests/test_backward_gpu.py -k d5 -x -q -s 2>&1 | tail -30
