"""lpm-b200: B200-native (sm_100a) NetVLAD learnable-pooling hot path of pomonam/LearnablePoolingMethods.

Python host code mirrors the reference's model/module interfaces; all arithmetic runs in hand-written
CUDA kernels behind the C-ABI of liblpm_b200.so (see include/lpm_b200.h).  No CPU/PyTorch fallback.
"""
__version__ = "0.1.0"
