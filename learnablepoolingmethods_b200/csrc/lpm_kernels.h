// Internal (C++) interfaces between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lpm_b200.h"

namespace lpm {

using GemmArgs = lpm_gemm_desc;

int gemm_f16(const GemmArgs& g, cudaStream_t st);
int gemm_pick_bn(int N);
int gemm_effective_splits(int K, int splits);

}  // namespace lpm
