// Optimiser step of the reference trainer on one flat fp32 buffer (train.py:321-336, utils.py:170-189):
//   g_t   = (grad + wd_t * p)                       wd_t: slim.l2_regularizer on the MoE weights
//   g_t  *= clip / max(||g_t||, clip)               per-tensor tf.clip_by_norm
//   Adam  (TF form: lr_t = lr * sqrt(1-b2^t)/(1-b1^t);  p -= lr_t * m / (sqrt(v) + eps))
// Multi-tensor: a chunk table maps fixed-size chunks of the flat buffer to tensors, so the whole step
// is three launches regardless of the number of variables.  A non-finite gradient norm (fp16
// loss-scale overflow) sets a device flag and the update is skipped.
#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

// chunk table entry: {tensor id, start offset (elements), length}
__global__ void __launch_bounds__(256) mt_sqnorm_kernel(const float* __restrict__ g, const float* __restrict__ p,
                                                        const int* __restrict__ table, const float* __restrict__ wd,
                                                        float* __restrict__ partial) {
  __shared__ float red[8];
  const int t = table[blockIdx.x * 3], len = table[blockIdx.x * 3 + 2];
  const long long start = (long long)(unsigned)table[blockIdx.x * 3 + 1] * 32;
  const float w = wd[t];
  float s = 0.f;
  for (int i = threadIdx.x; i < len; i += 256) {
    const float x = g[start + i] + w * p[start + i];
    s += x * x;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int i = 0; i < 8; ++i) r += red[i];
    partial[blockIdx.x] = r;
  }
}

__global__ void mt_clip_kernel(const float* __restrict__ partial, const int* __restrict__ chunk_begin, int n_tensors,
                               float clip, float* __restrict__ factor, float* __restrict__ norms, int* __restrict__ flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tensors) return;
  double s = 0.0;
  for (int c = chunk_begin[t]; c < chunk_begin[t + 1]; ++c) s += partial[c];
  const float n = (float)sqrt(s);
  norms[t] = n;
  if (!isfinite(n)) atomicExch(flag, 1);
  factor[t] = clip > 0.f ? clip / fmaxf(n, clip) : 1.f;
}

__global__ void __launch_bounds__(256) mt_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ m, float* __restrict__ v,
                                                      const int* __restrict__ table, const float* __restrict__ wd,
                                                      const float* __restrict__ factor, const int* __restrict__ flag,
                                                      float lr_t, float b1, float b2, float eps) {
  if (*flag) return;  // overflow: skip the step
  const int t = table[blockIdx.x * 3], len = table[blockIdx.x * 3 + 2];
  const long long start = (long long)(unsigned)table[blockIdx.x * 3 + 1] * 32;
  const float w = wd[t], f = factor[t];
  for (int i = threadIdx.x; i < len; i += 256) {
    const long long j = start + i;
    const float pj = p[j];
    const float gj = (g[j] + w * pj) * f;
    const float mj = b1 * m[j] + (1.f - b1) * gj;
    const float vj = b2 * v[j] + (1.f - b2) * gj * gj;
    m[j] = mj;
    v[j] = vj;
    p[j] = pj - lr_t * mj / (sqrtf(vj) + eps);
  }
}

int adam_clip_step(float* p, const float* g, float* m, float* v, const int* table, int n_chunks,
                   const int* chunk_begin, int n_tensors, const float* wd, float clip, float lr_t, float b1,
                   float b2, float eps, float* partial, float* factor, float* norms, int* flag, cudaStream_t st) {
  mt_sqnorm_kernel<<<n_chunks, 256, 0, st>>>(g, p, table, wd, partial);
  mt_clip_kernel<<<(n_tensors + 127) / 128, 128, 0, st>>>(partial, chunk_begin, n_tensors, clip, factor, norms, flag);
  mt_adam_kernel<<<n_chunks, 256, 0, st>>>(p, g, m, v, table, wd, factor, flag, lr_t, b1, b2, eps);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
