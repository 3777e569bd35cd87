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

// chunk table entry: {tensor id, start offset / 32, length, element offset inside the tensor / 32}
__global__ void __launch_bounds__(256) mt_sqnorm_kernel(const float* __restrict__ g, const float* __restrict__ p,
                                                        const int* __restrict__ table, const float* __restrict__ wd,
                                                        float* __restrict__ partial) {
  __shared__ float red[8];
  const int t = table[blockIdx.x * 4], len = table[blockIdx.x * 4 + 2];
  const long long start = (long long)(unsigned)table[blockIdx.x * 4 + 1] * 32;
  const float w = wd[t];
  float s = 0.f;
  if (w == 0.f && (len & 3) == 0) {      // segment starts are 32-element aligned: float4 loads are safe
    const float4* g4 = reinterpret_cast<const float4*>(g + start);
    for (int i = threadIdx.x; i < len / 4; i += 256) {
      const float4 x = __ldg(g4 + i);
      s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
  } else {
    for (int i = threadIdx.x; i < len; i += 256) {
      const float x = g[start + i] + w * p[start + i];
      s += x * x;
    }
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

// one warp per tensor
__global__ void __launch_bounds__(256) mt_clip_kernel(const float* __restrict__ partial, const int* __restrict__ chunk_begin,
                                                      int n_tensors, float clip, float* __restrict__ factor,
                                                      float* __restrict__ norms, int* __restrict__ flag) {
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (t >= n_tensors) return;
  double s = 0.0;
  for (int c = chunk_begin[t] + lane; c < chunk_begin[t + 1]; c += 32) s += partial[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const float n = (float)sqrt(s);
    norms[t] = n;
    if (!isfinite(n)) atomicExch(flag, 1);
    factor[t] = clip > 0.f ? clip / fmaxf(n, clip) : 1.f;
  }
}

// Adam update + (optional) fp16 operand shadow of the updated parameter: element e of tensor t (inner dimension
// sh_cols[t]) goes to sh_ptr[t][(e / cols) * sh_ld[t] + e % cols].  table entry: {tensor, start/32, len, offset/32}.
__global__ void __launch_bounds__(256) mt_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ m, float* __restrict__ v,
                                                      const int* __restrict__ table, const float* __restrict__ wd,
                                                      const float* __restrict__ factor, const int* __restrict__ flag,
                                                      const unsigned long long* __restrict__ sh_ptr,
                                                      const int* __restrict__ sh_cols, const long long* __restrict__ sh_ld,
                                                      float lr_t, float b1, float b2, float eps) {
  if (*flag) return;  // overflow: skip the step
  const int t = table[blockIdx.x * 4], len = table[blockIdx.x * 4 + 2];
  const long long start = (long long)(unsigned)table[blockIdx.x * 4 + 1] * 32;
  const long long eoff = (long long)(unsigned)table[blockIdx.x * 4 + 3] * 32;
  const float w = wd[t], f = factor[t];
  __half* sdst = sh_ptr ? reinterpret_cast<__half*>(sh_ptr[t]) : nullptr;
  const int cols = sdst ? sh_cols[t] : 1;
  const long long ld = sdst ? sh_ld[t] : 0;
  const int n4 = ((cols & 3) == 0 || sdst == nullptr) ? len / 4 : 0;   // float4 body only when rows stay 4-aligned
  float4* p4 = reinterpret_cast<float4*>(p + start);
  const float4* g4 = reinterpret_cast<const float4*>(g + start);
  float4* m4 = reinterpret_cast<float4*>(m + start);
  float4* v4 = reinterpret_cast<float4*>(v + start);
  for (int i = threadIdx.x; i < n4; i += 256) {
    float4 pj = p4[i], mj = m4[i], vj = v4[i];
    const float4 gr = __ldg(g4 + i);
    float gx = (gr.x + w * pj.x) * f, gy = (gr.y + w * pj.y) * f, gz = (gr.z + w * pj.z) * f, gw = (gr.w + w * pj.w) * f;
    mj.x = b1 * mj.x + (1.f - b1) * gx; mj.y = b1 * mj.y + (1.f - b1) * gy;
    mj.z = b1 * mj.z + (1.f - b1) * gz; mj.w = b1 * mj.w + (1.f - b1) * gw;
    vj.x = b2 * vj.x + (1.f - b2) * gx * gx; vj.y = b2 * vj.y + (1.f - b2) * gy * gy;
    vj.z = b2 * vj.z + (1.f - b2) * gz * gz; vj.w = b2 * vj.w + (1.f - b2) * gw * gw;
    pj.x -= lr_t * mj.x / (sqrtf(vj.x) + eps); pj.y -= lr_t * mj.y / (sqrtf(vj.y) + eps);
    pj.z -= lr_t * mj.z / (sqrtf(vj.z) + eps); pj.w -= lr_t * mj.w / (sqrtf(vj.w) + eps);
    m4[i] = mj; v4[i] = vj; p4[i] = pj;
    if (sdst) {
      const long long e = eoff + 4ll * i;
      const long long r = e / cols;
      const int c = (int)(e - r * cols);
      uint2 o;
      o.x = pack_half2(pj.x, pj.y); o.y = pack_half2(pj.z, pj.w);
      *reinterpret_cast<uint2*>(sdst + r * ld + c) = o;
    }
  }
  for (int i = n4 * 4 + threadIdx.x; i < len; i += 256) {
    const long long j = start + i;
    const float pj = p[j];
    const float gj = (g[j] + w * pj) * f;
    const float mj = b1 * m[j] + (1.f - b1) * gj;
    const float vj = b2 * v[j] + (1.f - b2) * gj * gj;
    const float pn = pj - lr_t * mj / (sqrtf(vj) + eps);
    m[j] = mj; v[j] = vj; p[j] = pn;
    if (sdst) {
      const long long e = eoff + i;
      const long long r = e / cols;
      sdst[r * ld + (e - r * cols)] = __float2half_rn(pn);
    }
  }
}

int adam_clip_step(float* p, const float* g, float* m, float* v, const int* table, int n_chunks,
                   const int* chunk_begin, int n_tensors, const float* wd, const unsigned long long* sh_ptr,
                   const int* sh_cols, const long long* sh_ld, float clip, float lr_t, float b1, float b2, float eps,
                   float* partial, float* factor, float* norms, int* flag, cudaStream_t st) {
  mt_sqnorm_kernel<<<n_chunks, 256, 0, st>>>(g, p, table, wd, partial);
  mt_clip_kernel<<<(n_tensors + 7) / 8, 256, 0, st>>>(partial, chunk_begin, n_tensors, clip, factor, norms, flag);
  mt_adam_kernel<<<n_chunks, 256, 0, st>>>(p, g, m, v, table, wd, factor, flag, sh_ptr, sh_cols, sh_ld, lr_t, b1, b2, eps);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
