// NetVladV2 ("attention-based cluster similarities") specific kernels:
//   * batch-norm statistics of the attention logits without materialising them (transformer_utils.py:646-654):
//     channel = key index j, statistics over (sample, head, query).  For one (sample, head):
//        sum_i  q_i.k_j      = (sum_i q_i) . k_j
//        sum_i (q_i.k_j)^2   = k_j^T (sum_i q_i q_i^T) k_j          (16 x 16 Gram matrix of the queries)
//   * per-column sum / sum-of-squares partials of an fp16 matrix (attention_bn, filter_bn, feed_output_bn)
//   * per-column affine in place (applies a folded batch norm), dropout with an explicit or generated mask
//   * fp16 d-major (reference flatten, index d*K + k) normalised descriptor through a tiled transpose
#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

__global__ void __launch_bounds__(256) mha_logit_stats_kernel(const __half* __restrict__ qkv, long long ld, int L, int Dm,
                                                              int H, float* __restrict__ partial) {
  extern __shared__ float smf[];
  float* sQ = smf;                 // [L][17]
  float* sK = sQ + L * 17;         // [L][17]
  float* sG = sK + L * 17;         // [16][16]
  float* sS = sG + 256;            // [16]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const __half* base = qkv + (long long)b * L * ld + h * 16;
  for (int i = threadIdx.x; i < L * 16; i += 256) {
    const int r = i >> 4, c = i & 15;
    sQ[r * 17 + c] = __half2float(base[(long long)r * ld + c]);
    sK[r * 17 + c] = __half2float(base[(long long)r * ld + Dm + c]);
  }
  __syncthreads();
  {
    const int a = threadIdx.x >> 4, c = threadIdx.x & 15;
    float g = 0.f, s = 0.f;
    for (int i = 0; i < L; ++i) {
      g += sQ[i * 17 + a] * sQ[i * 17 + c];
      if (c == 0) s += sQ[i * 17 + a];
    }
    sG[a * 16 + c] = g;
    if (c == 0) sS[a] = s;
  }
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * 2 * L;
  for (int j = threadIdx.x; j < L; j += 256) {
    float k[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) k[a] = sK[j * 17 + a];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      s1 += sS[a] * k[a];
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) t += sG[a * 16 + c] * k[c];
      s2 += k[a] * t;
    }
    out[j] = s1;
    out[L + j] = s2;
  }
}

// partial[chunk][0][c] = sum_r x[r][c], partial[chunk][1][c] = sum_r x[r][c]^2   (thread = 2 columns)
__global__ void __launch_bounds__(256) colstats_kernel(const __half* __restrict__ x, long long ld, long long rows, int C,
                                                       float* __restrict__ partial) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (c >= C) return;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(x + r * ld + c));
    s0 += v.x; s1 += v.y; q0 += v.x * v.x; q1 += v.y * v.y;
  }
  float* p = partial + (size_t)blockIdx.y * 2 * C;
  p[c] = s0; p[c + 1] = s1; p[C + c] = q0; p[C + c + 1] = q1;
}

// x[r][c] = x[r][c]*scale[c] + shift[c]  (in place, 8 columns per thread)
__global__ void __launch_bounds__(256) affine_cols_kernel(__half* __restrict__ x, long long rows, int C,
                                                          const float* __restrict__ scale, const float* __restrict__ shift) {
  const long long n8 = rows * C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 8) % C);
    uint4 v = reinterpret_cast<uint4*>(x)[i];
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      h[j] = __floats2half2_rn(f.x * scale[c + 2 * j] + shift[c + 2 * j], f.y * scale[c + 2 * j + 1] + shift[c + 2 * j + 1]);
    }
    reinterpret_cast<uint4*>(x)[i] = v;
  }
}

// tf.layers.dropout(rate): x *= keep/(1-rate).  keep comes from `mask_in` (fp16 0/1) or from a counter-based hash
// of (seed, element index); the mask actually used is written to `mask_out` when given.
__device__ __forceinline__ uint32_t hash32(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}
__global__ void __launch_bounds__(256) dropout_kernel(__half* __restrict__ x, long long n, const __half* __restrict__ mask_in,
                                                      __half* __restrict__ mask_out, unsigned long long seed, float rate) {
  const float inv_keep = 1.f / (1.f - rate);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float keep;
    if (mask_in) keep = __half2float(mask_in[i]);
    else keep = ((hash32(seed * 0x100000001B3ull + (uint64_t)i) >> 8) * (1.f / 16777216.f)) >= rate ? 1.f : 0.f;
    if (mask_out) mask_out[i] = __float2half_rn(keep);
    x[i] = __float2half_rn(__half2float(x[i]) * keep * inv_keep);
  }
}

// out[b][d*K + k] = fp16(z[b][k][d] * rscale[b][k])   (32 x 32 tiles through shared memory)
__global__ void vlad_dmajor_f16_kernel(const __half* __restrict__ z, const float* __restrict__ rscale, int K, int D,
                                       __half* __restrict__ out, long long out_stride) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, k0 = blockIdx.y * 32, d0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int k = k0 + j, d = d0 + threadIdx.x;
    if (k < K && d < D) tile[j][threadIdx.x] = __half2float(z[((size_t)b * K + k) * D + d]) * rscale[b * K + k];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int d = d0 + j, k = k0 + threadIdx.x;
    if (k < K && d < D) out[(size_t)b * out_stride + (size_t)d * K + k] = __float2half_rn(tile[threadIdx.x][j]);
  }
}

// ----------------------------------------------------------------------------------------------
static inline int grid_for_v2(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

int mha_logit_stats(const __half* qkv, long long ld, int B, int L, int Dm, int H, float* partial, cudaStream_t st) {
  LPM_REQUIRE(Dm / H == 16 && Dm % H == 0, "mha_logit_stats: head depth must be 16");
  const size_t smem = (size_t)(2 * L * 17 + 256 + 16) * sizeof(float);
  LPM_REQUIRE(smem <= 100 * 1024, "mha_logit_stats: sequence too long (%d)", L);
  static bool set = false;
  if (!set) { LPM_CUDA_CHECK(cudaFuncSetAttribute(mha_logit_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); set = true; }
  mha_logit_stats_kernel<<<B * H, 256, smem, st>>>(qkv, ld, L, Dm, H, partial);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int colstats_chunks(long long rows) {
  long long c = (rows + 127) / 128;
  return (int)(c > 64 ? 64 : (c < 1 ? 1 : c));
}

int colstats(const __half* x, long long ld, long long rows, int C, float* partial, cudaStream_t st) {
  LPM_REQUIRE(C % 2 == 0 && ld % 2 == 0, "colstats: column count and pitch must be even");
  dim3 grid((C / 2 + 255) / 256, colstats_chunks(rows));
  colstats_kernel<<<grid, 256, 0, st>>>(x, ld, rows, C, partial);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int affine_cols(__half* x, long long rows, int C, const float* scale, const float* shift, cudaStream_t st) {
  LPM_REQUIRE(C % 8 == 0, "affine_cols: column count must be a multiple of 8");
  affine_cols_kernel<<<grid_for_v2(rows * C / 8, 256), 256, 0, st>>>(x, rows, C, scale, shift);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int dropout_f16(__half* x, long long n, const __half* mask_in, __half* mask_out, unsigned long long seed, float rate,
                cudaStream_t st) {
  LPM_REQUIRE(rate >= 0.f && rate < 1.f, "dropout: rate must be in [0,1)");
  dropout_kernel<<<grid_for_v2(n, 256), 256, 0, st>>>(x, n, mask_in, mask_out, seed, rate);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int vlad_dmajor_f16(const __half* z, const float* rscale, int B, int K, int D, __half* out, long long out_stride,
                    cudaStream_t st) {
  dim3 grid((D + 31) / 32, (K + 31) / 32, B), block(32, 8);
  vlad_dmajor_f16_kernel<<<grid, block, 0, st>>>(z, rscale, K, D, out, out_stride);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
