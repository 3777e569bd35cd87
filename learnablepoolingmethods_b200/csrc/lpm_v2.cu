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

// Column-statistic kernels: a CTA owns 128 columns (64 threads x 2) and one chunk of rows; its four row lanes walk the
// chunk with a stride of 4 rows, four rows per lane in flight, and are combined through shared memory.  The grid is
// (C/128, chunks) with chunks chosen so that ~8 CTAs per SM are resident (colstats_chunks).
constexpr int CS_COLS = 128, CS_LANES = 4, CS_UNROLL = 4;

// partial[chunk][0][c] = sum_r x[r][c], partial[chunk][1][c] = sum_r x[r][c]^2
__global__ void __launch_bounds__(256) colstats_kernel(const __half* __restrict__ x, long long ld, long long rows, int C,
                                                       float* __restrict__ partial) {
  __shared__ float4 sh[CS_LANES][CS_COLS / 2];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int c = blockIdx.x * CS_COLS + tx * 2;
  const bool ok = c < C;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (ok) {
    for (long long r = r0 + ty; r < r1; r += CS_LANES * CS_UNROLL) {
      float2 v[CS_UNROLL];
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u) {
        const long long rr = r + u * CS_LANES;
        v[u] = rr < r1 ? __half22float2(*reinterpret_cast<const __half2*>(x + rr * ld + c)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u) { s0 += v[u].x; s1 += v[u].y; q0 += v[u].x * v[u].x; q1 += v[u].y * v[u].y; }
    }
  }
  sh[ty][tx] = make_float4(s0, s1, q0, q1);
  __syncthreads();
  if (ty == 0 && ok) {
    float4 t = sh[0][tx];
#pragma unroll
    for (int l = 1; l < CS_LANES; ++l) { const float4 o = sh[l][tx]; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
    float* p = partial + (size_t)blockIdx.y * 2 * C;
    p[c] = t.x; p[c + 1] = t.y; p[C + c] = t.z; p[C + c + 1] = t.w;
  }
}

// y[r][c] = x[r][c]*scale[c] + shift[c]  (y may alias x; 8 columns per thread)
__global__ void __launch_bounds__(256) affine_cols_kernel(const __half* x, __half* y, long long rows, int C,
                                                          const float* __restrict__ scale, const float* __restrict__ shift) {
  const long long n8 = rows * C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * 8) % C);
    uint4 v = reinterpret_cast<const uint4*>(x)[i];
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      h[j] = __floats2half2_rn(f.x * scale[c + 2 * j] + shift[c + 2 * j], f.y * scale[c + 2 * j + 1] + shift[c + 2 * j + 1]);
    }
    reinterpret_cast<uint4*>(y)[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// slim.batch_norm backward over the rows of an fp16 matrix (training statistics):
//   xhat = (x - p0[c]) * p1[c]            mode 0: x = BN input,  p0 = batch mean, p1 = rstd
//   xhat = (x - p0[c]) / p1[c]            mode 1: x = BN output, p0 = beta,       p1 = gamma
//   pass 1: partial[chunk][0][c] = sum dy, partial[chunk][1][c] = sum dy*xhat      (= dbeta, dgamma)
//   pass 2: dx = gamma*rstd*(dy - c1/N - xhat*c2/N), optionally masked by (x > 0) for a preceding ReLU
// ------------------------------------------------------------------------------------------------
// dy may be fp16 or fp32 (the batch-norm backward removes a large common-mode part of dy, so the producers that can
// afford it hand over fp32); `q` (optional, fp32 [rows/T][C]) is subtracted on the fly: dy_eff = dy - q[r/T][c].
template <typename TD>
__device__ __forceinline__ float2 ld2(const TD* p);
template <>
__device__ __forceinline__ float2 ld2<__half>(const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }
template <>
__device__ __forceinline__ float2 ld2<float>(const float* p) { return *reinterpret_cast<const float2*>(p); }

template <typename TD>
__global__ void __launch_bounds__(256) bn_bwd_stats_kernel(const TD* __restrict__ dy, long long ld_dy,
                                                           const float* __restrict__ q, int T,
                                                           const __half* __restrict__ x, long long ld_x, long long rows,
                                                           int C, const float* __restrict__ p0, const float* __restrict__ p1,
                                                           int mode, float* __restrict__ partial) {
  __shared__ float4 sh[CS_LANES][CS_COLS / 2];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int c = blockIdx.x * CS_COLS + tx * 2;
  const bool ok = c < C;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (ok) {
    const float a0 = p0[c], a1 = p0[c + 1];
    const float b0 = mode ? 1.f / p1[c] : p1[c], b1 = mode ? 1.f / p1[c + 1] : p1[c + 1];
    for (long long r = r0 + ty; r < r1; r += CS_LANES * CS_UNROLL) {
      float2 g[CS_UNROLL], v[CS_UNROLL];
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u) {
        const long long rr = r + u * CS_LANES;
        if (rr < r1) {
          g[u] = ld2<TD>(dy + rr * ld_dy + c);
          if (q) { const float2 qq = *reinterpret_cast<const float2*>(q + (rr / T) * C + c); g[u].x -= qq.x; g[u].y -= qq.y; }
          v[u] = __half22float2(*reinterpret_cast<const __half2*>(x + rr * ld_x + c));
        } else {
          g[u] = make_float2(0.f, 0.f);
          v[u] = make_float2(a0, a1);
        }
      }
#pragma unroll
      for (int u = 0; u < CS_UNROLL; ++u) {
        s0 += g[u].x; s1 += g[u].y;
        q0 += g[u].x * (v[u].x - a0) * b0; q1 += g[u].y * (v[u].y - a1) * b1;
      }
    }
  }
  sh[ty][tx] = make_float4(s0, s1, q0, q1);
  __syncthreads();
  if (ty == 0 && ok) {
    float4 t = sh[0][tx];
#pragma unroll
    for (int l = 1; l < CS_LANES; ++l) { const float4 o = sh[l][tx]; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
    float* p = partial + (size_t)blockIdx.y * 2 * C;
    p[c] = t.x; p[c + 1] = t.y; p[C + c] = t.z; p[C + c + 1] = t.w;
  }
}

template <typename TD>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const TD* __restrict__ dy, const float* __restrict__ q, int T,
                                                           __half* __restrict__ dx, const __half* __restrict__ x,
                                                           long long rows, int C, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ csum, float inv_n, int relu) {
  const long long n2 = rows * C / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 2;
    const long long r = e / C;
    const int c0 = (int)(e - r * C), c1 = c0 + 1;
    float2 g = ld2<TD>(dy + e);
    if (q) { const float2 qq = *reinterpret_cast<const float2*>(q + (r / T) * C + c0); g.x -= qq.x; g.y -= qq.y; }
    const float2 xv = __half22float2(*reinterpret_cast<const __half2*>(x + e));
    float o0 = gamma[c0] * rstd[c0] * (g.x - csum[c0] * inv_n - (xv.x - mean[c0]) * rstd[c0] * csum[C + c0] * inv_n);
    float o1 = gamma[c1] * rstd[c1] * (g.y - csum[c1] * inv_n - (xv.y - mean[c1]) * rstd[c1] * csum[C + c1] * inv_n);
    if (relu) { if (!(xv.x > 0.f)) o0 = 0.f; if (!(xv.y > 0.f)) o1 = 0.f; }
    *reinterpret_cast<__half2*>(dx + e) = __floats2half2_rn(o0, o1);
  }
}

// dy16[r][k] = fp16(G[r][k] - q[r / T][k])      (gradient of the aggregation wrt the supplied assignments)
__global__ void __launch_bounds__(256) sub_q_cast_kernel(const float* __restrict__ G, const float* __restrict__ q,
                                                         long long rows, int T, int K, __half* __restrict__ out) {
  const long long n = rows * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / K;
    const int k = (int)(i - r * K);
    out[i] = __float2half_rn(G[i] - q[(r / T) * K + k]);
  }
}

// out[b][k][d] = in[b*in_stride + d*K + k]   (gradient of the d-major flatten back to cluster-major)
__global__ void dmajor_to_kmajor_f16_kernel(const __half* __restrict__ in, long long in_stride, int K, int D,
                                            __half* __restrict__ out) {
  __shared__ __half tile[32][34];
  const int b = blockIdx.z, d0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int d = d0 + j, k = k0 + threadIdx.x;
    if (d < D && k < K) tile[j][threadIdx.x] = in[(size_t)b * in_stride + (size_t)d * K + k];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int k = k0 + j, d = d0 + threadIdx.x;
    if (d < D && k < K) out[((size_t)b * K + k) * D + d] = tile[threadIdx.x][j];
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
__global__ void __launch_bounds__(256) dropout_kernel(const __half* x, __half* y, long long n, const __half* __restrict__ mask_in,
                                                      __half* __restrict__ mask_out, unsigned long long seed,
                                                      const unsigned long long* __restrict__ seed_dev, float rate) {
  // effective seed = seed + *seed_dev: the device part lets a captured CUDA graph draw a fresh mask on every replay
  if (seed_dev) seed += *seed_dev;
  const float inv_keep = 1.f / (1.f - rate);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float keep;
    if (mask_in) keep = __half2float(mask_in[i]);
    else keep = ((hash32(seed * 0x100000001B3ull + (uint64_t)i) >> 8) * (1.f / 16777216.f)) >= rate ? 1.f : 0.f;
    if (mask_out) mask_out[i] = __float2half_rn(keep);
    y[i] = __float2half_rn(__half2float(x[i]) * keep * inv_keep);      // y may be x (in place)
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


// 64 x 64 fp16 tile transpose with 16-byte global accesses on both sides: dst[c][r] = src[r][c] * (row_scale ? row_scale[r] : 1).
// src: [R][C] (row stride lds), dst: [C][R] (row stride ldd); R, C multiples of 8, 16-byte aligned rows.  blockIdx.z = batch.
__global__ void __launch_bounds__(256) transpose64_f16_kernel(const __half* __restrict__ src, long long lds, long long src_batch,
                                                              const float* __restrict__ row_scale, long long scale_batch,
                                                              int R, int C, __half* __restrict__ dst, long long ldd,
                                                              long long dst_batch) {
  __shared__ __half tile[64][64 + 8];
  const int b = blockIdx.z, r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const __half* s = src + (size_t)b * src_batch;
  __half* d = dst + (size_t)b * dst_batch;
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int r = i >> 3, ch = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r0 + r < R && c0 + ch * 8 < C) {
      v = __ldg(reinterpret_cast<const uint4*>(s + (size_t)(r0 + r) * lds + c0 + ch * 8));
      if (row_scale) {
        const float f = __ldg(row_scale + (size_t)b * scale_batch + r0 + r);
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(h[j]); h[j] = __floats2half2_rn(t.x * f, t.y * f); }
      }
    }
    *reinterpret_cast<uint4*>(&tile[r][ch * 8]) = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int c = i & 63, ch = i >> 6;            // consecutive threads -> consecutive tile columns: conflict-free reads
    if (c0 + c < C && r0 + ch * 8 < R) {
      __align__(16) __half o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = tile[ch * 8 + j][c];
      *reinterpret_cast<uint4*>(d + (size_t)(c0 + c) * ldd + r0 + ch * 8) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

static inline bool vec16_ok(const void* p, long long ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 8 == 0; }

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

int colstats_chunks(long long rows, int C) {
  // ~8 CTAs per SM in total, at least 16 rows per chunk, at most 512 chunks (the partials are reduced by colsum_final)
  const long long gx = (C + CS_COLS - 1) / CS_COLS;
  long long c = (8ll * num_sms() + gx - 1) / gx;
  const long long by_rows = (rows + 15) / 16;
  if (c > by_rows) c = by_rows;
  if (c > 512) c = 512;
  return (int)(c < 1 ? 1 : c);
}

int colstats(const __half* x, long long ld, long long rows, int C, float* partial, cudaStream_t st) {
  LPM_REQUIRE(C % 2 == 0 && ld % 2 == 0, "colstats: column count and pitch must be even");
  dim3 grid((C + CS_COLS - 1) / CS_COLS, colstats_chunks(rows, C));
  colstats_kernel<<<grid, 256, 0, st>>>(x, ld, rows, C, partial);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int affine_cols(const __half* x, __half* y, long long rows, int C, const float* scale, const float* shift,
                cudaStream_t st) {
  LPM_REQUIRE(C % 8 == 0, "affine_cols: column count must be a multiple of 8");
  affine_cols_kernel<<<grid_for_v2(rows * C / 8, 256), 256, 0, st>>>(x, y, rows, C, scale, shift);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int bn_bwd_stats(const void* dy, int dy_f32, long long ld_dy, const float* q, int T, const __half* x, long long ld_x,
                 long long rows, int C, const float* p0, const float* p1, int mode, float* partial, cudaStream_t st) {
  LPM_REQUIRE(C % 2 == 0 && ld_dy % 2 == 0 && ld_x % 2 == 0, "bn_bwd_stats: even column count / pitches required");
  dim3 grid((C + CS_COLS - 1) / CS_COLS, colstats_chunks(rows, C));
  if (dy_f32) bn_bwd_stats_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(dy), ld_dy, q, T, x, ld_x, rows, C, p0, p1, mode, partial);
  else bn_bwd_stats_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(dy), ld_dy, q, T, x, ld_x, rows, C, p0, p1, mode, partial);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int bn_bwd_apply(const void* dy, int dy_f32, const float* q, int T, __half* dx, const __half* x, long long rows, int C,
                 const float* mean, const float* rstd, const float* gamma, const float* csum, int relu, cudaStream_t st) {
  LPM_REQUIRE(C % 2 == 0, "bn_bwd_apply: column count must be even");
  const int grid = grid_for_v2(rows * C / 2, 256);
  if (dy_f32) bn_bwd_apply_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(dy), q, T, dx, x, rows, C, mean, rstd, gamma, csum, 1.f / (float)rows, relu);
  else bn_bwd_apply_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(dy), q, T, dx, x, rows, C, mean, rstd, gamma, csum, 1.f / (float)rows, relu);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int sub_q_cast(const float* G, const float* q, long long rows, int T, int K, __half* out, cudaStream_t st) {
  sub_q_cast_kernel<<<grid_for_v2(rows * K, 256), 256, 0, st>>>(G, q, rows, T, K, out);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int dmajor_to_kmajor_f16(const __half* in, long long in_stride, int B, int K, int D, __half* out, cudaStream_t st) {
  if (K % 8 == 0 && D % 8 == 0 && vec16_ok(in, in_stride) && vec16_ok(out, D)) {
    // src = in[b] viewed as [D][K], dst = out[b] as [K][D]
    transpose64_f16_kernel<<<dim3((K + 63) / 64, (D + 63) / 64, B), 256, 0, st>>>(in, K, in_stride, nullptr, 0, D, K, out, D,
                                                                                  (long long)K * D);
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  dim3 grid((K + 31) / 32, (D + 31) / 32, B), block(32, 8);
  dmajor_to_kmajor_f16_kernel<<<grid, block, 0, st>>>(in, in_stride, K, D, out);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int dropout_f16(__half* x, __half* out, long long n, const __half* mask_in, __half* mask_out, unsigned long long seed,
                const unsigned long long* seed_dev, float rate, cudaStream_t st) {
  LPM_REQUIRE(rate >= 0.f && rate < 1.f, "dropout: rate must be in [0,1)");
  dropout_kernel<<<grid_for_v2(n, 256), 256, 0, st>>>(x, out ? out : x, n, mask_in, mask_out, seed, seed_dev, rate);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int vlad_dmajor_f16(const __half* z, const float* rscale, int B, int K, int D, __half* out, long long out_stride,
                    cudaStream_t st) {
  if (K % 8 == 0 && D % 8 == 0 && vec16_ok(z, D) && vec16_ok(out, out_stride)) {
    // src = z[b] as [K][D] scaled per row by rscale[b][k], dst = out[b] viewed as [D][K]
    transpose64_f16_kernel<<<dim3((D + 63) / 64, (K + 63) / 64, B), 256, 0, st>>>(z, D, (long long)K * D, rscale, K, K, D, out, K,
                                                                                  out_stride);
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  dim3 grid((D + 31) / 32, (K + 31) / 32, B), block(32, 8);
  vlad_dmajor_f16_kernel<<<grid, block, 0, st>>>(z, rscale, K, D, out, out_stride);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
