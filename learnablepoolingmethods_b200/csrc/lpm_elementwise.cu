// HBM-bound kernels of the hot path: frame sampling + input batch-norm, batch-norm finalisation,
// joint-axis layer-norm (+ residual), split-K reduction, context gating, MoE mixing, cross-entropy,
// fp32 -> fp16 parameter shadows.  Vectorised, coalesced, grid sized in multiples of the SM count.
#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

__device__ __forceinline__ float block_sum(float v, float* red) {
  // red: >= 32 floats of shared memory
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}

// ------------------------------------------------------------------------------------------------
// a2 + a3: SampleUniformFrames (model_utils.py:101-122) + input_bn (frame_level_models.py:2265-2271)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sample_index(int i, float step, int nf, int max_frames) {
  // int32( fl32(fl32(i*step) * fl32(nf)) ), truncation toward zero: the TF float32 arithmetic
  const float g = __fmul_rn(static_cast<float>(i), step);
  int idx = __float2int_rz(__fmul_rn(g, static_cast<float>(nf)));
  return min(max(idx, 0), max_frames - 1);
}

// partial[blockIdx][0][c] = sum, partial[blockIdx][1][c] = sum of squares over this block's rows
// CODES: x holds the YT8M uint8 codes; the frame is dequantised (utils.py:28-43, readers.py:185-193) and L2-normalised
// over all F features (train.py:264) on the fly, i.e. the ingest side of the boundary is fused into the gather.
// One block per frame row: 256 threads x (4 + 4) columns, F <= 2048.
struct FrameQuant { float scalar, bias; };

template <bool CODES>
__device__ __forceinline__ void load_frame(const void* __restrict__ x, size_t frame, int F, FrameQuant qz, float* red,
                                           int parity, float4 (&v)[2]) {
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = threadIdx.x * 4 + j * 1024;
    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < F) {
      if (CODES) {
        const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(static_cast<const uint8_t*>(x) + frame * F + c));
        v[j] = make_float4(fmaf((float)u.x, qz.scalar, qz.bias), fmaf((float)u.y, qz.scalar, qz.bias),
                           fmaf((float)u.z, qz.scalar, qz.bias), fmaf((float)u.w, qz.scalar, qz.bias));
        ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
      } else {
        v[j] = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(x) + frame * F + c));
      }
    }
  }
  if (CODES) {
    // tf.nn.l2_normalize over the feature axis: x * rsqrt(max(sum x^2, 1e-12)); block reduction, one barrier per
    // frame (the 8 warp partials alternate between two slots)
    ss = warp_sum(ss);
    float* slot = red + parity * 8;
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += slot[w];
    const float rn = rsqrtf(fmaxf(tot, 1e-12f));
#pragma unroll
    for (int j = 0; j < 2; ++j) { v[j].x *= rn; v[j].y *= rn; v[j].z *= rn; v[j].w *= rn; }
  }
}

template <bool CODES>
__global__ void __launch_bounds__(256) sample_stats_kernel(const void* __restrict__ x,
                                                           const int* __restrict__ num_frames, int B,
                                                           int max_frames, int F, int T, float step, FrameQuant qz,
                                                           const int* __restrict__ frame_index,
                                                           float* __restrict__ partial) {
  __shared__ float red[16];
  const int rows = B * T;
  float4 s[2], q[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) { s[j] = make_float4(0, 0, 0, 0); q[j] = make_float4(0, 0, 0, 0); }
  int parity = 0;
  for (int r = blockIdx.x; r < rows; r += gridDim.x, parity ^= 1) {
    const int b = r / T, i = r - b * T;
    // explicit indices (random frame sampling, model_utils.py:26-73) or the uniform rule (model_utils.py:101-122)
    const int idx = frame_index ? min(max(__ldg(frame_index + r), 0), max_frames - 1)
                                : sample_index(i, step, __ldg(num_frames + b), max_frames);
    float4 v[2];
    load_frame<CODES>(x, (size_t)b * max_frames + idx, F, qz, red, parity, v);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      s[j].x += v[j].x; s[j].y += v[j].y; s[j].z += v[j].z; s[j].w += v[j].w;
      q[j].x += v[j].x * v[j].x; q[j].y += v[j].y * v[j].y; q[j].z += v[j].z * v[j].z; q[j].w += v[j].w * v[j].w;
    }
  }
  float* ps = partial + (size_t)blockIdx.x * 2 * F;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = threadIdx.x * 4 + j * 1024;
    if (c < F) {
      *reinterpret_cast<float4*>(ps + c) = s[j];
      *reinterpret_cast<float4*>(ps + F + c) = q[j];
    }
  }
}

// y[r][c] = fp16( x[b, idx(b,i), c] * scale[c] + shift[c] )
template <bool CODES>
__global__ void __launch_bounds__(256) sample_apply_kernel(const void* __restrict__ x,
                                                           const int* __restrict__ num_frames, int B,
                                                           int max_frames, int F, int T, float step, FrameQuant qz,
                                                           const int* __restrict__ frame_index,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           __half* __restrict__ y, int split_col,
                                                           __half* __restrict__ y2) {
  // y2 == null: one [rows][F] matrix.  Otherwise columns [0, split_col) go to y ([rows][split_col]) and the rest
  // to y2 ([rows][F - split_col]): contiguous per-modality matrices for NetVladV2's residual / layer norm.
  __shared__ float red[16];
  const int rows = B * T;
  int parity = 0;
  for (int r = blockIdx.x; r < rows; r += gridDim.x, parity ^= 1) {
    const int b = r / T, i = r - b * T;
    // explicit indices (random frame sampling, model_utils.py:26-73) or the uniform rule (model_utils.py:101-122)
    const int idx = frame_index ? min(max(__ldg(frame_index + r), 0), max_frames - 1)
                                : sample_index(i, step, __ldg(num_frames + b), max_frames);
    float4 vv[2];
    load_frame<CODES>(x, (size_t)b * max_frames + idx, F, qz, red, parity, vv);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = threadIdx.x * 4 + j * 1024;
      if (c >= F) continue;
      __half* dst = y2 == nullptr ? y + (size_t)r * F : (c < split_col ? y + (size_t)r * split_col : y2 + (size_t)r * (F - split_col) - split_col);
      const float4 v = vv[j];
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c));
      uint2 o;
      o.x = pack_half2(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y));
      o.y = pack_half2(fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
      *reinterpret_cast<uint2*>(dst + c) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-per-frame versions of the two kernels above for feature sizes that are multiples of 128 (rgb 1024 + audio 128 =
// 9 x 128): a lane owns columns {j*128 + lane*4 .. +3}, every warp-level access is 128 (codes) / 512 (fp32) contiguous
// bytes, the per-frame L2 norm is a warp shuffle instead of a block barrier, and eight frames are in flight per CTA.
// The block-per-frame kernels spent their time in that barrier (40 us for 23.6 MB of codes).
// ------------------------------------------------------------------------------------------------
template <bool CODES, int NJ>
__device__ __forceinline__ void load_frame_warp(const void* __restrict__ x, size_t frame, int F, FrameQuant qz, int lane,
                                                float4 (&v)[NJ]) {
  if (CODES) {
    uchar4 u[NJ];
    const uint8_t* p = static_cast<const uint8_t*>(x) + frame * F + lane * 4;
#pragma unroll
    for (int j = 0; j < NJ; ++j) u[j] = __ldg(reinterpret_cast<const uchar4*>(p + j * 128));
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      v[j] = make_float4(fmaf((float)u[j].x, qz.scalar, qz.bias), fmaf((float)u[j].y, qz.scalar, qz.bias),
                         fmaf((float)u[j].z, qz.scalar, qz.bias), fmaf((float)u[j].w, qz.scalar, qz.bias));
      ss += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    }
    ss = warp_sum(ss);                                           // tf.nn.l2_normalize over the feature axis (train.py:264)
    const float rn = rsqrtf(fmaxf(ss, 1e-12f));
#pragma unroll
    for (int j = 0; j < NJ; ++j) { v[j].x *= rn; v[j].y *= rn; v[j].z *= rn; v[j].w *= rn; }
  } else {
    const float* p = static_cast<const float*>(x) + frame * F + lane * 4;
#pragma unroll
    for (int j = 0; j < NJ; ++j) v[j] = __ldg(reinterpret_cast<const float4*>(p + j * 128));
  }
}

template <bool CODES, int NJ>
__global__ void __launch_bounds__(256) sample_stats_warp_kernel(const void* __restrict__ x, const int* __restrict__ num_frames,
                                                                int B, int max_frames, int F, int T, float step, FrameQuant qz,
                                                                const int* __restrict__ frame_index, float* __restrict__ partial) {
  __shared__ float sh[2 * NJ * 128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = B * T;
  float4 s[NJ], q[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) { s[j] = make_float4(0, 0, 0, 0); q[j] = make_float4(0, 0, 0, 0); }
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const int b = r / T, i = r - b * T;
    const int idx = frame_index ? min(max(__ldg(frame_index + r), 0), max_frames - 1)
                                : sample_index(i, step, __ldg(num_frames + b), max_frames);
    float4 v[NJ];
    load_frame_warp<CODES, NJ>(x, (size_t)b * max_frames + idx, F, qz, lane, v);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      s[j].x += v[j].x; s[j].y += v[j].y; s[j].z += v[j].z; s[j].w += v[j].w;
      q[j].x += v[j].x * v[j].x; q[j].y += v[j].y * v[j].y; q[j].z += v[j].z * v[j].z; q[j].w += v[j].w * v[j].w;
    }
  }
  // the eight warps add their column partials in warp order (deterministic)
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float4* ps = reinterpret_cast<float4*>(sh + j * 128 + lane * 4);
        float4* pq = reinterpret_cast<float4*>(sh + NJ * 128 + j * 128 + lane * 4);
        if (w == 0) { *ps = s[j]; *pq = q[j]; }
        else {
          float4 a = *ps, c = *pq;
          a.x += s[j].x; a.y += s[j].y; a.z += s[j].z; a.w += s[j].w;
          c.x += q[j].x; c.y += q[j].y; c.z += q[j].z; c.w += q[j].w;
          *ps = a; *pq = c;
        }
      }
    }
    __syncthreads();
  }
  float* pp = partial + (size_t)blockIdx.x * 2 * F;
  for (int i = threadIdx.x; i < 2 * F; i += 256) pp[i] = sh[i];
}

template <bool CODES, int NJ>
__global__ void __launch_bounds__(256) sample_apply_warp_kernel(const void* __restrict__ x, const int* __restrict__ num_frames,
                                                                int B, int max_frames, int F, int T, float step, FrameQuant qz,
                                                                const int* __restrict__ frame_index,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                __half* __restrict__ y, int split_col, __half* __restrict__ y2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = B * T;
  float4 sc[NJ], sf[NJ];                                          // this lane's slice of the folded input_bn affine
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    sc[j] = __ldg(reinterpret_cast<const float4*>(scale + j * 128 + lane * 4));
    sf[j] = __ldg(reinterpret_cast<const float4*>(shift + j * 128 + lane * 4));
  }
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const int b = r / T, i = r - b * T;
    const int idx = frame_index ? min(max(__ldg(frame_index + r), 0), max_frames - 1)
                                : sample_index(i, step, __ldg(num_frames + b), max_frames);
    float4 v[NJ];
    load_frame_warp<CODES, NJ>(x, (size_t)b * max_frames + idx, F, qz, lane, v);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = j * 128 + lane * 4;
      __half* dst = y2 == nullptr ? y + (size_t)r * F + c
                                  : (c < split_col ? y + (size_t)r * split_col + c : y2 + (size_t)r * (F - split_col) + (c - split_col));
      uint2 o;
      o.x = pack_half2(fmaf(v[j].x, sc[j].x, sf[j].x), fmaf(v[j].y, sc[j].y, sf[j].y));
      o.y = pack_half2(fmaf(v[j].z, sc[j].z, sf[j].z), fmaf(v[j].w, sc[j].w, sf[j].w));
      *reinterpret_cast<uint2*>(dst) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// slim.batch_norm finalisation: partial sums -> mean / biased var -> affine (scale, shift); moving
// statistics update (decay 0.999; Bessel-corrected variance when `bessel`).  Inference: affine from
// the moving statistics.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psq,
                                                          int P, long long pstride, int C, double count,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                                          float decay, float eps, int bessel, int training,
                                                          float* __restrict__ scale, float* __restrict__ shift,
                                                          float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  // 32 channels x 32 partial-row lanes per block (the reduction is pure load latency: keep many loads in flight)
  __shared__ double red[2][32][33];
  const int cx = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double s = 0.0, q = 0.0;
  if (training && c < C) {
    int p = ky;
    for (; p + 32 < P; p += 64) {
      const float s0 = psum[(size_t)p * pstride + c], s1 = psum[(size_t)(p + 32) * pstride + c];
      const float q0 = psq[(size_t)p * pstride + c], q1 = psq[(size_t)(p + 32) * pstride + c];
      s += (double)s0 + (double)s1;
      q += (double)q0 + (double)q1;
    }
    for (; p < P; p += 32) {
      s += static_cast<double>(psum[(size_t)p * pstride + c]);
      q += static_cast<double>(psq[(size_t)p * pstride + c]);
    }
  }
  red[0][ky][cx] = s;
  red[1][ky][cx] = q;
  __syncthreads();
  if (ky != 0 || c >= C) return;
  float mean, var;
  if (training) {
    s = 0.0; q = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) { s += red[0][k][cx]; q += red[1][k][cx]; }
    const double m = s / count;
    double v = q / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = static_cast<float>(m);
    var = static_cast<float>(v);
    if (moving_mean != nullptr) {
      const double corr = (bessel && count > 1.0) ? count / (count - 1.0) : 1.0;
      moving_mean[c] = moving_mean[c] * decay + mean * (1.f - decay);
      moving_var[c] = moving_var[c] * decay + static_cast<float>(v * corr) * (1.f - decay);
    }
  } else {
    mean = moving_mean[c];
    var = moving_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  scale[c] = g * rstd;
  shift[c] = bt - mean * g * rstd;
  if (save_mean) save_mean[c] = mean;
  if (save_rstd) save_rstd[c] = rstd;
}

// ------------------------------------------------------------------------------------------------
// Joint-axis layer norm (tf.contrib.layers.layer_norm, begin_norm_axis=1) with fused residual:
//   u = a + b * row_scale ;  y = (u - mean_b) * rstd_b * gamma[d] + beta[d]
// pass 1 writes u (fp16, in place over a) and per-(sample, chunk) partial sums;
// pass 2 reduces the partials (fixed order) and normalises.  transformer_utils.py:406-411,712-713.
// ------------------------------------------------------------------------------------------------
constexpr int LN_CHUNKS = 32;

__global__ void __launch_bounds__(256) ln_stats_kernel(const __half* a, const __half* __restrict__ b,
                                                       const float* __restrict__ b_row_scale, __half* u_out,
                                                       int rows, int D, long long a_sample_stride,
                                                       long long b_sample_stride, float* __restrict__ partial) {
  __shared__ float red[32];
  const int sample = blockIdx.y, chunk = blockIdx.x;
  const long long n8 = (long long)rows * D / 8;
  const long long per = (n8 + LN_CHUNKS - 1) / LN_CHUNKS;
  const long long i0 = chunk * per, i1 = min(n8, i0 + per);
  const uint4* pa = reinterpret_cast<const uint4*>(a + sample * a_sample_stride);
  uint4* pu = reinterpret_cast<uint4*>(u_out + sample * a_sample_stride);   // u_out may alias a
  const uint4* pb = b ? reinterpret_cast<const uint4*>(b + sample * b_sample_stride) : nullptr;
  const bool copy_only = (pb == nullptr) && (u_out != a);
  float s = 0.f, q = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    uint4 va = pa[i];
    __half2* ha = reinterpret_cast<__half2*>(&va);
    float u[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(ha[j]); u[2 * j] = f.x; u[2 * j + 1] = f.y; }
    if (pb) {
      const uint4 vb = __ldg(pb + i);
      const __half2* hb = reinterpret_cast<const __half2*>(&vb);
      const float rs = b_row_scale ? __ldg(b_row_scale + (long long)sample * rows + (i * 8) / D) : 1.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(hb[j]); u[2 * j] += f.x * rs; u[2 * j + 1] += f.y * rs; }
#pragma unroll
      for (int j = 0; j < 4; ++j) ha[j] = __floats2half2_rn(u[2 * j], u[2 * j + 1]);
      pu[i] = va;
      // statistics of the stored (fp16-rounded) u, which is what pass 2 normalises
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(ha[j]); u[2 * j] = f.x; u[2 * j + 1] = f.y; }
    }
    if (copy_only) pu[i] = va;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += u[j]; q += u[j] * u[j]; }
  }
  s = block_sum(s, red);
  q = block_sum(q, red);
  if (threadIdx.x == 0) {
    partial[((size_t)sample * LN_CHUNKS + chunk) * 2 + 0] = s;
    partial[((size_t)sample * LN_CHUNKS + chunk) * 2 + 1] = q;
  }
}

__global__ void __launch_bounds__(256) ln_apply_kernel(const __half* __restrict__ u, int rows, int D,
                                                       long long u_sample_stride, const float* __restrict__ partial,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, __half* __restrict__ y, long long y_sample_stride,
                                                       float* __restrict__ save_mean_rstd) {
  const int sample = blockIdx.y, chunk = blockIdx.x;
  double ds = 0.0, dq = 0.0;
  for (int c = 0; c < LN_CHUNKS; ++c) {
    ds += partial[((size_t)sample * LN_CHUNKS + c) * 2 + 0];
    dq += partial[((size_t)sample * LN_CHUNKS + c) * 2 + 1];
  }
  const double n = (double)rows * D;
  const double m = ds / n;
  double var = dq / n - m * m;
  if (var < 0.0) var = 0.0;
  const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)eps));
  if (save_mean_rstd && chunk == 0 && threadIdx.x == 0) {
    save_mean_rstd[sample * 2 + 0] = mean;
    save_mean_rstd[sample * 2 + 1] = rstd;
  }
  const long long n8 = (long long)rows * D / 8;
  const long long per = (n8 + gridDim.x - 1) / gridDim.x;
  const long long i0 = chunk * per, i1 = min(n8, i0 + per);
  const uint4* pu = reinterpret_cast<const uint4*>(u + sample * u_sample_stride);
  uint4* py = reinterpret_cast<uint4*>(y + sample * y_sample_stride);
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const uint4 vu = __ldg(pu + i);
    const __half2* hu = reinterpret_cast<const __half2*>(&vu);
    const int d = (int)((i * 8) % D);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + d));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + d + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + d));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + d + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint4 vo;
    __half2* ho = reinterpret_cast<__half2*>(&vo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(hu[j]);
      ho[j] = __floats2half2_rn((f.x - mean) * rstd * gg[2 * j] + bb[2 * j],
                                (f.y - mean) * rstd * gg[2 * j + 1] + bb[2 * j + 1]);
    }
    py[i] = vo;
  }
}

// ------------------------------------------------------------------------------------------------
// split-K reduction + bias (+ReLU):  out[r][c] = act( sum_s part[s][r][c] + bias[c] )  -> fp32 and/or fp16
// (hidden projection, frame_level_models.py:2319,2329-2334, and split-K weight gradients)
// ------------------------------------------------------------------------------------------------
// part2 (optional): a second set of partials summed in (the low-order pass of a split-precision product).  split3: out16
// is [n / cols][3 * cols] = [hi | lo | hi] of the result (the split-precision operand of the next product).
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int splits,
                                                            long long split_stride, const float* __restrict__ part2,
                                                            int splits2, long long split_stride2, long long n, int cols,
                                                            const float* __restrict__ bias, int relu, float alpha,
                                                            int accumulate, float* __restrict__ out32,
                                                            __half* __restrict__ out16, int split3) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // four independent partial sums: the loop is load-latency bound (up to 148 splits), order stays fixed
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 3 < splits; k += 4) {
      s0 += part[k * split_stride + i]; s1 += part[(k + 1) * split_stride + i];
      s2 += part[(k + 2) * split_stride + i]; s3 += part[(k + 3) * split_stride + i];
    }
    for (; k < splits; ++k) s0 += part[k * split_stride + i];
    float s = (s0 + s1) + (s2 + s3);
    if (part2 != nullptr) {
      float t0 = 0.f, t1 = 0.f;
      for (k = 0; k + 1 < splits2; k += 2) { t0 += part2[k * split_stride2 + i]; t1 += part2[(k + 1) * split_stride2 + i]; }
      for (; k < splits2; ++k) t0 += part2[k * split_stride2 + i];
      s += t0 + t1;
    }
    s *= alpha;
    if (bias) s += __ldg(bias + (i % cols));
    if (relu) s = fmaxf(s, 0.f);
    if (out32) { if (accumulate) s += out32[i]; out32[i] = s; }
    if (out16) {
      const __half hi = __float2half_rn(s);
      if (split3) {
        const long long r = i / cols;
        __half* d = out16 + r * 3 * cols + (i - r * cols);
        d[0] = hi; d[cols] = __float2half_rn(s - __half2float(hi)); d[2 * cols] = hi;
      } else {
        out16[i] = hi;
      }
    }
  }
}

// Many partials, few outputs (the hidden projection: 74-222 partials of an [80, 512] result): a block owns 32 consecutive
// outputs, its 8 warps each sum every 8th partial (coalesced 128-byte rows, four loads in flight), and warp 0 adds the
// eight sums in warp order (deterministic).  The one-thread-per-output kernel above walks the partials serially.
__global__ void __launch_bounds__(256) splitk_reduce_wide_kernel(const float* __restrict__ part, int splits,
                                                                 long long split_stride, const float* __restrict__ part2,
                                                                 int splits2, long long split_stride2, long long n, int cols,
                                                                 const float* __restrict__ bias, int relu, float alpha,
                                                                 int accumulate, float* __restrict__ out32,
                                                                 __half* __restrict__ out16, int split3) {
  __shared__ float sh[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < n) {
    int k = w;
    for (; k + 24 < splits; k += 32) {
      s0 += part[k * split_stride + i]; s1 += part[(k + 8) * split_stride + i];
      s2 += part[(k + 16) * split_stride + i]; s3 += part[(k + 24) * split_stride + i];
    }
    for (; k < splits; k += 8) s0 += part[k * split_stride + i];
    if (part2 != nullptr)
      for (k = w; k < splits2; k += 8) s1 += part2[k * split_stride2 + i];
  }
  sh[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && i < n) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) s += sh[ww][lane];
    s *= alpha;
    if (bias) s += __ldg(bias + (i % cols));
    if (relu) s = fmaxf(s, 0.f);
    if (out32) { if (accumulate) s += out32[i]; out32[i] = s; }
    if (out16) {
      const __half hi = __float2half_rn(s);
      if (split3) {
        const long long r = i / cols;
        __half* d = out16 + r * 3 * cols + (i - r * cols);
        d[0] = hi; d[cols] = __float2half_rn(s - __half2float(hi)); d[2 * cols] = hi;
      } else {
        out16[i] = hi;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Context gating (frame_level_models.py:2342-2368): gates = BN_batch(g [- diag(Wg) * act]) ;
// act *= sigmoid(gates).  One thread per hidden unit, loops over the (small) batch.
// ------------------------------------------------------------------------------------------------
// g_splits > 1: g holds split-K partials [g_splits][B][H] (g_split_stride apart) of the gate product; they are summed here
// (fixed order) and the sum is stored to g_sum for the backward.  split3: out16 is [B][3H] = [hi | lo | hi].
__global__ void __launch_bounds__(1024) gating_fwd_kernel(const float* __restrict__ act, const float* __restrict__ g, int g_splits,
                                  long long g_split_stride, float* __restrict__ g_sum, int split3, int B, int H,
                                  const float* __restrict__ wg_diag, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ moving_mean,
                                  float* __restrict__ moving_var, float decay, float eps, int training,
                                  float* __restrict__ out32, __half* __restrict__ out16,
                                  float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  // 32 hidden units x 32 batch lanes per block
  __shared__ double red[2][32][33];
  __shared__ float sm[2][32];
  const int cx = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool ok = c < H;
  const float dg = (wg_diag && ok) ? wg_diag[c] : 0.f;
  if (g_splits > 1) {
    if (ok)
      for (int b = ky; b < B; b += 32) {
        float t = 0.f;
        for (int k = 0; k < g_splits; ++k) t += g[k * g_split_stride + (size_t)b * H + c];
        g_sum[(size_t)b * H + c] = t;
      }
    g = g_sum;                 // a thread re-reads only the elements it wrote itself
  }
  if (training) {
    double s = 0.0, q = 0.0;
    if (ok)
      for (int b = ky; b < B; b += 32) {
        const float v = g[(size_t)b * H + c] - dg * act[(size_t)b * H + c];
        s += v; q += (double)v * v;
      }
    red[0][ky][cx] = s; red[1][ky][cx] = q;
    __syncthreads();
    if (ky == 0 && ok) {
      s = 0.0; q = 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k) { s += red[0][k][cx]; q += red[1][k][cx]; }
      const double m = s / B;
      double vv = q / B - m * m;
      if (vv < 0.0) vv = 0.0;
      const double corr = B > 1 ? (double)B / (B - 1) : 1.0;
      moving_mean[c] = moving_mean[c] * decay + (float)m * (1.f - decay);
      moving_var[c] = moving_var[c] * decay + (float)(vv * corr) * (1.f - decay);
      sm[0][cx] = (float)m; sm[1][cx] = (float)vv;
    }
    __syncthreads();
  } else if (ky == 0 && ok) {
    sm[0][cx] = moving_mean[c]; sm[1][cx] = moving_var[c];
  }
  if (!training) __syncthreads();
  if (!ok) return;
  const float mean = sm[0][cx], var = sm[1][cx];
  const float rstd = rsqrtf(var + eps);
  if (save_mean && ky == 0) { save_mean[c] = mean; save_rstd[c] = rstd; }
  const float sc = gamma[c] * rstd, sh = beta[c] - mean * sc;
  for (int b = ky; b < B; b += 32) {
    const float a = act[(size_t)b * H + c];
    const float v = (g[(size_t)b * H + c] - dg * a) * sc + sh;
    const float o = a / (1.f + __expf(-v));
    out32[(size_t)b * H + c] = o;
    if (out16) {
      const __half hi = __float2half_rn(o);
      if (split3) {
        __half* d = out16 + (size_t)b * 3 * H + c;
        d[0] = hi; d[H] = __float2half_rn(o - __half2float(hi)); d[2 * H] = hi;
      } else {
        out16[(size_t)b * H + c] = hi;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// MoE mixing (video_level_models.py:116-126): logits [B][ld] = [gates V*(M+1) | experts V*M]
//   p[b,v] = sum_m softmax(gate[b,v,:])[m] * sigmoid(expert[b,v,m]);  experts start at column expert_off
// ------------------------------------------------------------------------------------------------
__global__ void moe_mix_kernel(const float* __restrict__ logits, long long ld, int B, int V, int M,
                               int expert_off, float* __restrict__ pred) {
  const long long n = (long long)B * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / V), v = (int)(i - (long long)b * V);
    const float* gl = logits + b * ld + (long long)v * (M + 1);
    const float* el = logits + b * ld + expert_off + (long long)v * M;
    float mx = gl[0];
    for (int m = 1; m <= M; ++m) mx = fmaxf(mx, gl[m]);
    float den = 0.f, num = 0.f;
    for (int m = 0; m <= M; ++m) {
      const float e = __expf(gl[m] - mx);
      den += e;
      if (m < M) num += e / (1.f + __expf(-el[m]));
    }
    pred[i] = num / den;
  }
}

// ------------------------------------------------------------------------------------------------
// CrossEntropyLoss (losses.py:44-51): per-sample sums -> loss = mean_b ; labels as uint8 {0,1}
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) xent_rows_kernel(const float* __restrict__ pred,
                                                        const uint8_t* __restrict__ labels, int V,
                                                        float* __restrict__ row_loss) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float p = pred[(size_t)b * V + v];
    const float y = labels[(size_t)b * V + v] ? 1.f : 0.f;
    s -= y * logf(p + 1e-5f) + (1.f - y) * logf(1.f - p + 1e-5f);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) row_loss[b] = s;
}
__global__ void mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i];
    *out = (float)(s / n);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 -> fp16 2-D copy with destination leading dimension / zero column padding (parameter shadows)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_2d_kernel(const float* __restrict__ src, long long ld_src, int rows,
                                                      int cols, __half* __restrict__ dst, long long ld_dst,
                                                      int cols_dst) {
  const long long n = (long long)rows * cols_dst;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols_dst), c = (int)(i - (long long)r * cols_dst);
    dst[r * ld_dst + c] = __float2half_rn(c < cols ? src[r * ld_src + c] : 0.f);
  }
}
// fast path: cols == cols_dst, multiple of 8, all strides / bases 16-byte friendly
__global__ void __launch_bounds__(256) cast_2d_vec_kernel(const float* __restrict__ src, long long ld_src, int rows,
                                                          int cols, __half* __restrict__ dst, long long ld_dst) {
  const int c8 = cols / 8;
  const long long n = (long long)rows * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c8;
    const int c = (int)(i - r * c8) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + r * ld_src + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + r * ld_src + c + 4));
    uint4 o;
    o.x = pack_half2(a.x, a.y); o.y = pack_half2(a.z, a.w); o.z = pack_half2(b.x, b.y); o.w = pack_half2(b.z, b.w);
    *reinterpret_cast<uint4*>(dst + r * ld_dst + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// NetVLAD descriptor finalisation (frame_level_models.py:2819-2822): z[b,k,:] (un-normalised, fp16)
// times rscale[b,k] -> fp32 [B][D*K] d-major (reference flatten) or [B][K][D].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vlad_finalize_kernel(const __half* __restrict__ z, const float* __restrict__ rscale,
                                                            int B, int K, int D, int d_major, float* __restrict__ out) {
  const long long n = (long long)B * K * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / ((long long)K * D));
    const int rem = (int)(i - (long long)b * K * D);
    int k, d;
    if (d_major) { d = rem / K; k = rem - d * K; } else { k = rem / D; d = rem - k * D; }
    out[i] = __half2float(z[((long long)b * K + k) * D + d]) * rscale[b * K + k];
  }
}

// y[r][:] = fp16(x[r][:] * row_scale[r])   (materialises the normalised VLAD descriptor for training)
__global__ void __launch_bounds__(256) scale_rows_kernel(const __half* __restrict__ x, const float* __restrict__ rs,
                                                         long long rows, int D, __half* __restrict__ y) {
  const long long n8 = rows * D / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float sc = __ldg(rs + (i * 8) / D);
    uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); h[j] = __floats2half2_rn(f.x * sc, f.y * sc); }
    reinterpret_cast<uint4*>(y)[i] = v;
  }
}

// fp32 2-D transpose (parameter layout preparation, e.g. cluster_weights2 [D][K] -> [K][D])
__global__ void transpose_2d_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst,
                                    __half* __restrict__ dst16) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) {
      const float v = tile[threadIdx.x][j];
      if (dst != nullptr) dst[(size_t)c * rows + r] = v;
      if (dst16 != nullptr) dst16[(size_t)c * rows + r] = __float2half_rn(v);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------------
static inline int grid_for(long long n, int threads, int per_sm = 8) {
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)num_sms() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int sample_stats_blocks() { return num_sms() * 4; }

static FrameQuant frame_quant(float qmax, float qmin) {
  // utils.py:39-43: scalar = range / 255, bias = range / 512 + min
  const float range = qmax - qmin;
  return FrameQuant{range / 255.0f, range / 512.0f + qmin};
}

int sample_stats(const void* x, int codes, float qmax, float qmin, const int* nf, const int* frame_index, int B,
                 int max_frames, int F, int T, float* partial, cudaStream_t st) {
  LPM_REQUIRE(F % 4 == 0 && F <= 2048, "sample_stats: feature size must be a multiple of 4 and <= 2048 (got %d)", F);
  LPM_REQUIRE(!codes || qmax > qmin, "sample_stats: max_quantized_value must exceed min_quantized_value");
  const FrameQuant qz = frame_quant(qmax, qmin);
  if (F == 9 * 128) {          // rgb 1024 + audio 128: warp-per-frame kernels
    if (codes) sample_stats_warp_kernel<true, 9><<<sample_stats_blocks(), 256, 0, st>>>(x, nf, B, max_frames, F, T, 1.0f / (float)T, qz, frame_index, partial);
    else sample_stats_warp_kernel<false, 9><<<sample_stats_blocks(), 256, 0, st>>>(x, nf, B, max_frames, F, T, 1.0f / (float)T, qz, frame_index, partial);
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  if (codes) sample_stats_kernel<true><<<sample_stats_blocks(), 256, 0, st>>>(x, nf, B, max_frames, F, T, 1.0f / (float)T, qz, frame_index, partial);
  else sample_stats_kernel<false><<<sample_stats_blocks(), 256, 0, st>>>(x, nf, B, max_frames, F, T, 1.0f / (float)T, qz, frame_index, partial);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int sample_apply(const void* x, int codes, float qmax, float qmin, const int* nf, const int* frame_index, int B,
                 int max_frames, int F, int T, const float* scale, const float* shift, __half* y, int split_col,
                 __half* y2, cudaStream_t st) {
  LPM_REQUIRE(F % 4 == 0 && F <= 2048, "sample_apply: feature size must be a multiple of 4 and <= 2048 (got %d)", F);
  LPM_REQUIRE(!codes || qmax > qmin, "sample_apply: max_quantized_value must exceed min_quantized_value");
  const FrameQuant qz = frame_quant(qmax, qmin);
  LPM_REQUIRE(y2 == nullptr || (split_col % 4 == 0 && split_col > 0 && split_col < F), "sample_apply: bad split column");
  if (F == 9 * 128) {
    int gw = (B * T + 7) / 8;
    if (gw > num_sms() * 8) gw = num_sms() * 8;
    if (codes) sample_apply_warp_kernel<true, 9><<<gw, 256, 0, st>>>(x, nf, B, max_frames, F, T, 1.0f / (float)T, qz, frame_index, scale, shift, y, split_col, y2);
    else sample_apply_warp_kernel<false, 9><<<gw, 256, 0, st>>>(x, nf, B, max_frames, F, T, 1.0f / (float)T, qz, frame_index, scale, shift, y, split_col, y2);
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  int grid = B * T < num_sms() * 8 ? B * T : num_sms() * 8;
  if (codes) sample_apply_kernel<true><<<grid, 256, 0, st>>>(x, nf, B, max_frames, F, T, 1.0f / (float)T, qz, frame_index, scale, shift, y, split_col, y2);
  else sample_apply_kernel<false><<<grid, 256, 0, st>>>(x, nf, B, max_frames, F, T, 1.0f / (float)T, qz, frame_index, scale, shift, y, split_col, y2);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int bn_finalize(const float* psum, const float* psq, int P, long long pstride, int C, double count,
                const float* gamma, const float* beta, float* mm, float* mv, float decay, float eps, int bessel,
                int training, float* scale, float* shift, float* save_mean, float* save_rstd, cudaStream_t st) {
  bn_finalize_kernel<<<(C + 31) / 32, 1024, 0, st>>>(psum, psq, P, pstride, C, count, gamma, beta, mm, mv, decay,
                                                      eps, bessel, training, scale, shift, save_mean, save_rstd);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int layernorm_joint(__half* a, const __half* b, const float* b_row_scale, __half* u_out, int B, int rows, int D,
                    long long a_stride, long long b_stride, const float* gamma, const float* beta, float eps,
                    __half* y, long long y_stride, float* partial, float* save_mean_rstd, cudaStream_t st) {
  if (u_out == nullptr) u_out = a;
  LPM_REQUIRE(D % 8 == 0 && a_stride % 8 == 0 && b_stride % 8 == 0 && y_stride % 8 == 0,
              "layernorm_joint: D and sample strides must be multiples of 8");
  dim3 g1(LN_CHUNKS, B);
  ln_stats_kernel<<<g1, 256, 0, st>>>(a, b, b_row_scale, u_out, rows, D, a_stride, b_stride, partial);
  LPM_CUDA_CHECK(cudaGetLastError());
  int chunks = (num_sms() * 8 + B - 1) / B;
  const long long n8 = (long long)rows * D / 8;
  if (chunks > (n8 + 255) / 256) chunks = (int)((n8 + 255) / 256);
  if (chunks < 1) chunks = 1;
  dim3 g2(chunks, B);
  ln_apply_kernel<<<g2, 256, 0, st>>>(u_out, rows, D, a_stride, partial, gamma, beta, eps, y, y_stride, save_mean_rstd);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int splitk_reduce(const float* part, int splits, long long split_stride, const float* part2, int splits2, long long split_stride2,
                  long long n, int cols, const float* bias, int relu, float alpha, int accumulate, float* out32, __half* out16,
                  int split3, cudaStream_t st) {
  if (splits + splits2 >= 32 && n <= (1 << 20)) {
    splitk_reduce_wide_kernel<<<(unsigned)((n + 31) / 32), 256, 0, st>>>(part, splits, split_stride, part2, splits2, split_stride2, n,
                                                                        cols, bias, relu, alpha, accumulate, out32, out16, split3);
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  splitk_reduce_kernel<<<grid_for(n, 256), 256, 0, st>>>(part, splits, split_stride, part2, splits2, split_stride2, n, cols, bias,
                                                         relu, alpha, accumulate, out32, out16, split3);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int gating_fwd(const float* act, const float* g, int g_splits, long long g_split_stride, float* g_sum, int split3, int B, int H,
               const float* wg_diag, const float* gamma, const float* beta, float* mm, float* mv, float decay, float eps,
               int training, float* out32, __half* out16, float* save_mean, float* save_rstd, cudaStream_t st) {
  gating_fwd_kernel<<<(H + 31) / 32, 1024, 0, st>>>(act, g, g_splits, g_split_stride, g_sum, split3, B, H, wg_diag, gamma, beta, mm,
                                                  mv, decay, eps, training, out32, out16, save_mean, save_rstd);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int moe_mix(const float* logits, long long ld, int B, int V, int M, int expert_off, float* pred, cudaStream_t st) {
  moe_mix_kernel<<<grid_for((long long)B * V, 256), 256, 0, st>>>(logits, ld, B, V, M, expert_off, pred);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int xent_loss(const float* pred, const uint8_t* labels, int B, int V, float* row_loss, float* loss, cudaStream_t st) {
  xent_rows_kernel<<<B, 256, 0, st>>>(pred, labels, V, row_loss);
  mean_kernel<<<1, 32, 0, st>>>(row_loss, B, loss);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int cast_2d(const float* src, long long ld_src, int rows, int cols, __half* dst, long long ld_dst, int cols_dst,
            cudaStream_t st) {
  const bool vec = cols == cols_dst && cols % 8 == 0 && ld_src % 4 == 0 && ld_dst % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  if (vec)
    cast_2d_vec_kernel<<<grid_for((long long)rows * cols / 8, 256), 256, 0, st>>>(src, ld_src, rows, cols, dst, ld_dst);
  else
    cast_2d_kernel<<<grid_for((long long)rows * cols_dst, 256), 256, 0, st>>>(src, ld_src, rows, cols, dst, ld_dst, cols_dst);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int scale_rows(const __half* x, const float* rs, long long rows, int D, __half* y, cudaStream_t st) {
  LPM_REQUIRE(D % 8 == 0, "scale_rows: D must be a multiple of 8");
  scale_rows_kernel<<<grid_for(rows * D / 8, 256), 256, 0, st>>>(x, rs, rows, D, y);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

// Split-precision operands for the head (context gating + MoE; frame_level_models.py:2342-2368, video_level_models.py:86-126):
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi), i.e. ~22 significant bits in two fp16 terms.  The product of two such
// operands, A W ~= A_hi W_hi + A_lo W_hi + A_hi W_lo, is ONE tcgen05 GEMM over a 3x longer reduction:
//   activations  [rows, cols] fp32 -> [rows, 3*cols] fp16 = [ hi | lo | hi ]             (mode 0: along the columns)
//   weights      [rows, cols] fp32 -> rows [rows, 2*rows) = hi, [2*rows, 3*rows) = lo     (mode 1: along the rows; rows
//                [0, rows) of the destination, the plain fp16 copy, are maintained by the optimiser's shadow refresh and
//                are rewritten here too so that the three blocks are always consistent)
// These few small products decide the predictions: with fp16 operands their error dominates the sigmoid outputs of a
// trained model (numerics), with split operands the head is accurate to fp32 level at ~10 us per forward.
__global__ void __launch_bounds__(256) split_hi_lo_kernel(const float* __restrict__ src, long long ld_src, int rows, int cols,
                                                          __half* __restrict__ dst, long long ld_dst, int mode) {
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float x = src[(long long)r * ld_src + c];
    const __half hi = __float2half_rn(x);
    const __half lo = __float2half_rn(x - __half2float(hi));
    if (mode == 0) {
      __half* d = dst + (long long)r * ld_dst + c;
      d[0] = hi; d[cols] = lo; d[2 * cols] = hi;
    } else if (mode == 1) {
      __half* d = dst + (long long)r * ld_dst + c;
      d[0] = hi; d[(long long)rows * ld_dst] = hi; d[2ll * rows * ld_dst] = lo;
    } else {
      dst[(long long)r * ld_dst + c] = lo;
    }
  }
}

// cols % 8 == 0, 16-byte aligned rows: 8 elements per thread, 32-byte loads and 16-byte stores
__global__ void __launch_bounds__(256) split_hi_lo_vec_kernel(const float* __restrict__ src, long long ld_src, int rows, int cols,
                                                              __half* __restrict__ dst, long long ld_dst, int mode) {
  const int c8 = cols / 8;
  const long long n = (long long)rows * c8;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / c8), c = (int)(i - (long long)r * c8) * 8;
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(src + (long long)r * ld_src + c));
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(src + (long long)r * ld_src + c + 4));
    const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half h0 = __float2half_rn(x[2 * j]), h1 = __float2half_rn(x[2 * j + 1]);
      hi[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      lo[j] = pack_half2(x[2 * j] - __half2float(h0), x[2 * j + 1] - __half2float(h1));
    }
    const uint4 vh = make_uint4(hi[0], hi[1], hi[2], hi[3]), vl = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    __half* d = dst + (long long)r * ld_dst + c;
    if (mode == 0) {
      *reinterpret_cast<uint4*>(d) = vh; *reinterpret_cast<uint4*>(d + cols) = vl; *reinterpret_cast<uint4*>(d + 2 * cols) = vh;
    } else if (mode == 1) {
      *reinterpret_cast<uint4*>(d) = vh; *reinterpret_cast<uint4*>(d + (long long)rows * ld_dst) = vh;
      *reinterpret_cast<uint4*>(d + 2ll * rows * ld_dst) = vl;
    } else {
      *reinterpret_cast<uint4*>(d) = vl;
    }
  }
}

// cols % 2 == 0 (the MoE weights: 3 * 3862 and 2 * 3862 columns): two elements per thread, 8-byte loads, 4-byte stores
__global__ void __launch_bounds__(256) split_hi_lo_pair_kernel(const float* __restrict__ src, long long ld_src, int rows, int cols,
                                                               __half* __restrict__ dst, long long ld_dst, int mode) {
  const int c2 = cols / 2;
  const long long n = (long long)rows * c2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / c2), c = (int)(i - (long long)r * c2) * 2;
    const float2 x = __ldg(reinterpret_cast<const float2*>(src + (long long)r * ld_src + c));
    const __half h0 = __float2half_rn(x.x), h1 = __float2half_rn(x.y);
    const uint32_t vh = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    const uint32_t vl = pack_half2(x.x - __half2float(h0), x.y - __half2float(h1));
    __half* d = dst + (long long)r * ld_dst + c;
    if (mode == 0) {
      *reinterpret_cast<uint32_t*>(d) = vh; *reinterpret_cast<uint32_t*>(d + cols) = vl; *reinterpret_cast<uint32_t*>(d + 2 * cols) = vh;
    } else if (mode == 1) {
      *reinterpret_cast<uint32_t*>(d) = vh; *reinterpret_cast<uint32_t*>(d + (long long)rows * ld_dst) = vh;
      *reinterpret_cast<uint32_t*>(d + 2ll * rows * ld_dst) = vl;
    } else {
      *reinterpret_cast<uint32_t*>(d) = vl;
    }
  }
}

int split_hi_lo(const float* src, long long ld_src, int rows, int cols, __half* dst, long long ld_dst, int mode, cudaStream_t st) {
  const long long n = (long long)rows * cols;
  if (cols % 8 != 0 && cols % 2 == 0 && ld_src % 2 == 0 && ld_dst % 2 == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
    long long blocks = (n / 2 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_hi_lo_pair_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, ld_src, rows, cols, dst, ld_dst, mode);
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  if (cols % 8 == 0 && ld_src % 4 == 0 && ld_dst % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    long long blocks = (n / 8 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_hi_lo_vec_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, ld_src, rows, cols, dst, ld_dst, mode);
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_hi_lo_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, ld_src, rows, cols, dst, ld_dst, mode);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int transpose_2d(const float* src, int rows, int cols, float* dst, __half* dst16, cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_2d_kernel<<<grid, block, 0, st>>>(src, rows, cols, dst, dst16);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int vlad_finalize(const __half* z, const float* rscale, int B, int K, int D, int d_major, float* out, cudaStream_t st) {
  vlad_finalize_kernel<<<grid_for((long long)B * K * D, 256), 256, 0, st>>>(z, rscale, B, K, D, d_major, out);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
