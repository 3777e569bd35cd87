// Backward kernels of the HBM-bound stages (autodiff of the reference lines each forward kernel cites):
// cross-entropy / MoE mixing, context gating, joint layer norm, NetVLAD normalisation, soft-assignment
// softmax + cluster batch norm, cluster-centre and input-batch-norm parameter gradients, column sums.
// Activation gradients travel as fp16 scaled by `loss_scale`; parameter gradients leave in fp32, unscaled.
#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

static inline int grid_for_b(long long n, int threads, int per_sm = 8) {
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)num_sms() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------
// losses.py:44-51 backward: dL/dpred = -(y/(p+eps) - (1-y)/(1-p+eps)) * gscale   (gscale = dL/B)
// ------------------------------------------------------------------------------------------------
__global__ void xent_bwd_kernel(const float* __restrict__ pred, const uint8_t* __restrict__ labels, long long n,
                                float gscale, const float* __restrict__ upstream, float* __restrict__ dpred) {
  if (upstream != nullptr) gscale *= *upstream;     // dLoss of the autograd edge, left on the device (no host sync)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float p = pred[i];
    dpred[i] = labels[i] ? -gscale / (p + 1e-5f) : gscale / (1.f - p + 1e-5f);
  }
}

// ------------------------------------------------------------------------------------------------
// video_level_models.py:116-126 backward -> fp16 logit gradients (x loss_scale), padding columns zeroed
// ------------------------------------------------------------------------------------------------
__global__ void moe_mix_bwd_kernel(const float* __restrict__ logits, long long ld, int B, int V, int M,
                                   int expert_off, const float* __restrict__ dpred, float loss_scale,
                                   __half* __restrict__ dl, long long ldo, int ncols) {
  const long long n = (long long)B * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / V), v = (int)(i - (long long)b * V);
    const float* gl = logits + b * ld + (long long)v * (M + 1);
    const float* el = logits + b * ld + expert_off + (long long)v * M;
    __half* dgl = dl + b * ldo + (long long)v * (M + 1);
    __half* del = dl + b * ldo + expert_off + (long long)v * M;
    float mx = gl[0];
    for (int m = 1; m <= M; ++m) mx = fmaxf(mx, gl[m]);
    float den = 0.f;
    for (int m = 0; m <= M; ++m) den += __expf(gl[m] - mx);
    float p = 0.f;
    for (int m = 0; m < M; ++m) p += (__expf(gl[m] - mx) / den) / (1.f + __expf(-el[m]));
    const float dp = dpred[i] * loss_scale;
    for (int m = 0; m <= M; ++m) {
      const float g = __expf(gl[m] - mx) / den;
      const float e = m < M ? 1.f / (1.f + __expf(-el[m])) : 0.f;
      dgl[m] = __float2half_rn(g * (e - p) * dp);
      if (m < M) del[m] = __float2half_rn(g * e * (1.f - e) * dp);
    }
    if (v == 0) {
      for (int c = V * (M + 1); c < expert_off; ++c) dl[b * ldo + c] = __float2half_rn(0.f);
      for (int c = expert_off + V * M; c < ncols; ++c) dl[b * ldo + c] = __float2half_rn(0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// column sums: out[c] (+)= alpha * sum_r x[r][c]      (bias gradients)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ x, long long ld, long long rows,
                                                             int cols, float* __restrict__ partial) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += to_f<T>(x[r * ld + c]);
  partial[(size_t)blockIdx.y * cols + c] = s;
}
// 32 columns x 32 partial-row lanes per block: the partial rows are summed in parallel (four independent loads in
// flight per thread: these reductions are pure load latency), then across lanes
__global__ void __launch_bounds__(1024) colsum_final_kernel(const float* __restrict__ partial, int chunks,
                                                            long long pstride, int cols, float alpha, int accumulate,
                                                            float* __restrict__ out) {
  __shared__ double red[32][33];
  const int cx = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double s = 0.0;
  if (c < cols) {
    int k = ky;
    for (; k + 96 < chunks; k += 128) {
      const float v0 = partial[(size_t)k * pstride + c], v1 = partial[(size_t)(k + 32) * pstride + c];
      const float v2 = partial[(size_t)(k + 64) * pstride + c], v3 = partial[(size_t)(k + 96) * pstride + c];
      s += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
    }
    for (; k < chunks; k += 32) s += partial[(size_t)k * pstride + c];
  }
  red[ky][cx] = s;
  __syncthreads();
  if (ky == 0 && c < cols) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[k][cx];
    const float r = (float)(t * alpha);
    out[c] = accumulate ? out[c] + r : r;
  }
}

// ------------------------------------------------------------------------------------------------
// Context gating backward (frame_level_models.py:2342-2368): out = act * sigmoid(BN_batch(g))
//   dact (direct path, fp32, still scaled), dg (fp16, scaled), dgamma / dbeta (unscaled)
// 32 hidden units x 32 batch lanes per block.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) gating_bwd_kernel(const float* __restrict__ act, const float* __restrict__ g, int B, int H,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                  const float* __restrict__ dout, float inv_scale, float* __restrict__ dact,
                                  __half* __restrict__ dg, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                  const float* __restrict__ wg_diag, float* __restrict__ ddiag) {
  // wg_diag != null (--gating_remove_diag, :2349-2352): the batch norm saw g - diag*act; then additionally
  //   dact -= diag * dv   and   ddiag[c] = -sum_b dv[b,c] * act[b,c]   (dv = gradient at the batch-norm input)
  __shared__ double red[2][32][33];
  __shared__ float sm[2][32];
  const int cx = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool ok = c < H;
  const float mu = ok ? mean[c] : 0.f, rs = ok ? rstd[c] : 0.f, ga = ok ? gamma[c] : 0.f, be = ok ? beta[c] : 0.f;
  const float dgc = (wg_diag && ok) ? wg_diag[c] : 0.f;
  double s1 = 0.0, s2 = 0.0;
  if (ok)
    for (int b = ky; b < B; b += 32) {
      const float xh = (g[(size_t)b * H + c] - dgc * act[(size_t)b * H + c] - mu) * rs;
      const float sg = 1.f / (1.f + __expf(-(xh * ga + be)));
      const float a = act[(size_t)b * H + c], d = dout[(size_t)b * H + c];
      const float dv = d * a * sg * (1.f - sg);
      dact[(size_t)b * H + c] = d * sg;
      s1 += dv;
      s2 += (double)dv * xh;
    }
  red[0][ky][cx] = s1; red[1][ky][cx] = s2;
  __syncthreads();
  if (ky == 0 && ok) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) { t1 += red[0][k][cx]; t2 += red[1][k][cx]; }
    dgamma[c] = (float)t2 * inv_scale;
    dbeta[c] = (float)t1 * inv_scale;
    sm[0][cx] = (float)(t1 / B); sm[1][cx] = (float)(t2 / B);
  }
  __syncthreads();
  const float m1 = sm[0][cx], m2 = sm[1][cx];
  double sd = 0.0;
  if (ok)
    for (int b = ky; b < B; b += 32) {
      const float a = act[(size_t)b * H + c];
      const float xh = (g[(size_t)b * H + c] - dgc * a - mu) * rs;
      const float sg = 1.f / (1.f + __expf(-(xh * ga + be)));
      const float dv = dout[(size_t)b * H + c] * a * sg * (1.f - sg);
      const float gv = ga * rs * (dv - m1 - xh * m2);
      dg[(size_t)b * H + c] = __float2half_rn(gv);
      if (wg_diag) {
        dact[(size_t)b * H + c] -= dgc * gv;       // written by this same thread in the first loop
        sd += (double)gv * a;
      }
    }
  if (ddiag == nullptr) return;
  __syncthreads();
  red[0][ky][cx] = sd;
  __syncthreads();
  if (ky == 0 && ok) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[0][k][cx];
    ddiag[c] = -(float)t * inv_scale;
  }
}

// ------------------------------------------------------------------------------------------------
// Joint layer norm backward.  Thread owns 8 fixed columns; a CTA walks a row range of one sample.
//   pass 1: per-sample c1 = sum dy*gamma, c2 = sum dy*gamma*xhat ; per-column sum dy*xhat, sum dy
//   pass 2: du = rstd*(dy*gamma - c1/N - xhat*c2/N) ; optional du_masked = du o (mask>0) (ReLU backward of the
//           pre-residual branch) ; optional column sums of du_masked (or du when there is no mask)
// ------------------------------------------------------------------------------------------------
// 7 row chunks per sample: 560 CTAs at batch 80 = one resident wave of 4 CTAs per SM on 148 SMs (4 chunks = 320 CTAs
// ran as 2.2 waves with a 16 %-full tail)
constexpr int LNB_CHUNKS = 7;

__device__ __forceinline__ void load8(const __half* p, float* f) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ void load8_regs(const uint4& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ void store8(__half* p, const float* f) {
  uint4 v;
  v.x = pack_half2(f[0], f[1]); v.y = pack_half2(f[2], f[3]); v.z = pack_half2(f[4], f[5]); v.w = pack_half2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = v;
}

__global__ void __launch_bounds__(256) ln_bwd1_kernel(const __half* __restrict__ u, const __half* __restrict__ dy,
                                                      long long dy_stride, int rows, int D,
                                                      const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                                      float* __restrict__ part_sample, float* __restrict__ part_cols) {
  extern __shared__ float sh[];  // [rpi][2][D] column partials + 64 reduction floats
  const int sample = blockIdx.y, chunk = blockIdx.x;
  const int tpr = D / 8, rpi = 256 / tpr;
  const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
  const int per = (rows + LNB_CHUNKS - 1) / LNB_CHUNKS;
  const int r0 = chunk * per, r1 = min(rows, r0 + per);
  const float mu = mean_rstd[sample * 2], rs = mean_rstd[sample * 2 + 1];
  float ga[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ga[j] = gamma[tc * 8 + j];
  float c1 = 0.f, c2 = 0.f, gacc[8], bacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gacc[j] = bacc[j] = 0.f;
  const __half* us = u + (size_t)sample * rows * D;
  const __half* ds = dy + (size_t)sample * dy_stride;
#pragma unroll 2
  for (int r = r0 + tr; r < r1; r += rpi) {
    float uu[8], dd[8];
    load8(us + (size_t)r * D + tc * 8, uu);
    load8(ds + (size_t)r * D + tc * 8, dd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (uu[j] - mu) * rs;
      c1 += dd[j] * ga[j];
      c2 += dd[j] * ga[j] * xh;
      gacc[j] += dd[j] * xh;
      bacc[j] += dd[j];
    }
  }
  float* red = sh + rpi * 2 * D;
  c1 = warp_sum(c1); c2 = warp_sum(c2);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[w] = c1; red[8 + w] = c2; }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[(tr * 2 + 0) * D + tc * 8 + j] = gacc[j];
    sh[(tr * 2 + 1) * D + tc * 8 + j] = bacc[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[i]; b += red[8 + i]; }
    part_sample[((size_t)sample * LNB_CHUNKS + chunk) * 2 + 0] = a;
    part_sample[((size_t)sample * LNB_CHUNKS + chunk) * 2 + 1] = b;
  }
  float* pc = part_cols + ((size_t)sample * LNB_CHUNKS + chunk) * 2 * D;
  for (int i = threadIdx.x; i < 2 * D; i += 256) {
    float s = 0.f;
    for (int k = 0; k < rpi; ++k) s += sh[k * 2 * D + i];
    pc[i] = s;
  }
}

__global__ void __launch_bounds__(256) ln_bwd2_kernel(const __half* __restrict__ u, const __half* __restrict__ dy,
                                                      long long dy_stride, int rows, int D,
                                                      const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                                      const float* __restrict__ part_sample, const __half* __restrict__ mask,
                                                      __half* __restrict__ du, __half* __restrict__ du_masked,
                                                      float* __restrict__ part_cols) {
  extern __shared__ float sh[];  // [rpi][D] column partials (only when part_cols != null)
  const int sample = blockIdx.y, chunk = blockIdx.x;
  const int tpr = D / 8, rpi = 256 / tpr;
  const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
  const int per = (rows + LNB_CHUNKS - 1) / LNB_CHUNKS;
  const int r0 = chunk * per, r1 = min(rows, r0 + per);
  const float mu = mean_rstd[sample * 2], rs = mean_rstd[sample * 2 + 1];
  float c1 = 0.f, c2 = 0.f;
  for (int k = 0; k < LNB_CHUNKS; ++k) {
    c1 += part_sample[((size_t)sample * LNB_CHUNKS + k) * 2 + 0];
    c2 += part_sample[((size_t)sample * LNB_CHUNKS + k) * 2 + 1];
  }
  const float invn = 1.f / ((float)rows * (float)D);
  c1 *= invn; c2 *= invn;
  float ga[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ga[j] = gamma[tc * 8 + j]; acc[j] = 0.f; }
  const size_t sbase = (size_t)sample * rows * D;
  const __half* ds = dy + (size_t)sample * dy_stride;
#pragma unroll 2
  for (int r = r0 + tr; r < r1; r += rpi) {
    float uu[8], dd[8], mm[8], o[8];
    const size_t off = (size_t)r * D + tc * 8;
    load8(u + sbase + off, uu);
    load8(ds + off, dd);
    if (mask) load8(mask + sbase + off, mm);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (uu[j] - mu) * rs;
      o[j] = rs * (dd[j] * ga[j] - c1 - xh * c2);
    }
    store8(du + sbase + off, o);
    if (mask) {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (!(mm[j] > 0.f)) o[j] = 0.f;
      store8(du_masked + sbase + off, o);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += o[j];
  }
  if (part_cols != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[tr * D + tc * 8 + j] = acc[j];
    __syncthreads();
    float* pc = part_cols + ((size_t)sample * LNB_CHUNKS + chunk) * D;
    for (int i = threadIdx.x; i < D; i += 256) {
      float s = 0.f;
      for (int k = 0; k < rpi; ++k) s += sh[k * D + i];
      pc[i] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// NetVLAD normalisation backward (frame_level_models.py:2819-2822): vhat_k = z_k * rs_k, rs_k = 1/(|z_k| sqrt(K)).
//   dz_k = rs_k * (dvhat_k - z_k (z_k . dvhat_k)/|z_k|^2)   (the global-norm Jacobian vanishes: every
//   intra-normalised row has unit norm, so the global norm is the constant sqrt(K))
//   q[row] = dz_k . C[:,k]   (needed by the soft-assignment backward)
// One warp per (video, cluster) row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vlad_norm_bwd_kernel(const __half* __restrict__ z, const float* __restrict__ rs,
                                                            const __half* __restrict__ dvh, long long rows, int K, int D,
                                                            const float* __restrict__ centers_t, __half* __restrict__ dz,
                                                            float* __restrict__ q) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const __half* zr = z + r * D;
    const __half* dr = dvh + r * D;
    float n2 = 0.f, dot = 0.f;
    for (int i = lane * 8; i < D; i += 256) {
      float a[8], b[8];
      load8(zr + i, a); load8(dr + i, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) { n2 += a[j] * a[j]; dot += a[j] * b[j]; }
    }
    n2 = warp_sum(n2); dot = warp_sum(dot);
    const float sc = rs[r], proj = dot / fmaxf(n2, 1e-12f);
    const float* cr = centers_t + (size_t)(r % K) * D;
    float qq = 0.f;
    for (int i = lane * 8; i < D; i += 256) {
      float a[8], b[8], o[8];
      load8(zr + i, a); load8(dr + i, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j] = sc * (b[j] - a[j] * proj); qq += o[j] * cr[i + j]; }
      store8(dz + r * D + i, o);
    }
    qq = warp_sum(qq);
    if (lane == 0) q[r] = qq;
  }
}

// ------------------------------------------------------------------------------------------------
// Soft-assignment backward (softmax + cluster_bn), frame_level_models.py:2781-2803 autodiff.
//   pass 1 (warp per frame row): dA = G - q[b,:]; dShat = A*(dA - sum_k A dA); column partials of
//          dShat and dShat*shat (shat = (S-mean)*rstd)          G = X dV^T (fp32), S = X Wc (fp16)
//   pass 2: dS = gamma*rstd*(dShat - c1/N - shat*c2/N)  (training batch norm), in place
// ------------------------------------------------------------------------------------------------
template <int J>   // clusters per lane: K <= 32*J
__global__ void __launch_bounds__(256) assign_bwd1_kernel(const float* __restrict__ G, const __half* __restrict__ A,
                                                          const float* __restrict__ q, const __half* __restrict__ S,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          long long rows, int T, int K, __half* __restrict__ dsh,
                                                          float* __restrict__ partial) {
  extern __shared__ float sh[];  // [8 warps][2][K]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * 8 + w, nwarps = (long long)gridDim.x * 8;
  float c1[J], c2[J];
#pragma unroll
  for (int j = 0; j < J; ++j) c1[j] = c2[j] = 0.f;
  for (long long r = warp; r < rows; r += nwarps) {
    const long long b = r / T;
    float a[J], da[J];
    float inner = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int k = lane + 32 * j;
      if (k < K) {
        a[j] = __half2float(A[r * K + k]);
        da[j] = G[r * K + k] - q[b * K + k];
        inner += a[j] * da[j];
      }
    }
    inner = warp_sum(inner);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int k = lane + 32 * j;
      if (k < K) {
        const float d = a[j] * (da[j] - inner);
        const float shat = (__half2float(S[r * K + k]) - mean[k]) * rstd[k];
        dsh[r * K + k] = __float2half_rn(d);
        c1[j] += d;
        c2[j] += d * shat;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int k = lane + 32 * j;
    if (k < K) { sh[(w * 2 + 0) * K + k] = c1[j]; sh[(w * 2 + 1) * K + k] = c2[j]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * K; i += 256) {
    float s = 0.f;
    for (int ww = 0; ww < 8; ++ww) s += sh[ww * 2 * K + i];
    partial[(size_t)blockIdx.x * 2 * K + i] = s;
  }
}

// K == 256 (the rgb pool of config 1): a lane owns 8 CONSECUTIVE clusters, so A / S / dShat move as 16-byte and G as two
// 16-byte accesses per lane (the strided version above issues 2-byte loads: 64 bytes per warp request).  Same arithmetic,
// same partial layout.
__global__ void __launch_bounds__(256) assign_bwd1_k256_kernel(const float* __restrict__ G, const __half* __restrict__ A,
                                                               const float* __restrict__ q, const __half* __restrict__ S,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               long long rows, int T, __half* __restrict__ dsh,
                                                               float* __restrict__ partial) {
  constexpr int K = 256;
  extern __shared__ float sh[];  // [8 warps][2][K]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * 8 + w, nwarps = (long long)gridDim.x * 8;
  const int k0 = lane * 8;
  float mu[8], rs[8], c1[8], c2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { mu[j] = mean[k0 + j]; rs[j] = rstd[k0 + j]; c1[j] = c2[j] = 0.f; }
#pragma unroll 2
  for (long long r = warp; r < rows; r += nwarps) {
    const long long b = r / T;
    const uint4 va = __ldg(reinterpret_cast<const uint4*>(A + r * K + k0));
    const uint4 vs = __ldg(reinterpret_cast<const uint4*>(S + r * K + k0));
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(G + r * K + k0)), g1 = __ldg(reinterpret_cast<const float4*>(G + r * K + k0 + 4));
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(q + b * K + k0)), q1 = __ldg(reinterpret_cast<const float4*>(q + b * K + k0 + 4));
    float a[8], sv[8], da[8];
    load8_regs(va, a); load8_regs(vs, sv);
    da[0] = g0.x - q0.x; da[1] = g0.y - q0.y; da[2] = g0.z - q0.z; da[3] = g0.w - q0.w;
    da[4] = g1.x - q1.x; da[5] = g1.y - q1.y; da[6] = g1.z - q1.z; da[7] = g1.w - q1.w;
    float inner = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) inner += a[j] * da[j];
    inner = warp_sum(inner);
    float d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d[j] = a[j] * (da[j] - inner);
      const float shat = (sv[j] - mu[j]) * rs[j];
      c1[j] += d[j];
      c2[j] += d[j] * shat;
    }
    store8(dsh + r * K + k0, d);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sh[(w * 2 + 0) * K + k0 + j] = c1[j]; sh[(w * 2 + 1) * K + k0 + j] = c2[j]; }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * K; i += 256) {
    float s = 0.f;
    for (int ww = 0; ww < 8; ++ww) s += sh[ww * 2 * K + i];
    partial[(size_t)blockIdx.x * 2 * K + i] = s;
  }
}

__global__ void __launch_bounds__(256) assign_bwd2_kernel(__half* __restrict__ dsh, const __half* __restrict__ S,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ gamma, const float* __restrict__ csum,
                                                          long long rows, int K, float inv_n) {
  // csum: [2][K] totals of dShat and dShat*shat (still loss-scaled)
  const long long n = rows * K;
  if (K % 8 == 0 && K <= 512 && ((reinterpret_cast<uintptr_t>(dsh) | reinterpret_cast<uintptr_t>(S)) & 15) == 0) {
    // eight clusters per thread, 16-byte accesses; the five per-cluster coefficients are folded into three and kept in
    // shared memory (the scalar loop below re-reads five global arrays per element, which saturates L1):
    //   dS = a_k * (d - b_k) - c_k * (s - mean_k)   with a = gamma*rstd, b = c1/N, c = a*rstd*c2/N
    __shared__ __align__(16) float co[4][512];
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const float a = gamma[k] * rstd[k];
      co[0][k] = a;
      co[1][k] = csum[k] * inv_n;
      co[2][k] = rstd[k] * csum[K + k] * inv_n;
      co[3][k] = mean[k];
    }
    __syncthreads();
    for (long long i8 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i8 < n / 8; i8 += (long long)gridDim.x * blockDim.x) {
      const int k0 = (int)((i8 * 8) % K);
      float sv[8], dv[8], o[8];
      load8(S + i8 * 8, sv);
      load8_regs(*reinterpret_cast<const uint4*>(dsh + i8 * 8), dv);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 a4 = *reinterpret_cast<const float4*>(&co[0][k0 + 4 * h]), b4 = *reinterpret_cast<const float4*>(&co[1][k0 + 4 * h]);
        const float4 c4 = *reinterpret_cast<const float4*>(&co[2][k0 + 4 * h]), m4 = *reinterpret_cast<const float4*>(&co[3][k0 + 4 * h]);
        const float aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
        const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // same value as gamma*rstd*(d - c1/N - shat*c2/N), shat = (s - mean)*rstd, regrouped
          const float shat_c = (sv[4 * h + j] - mm[j]) * cc[j];
          o[4 * h + j] = aa[j] * (dv[4 * h + j] - bb[j] - shat_c);
        }
      }
      store8(dsh + i8 * 8, o);
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const float shat = (__half2float(S[i]) - mean[k]) * rstd[k];
    const float d = __half2float(dsh[i]);
    dsh[i] = __float2half_rn(gamma[k] * rstd[k] * (d - csum[k] * inv_n - shat * csum[K + k] * inv_n));
  }
}

// ------------------------------------------------------------------------------------------------
// cluster_weights2 and input_bn parameter gradients from per-(k,d) batch reductions:
//   s1 = sum_b a_sum[b,k] dV[b,k,d] ;  s2 = sum_b dV[b,k,d] Z[b,k,d]
//   dCt[k][d] = -s1 ;  E[k][d] = s2 + s1*(C[d,k] - beta_in[d])       (both unscaled)
// then per feature d:  dgamma_in[d] = (sum_k Wc[d,k] dWc[d,k] + sum_k E[k][d]) / gamma_in[d]
//                      dbeta_in[d]  = -sum_k dCt[k][d]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) center_bwd_kernel(const __half* __restrict__ dV, const __half* __restrict__ Z,
                                                         const float* __restrict__ a_sum, int B, int K, int D,
                                                         const float* __restrict__ centers_t,
                                                         const float* __restrict__ beta_in, float inv_scale,
                                                         float* __restrict__ dCt, float* __restrict__ E) {
  const int k = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float s1 = 0.f, s2 = 0.f;
  for (int b = 0; b < B; ++b) {
    const size_t i = ((size_t)b * K + k) * D + d;
    const float dv = __half2float(dV[i]);
    s1 += a_sum[b * K + k] * dv;
    s2 += dv * __half2float(Z[i]);
  }
  s1 *= inv_scale; s2 *= inv_scale;
  dCt[(size_t)k * D + d] = -s1;
  E[(size_t)k * D + d] = s2 + s1 * (centers_t[(size_t)k * D + d] - beta_in[d]);
}

// D % 8 == 0: a thread owns 8 consecutive features of one cluster (16-byte loads of dV / Z); the 256 threads of a block
// are `cgs` column groups x 256 / cgs video lanes (a lane walks every (256 / cgs)-th video), the lanes' partial sums are
// combined through shared memory in lane order (deterministic).  grid = (ceil(D / 8 / cgs), K).
__global__ void __launch_bounds__(256) center_bwd_vec_kernel(const __half* __restrict__ dV, const __half* __restrict__ Z,
                                                             const float* __restrict__ a_sum, int B, int K, int D, int cgs,
                                                             const float* __restrict__ centers_t,
                                                             const float* __restrict__ beta_in, float inv_scale,
                                                             float* __restrict__ dCt, float* __restrict__ E) {
  __shared__ float sh[2][256 * 8];
  const int lanes = 256 / cgs;
  const int cg = threadIdx.x % cgs, vl = threadIdx.x / cgs;
  const int k = blockIdx.y;
  const int d0 = (blockIdx.x * cgs + cg) * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (d0 < D) {
#pragma unroll 4
    for (int b = vl; b < B; b += lanes) {
      const size_t i = ((size_t)b * K + k) * D + d0;
      float dv[8], zz[8];
      load8(dV + i, dv); load8(Z + i, zz);
      const float as = __ldg(a_sum + b * K + k);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] = fmaf(as, dv[j], s1[j]); s2[j] = fmaf(dv[j], zz[j], s2[j]); }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sh[0][(vl * cgs + cg) * 8 + j] = s1[j]; sh[1][(vl * cgs + cg) * 8 + j] = s2[j]; }
  __syncthreads();
  // cgs * 8 output columns per block, one thread each (cgs * 8 <= 256)
  const int c = threadIdx.x;
  if (c < cgs * 8) {
    const int d = blockIdx.x * cgs * 8 + c;
    if (d < D) {
      float a = 0.f, b2 = 0.f;
      for (int l = 0; l < lanes; ++l) { a += sh[0][l * cgs * 8 + c]; b2 += sh[1][l * cgs * 8 + c]; }
      a *= inv_scale; b2 *= inv_scale;
      dCt[(size_t)k * D + d] = -a;
      E[(size_t)k * D + d] = b2 + a * (centers_t[(size_t)k * D + d] - beta_in[d]);
    }
  }
}

__global__ void __launch_bounds__(256) input_bn_grad_kernel(const float* __restrict__ Wc, const float* __restrict__ dWc,
                                                            const float* __restrict__ dCt, const float* __restrict__ E,
                                                            int D, int K, const float* __restrict__ gamma_in,
                                                            float* __restrict__ dgamma_in, float* __restrict__ dbeta_in) {
  const int d = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (d >= D) return;
  float t = 0.f, sdc = 0.f;
  for (int k = lane; k < K; k += 32) {
    t += Wc[(size_t)d * K + k] * dWc[(size_t)d * K + k] + E[(size_t)k * D + d];
    sdc += dCt[(size_t)k * D + d];
  }
  t = warp_sum(t);
  sdc = warp_sum(sdc);
  if (lane == 0) {
    const float g = gamma_in[d];
    dgamma_in[d] = g != 0.f ? t / g : 0.f;
    dbeta_in[d] = -sdc;
  }
}

// fp32 -> fp16 elementwise (scaled activation gradients entering a GEMM)
__global__ void cast_scaled_kernel(const float* __restrict__ x, long long n, float alpha, __half* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(x[i] * alpha);
}

// ----------------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------------
int xent_bwd(const float* pred, const uint8_t* labels, long long n, float gscale, const float* upstream, float* dpred,
             cudaStream_t st) {
  xent_bwd_kernel<<<grid_for_b(n, 256), 256, 0, st>>>(pred, labels, n, gscale, upstream, dpred);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int moe_mix_bwd(const float* logits, long long ld, int B, int V, int M, int expert_off, const float* dpred,
                float loss_scale, __half* dl, long long ldo, int ncols, cudaStream_t st) {
  moe_mix_bwd_kernel<<<grid_for_b((long long)B * V, 256), 256, 0, st>>>(logits, ld, B, V, M, expert_off, dpred,
                                                                        loss_scale, dl, ldo, ncols);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int colsum_chunks(long long rows) {
  long long c = (rows + 255) / 256;
  if (c > 64) c = 64;
  if (c < 1) c = 1;
  return (int)c;
}

int colsum(const void* x, int is_f32, long long ld, long long rows, int cols, float alpha, int accumulate,
           float* partial, float* out, cudaStream_t st) {
  const int chunks = colsum_chunks(rows);
  dim3 grid((cols + 255) / 256, chunks);
  if (is_f32) colsum_partial_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), ld, rows, cols, partial);
  else colsum_partial_kernel<__half><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(x), ld, rows, cols, partial);
  colsum_final_kernel<<<(cols + 31) / 32, 1024, 0, st>>>(partial, chunks, cols, cols, alpha, accumulate, out);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int colsum_final(const float* partial, int chunks, long long pstride, int cols, float alpha, int accumulate,
                 float* out, cudaStream_t st) {
  colsum_final_kernel<<<(cols + 31) / 32, 1024, 0, st>>>(partial, chunks, pstride, cols, alpha, accumulate, out);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int gating_bwd(const float* act, const float* g, int B, int H, const float* gamma, const float* beta,
               const float* mean, const float* rstd, const float* dout, float inv_scale, float* dact, __half* dg,
               float* dgamma, float* dbeta, const float* wg_diag, float* ddiag, cudaStream_t st) {
  LPM_REQUIRE((wg_diag == nullptr) == (ddiag == nullptr), "gating_bwd: wg_diag and ddiag go together");
  gating_bwd_kernel<<<(H + 31) / 32, 1024, 0, st>>>(act, g, B, H, gamma, beta, mean, rstd, dout, inv_scale, dact, dg,
                                                  dgamma, dbeta, wg_diag, ddiag);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int ln_bwd_chunks() { return LNB_CHUNKS; }

int layernorm_joint_bwd(const __half* u, const __half* dy, long long dy_stride, int B, int rows, int D,
                        const float* mean_rstd, const float* gamma, const __half* mask, __half* du,
                        __half* du_masked, float* part_sample, float* part_cols, float* part_cols_du,
                        cudaStream_t st) {
  LPM_REQUIRE(!mask || du_masked, "layernorm_joint_bwd: mask needs du_masked");
  LPM_REQUIRE(D % 8 == 0 && D / 8 <= 256 && 256 % (D / 8) == 0, "layernorm_joint_bwd: D/8 must divide 256 (D=%d)", D);
  LPM_REQUIRE(dy_stride % 8 == 0, "layernorm_joint_bwd: dy stride must be a multiple of 8");
  const int rpi = 256 / (D / 8);
  dim3 grid(LNB_CHUNKS, B);
  const size_t sm1 = (size_t)(rpi * 2 * D + 64) * sizeof(float);
  const size_t sm2 = (size_t)(rpi * D) * sizeof(float);
  static bool set = false;
  if (!set) {
    LPM_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    LPM_CUDA_CHECK(cudaFuncSetAttribute(ln_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    set = true;
  }
  LPM_REQUIRE(sm1 <= 100 * 1024, "layernorm_joint_bwd: D too large");
  ln_bwd1_kernel<<<grid, 256, sm1, st>>>(u, dy, dy_stride, rows, D, mean_rstd, gamma, part_sample, part_cols);
  ln_bwd2_kernel<<<grid, 256, part_cols_du ? sm2 : 0, st>>>(u, dy, dy_stride, rows, D, mean_rstd, gamma, part_sample,
                                                            mask, du, du_masked, part_cols_du);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int vlad_norm_bwd(const __half* z, const float* rs, const __half* dvh, long long rows, int K, int D,
                  const float* centers_t, __half* dz, float* q, cudaStream_t st) {
  LPM_REQUIRE(D % 8 == 0, "vlad_norm_bwd: D must be a multiple of 8");
  vlad_norm_bwd_kernel<<<grid_for_b(rows * 32, 256), 256, 0, st>>>(z, rs, dvh, rows, K, D, centers_t, dz, q);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int assign_bwd_blocks() { return num_sms() * 2; }

int assign_bwd1(const float* G, const __half* A, const float* q, const __half* S, const float* mean,
                const float* rstd, long long rows, int T, int K, __half* dsh, float* partial, cudaStream_t st) {
  LPM_REQUIRE(K <= 512, "assign_bwd1: K must be <= 512");
  const bool al16 = ((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(S) |
                      reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(dsh)) & 15) == 0;
  if (K == 256 && al16)
    assign_bwd1_k256_kernel<<<assign_bwd_blocks(), 256, (size_t)16 * K * sizeof(float), st>>>(G, A, q, S, mean, rstd, rows, T, dsh,
                                                                                             partial);
  else if (K > 256)
    assign_bwd1_kernel<16><<<assign_bwd_blocks(), 256, (size_t)16 * K * sizeof(float), st>>>(G, A, q, S, mean, rstd, rows, T, K,
                                                                                             dsh, partial);
  else
  assign_bwd1_kernel<8><<<assign_bwd_blocks(), 256, (size_t)16 * K * sizeof(float), st>>>(G, A, q, S, mean, rstd, rows, T, K,
                                                                                        dsh, partial);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int assign_bwd2(__half* dsh, const __half* S, const float* mean, const float* rstd, const float* gamma,
                const float* csum, long long rows, int K, cudaStream_t st) {
  assign_bwd2_kernel<<<grid_for_b(rows * K, 256), 256, 0, st>>>(dsh, S, mean, rstd, gamma, csum, rows, K,
                                                                1.f / (float)rows);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int center_bwd(const __half* dV, const __half* Z, const float* a_sum, int B, int K, int D, const float* centers_t,
               const float* beta_in, float inv_scale, float* dCt, float* E, cudaStream_t st) {
  if (D % 8 == 0 && ((reinterpret_cast<uintptr_t>(dV) | reinterpret_cast<uintptr_t>(Z)) & 15) == 0) {
    int cgs = 32;                                  // column groups per block (power of two <= 32 that covers D / 8)
    while (cgs > 1 && cgs > D / 8) cgs >>= 1;
    dim3 grid((D / 8 + cgs - 1) / cgs, K);
    center_bwd_vec_kernel<<<grid, 256, 0, st>>>(dV, Z, a_sum, B, K, D, cgs, centers_t, beta_in, inv_scale, dCt, E);
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  dim3 grid((D + 255) / 256, K);
  center_bwd_kernel<<<grid, 256, 0, st>>>(dV, Z, a_sum, B, K, D, centers_t, beta_in, inv_scale, dCt, E);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int input_bn_grad(const float* Wc, const float* dWc, const float* dCt, const float* E, int D, int K,
                  const float* gamma_in, float* dgamma_in, float* dbeta_in, cudaStream_t st) {
  input_bn_grad_kernel<<<(D + 7) / 8, 256, 0, st>>>(Wc, dWc, dCt, E, D, K, gamma_in, dgamma_in, dbeta_in);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int cast_scaled(const float* x, long long n, float alpha, __half* y, cudaStream_t st) {
  cast_scaled_kernel<<<grid_for_b(n, 256), 256, 0, st>>>(x, n, alpha, y);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
