// Head variants behind the reference's flags (frame_level_models.py:2319-2340, 2349-2352):
//   --netvlad_relu        : activation = relu6( slim.batch_norm(vlad x hidden1_weights, scope="hidden1_bn") )  (no bias)
//   --gating_remove_diag  : the gradient of diag(gating_weights_2) joins the dense weight gradient (lpm_add_diag)
// [B, H] matrices with B = tower batch (tens to hundreds of rows): 32 hidden units x 32 batch lanes per block.
#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

__global__ void __launch_bounds__(1024) hidden_bn_relu6_fwd_kernel(const float* __restrict__ x, int B, int H,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                                                   float decay, float eps, int training, int relu6,
                                                                   float* __restrict__ out32, __half* __restrict__ out16,
                                                                   float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  __shared__ double red[2][32][33];
  __shared__ float sm[2][32];
  const int cx = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool ok = c < H;
  if (training) {
    double s = 0.0, q = 0.0;
    if (ok)
      for (int b = ky; b < B; b += 32) { const float v = x[(size_t)b * H + c]; s += v; q += (double)v * v; }
    red[0][ky][cx] = s; red[1][ky][cx] = q;
    __syncthreads();
    if (ky == 0 && ok) {
      s = 0.0; q = 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k) { s += red[0][k][cx]; q += red[1][k][cx]; }
      const double m = s / B;
      double vv = q / B - m * m;
      if (vv < 0.0) vv = 0.0;
      const double corr = B > 1 ? (double)B / (B - 1) : 1.0;      // fused rank-2 path: Bessel-corrected moving variance
      moving_mean[c] = moving_mean[c] * decay + (float)m * (1.f - decay);
      moving_var[c] = moving_var[c] * decay + (float)(vv * corr) * (1.f - decay);
      sm[0][cx] = (float)m; sm[1][cx] = (float)vv;
    }
  } else if (ky == 0 && ok) {
    sm[0][cx] = moving_mean[c]; sm[1][cx] = moving_var[c];
  }
  __syncthreads();
  if (!ok) return;
  const float mean = sm[0][cx], rstd = rsqrtf(sm[1][cx] + eps);
  if (save_mean && ky == 0) { save_mean[c] = mean; save_rstd[c] = rstd; }
  const float sc = gamma[c] * rstd, sh = beta[c] - mean * sc;
  for (int b = ky; b < B; b += 32) {
    float v = fmaf(x[(size_t)b * H + c], sc, sh);
    if (relu6) v = fminf(fmaxf(v, 0.f), 6.f);
    out32[(size_t)b * H + c] = v;
    if (out16) out16[(size_t)b * H + c] = __float2half_rn(v);
  }
}

// dy: gradient at the relu6 output (fp32, loss-scaled); y: the forward output (the relu6 mask is 0 < y < 6);
// dx (in place over dy allowed) = gamma*rstd*(dv - mean(dv) - xhat*mean(dv*xhat)), dv = dy o mask; dgamma/dbeta unscaled
__global__ void __launch_bounds__(1024) hidden_bn_relu6_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                   const float* __restrict__ dy, int B, int H,
                                                                   const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, int relu6, float inv_scale,
                                                                   float* __restrict__ dx, float* __restrict__ dgamma,
                                                                   float* __restrict__ dbeta) {
  __shared__ double red[2][32][33];
  __shared__ float sm[2][32];
  const int cx = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool ok = c < H;
  const float mu = ok ? mean[c] : 0.f, rs = ok ? rstd[c] : 0.f, ga = ok ? gamma[c] : 0.f;
  double s1 = 0.0, s2 = 0.0;
  if (ok)
    for (int b = ky; b < B; b += 32) {
      const size_t o = (size_t)b * H + c;
      const float yy = y[o];
      const float dv = (!relu6 || (yy > 0.f && yy < 6.f)) ? dy[o] : 0.f;
      s1 += dv;
      s2 += (double)dv * ((x[o] - mu) * rs);
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
  if (!ok) return;
  const float m1 = sm[0][cx], m2 = sm[1][cx];
  for (int b = ky; b < B; b += 32) {
    const size_t o = (size_t)b * H + c;
    const float yy = y[o];
    const float dv = (!relu6 || (yy > 0.f && yy < 6.f)) ? dy[o] : 0.f;
    dx[o] = ga * rs * (dv - m1 - (x[o] - mu) * rs * m2);
  }
}

__global__ void add_diag_kernel(float* __restrict__ m, int n, long long ld, const float* __restrict__ d, float alpha) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) m[(size_t)i * ld + i] += alpha * d[i];
}

// caller prelude (train.py:262-264, eval.py:140-143, export_model.py:91-92): tf.nn.l2_normalize(model_input, 2) on fp32
// frames; one warp per frame row, two passes over the row (the second one hits L1/L2).  y may alias x.
__global__ void __launch_bounds__(256) l2_normalize_rows_kernel(const float* __restrict__ x, long long rows, int F,
                                                                float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * F);
    float ss = 0.f;
    for (int i = lane; i < F / 4; i += 32) {
      const float4 v = xr[i];
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    const float rn = rsqrtf(fmaxf(ss, 1e-12f));
    float4* yr = reinterpret_cast<float4*>(y + r * F);
    for (int i = lane; i < F / 4; i += 32) {
      float4 v = xr[i];
      v.x *= rn; v.y *= rn; v.z *= rn; v.w *= rn;
      yr[i] = v;
    }
  }
}

int l2_normalize_rows(const float* x, long long rows, int F, float* y, cudaStream_t st) {
  LPM_REQUIRE(F % 4 == 0 && F > 0, "l2_normalize_rows: feature size must be a multiple of 4 (got %d)", F);
  long long blocks = (rows + 7) / 8;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  l2_normalize_rows_kernel<<<(int)blocks, 256, 0, st>>>(x, rows, F, y);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int hidden_bn_relu6_fwd(const float* x, int B, int H, const float* gamma, const float* beta, float* mm, float* mv,
                        float decay, float eps, int training, int relu6, float* out32, __half* out16, float* save_mean,
                        float* save_rstd, cudaStream_t st) {
  hidden_bn_relu6_fwd_kernel<<<(H + 31) / 32, 1024, 0, st>>>(x, B, H, gamma, beta, mm, mv, decay, eps, training, relu6, out32,
                                                            out16, save_mean, save_rstd);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int hidden_bn_relu6_bwd(const float* x, const float* y, const float* dy, int B, int H, const float* gamma, const float* mean,
                        const float* rstd, int relu6, float inv_scale, float* dx, float* dgamma, float* dbeta, cudaStream_t st) {
  hidden_bn_relu6_bwd_kernel<<<(H + 31) / 32, 1024, 0, st>>>(x, y, dy, B, H, gamma, mean, rstd, relu6, inv_scale, dx, dgamma, dbeta);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int add_diag(float* m, int n, long long ld, const float* d, float alpha, cudaStream_t st) {
  add_diag_kernel<<<(n + 255) / 256, 256, 0, st>>>(m, n, ld, d, alpha);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
