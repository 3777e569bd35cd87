// Baseline NetVLAD (WillowModelReg / NetVladOrthoReg) pieces that the NetVladV1/V2 path lacks:
//   * random frame sampling indices (model_utils.py:26-73), consumed by the gather kernels through `frame_index`
//   * the orthogonal regulariser on the cluster centres (module_utils.py:55-90): value and gradient, fp32 on the
//     CUDA cores (K x K x D = 67 MFLOP for rgb: launch-bound, the tensor cores would only add rounding to a sign())
#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

// ------------------------------------------------------------------------------------------------
// counter-based uniform [0,1): splitmix64 of (seed, counter) -> 24 random mantissa bits
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long ctr) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ctr + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// mode 0: SampleRandomFrames   idx[b,i] = int32( u[b,i] * fl32(nf[b]) )                        (model_utils.py:54-73)
// mode 1: SampleRandomSequence start = int32( u[b] * fl32(max(nf-T,0)+1) ); idx = min(start+i, nf-1)    (:26-51)
__global__ void __launch_bounds__(256) random_index_kernel(const int* __restrict__ nf, const float* __restrict__ uniform,
                                                           unsigned long long seed, int B, int T, int max_frames,
                                                           int mode, int* __restrict__ idx) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * T) return;
  const int b = r / T, i = r - b * T;
  const int n = __ldg(nf + b);
  int v;
  if (mode == 0) {
    const float u = uniform ? __ldg(uniform + r) : uniform01(seed, (unsigned long long)r);
    v = __float2int_rz(__fmul_rn(u, (float)n));
  } else {
    const float u = uniform ? __ldg(uniform + b) : uniform01(seed, (unsigned long long)b);
    const float span = fmaxf((float)n - (float)T, 0.f) + 1.f;
    v = min(__float2int_rz(__fmul_rn(u, span)) + i, n - 1);
  }
  idx[r] = min(max(v, 0), max_frames - 1);   // tf.gather_nd would fault on nf = 0; frame 0 (zero padding) is read instead
}

int random_frame_index(const int* nf, const float* uniform, unsigned long long seed, int B, int T, int max_frames,
                       int mode, int* idx, cudaStream_t st) {
  LPM_REQUIRE(mode == 0 || mode == 1, "random_frame_index: mode must be 0 (frames) or 1 (sequence)");
  random_index_kernel<<<(B * T + 255) / 256, 256, 0, st>>>(nf, uniform, seed, B, T, max_frames, mode, idx);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

// ------------------------------------------------------------------------------------------------
// orthogonal regulariser  R(W) = scale * sum_ij | (N^T N - I)_ij |,  N = l2_normalize(W [D][K], axis=1)
//   dR/dN = scale * N (S + S^T),  S = sign(N^T N - I);   dW_d = (dN_d - N_d (N_d . dN_d)) * rn_d
// workspace (floats): N [D*K] | dN [D*K] | M [K*K] | rn [D] | partial [tiles]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ortho_rownorm_kernel(const float* __restrict__ w, int D, int K, float* __restrict__ n,
                                                            float* __restrict__ rn) {
  const int d = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (d >= D) return;
  float ss = 0.f;
  for (int k = lane; k < K; k += 32) { const float v = w[(size_t)d * K + k]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float r = rsqrtf(fmaxf(ss, 1e-12f));
  if (lane == 0) rn[d] = ss >= 1e-12f ? r : -r;          // sign bit marks the clamped branch (constant scale)
  for (int k = lane; k < K; k += 32) n[(size_t)d * K + k] = w[(size_t)d * K + k] * r;
}

// 64 x 64 output tile per block, 4 x 4 outputs per thread, BK-deep slices through shared memory (fp32 FMA).  These
// products are small and latency-bound (one global round trip per slice), hence the deep slices.
constexpr int BK = 32;
struct Tile64 {
  float a[BK][64 + 4];
  float b[BK][64 + 4];
};

// Partial Gram tile over one slice of the contraction: Gp[z] = N[d in slice z]^T N[d in slice z].  The contraction is
// split over gridDim.z so that the (K/64)^2 output tiles still fill the GPU (16 tiles for K = 256) and no block walks
// more than a few global-load round trips (a single-block-per-tile version was load-latency bound: 160 us).
__global__ void __launch_bounds__(256) ortho_gram_kernel(const float* __restrict__ n, int D, int K, int d_per_split,
                                                         float* __restrict__ Gp) {
  __shared__ Tile64 t;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int d_begin = blockIdx.z * d_per_split, d_end = min(D, d_begin + d_per_split);
  float acc[4][4] = {};
  for (int d0 = d_begin; d0 < d_end; d0 += BK) {
    for (int e = threadIdx.x; e < BK * 64; e += 256) {
      const int dd = e >> 6, c = e & 63, d = d0 + dd;
      t.a[dd][c] = (d < d_end && i0 + c < K) ? n[(size_t)d * K + i0 + c] : 0.f;
      t.b[dd][c] = (d < d_end && j0 + c < K) ? n[(size_t)d * K + j0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int dd = 0; dd < BK; ++dd) {
      const float4 av = *reinterpret_cast<const float4*>(&t.a[dd][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&t.b[dd][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a4[p], b4[q], acc[p][q]);
    }
    __syncthreads();
  }
  float* g = Gp + (size_t)blockIdx.z * K * K;
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + ty * 4 + p, j = j0 + tx * 4 + q;
      if (i < K && j < K) g[(size_t)i * K + j] = acc[p][q];
    }
}

// G = sum of the slices (fixed order); S = sign(G - I); per-block sums of |G - I|
__global__ void __launch_bounds__(256) ortho_sign_kernel(const float* __restrict__ Gp, int splits, int K, float* __restrict__ S,
                                                         float* __restrict__ partial) {
  __shared__ float red[8];
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  float a = 0.f;
  if (e < (long long)K * K) {
    float g = 0.f;
    for (int z = 0; z < splits; ++z) g += Gp[(size_t)z * K * K + e];
    const int i = (int)(e / K), j = (int)(e - (long long)i * K);
    g -= (i == j ? 1.f : 0.f);
    a = fabsf(g);
    S[e] = g > 0.f ? 1.f : (g < 0.f ? -1.f : 0.f);
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

// dN tile = N (S + S^T) = 2 N S   [D x K]   (G[i][j] and G[j][i] are the same products summed in the same order, so S is
// exactly symmetric and the transposed, uncoalesced read of S is not needed)
__global__ void __launch_bounds__(256) ortho_dn_kernel(const float* __restrict__ n, const float* __restrict__ S, int D, int K,
                                                       float* __restrict__ dn) {
  __shared__ Tile64 t;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int r0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int i0 = 0; i0 < K; i0 += BK) {
    for (int e = threadIdx.x; e < BK * 64; e += 256) {
      const int c = e / BK, ii = e % BK;                 // a: N[r0 + c][i0 + ii] (coalesced along i), stored [ii][c]
      t.a[ii][c] = (r0 + c < D && i0 + ii < K) ? n[(size_t)(r0 + c) * K + i0 + ii] : 0.f;
      const int i2 = e >> 6, c2 = e & 63, i = i0 + i2, j = j0 + c2;
      t.b[i2][c2] = (i < K && j < K) ? 2.f * S[(size_t)i * K + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int ii = 0; ii < BK; ++ii) {
      const float4 av = *reinterpret_cast<const float4*>(&t.a[ii][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&t.b[ii][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a4[p], b4[q], acc[p][q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int d = r0 + ty * 4 + p, j = j0 + tx * 4 + q;
      if (d < D && j < K) dn[(size_t)d * K + j] = acc[p][q];
    }
}

// back through the row normalisation; accumulates grad_scale * dW into dw and writes the regulariser value
__global__ void __launch_bounds__(256) ortho_finish_kernel(const float* __restrict__ n, const float* __restrict__ dn,
                                                           const float* __restrict__ rn, int D, int K, float scale,
                                                           float grad_scale, int accumulate, float* __restrict__ dw,
                                                           const float* __restrict__ partial, int n_partial,
                                                           float* __restrict__ value) {
  const int d = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0 && value != nullptr) {
    double t = 0.0;
    for (int i = 0; i < n_partial; ++i) t += partial[i];
    *value = scale * (float)t;
  }
  if (d >= D || dw == nullptr) return;
  float dot = 0.f;
  for (int k = lane; k < K; k += 32) dot = fmaf(n[(size_t)d * K + k], dn[(size_t)d * K + k], dot);
  dot = warp_sum(dot);
  const float r = rn[d];
  const float rr = fabsf(r);
  if (r < 0.f) dot = 0.f;                                // clamped norm: N = W * const
  for (int k = lane; k < K; k += 32) {
    const size_t o = (size_t)d * K + k;
    const float g = scale * grad_scale * (dn[o] - n[o] * dot) * rr;
    dw[o] = accumulate ? dw[o] + g : g;
  }
}

static int ortho_splits(int D) { int sp = (D + 63) / 64; return sp > 16 ? 16 : (sp < 1 ? 1 : sp); }

unsigned long long ortho_reg_workspace_bytes(int D, int K) {
  const unsigned long long nb = ((unsigned long long)K * K + 255) / 256;
  return sizeof(float) * (2ull * D * K + (unsigned long long)K * K + D + nb + (unsigned long long)ortho_splits(D) * K * K);
}

int ortho_reg(const float* w, int D, int K, float scale, float grad_scale, int accumulate, float* value, float* dw,
              float* ws, unsigned long long ws_bytes, cudaStream_t st) {
  LPM_REQUIRE(D > 0 && K > 0, "ortho_reg: bad shape %d x %d", D, K);
  if (ws_bytes < ortho_reg_workspace_bytes(D, K))
    return fail(LPM_ERR_WORKSPACE, "ortho_reg: workspace too small (%llu < %llu bytes)", ws_bytes, ortho_reg_workspace_bytes(D, K));
  const int nb = (int)(((long long)K * K + 255) / 256), splits = ortho_splits(D);
  float* n = ws;
  float* dn = n + (size_t)D * K;
  float* S = dn + (size_t)D * K;
  float* rn = S + (size_t)K * K;
  float* partial = rn + D;
  float* Gp = partial + nb;
  const int kt = (K + 63) / 64;
  const int d_per_split = ((D + splits - 1) / splits + BK - 1) / BK * BK;
  ortho_rownorm_kernel<<<(D + 7) / 8, 256, 0, st>>>(w, D, K, n, rn);
  ortho_gram_kernel<<<dim3(kt, kt, splits), 256, 0, st>>>(n, D, K, d_per_split, Gp);
  ortho_sign_kernel<<<nb, 256, 0, st>>>(Gp, splits, K, S, partial);
  if (dw != nullptr) ortho_dn_kernel<<<dim3(kt, (D + 63) / 64), 256, 0, st>>>(n, S, D, K, dn);
  ortho_finish_kernel<<<(D + 7) / 8, 256, 0, st>>>(n, dn, rn, D, K, scale, grad_scale, accumulate, dw, partial, nb, value);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
