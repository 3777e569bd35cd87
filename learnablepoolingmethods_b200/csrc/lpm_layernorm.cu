// Joint-axis layer norm with fused residual as ONE pass over HBM, for one or two chained norms
// (tf.contrib.layers.layer_norm with begin_norm_axis=1: moments over all [rows, D] elements of a sample;
//  transformer_utils.py:406-411 and 712-713):
//
//   stage 1   u1 = a + b * row_scale          y1 = (u1 - mean1) * rstd1 * gamma1 + beta1
//   stage 2   u2 = y1 + b                     y2 = (u2 - mean2) * rstd2 * gamma2 + beta2      (optional)
//
// The second stage is exactly the tail of the encoder block: FeedForwardNetwork ends with LN(f + h1)
// (:712-713) and TransformerEncoder.forward applies LN(. + h1) again (:410-411) with the same residual.
//
// A thread-block cluster of up to 8 CTAs owns one sample (262 144 elements = 512 KB of fp16 at config 1).
// Each CTA keeps its slice of u in shared memory, the cluster exchanges the (sum, sum of squares) pairs
// through distributed shared memory, and the normalised result is written straight from shared memory:
// a and b are read once, y written once (u1 / u2 are stored only when the backward needs them).  The
// two-kernel path it replaces read a, b, wrote u, re-read u and wrote y, per norm.
#include <cooperative_groups.h>

#include <cstdlib>

#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace cg = cooperative_groups;

namespace lpm {

struct LnChainParams {
  const __half* a; long long a_stride;
  const __half* b; long long b_stride;
  const float* b_row_scale;        // [B][rows] or null
  int rows, D;
  long long n8;                    // rows * D / 8
  int per;                         // 16-byte pieces per CTA
  int stage_cnt;                   // pieces [0, stage_cnt) of the slice live in shared memory, the rest are recomputed from a / b
  float eps;
  const float* gamma1; const float* beta1;
  __half* u1_out; long long u1_stride; float* stats1;
  const float* gamma2; const float* beta2;      // null -> single norm
  __half* u2_out; long long u2_stride; float* stats2;
  __half* y; long long y_stride;
  __half* y_lo;                    // optional: fp16(y - fp16(y)) with the same layout (split-precision operand of the
                                   // hidden projection, frame_level_models.py:2319), or null
};


__device__ __forceinline__ void ln_unpack(const uint4& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ uint4 ln_pack(const float* f) {
  uint4 v;
  v.x = pack_half2(f[0], f[1]); v.y = pack_half2(f[2], f[3]); v.z = pack_half2(f[4], f[5]); v.w = pack_half2(f[6], f[7]);
  return v;
}

// low-order halves of 8 fp32 values whose high-order fp16 halves are `hi`
__device__ __forceinline__ uint4 ln_pack_lo(const float* f, const uint4& hi) {
  float h[8];
  ln_unpack(hi, h);
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = f[j] - h[j];
  return ln_pack(h);
}

// (sum, sumsq) of this CTA -> cluster-wide (mean, rstd), reduced in rank order (deterministic)
template <int NW>
__device__ __forceinline__ float2 ln_cluster_moments(float s, float q, float* red, float2* slot, double n, float eps) {
  cg::cluster_group cluster = cg::this_cluster();
  s = warp_sum(s); q = warp_sum(q);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[w] = s; red[8 + w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) { ts += red[i]; tq += red[8 + i]; }
    *slot = make_float2(ts, tq);
  }
  cluster.sync();                                  // every CTA's slot is written and visible cluster-wide
  if (threadIdx.x == 0) {
    double ds = 0.0, dq = 0.0;
    const unsigned cs = cluster.num_blocks();
    for (unsigned r = 0; r < cs; ++r) {
      const float2 v = *cluster.map_shared_rank(slot, r);
      ds += v.x; dq += v.y;
    }
    const double m = ds / n;
    double var = dq / n - m * m;
    if (var < 0.0) var = 0.0;
    red[16] = (float)m;
    red[17] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  return make_float2(red[16], red[17]);
}

// STAGE = true: this CTA's slice of u lives in shared memory between the passes (64 KB per CTA at config 1: three CTAs per
// SM, 640 CTAs = 1.45 waves).  STAGE = false: no shared-memory slice -- the later passes re-read a and b, which are still in
// the 126 MB L2 (a sample is 1 MB, a resident wave of samples < 100 MB), and recompute u: every CTA of the launch is
// resident at once and HBM still sees one pass.  Same arithmetic, same rounding points (u and y1 are rounded to fp16
// exactly where the staged version stores them).
template <bool STAGE, int NT, int DEPTH, int MINB>
__global__ void __launch_bounds__(NT, MINB) ln_chain_kernel(const LnChainParams p) {
  extern __shared__ __align__(16) uint8_t ln_sm[];
  uint4* sU = reinterpret_cast<uint4*>(ln_sm);                 // this CTA's slice of u (fp16), STAGE only
  __shared__ float red[32];
  __shared__ float2 slots[2];
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks();
  const int sample = blockIdx.x / cs, rank = (int)cluster.block_rank();
  const long long i0 = (long long)rank * p.per;
  const int cnt = (int)max(0ll, min((long long)p.per, p.n8 - i0));
  const uint4* pa = reinterpret_cast<const uint4*>(p.a + sample * p.a_stride) + i0;
  const uint4* pb = reinterpret_cast<const uint4*>(p.b + sample * p.b_stride) + i0;
  const float* rsc = p.b_row_scale ? p.b_row_scale + (long long)sample * p.rows : nullptr;
  uint4* pu1 = p.u1_out ? reinterpret_cast<uint4*>(p.u1_out + sample * p.u1_stride) + i0 : nullptr;
  const double n = (double)p.rows * p.D;

  // ---- pass A: u1 = a + b*rs -> shared memory (+ global when the backward wants it), moments ------------------
  float s = 0.f, q = 0.f;
  const int stage_cnt = STAGE ? p.stage_cnt : 0;
  for (int base = 0; base < cnt; base += DEPTH * NT) {
    uint4 va[DEPTH], vb[DEPTH];
#pragma unroll
    for (int k = 0; k < DEPTH; ++k) {
      const int i = base + k * NT + threadIdx.x;
      if (i < cnt) { va[k] = pa[i]; vb[k] = __ldg(pb + i); }
    }
#pragma unroll
    for (int k = 0; k < DEPTH; ++k) {
      const int i = base + k * NT + threadIdx.x;
      if (i < cnt) {
        float fa[8], fb[8];
        ln_unpack(va[k], fa); ln_unpack(vb[k], fb);
        const float rs = rsc ? __ldg(rsc + ((i0 + i) * 8) / p.D) : 1.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) fa[j] += fb[j] * rs;
        const uint4 vu = ln_pack(fa);
        if (i < stage_cnt) sU[i] = vu;
        if (pu1) pu1[i] = vu;
        ln_unpack(vu, fa);                       // moments of the stored (fp16-rounded) values
#pragma unroll
        for (int j = 0; j < 8; ++j) { s += fa[j]; q += fa[j] * fa[j]; }
      }
    }
  }
  const float2 mr1 = ln_cluster_moments<NT / 32>(s, q, red, &slots[0], n, p.eps);
  if (p.stats1 && rank == 0 && threadIdx.x == 0) { p.stats1[sample * 2] = mr1.x; p.stats1[sample * 2 + 1] = mr1.y; }

  // u1 piece i of this CTA's slice: from shared memory, or recomputed from a and b (L2 hits) with the same rounding
  auto load_u1 = [&](int i, float* f) {
    if (i < stage_cnt) { ln_unpack(sU[i], f); return; }
    float fb[8];
    ln_unpack(__ldg(pa + i), f); ln_unpack(__ldg(pb + i), fb);
    const float rs = rsc ? __ldg(rsc + ((i0 + i) * 8) / p.D) : 1.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += fb[j] * rs;
    ln_unpack(ln_pack(f), f);
  };
  const bool two = p.gamma2 != nullptr;
  uint4* py = reinterpret_cast<uint4*>(p.y + sample * p.y_stride) + i0;
  uint4* pyl = p.y_lo ? reinterpret_cast<uint4*>(p.y_lo + sample * p.y_stride) + i0 : nullptr;
  if (!two) {
    // ---- pass B (single norm): y = LN(u1) ----------------------------------------------------------------------
#pragma unroll 2
    for (int i = threadIdx.x; i < cnt; i += NT) {
      float f[8];
      load_u1(i, f);
      const int d = (int)(((i0 + i) * 8) % p.D);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma1 + d)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma1 + d + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta1 + d)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta1 + d + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mr1.x) * mr1.y * gg[j] + bb[j];
      const uint4 hi = ln_pack(f);
      py[i] = hi;
      if (pyl) pyl[i] = ln_pack_lo(f, hi);
    }
    cluster.sync();                                // no CTA may exit while a peer can still read its slot
    return;
  }
  // ---- pass B (chained): y1 = LN(u1) (rounded to fp16 as the unfused path stores it); u2 = y1 + b -------------
  uint4* pu2 = p.u2_out ? reinterpret_cast<uint4*>(p.u2_out + sample * p.u2_stride) + i0 : nullptr;
  s = 0.f; q = 0.f;
  for (int base = 0; base < cnt; base += DEPTH * NT) {
    uint4 vb[DEPTH];
#pragma unroll
    for (int k = 0; k < DEPTH; ++k) {
      const int i = base + k * NT + threadIdx.x;
      if (i < cnt) vb[k] = __ldg(pb + i);          // second use of the residual: an L2 hit
    }
#pragma unroll
    for (int k = 0; k < DEPTH; ++k) {
      const int i = base + k * NT + threadIdx.x;
      if (i < cnt) {
        float f[8], fb[8];
        load_u1(i, f); ln_unpack(vb[k], fb);
        const int d = (int)(((i0 + i) * 8) % p.D);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma1 + d)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma1 + d + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta1 + d)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta1 + d + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (f[j] - mr1.x) * mr1.y * gg[j] + bb[j];
        ln_unpack(ln_pack(f), f);                  // y1 as fp16
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += fb[j];
        const uint4 vu = ln_pack(f);
        if (i < stage_cnt) sU[i] = vu;
        if (pu2) pu2[i] = vu;
        ln_unpack(vu, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s += f[j]; q += f[j] * f[j]; }
      }
    }
  }
  const float2 mr2 = ln_cluster_moments<NT / 32>(s, q, red, &slots[1], n, p.eps);
  if (p.stats2 && rank == 0 && threadIdx.x == 0) { p.stats2[sample * 2] = mr2.x; p.stats2[sample * 2 + 1] = mr2.y; }
  // ---- pass C: y2 = LN(u2) ---------------------------------------------------------------------------------------
#pragma unroll 2
  for (int i = threadIdx.x; i < cnt; i += NT) {
    float f[8];
    if (i < stage_cnt) {
      ln_unpack(sU[i], f);
    } else {
      // u2 = fp16(fp16(LN1(u1)) + b), recomputed
      float fb[8];
      load_u1(i, f); ln_unpack(__ldg(pb + i), fb);
      const int d1 = (int)(((i0 + i) * 8) % p.D);
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(p.gamma1 + d1)), h1 = __ldg(reinterpret_cast<const float4*>(p.gamma1 + d1 + 4));
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(p.beta1 + d1)), c1 = __ldg(reinterpret_cast<const float4*>(p.beta1 + d1 + 4));
      const float hg[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
      const float hb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mr1.x) * mr1.y * hg[j] + hb[j];
      ln_unpack(ln_pack(f), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += fb[j];
      ln_unpack(ln_pack(f), f);
    }
    const int d = (int)(((i0 + i) * 8) % p.D);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma2 + d)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma2 + d + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta2 + d)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta2 + d + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (f[j] - mr2.x) * mr2.y * gg[j] + bb[j];
    const uint4 hi = ln_pack(f);
    py[i] = hi;
    if (pyl) pyl[i] = ln_pack_lo(f, hi);
  }
  cluster.sync();
}

// cluster size and per-CTA slice for a sample of n8 16-byte pieces; returns false when the slice does not fit
static bool ln_chain_plan(long long n8, int* cs, int* per) {
  int c = 8;
  while (c > 1 && n8 / c < 1024) c >>= 1;          // at least 8192 elements per CTA before splitting further
  const long long pr = (n8 + c - 1) / c;
  if (pr * 16 > 200 * 1024) return false;
  *cs = c; *per = (int)pr;
  return true;
}

int layernorm_chain_supported(int rows, int D) {
  int cs, per;
  return (D % 8 == 0 && rows > 0) ? (ln_chain_plan((long long)rows * D / 8, &cs, &per) ? 1 : 0) : 0;
}

int layernorm_chain_fwd(const __half* a, long long a_stride, const __half* b, long long b_stride, const float* b_row_scale,
                        int B, int rows, int D, float eps, const float* gamma1, const float* beta1, __half* u1_out,
                        long long u1_stride, float* stats1, const float* gamma2, const float* beta2, __half* u2_out,
                        long long u2_stride, float* stats2, __half* y, long long y_stride, __half* y_lo, cudaStream_t st) {
  LPM_REQUIRE(D % 8 == 0 && a_stride % 8 == 0 && b_stride % 8 == 0 && y_stride % 8 == 0 && u1_stride % 8 == 0 && u2_stride % 8 == 0,
              "layernorm_chain: D and sample strides must be multiples of 8");
  LPM_REQUIRE(a && b && gamma1 && beta1 && y && (gamma2 == nullptr) == (beta2 == nullptr), "layernorm_chain: bad arguments");
  LnChainParams p{};
  p.a = a; p.a_stride = a_stride; p.b = b; p.b_stride = b_stride; p.b_row_scale = b_row_scale;
  p.rows = rows; p.D = D; p.n8 = (long long)rows * D / 8; p.eps = eps;
  p.gamma1 = gamma1; p.beta1 = beta1; p.u1_out = u1_out; p.u1_stride = u1_stride; p.stats1 = stats1;
  p.gamma2 = gamma2; p.beta2 = beta2; p.u2_out = u2_out; p.u2_stride = u2_stride; p.stats2 = stats2;
  p.y = y; p.y_stride = y_stride; p.y_lo = y_lo;
  int cs = 1;
  if (!ln_chain_plan(p.n8, &cs, &p.per)) return fail(LPM_ERR_ARG, "layernorm_chain: sample of %d x %d does not fit", rows, D);
  // Measured at config 1 (gpurun r2n): staged 58 us per launch (1.45 waves of 3 CTAs per SM), unstaged 72-93 us although the
  // whole launch is resident -- the second and third read of a and b cost more L2 traffic than the tail wave.  Staged is the
  // default; LPM_LN_STAGE=0 selects the L2 re-read variant (measurement switch).
  static const bool unstaged = getenv("LPM_LN_STAGE") != nullptr && getenv("LPM_LN_STAGE")[0] == '0';
  const bool stage = !unstaged || a == u1_out;            // in-place u1 (a aliased): the re-read would see u1, not a
  // Hybrid (LPM_LN_HYBRID=1, experiment): when the fully staged launch does not fit in one wave (three 64 KB CTAs per SM at
  // 74 registers: 640 CTAs = 1.45 waves at config 1), stage only 40 KB of each slice, recompute the rest from L2, and cap the
  // registers for five CTAs per SM so that every cluster of the launch is resident at once.
  static const bool hybrid_on = getenv("LPM_LN_HYBRID") != nullptr && getenv("LPM_LN_HYBRID")[0] == '1';
  const bool hybrid = hybrid_on && stage && a != u1_out && (long long)B * cs > 3ll * num_sms() && (long long)B * cs <= 5ll * num_sms() &&
                      (size_t)p.per * 16 > 40 * 1024;
  p.stage_cnt = hybrid ? 40 * 1024 / 16 : p.per;
  const size_t smem = stage ? (size_t)p.stage_cnt * 16 : 0;
  static size_t attr = 0;
  if (smem > attr) {
    LPM_CUDA_CHECK(cudaFuncSetAttribute(ln_chain_kernel<true, 256, 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LPM_CUDA_CHECK(cudaFuncSetAttribute(ln_chain_kernel<true, 256, 2, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 44 * 1024));
    attr = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(B * cs));
  cfg.blockDim = dim3(stage ? 256 : 128);     // unstaged: 128-thread CTAs, six per SM by registers -> the whole launch resident
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (hybrid) LPM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, ln_chain_kernel<true, 256, 2, 5>, p));
  else if (stage) LPM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, ln_chain_kernel<true, 256, 4, 3>, p));
  else LPM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, ln_chain_kernel<false, 128, 4, 6>, p));
  return LPM_OK;
}

}  // namespace lpm
