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

// p -= lr_t * m / (sqrt(v) + eps) with the approximate MUFU forms (sqrt.approx / rcp.approx, <= 2 ulp each):
// the update is ~1e-4 of |p|, so the difference from the IEEE sequences is far below one ulp of the parameter,
// and the dependent-instruction chain per element drops from ~20 to 4.
__device__ __forceinline__ float adam_update(float m, float v, float lr_t, float eps) {
  float s, r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(v));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s + eps));
  return lr_t * m * r;
}

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
                                                      int tensor0, int n_tensors, float clip, float* __restrict__ factor,
                                                      float* __restrict__ norms, int* __restrict__ flag) {
  const int t = tensor0 + blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (t >= tensor0 + n_tensors) return;
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
                                                      float lr_t, const float* __restrict__ lr_dev, float b1, float b2, float eps) {
  if (*flag) return;  // overflow: skip the step
  if (lr_dev != nullptr) lr_t = *lr_dev;   // step size kept in device memory (CUDA-graph replays)
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
  // two float4 per array and iteration, every load issued before the first use (8 x 16 B in flight per thread: the 23 M
  // non-factored parameters stream at HBM speed instead of 3.8 TB/s); p / m / v are touched once per step -> streaming
  // (evict-first) accesses, the gradient keeps its default policy (the norm pass just read it)
  auto update4 = [&](int i, float4 pj, float4 mj, float4 vj, const float4 gr) {
    float gx = (gr.x + w * pj.x) * f, gy = (gr.y + w * pj.y) * f, gz = (gr.z + w * pj.z) * f, gw = (gr.w + w * pj.w) * f;
    mj.x = b1 * mj.x + (1.f - b1) * gx; mj.y = b1 * mj.y + (1.f - b1) * gy;
    mj.z = b1 * mj.z + (1.f - b1) * gz; mj.w = b1 * mj.w + (1.f - b1) * gw;
    vj.x = b2 * vj.x + (1.f - b2) * gx * gx; vj.y = b2 * vj.y + (1.f - b2) * gy * gy;
    vj.z = b2 * vj.z + (1.f - b2) * gz * gz; vj.w = b2 * vj.w + (1.f - b2) * gw * gw;
    pj.x -= adam_update(mj.x, vj.x, lr_t, eps); pj.y -= adam_update(mj.y, vj.y, lr_t, eps);
    pj.z -= adam_update(mj.z, vj.z, lr_t, eps); pj.w -= adam_update(mj.w, vj.w, lr_t, eps);
    __stcs(m4 + i, mj); __stcs(v4 + i, vj); __stcs(p4 + i, pj);
    if (sdst) {
      const long long e = eoff + 4ll * i;
      const long long r = e / cols;
      const int c = (int)(e - r * cols);
      uint2 o;
      o.x = pack_half2(pj.x, pj.y); o.y = pack_half2(pj.z, pj.w);
      *reinterpret_cast<uint2*>(sdst + r * ld + c) = o;
    }
  };
  int i = threadIdx.x;
  for (; i + 256 < n4; i += 512) {
    const float4 pa = __ldcs(p4 + i), ma = __ldcs(m4 + i), va = __ldcs(v4 + i), ga = __ldg(g4 + i);
    const float4 pb = __ldcs(p4 + i + 256), mb = __ldcs(m4 + i + 256), vb = __ldcs(v4 + i + 256), gb = __ldg(g4 + i + 256);
    update4(i, pa, ma, va, ga);
    update4(i + 256, pb, mb, vb, gb);
  }
  for (; i < n4; i += 256) update4(i, __ldcs(p4 + i), __ldcs(m4 + i), __ldcs(v4 + i), __ldg(g4 + i));
  for (int i = n4 * 4 + threadIdx.x; i < len; i += 256) {
    const long long j = start + i;
    const float pj = p[j];
    const float gj = (g[j] + w * pj) * f;
    const float mj = b1 * m[j] + (1.f - b1) * gj;
    const float vj = b2 * v[j] + (1.f - b2) * gj * gj;
    const float pn = pj - adam_update(mj, vj, lr_t, eps);
    m[j] = mj; v[j] = vj; p[j] = pn;
    if (sdst) {
      const long long e = eoff + i;
      const long long r = e / cols;
      sdst[r * ld + (e - r * cols)] = __float2half_rn(pn);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Rank-R ("factored") weight-gradient Adam for the hidden projection (frame_level_models.py:2314-2319;
// train.py:321-336).  The gradient of W[Kd][N] is the rank-R product dW = alpha * A^T G with A = the layer
// input [R][Kd] (the VLAD descriptor) and G = the output gradient [R][N], R = tower batch.  W is 138 M
// parameters at config 1, so writing dW (4 B/param), re-reading it for the clip norm and again for Adam
// is 1.7 GB of HBM traffic per step.  Instead:
//   * the clip norm comes from two R x R Gram matrices:  ||A^T G||_F^2 = sum_ij (A A^T)_ij (G G^T)_ij
//   * this kernel recomputes each 128 x N panel of dW on the warp-level tensor path (K = R <= 128: the
//     product is 2 % of the kernel's time, the rest is the m/v/w stream) and applies clip + Adam +
//     fp16-shadow refresh in registers.  dW never exists in memory.
// Column permutation: four m16n8 accumulator tiles are interleaved so that a thread owns two runs of 4 consecutive
// columns of a row and every 16-byte access of a quad lands in one fully used 64-byte run.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ra_ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ra_mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int RA_ROWS = 64;         // rows of W per panel (four 16-row blocks)
constexpr int RA_MAX_KS = 8;        // R <= 128

__device__ __forceinline__ void ra_cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int bytes = valid ? 16 : 0;   // src-size 0 -> the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ra_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ra_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// G [R][N] -> Gp [Rp][N] with the columns of every 32-column group in accumulator order (see rank_adam_kernel) and
// zero rows beyond R.  Physical column p -> accumulator tile s = 2*(p/16) + (p%4)/2, column n = 2*((p%16)/4) + p%2:
// a thread then owns columns 4t..4t+3 and 16+4t..16+4t+3 of its row, so every 16-byte access of a quad lands in one
// contiguous, fully used 64-byte run.
__global__ void __launch_bounds__(256) rank_permute_kernel(const __half* __restrict__ G, long long ldg, int R, int Rp, int N,
                                                           __half* __restrict__ Gp) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= Rp * (N / 8)) return;
  const int r = c / (N / 8), ch = c - r * (N / 8);        // physical columns ch*8 .. ch*8+7
  uint4 val = make_uint4(0, 0, 0, 0);
  if (r < R) val = __ldg(reinterpret_cast<const uint4*>(G + (long long)r * ldg + ch * 8));
  const int q = ch >> 2, hi = (ch >> 1) & 1, c2 = ch & 1;
  uint32_t* dst = reinterpret_cast<uint32_t*>(Gp + (size_t)r * N + q * 32) + 8 * hi + 2 * c2;
  dst[0] = val.x; dst[4] = val.y; dst[1] = val.z; dst[5] = val.w;
}

// Persistent: one CTA per SM walks the 64-row panels of W in 16-row blocks; within a block the warps take
// interleaved 32-column groups, so the CTA streams one contiguous 16 x N block of w / m / v at a time.
// Gp stays resident in shared memory; the A panel of the next iteration arrives by cp.async while the current one
// is processed.  The kernel is bound by bytes in flight, not by arithmetic: saturating HBM with this read-modify-
// write stream needs ~100 KB outstanding per SM, more than the register file can stage next to the MMA fragments.
// So every thread owns a private 192-byte slot of shared memory into which the w / m / v pieces of its NEXT
// (block, column group) item are copied asynchronously while it works on the current one.
__global__ void __launch_bounds__(512, 1) rank_adam_kernel(const __half* __restrict__ A, long long lda,
                                                           const __half* __restrict__ Gp, int R,
                                                           long long Kd, int N, float alpha,
                                                           const float* __restrict__ factor, const int* __restrict__ flag,
                                                           float* __restrict__ w, float* __restrict__ m,
                                                           float* __restrict__ v, __half* __restrict__ w16, long long ldw16,
                                                           float lr_t, const float* __restrict__ lr_dev, float b1, float b2, float eps) {
  if (*flag) return;   // non-finite gradient norm somewhere: the whole step is skipped
  if (lr_dev != nullptr) lr_t = *lr_dev;
  extern __shared__ __align__(16) uint8_t ra_sm[];
  const int Rp = (R + 15) & ~15;
  const int gs = N + 8, as = RA_ROWS + 8;                 // padded row strides (halves): conflict-free ldmatrix
  const int nthr = blockDim.x, nwarps = nthr >> 5;
  uint4* sSlot = reinterpret_cast<uint4*>(ra_sm);         // [12][nthr] 16-byte pieces (thread-private staging)
  __half* sG = reinterpret_cast<__half*>(ra_sm + (size_t)12 * nthr * 16);   // [Rp][gs], accumulator column order
  __half* sAbuf = sG + (size_t)Rp * gs;                   // 2 x [Rp][as]
  const int npanels = (int)((Kd + RA_ROWS - 1) / RA_ROWS);
  const int nq = N / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nks = Rp / 16;
  const int j = lane >> 3, i = lane & 7;
  const int g = lane >> 2, t = lane & 3;
  auto stage_a = [&](int panel, int buf) {
    const long long r0 = (long long)panel * RA_ROWS;
    __half* dst = sAbuf + (size_t)buf * Rp * as;
    for (int c = threadIdx.x; c < Rp * (RA_ROWS / 8); c += nthr) {
      const int r = c / (RA_ROWS / 8), ch = c - r * (RA_ROWS / 8);
      const bool ok = r < R && r0 + ch * 8 < Kd;
      ra_cp_async16(dst + (size_t)r * as + ch * 8, ok ? A + (long long)r * lda + r0 + ch * 8 : A, ok);
    }
  };
  // this thread's pieces of item (panel, sb, q): rows r_lo / r_lo + 8, columns q*32 + 4t (+16), of w, m, v
  auto prefetch_item = [&](int panel, int sb, int q) {
    const long long r_lo = (long long)panel * RA_ROWS + sb * 16 + g;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long r = r_lo + 8 * h;
      const long long off = (r < Kd ? r : r_lo) * N + q * 32 + t * 4;
      ra_cp_async16(&sSlot[(h * 6 + 0) * nthr + threadIdx.x], w + off, true);
      ra_cp_async16(&sSlot[(h * 6 + 1) * nthr + threadIdx.x], w + off + 16, true);
      ra_cp_async16(&sSlot[(h * 6 + 2) * nthr + threadIdx.x], m + off, true);
      ra_cp_async16(&sSlot[(h * 6 + 3) * nthr + threadIdx.x], m + off + 16, true);
      ra_cp_async16(&sSlot[(h * 6 + 4) * nthr + threadIdx.x], v + off, true);
      ra_cp_async16(&sSlot[(h * 6 + 5) * nthr + threadIdx.x], v + off + 16, true);
    }
  };
  for (int c = threadIdx.x; c < Rp * (N / 8); c += nthr) {
    const int r = c / (N / 8), ch = c - r * (N / 8);
    ra_cp_async16(sG + (size_t)r * gs + ch * 8, Gp + (size_t)r * N + ch * 8, true);
  }
  if ((int)blockIdx.x < npanels) {
    stage_a(blockIdx.x, 0);
    if (warp < nq) prefetch_item(blockIdx.x, 0, warp);
  }
  ra_cp_commit();
  const float f = alpha * factor[0];
  int buf = 0;
  for (int panel = blockIdx.x; panel < npanels; panel += gridDim.x, buf ^= 1) {
    ra_cp_wait<0>();
    __syncthreads();          // sG / sA[buf] complete for everyone; every warp has left sA[buf ^ 1]
    const bool more_panels = panel + (int)gridDim.x < npanels;
    if (more_panels) stage_a(panel + gridDim.x, buf ^ 1);   // committed with the first item's prefetch group
    const __half* sA = sAbuf + (size_t)buf * Rp * as;
    const long long row0 = (long long)panel * RA_ROWS;
    int nsb = RA_ROWS / 16;
    if (row0 + RA_ROWS > Kd) nsb = (int)((Kd - row0 + 15) / 16);
    for (int sb = 0; sb < nsb; ++sb) {
      const long long r_lo = row0 + sb * 16 + g;
      for (int q = warp; q < nq; q += nwarps) {
        // ---- this item's w / m / v pieces: private slot -> registers -----------------------------------------
        ra_cp_wait<0>();
        uint4 pc[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) pc[k] = sSlot[k * nthr + threadIdx.x];
        // ---- next item's pieces -> private slot (asynchronous, lands while this item is processed) -----------
        {
          int qn = q + nwarps, sbn = sb, pn = panel;
          if (qn >= nq) { qn = warp; if (++sbn == nsb) { sbn = 0; pn = panel + gridDim.x; } }
          if (pn < npanels) prefetch_item(pn, sbn, qn);
          ra_cp_commit();
        }
        // ---- gradient tile on the tensor cores -----------------------------------------------------------------
        float acc[4][4];
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < RA_MAX_KS; ++ks) {
          if (ks < nks) {
            uint32_t af[4];
            const __half* asrc = sA + (size_t)(ks * 16 + (j >> 1) * 8 + i) * as + sb * 16 + (j & 1) * 8;
            ra_ldsm_x4_t(smem_u32(asrc), af[0], af[1], af[2], af[3]);
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
              uint32_t b0, b1, b2, b3;
              const __half* src = sG + (size_t)(ks * 16 + (j & 1) * 8 + i) * gs + q * 32 + 8 * (2 * sp + (j >> 1));
              ra_ldsm_x4_t(smem_u32(src), b0, b1, b2, b3);
              ra_mma16816(acc[2 * sp], af, b0, b1);
              ra_mma16816(acc[2 * sp + 1], af, b2, b3);
            }
          }
        }
        // ---- clip + Adam + fp16 shadow ------------------------------------------------------------------------
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const long long r = r_lo + 8 * h;
          if (r >= Kd) continue;
          const long long off = r * N + q * 32 + t * 4;
          float gr[8];
#pragma unroll
          for (int s = 0; s < 4; ++s) { gr[2 * s] = acc[s][2 * h] * f; gr[2 * s + 1] = acc[s][2 * h + 1] * f; }
          float* wv = reinterpret_cast<float*>(&pc[h * 6 + 0]);     // 8 consecutive floats: pieces 0,1 / 2,3 / 4,5
          float* mv = reinterpret_cast<float*>(&pc[h * 6 + 2]);
          float* vv = reinterpret_cast<float*>(&pc[h * 6 + 4]);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            mv[e] = b1 * mv[e] + (1.f - b1) * gr[e];
            vv[e] = b2 * vv[e] + (1.f - b2) * gr[e] * gr[e];
            wv[e] -= adam_update(mv[e], vv[e], lr_t, eps);
          }
          *reinterpret_cast<uint4*>(m + off) = pc[h * 6 + 2];
          *reinterpret_cast<uint4*>(m + off + 16) = pc[h * 6 + 3];
          *reinterpret_cast<uint4*>(v + off) = pc[h * 6 + 4];
          *reinterpret_cast<uint4*>(v + off + 16) = pc[h * 6 + 5];
          *reinterpret_cast<uint4*>(w + off) = pc[h * 6 + 0];
          *reinterpret_cast<uint4*>(w + off + 16) = pc[h * 6 + 1];
          if (w16 != nullptr) {
            __half* dst = w16 + r * ldw16 + q * 32 + t * 4;
            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_half2(wv[0], wv[1]), pack_half2(wv[2], wv[3]));
            *reinterpret_cast<uint2*>(dst + 16) = make_uint2(pack_half2(wv[4], wv[5]), pack_half2(wv[6], wv[7]));
          }
        }
      }
    }
  }
  ra_cp_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Tiled variant of the factored Adam: the same arithmetic as rank_adam_kernel (bit-identical results) in many small,
// short-lived CTAs -- 16 rows of W each, 128 threads, 4 KB of shared memory, no persistent loop.  It is meant to run
// on a LOW-PRIORITY stream underneath the tensor-bound backward of the same step: its CTAs fit next to a resident
// GEMM CTA (which owns ~200 KB of shared memory but only half of the register file) and fill the SMs that small
// grids and kernel tails leave idle, so the 3.6 GB optimiser stream of hidden1_weights leaves the critical path.
// The B fragments (G^T) come pre-arranged in fragment order from global memory (80 KB, L2-resident) instead of a
// shared-memory copy per CTA; the w / m / v pieces are plain 16-byte loads issued before the tensor work.
//   Gf layout: [q = N/32][ks = Rp/16][hf = 2][lane = 32] uint4; hf selects accumulator tiles {2hf, 2hf+1}:
//   {b0(s), b1(s), b0(s+1), b1(s+1)} with b0 = {G[16ks+2t][c], G[16ks+2t+1][c]}, b1 = rows +8,
//   c = 32q + 16(s/2) + 4(g/2) + 2(s%2) + g%2  (g = lane/4, t = lane%4: the column order of rank_permute_kernel).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rank_gfrag_kernel(const __half* __restrict__ G, long long ldg, int R, int nks, int N,
                                                         uint4* __restrict__ Gf) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int total = (N / 32) * nks * 64;
  if (idx >= total) return;
  const int lane = idx & 31, hf = (idx >> 5) & 1, rest = idx >> 6;
  const int ks = rest % nks, q = rest / nks;
  const int g = lane >> 2, t = lane & 3;
  uint32_t out[4];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int s = 2 * hf + u;
    const int col = q * 32 + 16 * (s >> 1) + 4 * (g >> 1) + 2 * (s & 1) + (g & 1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k0 = ks * 16 + 2 * t + 8 * h;
      const __half lo = k0 < R ? G[(long long)k0 * ldg + col] : __float2half_rn(0.f);
      const __half hi = k0 + 1 < R ? G[(long long)(k0 + 1) * ldg + col] : __float2half_rn(0.f);
      out[2 * u + h] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
    }
  }
  Gf[idx] = make_uint4(out[0], out[1], out[2], out[3]);
}

constexpr int RT_AS = 24;   // padded row stride (halves) of the staged A tile [Rp][16]

template <int NKS>
__global__ void __launch_bounds__(128, 4) rank_adam_tile_kernel(const __half* __restrict__ A, long long lda,
                                                             const uint4* __restrict__ Gf, int R, long long Kd, int N,
                                                             float alpha, const float* __restrict__ factor,
                                                             const int* __restrict__ flag, float* __restrict__ w,
                                                             float* __restrict__ m, float* __restrict__ v,
                                                             __half* __restrict__ w16, long long ldw16, float lr_t,
                                                             const float* __restrict__ lr_dev, float b1, float b2, float eps) {
  if (*flag) return;
  if (lr_dev != nullptr) lr_t = *lr_dev;
  __shared__ __align__(16) __half sA[16 * NKS * RT_AS];
  constexpr int nks = NKS, Rp = 16 * NKS;
  const int nq = N / 32;
  const long long row0 = (long long)blockIdx.x * 16;
  for (int c = threadIdx.x; c < Rp * 2; c += 128) {
    const int r = c >> 1, ch = c & 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < R && row0 + ch * 8 < Kd) val = __ldg(reinterpret_cast<const uint4*>(A + (long long)r * lda + row0 + ch * 8));
    *reinterpret_cast<uint4*>(sA + r * RT_AS + ch * 8) = val;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = lane >> 3, i = lane & 7, g = lane >> 2, t = lane & 3;
  uint32_t af[NKS][4];
#pragma unroll
  for (int ks = 0; ks < NKS; ++ks) {
    const __half* asrc = sA + (ks * 16 + (j >> 1) * 8 + i) * RT_AS + (j & 1) * 8;
    ra_ldsm_x4_t(smem_u32(asrc), af[ks][0], af[ks][1], af[ks][2], af[ks][3]);
  }
  const float f = alpha * factor[0];
  const long long r_lo = row0 + g;
  for (int q = blockIdx.y * 4 + warp; q < nq; q += 4 * gridDim.y) {   // gridDim.y column splits: shorter-lived CTAs
    uint4 pc[12];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long r = r_lo + 8 * h;
      const long long off = (r < Kd ? r : r_lo) * N + q * 32 + t * 4;
      pc[h * 6 + 0] = *reinterpret_cast<const uint4*>(w + off);
      pc[h * 6 + 1] = *reinterpret_cast<const uint4*>(w + off + 16);
      pc[h * 6 + 2] = *reinterpret_cast<const uint4*>(m + off);
      pc[h * 6 + 3] = *reinterpret_cast<const uint4*>(m + off + 16);
      pc[h * 6 + 4] = *reinterpret_cast<const uint4*>(v + off);
      pc[h * 6 + 5] = *reinterpret_cast<const uint4*>(v + off + 16);
    }
    float acc[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.f;
    const uint4* gq = Gf + (size_t)q * nks * 64 + lane;
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      const uint4 b01 = __ldg(gq + ks * 64), b23 = __ldg(gq + ks * 64 + 32);
      ra_mma16816(acc[0], af[ks], b01.x, b01.y);
      ra_mma16816(acc[1], af[ks], b01.z, b01.w);
      ra_mma16816(acc[2], af[ks], b23.x, b23.y);
      ra_mma16816(acc[3], af[ks], b23.z, b23.w);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long r = r_lo + 8 * h;
      if (r >= Kd) continue;
      const long long off = r * N + q * 32 + t * 4;
      float gr[8];
#pragma unroll
      for (int s = 0; s < 4; ++s) { gr[2 * s] = acc[s][2 * h] * f; gr[2 * s + 1] = acc[s][2 * h + 1] * f; }
      float* wv = reinterpret_cast<float*>(&pc[h * 6 + 0]);
      float* mv = reinterpret_cast<float*>(&pc[h * 6 + 2]);
      float* vv = reinterpret_cast<float*>(&pc[h * 6 + 4]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        mv[e] = b1 * mv[e] + (1.f - b1) * gr[e];
        vv[e] = b2 * vv[e] + (1.f - b2) * gr[e] * gr[e];
        wv[e] -= adam_update(mv[e], vv[e], lr_t, eps);
      }
      *reinterpret_cast<uint4*>(m + off) = pc[h * 6 + 2];
      *reinterpret_cast<uint4*>(m + off + 16) = pc[h * 6 + 3];
      *reinterpret_cast<uint4*>(v + off) = pc[h * 6 + 4];
      *reinterpret_cast<uint4*>(v + off + 16) = pc[h * 6 + 5];
      *reinterpret_cast<uint4*>(w + off) = pc[h * 6 + 0];
      *reinterpret_cast<uint4*>(w + off + 16) = pc[h * 6 + 1];
      if (w16 != nullptr) {
        __half* dst = w16 + r * ldw16 + q * 32 + t * 4;
        *reinterpret_cast<uint2*>(dst) = make_uint2(pack_half2(wv[0], wv[1]), pack_half2(wv[2], wv[3]));
        *reinterpret_cast<uint2*>(dst + 16) = make_uint2(pack_half2(wv[4], wv[5]), pack_half2(wv[6], wv[7]));
      }
    }
  }
}

// Start-of-step latch of the overflow flag: `skipped` counts the steps whose update was dropped (sticky, for
// reporting), `flag` is cleared so that one non-finite gradient norm skips exactly one optimiser step.
__global__ void step_begin_kernel(int* __restrict__ flag, int* __restrict__ skipped) {
  if (*flag) { *skipped += 1; *flag = 0; }
}

int step_begin(int* flag, int* skipped, cudaStream_t st) {
  step_begin_kernel<<<1, 1, 0, st>>>(flag, skipped);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

// norm = alpha * sqrt(sum_ij GA_ij * GG_ij);  factor = clip / max(norm, clip)  (tf.clip_by_norm, utils.py:181-188)
__global__ void __launch_bounds__(256) rank_grad_clip_kernel(const float* __restrict__ ga, const float* __restrict__ gg, int n,
                                                             float alpha, float clip, float* __restrict__ factor,
                                                             float* __restrict__ norm, int* __restrict__ flag) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)ga[i] * (double)gg[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int i = 0; i < 8; ++i) r += red[i];
    const float nrm = alpha * (float)sqrt(r > 0.0 ? r : 0.0);
    if (!isfinite(nrm) || !(r == r)) atomicExch(flag, 1);
    norm[0] = nrm;
    factor[0] = clip > 0.f ? clip / fmaxf(nrm, clip) : 1.f;
  }
}

int rank_grad_clip(const float* gram_a, const float* gram_g, int R, float alpha, float clip, float* factor, float* norm,
                   int* flag, cudaStream_t st) {
  rank_grad_clip_kernel<<<1, 256, 0, st>>>(gram_a, gram_g, R * R, alpha, clip, factor, norm, flag);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

static size_t rank_adam_smem(int R, int N) {
  const int Rp = (R + 15) & ~15, nthr = N >= 512 ? 512 : 256;
  return (size_t)12 * nthr * 16 + (size_t)Rp * (N + 8) * 2 + 2 * (size_t)Rp * (RA_ROWS + 8) * 2;
}
// 0 = this (rank, width) is not supported (the caller keeps the dense gradient path)
size_t rank_adam_workspace_bytes(int R, int N) {
  if (R < 1 || R > 16 * RA_MAX_KS || N < 32 || N % 32 != 0 || rank_adam_smem(R, N) > 226 * 1024) return 0;
  return (size_t)((R + 15) & ~15) * N * 2;     // either layout of G^T (permuted rows / fragment order): Rp x N halves
}

int rank_adam_step(const __half* a16, long long lda, const __half* g16, long long ldg, int R, long long Kd, int N,
                   float alpha, const float* factor, const int* flag, float* w, float* m, float* v, __half* w16,
                   long long ldw16, float lr_t, const float* lr_dev, int tiled, float b1, float b2, float eps, void* workspace,
                   size_t workspace_bytes, cudaStream_t st) {
  LPM_REQUIRE(R >= 1 && R <= 16 * RA_MAX_KS, "rank_adam_step: rank (tower batch) must be in [1,%d] (got %d)", 16 * RA_MAX_KS, R);
  LPM_REQUIRE(N % 32 == 0 && N >= 32, "rank_adam_step: output width must be a multiple of 32 (got %d)", N);
  LPM_REQUIRE(lda % 8 == 0 && ldg % 8 == 0 && ldw16 % 8 == 0 && Kd % 8 == 0, "rank_adam_step: strides must be multiples of 8");
  LPM_REQUIRE(rank_adam_workspace_bytes(R, N) > 0, "rank_adam_step: R=%d x N=%d does not fit in shared memory", R, N);
  if (workspace == nullptr || workspace_bytes < rank_adam_workspace_bytes(R, N))
    return fail(LPM_ERR_WORKSPACE, "rank_adam_step: workspace of %zu bytes required", rank_adam_workspace_bytes(R, N));
  const int Rp = (R + 15) & ~15;
  if (tiled) {
    const int nks = Rp / 16;
    uint4* gf = reinterpret_cast<uint4*>(workspace);
    rank_gfrag_kernel<<<((N / 32) * nks * 64 + 255) / 256, 256, 0, st>>>(g16, ldg, R, nks, N, gf);
    const long long nblk = (Kd + 15) / 16;
    LPM_REQUIRE(nblk <= 0x7fffffffLL, "rank_adam_step: too many row blocks");
    int ysplit = tiled < 1 ? 1 : tiled;
    if (ysplit > (N / 32 + 3) / 4) ysplit = (N / 32 + 3) / 4;
    const dim3 tgrid((unsigned)nblk, (unsigned)ysplit);
#define LPM_RT(NK) case NK: rank_adam_tile_kernel<NK><<<tgrid, 128, 0, st>>>(a16, lda, gf, R, Kd, N, alpha, factor, flag, w, m, \
                                                                           v, w16, ldw16, lr_t, lr_dev, b1, b2, eps); break;
    switch (nks) { LPM_RT(1) LPM_RT(2) LPM_RT(3) LPM_RT(4) LPM_RT(5) LPM_RT(6) LPM_RT(7) LPM_RT(8) default: break; }
#undef LPM_RT
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  const int nthr = N >= 512 ? 512 : 256;
  const size_t smem = rank_adam_smem(R, N);
  static size_t attr = 0;
  if (smem > attr) {
    LPM_CUDA_CHECK(cudaFuncSetAttribute(rank_adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  __half* gp = reinterpret_cast<__half*>(workspace);
  rank_permute_kernel<<<(Rp * (N / 8) + 255) / 256, 256, 0, st>>>(g16, ldg, R, Rp, N, gp);
  const int npanels = (int)((Kd + RA_ROWS - 1) / RA_ROWS);
  const int grid = npanels < num_sms() ? npanels : num_sms();
  rank_adam_kernel<<<grid, nthr, smem, st>>>(a16, lda, gp, R, Kd, N, alpha, factor, flag, w, m, v, w16, ldw16, lr_t, lr_dev, b1, b2, eps);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Split-phase clip + Adam for a tensor whose rows are SHARDED over the data-parallel ranks (hidden1_weights):
//   shard_sqnorm : sum of squares of this rank's gradient shard -> sumsq[0]      (then all-reduced by the caller)
//   shard_adam   : factor = clip / max(sqrt(sumsq), clip) (tf.clip_by_norm over the WHOLE tensor, utils.py:181-188),
//                  Adam + fp16 shadow on the shard with the multi-tensor kernel (table of one tensor)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) shard_sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int i = 0; i < 8; ++i) r += red[i];
    out[0] = (float)r;
  }
}

__global__ void shard_clip_kernel(const float* __restrict__ sumsq, float clip, float* __restrict__ factor,
                                  float* __restrict__ norm, int* __restrict__ flag) {
  const float n = sqrtf(fmaxf(sumsq[0], 0.f));
  if (!isfinite(n) || !(sumsq[0] == sumsq[0])) atomicExch(flag, 1);
  norm[0] = n;
  factor[0] = clip > 0.f ? clip / fmaxf(n, clip) : 1.f;
}

int shard_sqnorm(const float* g, const float* p, const int* table, int n_chunks, const float* wd1, float* partial,
                 float* sumsq, cudaStream_t st) {
  mt_sqnorm_kernel<<<n_chunks, 256, 0, st>>>(g, p, table, wd1, partial);
  shard_sumsq_final_kernel<<<1, 256, 0, st>>>(partial, n_chunks, sumsq);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int shard_adam(float* p, const float* g, float* m, float* v, const int* table, int n_chunks, const float* wd1,
               const float* sumsq, float clip, float* factor, float* norm, int* flag, const unsigned long long* sh_ptr,
               const int* sh_cols, const long long* sh_ld, float lr_t, float b1, float b2, float eps, cudaStream_t st) {
  shard_clip_kernel<<<1, 1, 0, st>>>(sumsq, clip, factor, norm, flag);
  mt_adam_kernel<<<n_chunks, 256, 0, st>>>(p, g, m, v, table, wd1, factor, flag, sh_ptr, sh_cols, sh_ld, lr_t, nullptr, b1, b2, eps);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

// chunk0 / tensor0: first chunk / tensor of the range to update (chunks chunk0 .. chunk0 + n_chunks - 1 must be exactly the
// chunks of tensors tensor0 .. tensor0 + n_tensors - 1); all arrays are indexed by ABSOLUTE chunk / tensor ids
int adam_clip_step(float* p, const float* g, float* m, float* v, const int* table, int n_chunks,
                   const int* chunk_begin, int n_tensors, const float* wd, const unsigned long long* sh_ptr,
                   const int* sh_cols, const long long* sh_ld, float clip, float lr_t, const float* lr_dev, float b1, float b2,
                   float eps, float* partial, float* factor, float* norms, int* flag, cudaStream_t st, int chunk0, int tensor0) {
  mt_sqnorm_kernel<<<n_chunks, 256, 0, st>>>(g, p, table + 4 * (size_t)chunk0, wd, partial + chunk0);
  mt_clip_kernel<<<(n_tensors + 7) / 8, 256, 0, st>>>(partial, chunk_begin, tensor0, n_tensors, clip, factor, norms, flag);
  mt_adam_kernel<<<n_chunks, 256, 0, st>>>(p, g, m, v, table + 4 * (size_t)chunk0, wd, factor, flag, sh_ptr, sh_cols, sh_ld, lr_t, lr_dev,
                                           b1, b2, eps);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
