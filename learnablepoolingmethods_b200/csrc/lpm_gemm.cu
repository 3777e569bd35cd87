// Generic fp16 x fp16 -> fp32 GEMM for sm_100a: TMA (128B swizzle) -> smem ring -> tcgen05.mma with
// TMEM accumulators (double buffered) -> fused epilogue.  Persistent, warp specialised:
//   warp 0   : TMA producer (one lane)
//   warp 1   : TMEM allocator + MMA issuer (one lane)
//   warps 2-9: epilogue, two groups of four (TMEM -> registers -> bias / row-scale / ReLU / masks -> swizzled
//              smem slab -> TMA store; the groups take alternate 64-column slabs)
// Either operand may be K-major or MN-major in memory, so NN / NT / TN products (forward, dX, dW)
// need no transposes.  Batched (3-D tensor maps) and split-K (fp32 partials) variants included.
//
// TWO = 1 (large products): a 2-CTA cluster works on a 256 x 256 tile with tcgen05.mma.cta_group::2.  Each CTA
// loads its own 128 rows of A and HALF of the B tile (128 columns), so the L2 -> shared-memory feed per CTA drops
// from 48 KB to 32 KB per k-block (the 1-CTA kernel is bound by that feed), and the ring holds 6 stages.  The
// leader CTA issues every MMA; completion is multicast to both CTAs' barriers; each CTA runs the unchanged
// epilogue on its own 128 x 256 accumulator.
//
// Replaces the tf.matmul / tf.layers.dense / slim.fully_connected call sites of the hot path:
//   frame_level_models.py:2319,2347  transformer_utils.py:559-561,583-585,701-711
//   video_level_models.py:86-114 and their autodiff transposes.
#include <cstdlib>
#include <cstring>

#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

constexpr int BM = 128;
constexpr int BK = 64;

struct GemmKernelParams {
  int M, N, K;
  int batch, splits;
  int m_tiles, n_tiles;
  int kb_total, kb_per_split;
  int a_batched, b_batched;
  // epilogue
  void* out;
  int out_f32;
  long long ldc, out_batch_stride, out_split_stride;
  const float* bias;
  const float* row_scale;
  long long row_scale_batch_stride;
  int relu;
  int accumulate;
  float alpha;
  float* stat_sum;
  float* stat_sq;
  const __half* mask;   // optional: v = mask > 0 ? v : 0   (ReLU backward fused into the producing GEMM)
  long long ld_mask;
  const __half* add1;   // optional fp16 addends (fused residual-gradient sums)
  const __half* add2;
  long long ld_add;
  int tma_store;        // 1: epilogue stages 128-byte-row slabs in smem and writes them with TMA
  int gated;            // run the context-gating tail (GemmKernelParams::tail) after the last tile
  lpm_gating_tail tail;
  int nt_fast;          // tile order: 0 = m-tiles fastest (concurrent tiles share the B tile), 1 = n-tiles fastest (they share A)
};

template <int BN, int STAGES>     // BN = B columns staged by ONE CTA per k-block
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SLAB_BYTES = BM * 128;                 // per epilogue group: 128 rows x 128 B (64 fp16 / 32 fp32 columns)
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;        // one staging slab per epilogue group
  static constexpr int BAR_OFF = STG_OFF + 2 * SLAB_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024;  // + alignment slack
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

// Epilogue math on one 32-column chunk held in registers: v = acc*rs (+bias) (ReLU) (+add1 +add2) (mask).
__device__ __forceinline__ void epi_chunk(const uint32_t* r, float* v, const GemmKernelParams& p, float rs, int col0,
                                          int row, bool row_ok, int bz) {
  const bool full = (col0 + 32 <= p.N);                    // warp-uniform
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * rs;
  if (p.bias != nullptr) {
    if (full) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + i);
        v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + min(col0 + i, p.N - 1));
    }
  }
  if (p.relu) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (p.add1 != nullptr || p.mask != nullptr) {
    const long long rofs = (long long)bz * p.out_batch_stride;
    if (p.add1 != nullptr && row_ok) {
      const __half* a1 = p.add1 + rofs + (long long)row * p.ld_add + col0;
      const __half* a2 = p.add2 ? p.add2 + rofs + (long long)row * p.ld_add + col0 : nullptr;
      if (full && (p.ld_add & 7) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 w = __ldg(reinterpret_cast<const uint4*>(a1) + i);
          const __half2* h = reinterpret_cast<const __half2*>(&w);
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); v[8 * i + 2 * j] += f.x; v[8 * i + 2 * j + 1] += f.y; }
        }
        if (a2 != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(a2) + i);
            const __half2* h = reinterpret_cast<const __half2*>(&w);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); v[8 * i + 2 * j] += f.x; v[8 * i + 2 * j + 1] += f.y; }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < p.N) {
            v[i] += __half2float(a1[i]);
            if (a2 != nullptr) v[i] += __half2float(a2[i]);
          }
      }
    }
    if (p.mask != nullptr && row_ok) {
      const __half* mk = p.mask + rofs + (long long)row * p.ld_mask + col0;
      if (full && (p.ld_mask & 7) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 w = __ldg(reinterpret_cast<const uint4*>(mk) + i);
          const __half2* h = reinterpret_cast<const __half2*>(&w);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            if (!(f.x > 0.f)) v[8 * i + 2 * j] = 0.f;
            if (!(f.y > 0.f)) v[8 * i + 2 * j + 1] = 0.f;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < p.N && !(__half2float(mk[i]) > 0.f)) v[i] = 0.f;
      }
    }
  }
}

// direct (non-TMA) store of one chunk: unaligned pitch / accumulate
__device__ __forceinline__ void epi_store_direct(const float* v, const GemmKernelParams& p, long long obase, int col0) {
  const bool full = (col0 + 32 <= p.N);
  if (p.out_f32) {
    float* o = reinterpret_cast<float*>(p.out) + obase + col0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 w = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        float4* dst = reinterpret_cast<float4*>(o) + i;
        if (p.accumulate) { const float4 old = *dst; w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w; }
        *dst = w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) o[i] = p.accumulate ? o[i] + v[i] : v[i];
    }
  } else {
    __half* o = reinterpret_cast<__half*>(p.out) + obase + col0;
    if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
        w.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
        w.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
        w.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
        reinterpret_cast<uint4*>(o)[i] = w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) o[i] = __float2half_rn(v[i]);
    }
  }
}


// ----------------------------------------------------------------------------------------------
// K3 tail: split-K reduction + bias + context gating, run by the 256 epilogue threads of every CTA after its last
// tile (lpm_gemm_splitk_gated_fwd, include/lpm_b200.h; frame_level_models.py:2319-2368).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier_one_thread(int* counter, int expected) {
  __threadfence();
  atomicAdd(counter, 1);
  uint32_t spins = 0;
  while (*reinterpret_cast<volatile int*>(counter) < expected) {
    __nanosleep(64);
    if (++spins > (1u << 24)) { printf("lpm: gated GEMM grid barrier timeout block=%d\n", (int)blockIdx.x); __trap(); }
  }
  __threadfence();
}
__device__ __forceinline__ void store_act16(__half* dst16, int split3, int B, int H, int r, int c, float v) {
  const __half hi = __float2half_rn(v);
  if (split3) {
    __half* d = dst16 + (size_t)r * 3 * H + c;
    d[0] = hi; d[H] = __float2half_rn(v - __half2float(hi)); d[2 * H] = hi;
  } else {
    dst16[(size_t)r * H + c] = hi;
  }
}

// tid: 0..255 (epilogue threads); scratch: >= 16 KB of shared memory (the idle operand ring)
__device__ void gating_tail(const GemmKernelParams& p, int tid, uint8_t* scratch) {
  const lpm_gating_tail& t = p.tail;
  const int B = p.M, H = p.N, G = (int)gridDim.x;
  int cw = (H + G - 1) / G;
  cw = (cw + 3) & ~3;                                   // columns per CTA, a multiple of 4
  const int c0 = (int)blockIdx.x * cw;
  const int nc = max(0, min(cw, H - c0));               // this CTA's columns (0 for the last CTAs)
  const int lane = tid & 31, warp = tid >> 5;
  float* sred = reinterpret_cast<float*>(scratch);      // [KG][256][4] partial sums of the k-groups
  float* sg = sred + 8 * 256 * 4;                       // [B][cw] gate pre-activations, then statistics
  float* sact = sg + (size_t)B * cw + 2 * cw + (size_t)H * 4;   // [B][cw] this CTA's slice of the hidden activation

  // ---- barrier 1: every partial tile of this launch is in global memory ----
  named_barrier(5, 256);
  if (tid == 0) grid_barrier_one_thread(t.counters + 0, G);
  named_barrier(5, 256);

  // ---- phase A: this CTA's CONTIGUOUS chunk of the [B x H] activation = bias + sum of the partials.  Consecutive threads
  //      read consecutive 16-byte pieces of a slab (full 128-byte lines; a column slice per CTA would touch a 32-byte
  //      sector per 16 bytes); the slabs are split over k-groups of threads and added in a fixed order ----
  const float* part = reinterpret_cast<const float*>(p.out);
  {
    const int n4 = B * H / 4;                            // float4 pieces of the activation (ldc == H: slabs are contiguous)
    const int per = (n4 + G - 1) / G;
    const int i0 = (int)blockIdx.x * per, npos = max(0, min(per, n4 - i0));
    int KG = npos > 0 ? 256 / npos : 1; if (KG > 8) KG = 8; if (KG < 1) KG = 1;
    for (int pbase = 0; pbase < npos; pbase += 256 / KG) {      // one pass unless the chunk exceeds 256 pieces
      const int np = min(256 / KG, npos - pbase);
      const int pos = tid % (256 / KG), kg = tid / (256 / KG);
      if (pos < np && kg < KG) {
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        auto sum_set = [&](const float* base, int n_slabs, long long stride, float4& acc) {
          const float4* src = reinterpret_cast<const float4*>(base) + i0 + pbase + pos;
          for (int k = kg; k < n_slabs; k += 8 * KG) {
            float4 x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int kk = k + u * KG;
              x[u] = kk < n_slabs ? __ldcg(src + (size_t)kk * (stride / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
          }
        };
        sum_set(part, p.splits, p.out_split_stride, a0);
        if (t.part2 != nullptr) sum_set(t.part2, t.splits2, t.split_stride2, a1);
        *reinterpret_cast<float4*>(sred + ((size_t)kg * 256 + pos) * 4) =
            make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
      }
      named_barrier(5, 256);
      if (tid < np) {
        float4 sum = *reinterpret_cast<const float4*>(sred + (size_t)tid * 4);
        for (int g = 1; g < KG; ++g) {
          const float4 x = *reinterpret_cast<const float4*>(sred + ((size_t)g * 256 + tid) * 4);
          sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
        }
        const int e = (i0 + pbase + tid) * 4;            // first element: row e / H, columns e % H .. +3 (H % 4 == 0)
        const int r = e / H, c = e - r * H;
        float v[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (t.bias) v[j] += __ldg(t.bias + c + j);
          if (t.act16) store_act16(reinterpret_cast<__half*>(t.act16), t.act_split3, B, H, r, c + j, v[j]);
        }
        *reinterpret_cast<float4*>(t.act32 + e) = make_float4(v[0], v[1], v[2], v[3]);
      }
      named_barrier(5, 256);
    }
  }

  // ---- barrier 2: the whole hidden activation is in global memory ----
  if (tid == 0) grid_barrier_one_thread(t.counters + 1, G);
  named_barrier(5, 256);

  if (nc > 0) {
    for (int i = tid; i < B * nc; i += 256) {           // this CTA's columns of the activation (for the statistics and the gate)
      const int r = i / nc, j = i - r * nc;
      sact[(size_t)r * cw + j] = __ldcg(t.act32 + (size_t)r * H + c0 + j);
    }
    // ---- phase B: g[r][j] = sum_k hidden[r][k] * Wg[k][c0 + j] in fp32: a warp per row, lanes over k.  The CTA's
    //      column slice of Wg (H x 4 floats per pass) is staged in shared memory: with the operand ring taking the SM's
    //      shared memory the L1 cannot hold its 512 lines, and every row would re-fetch them from L2 ----
    float* swg = sg + (size_t)B * cw + 2 * cw;          // [H][4]
    for (int cg = 0; cg < nc; cg += 4) {
      named_barrier(5, 256);
      for (int k = tid; k < H; k += 256)
        *reinterpret_cast<float4*>(swg + (size_t)k * 4) = __ldg(reinterpret_cast<const float4*>(t.wg + (size_t)k * t.ldwg + c0 + cg));
      named_barrier(5, 256);
      for (int r = warp; r < B; r += 8) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int kb = 0; kb < H; kb += 512) {           // 16 loads of the row in flight per lane (L2 round trips, not FMAs, cost)
          float h[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int k = kb + u * 32 + lane;
            h[u] = k < H ? __ldcg(t.act32 + (size_t)r * H + k) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int k = kb + u * 32 + lane;
            if (k < H) {
              const float4 w = *reinterpret_cast<const float4*>(swg + (size_t)k * 4);
              a0 = fmaf(h[u], w.x, a0); a1 = fmaf(h[u], w.y, a1); a2 = fmaf(h[u], w.z, a2); a3 = fmaf(h[u], w.w, a3);
            }
          }
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
        if (lane == 0) {
          float* d = sg + (size_t)r * cw + cg;
          d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
        }
      }
    }
    named_barrier(5, 256);
    // ---- batch norm over the rows of each owned column (frame_level_models.py:2354-2362), then the gate ----
    float* sstat = sg + (size_t)B * cw;                 // [cw][2]: mean, var
    for (int j = warp; j < nc; j += 8) {                 // a warp per column, lanes over the rows, fixed reduction order
      const int c = c0 + j;
      const float dg = t.wg_diag ? t.wg_diag[c] : 0.f;
      float mean, var;
      if (t.training) {
        double s = 0.0, q = 0.0;
        for (int r = lane; r < B; r += 32) {
          const float v = sg[(size_t)r * cw + j] - dg * sact[(size_t)r * cw + j];
          s += v; q += (double)v * v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        const double m = s / B;
        double vv = q / B - m * m;
        if (vv < 0.0) vv = 0.0;
        mean = (float)m; var = (float)vv;
        if (lane == 0) {
          const double corr = B > 1 ? (double)B / (B - 1) : 1.0;
          t.moving_mean[c] = t.moving_mean[c] * t.decay + (float)m * (1.f - t.decay);
          t.moving_var[c] = t.moving_var[c] * t.decay + (float)(vv * corr) * (1.f - t.decay);
        }
      } else {
        mean = t.moving_mean[c]; var = t.moving_var[c];
      }
      if (lane == 0) {
        sstat[2 * j] = mean; sstat[2 * j + 1] = var;
        if (t.save_mean) { t.save_mean[c] = mean; t.save_rstd[c] = rsqrtf(var + t.eps); }
      }
    }
    named_barrier(5, 256);
    for (int i = tid; i < B * nc; i += 256) {
      const int r = i / nc, j = i - r * nc, c = c0 + j;
      const float dg = t.wg_diag ? t.wg_diag[c] : 0.f;
      const float rstd = rsqrtf(sstat[2 * j + 1] + t.eps);
      const float sc = t.gamma[c] * rstd, sh = t.beta[c] - sstat[2 * j] * sc;
      const float a = sact[(size_t)r * cw + j];
      const float g = sg[(size_t)r * cw + j];
      if (t.g_sum) t.g_sum[(size_t)r * H + c] = g;
      const float v = (g - dg * a) * sc + sh;
      const float o = a / (1.f + __expf(-v));
      t.out32[(size_t)r * H + c] = o;
      if (t.out16) store_act16(reinterpret_cast<__half*>(t.out16), t.out_split3, B, H, r, c, o);
    }
  }
  // ---- leave the counters at zero for the next launch: the last CTA through resets them ----
  named_barrier(5, 256);
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(t.counters + 2, 1) == G - 1) {
      t.counters[0] = 0; t.counters[1] = 0; t.counters[2] = 0;
      __threadfence();
    }
  }
}

// ---- 2-CTA (cta_group::2) primitives ----
__device__ __forceinline__ uint32_t g2_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t g2_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void g2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into this CTA's shared memory, completion bytes signalled on a barrier of the CTA pair (cluster address)
__device__ __forceinline__ void g2_tma_load_3d(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                               int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr),
        "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void g2_umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the issued MMAs retire
__device__ __forceinline__ void g2_umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void g2_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void g2_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 22)) {
      printf("lpm: gemm pair mbarrier timeout block=%d thread=%d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
template <uint32_t kCols>
__device__ __forceinline__ void g2_tmem_alloc(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void g2_tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols) : "memory");
}

template <int BN, int STAGES, int A_MN, int B_MN, int TWO>
__global__ void __launch_bounds__(320, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_c, const GemmKernelParams p) {
  constexpr int BNL = TWO ? BN / 2 : BN;      // B columns staged by this CTA
  constexpr int MT = TWO ? 2 : 1;             // 128-row tiles per work item (one per CTA of the pair)
  using L = GemmSmem<BNL, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = warp_index_uniform();
  const int lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.tma_store) tma_prefetch_desc(&tmap_c);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8 * MT);       // the leader's barrier also collects the peer's epilogue warps
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if (TWO) g2_tmem_alloc<TMEM_COLS>(tmem_slot); else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) g2_cluster_sync();                  // both CTAs' barriers and TMEM exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                  // programmatic dependent launch: the set-up above ran under the predecessor
  const int crank = TWO ? (int)g2_ctarank() : 0;
  const int work_id = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int work_stride = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  const int m_units = (p.m_tiles + MT - 1) / MT;
  const int tiles_per_batch = m_units * p.n_tiles;
  const int total_tiles = tiles_per_batch * p.batch * p.splits;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = work_id; tile < total_tiles; tile += work_stride) {
        int t = tile, mt, nt;
        if (p.nt_fast) { nt = t % p.n_tiles; t /= p.n_tiles; mt = t % m_units; t /= m_units; }
        else           { mt = t % m_units;   t /= m_units;   nt = t % p.n_tiles; t /= p.n_tiles; }
        const int bz = t % p.batch;   t /= p.batch;
        const int sp = t;
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int m0 = (mt * MT + crank) * BM, n0 = nt * BN + crank * BNL;
        const int za = p.a_batched ? bz : 0, zb = p.b_batched ? bz : 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (!TWO) {
            mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
            if (A_MN == 0) {
              tma_load_3d(sa, &tmap_a, &full_bar[stage], kb * BK, m0, za);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_3d(sa + j * 8192, &tmap_a, &full_bar[stage], m0 + 64 * j, kb * BK, za);
            }
            if (B_MN == 0) {
              tma_load_3d(sb, &tmap_b, &full_bar[stage], kb * BK, n0, zb);
            } else {
#pragma unroll
              for (int j = 0; j < BNL / 64; ++j)
                tma_load_3d(sb + j * 8192, &tmap_b, &full_bar[stage], n0 + 64 * j, kb * BK, zb);
            }
          } else {
            // both CTAs' bytes complete on the LEADER's barrier, which expects the whole pair's stage
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
            const uint32_t fb = g2_mapa(smem_u32(&full_bar[stage]), 0);
            if (A_MN == 0) {
              g2_tma_load_3d(sa, &tmap_a, fb, kb * BK, m0, za);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                g2_tma_load_3d(sa + j * 8192, &tmap_a, fb, m0 + 64 * j, kb * BK, za);
            }
            if (B_MN == 0) {
              g2_tma_load_3d(sb, &tmap_b, fb, kb * BK, n0, zb);
            } else {
#pragma unroll
              for (int j = 0; j < BNL / 64; ++j)
                g2_tma_load_3d(sb + j * 8192, &tmap_b, fb, n0 + 64 * j, kb * BK, zb);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    // The whole warp runs the loop on warp-uniform values and an elected lane issues each instruction (umma_f16_w,
    // lpm_common.cuh): inside an `if (lane == 0)` region every tcgen05.mma was wrapped in a ~17-instruction uniform-
    // register waterfall, ~80 clk of issue per MMA.
    if (crank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM * MT, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t smem_base = smem_u32(smem);
      for (int tile = work_id; tile < total_tiles; tile += work_stride) {
        const int sp = tile / (tiles_per_batch * p.batch);
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        if (TWO) g2_wait_cluster(&tempty_bar[acc], acc_phase ^ 1); else mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
          const uint32_t sb = sa + L::A_BYTES;
          const uint64_t a0 = A_MN ? umma_smem_desc(sa, 8192, 1024) : umma_smem_desc(sa, 16, 1024);
          const uint64_t b0 = B_MN ? umma_smem_desc(sb, 8192, 1024) : umma_smem_desc(sb, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks) {
              const uint64_t adesc = a0 + ((A_MN ? ks * 2048 : ks * 32) >> 4);
              const uint64_t bdesc = b0 + ((B_MN ? ks * 2048 : ks * 32) >> 4);
              if (TWO) g2_umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
              else umma_f16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs of a pair) once these MMAs retire
            if (TWO) g2_umma_commit(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (elect_one()) { if (TWO) g2_umma_commit(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]); }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------- epilogue -----------------------------------
    // 8 warps = 2 groups x 4 TMEM lane quarters; the groups take alternate 64-column slabs of the tile.
    const int quarter = warp & 3;            // TMEM sub-partition this warp may read
    const int grp = (warp - 2) >> 2;         // 0 / 1
    const bool leader = ((warp - 2) & 3) == 0 && lane == 0;
    const int bar_a = 1 + 2 * grp, bar_b = 2 + 2 * grp;
    uint8_t* slab = smem + L::STG_OFF + grp * L::SLAB_BYTES;
    const int trow = quarter * 32 + lane;    // row inside the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader[2] = {TWO ? g2_mapa(smem_u32(&tempty_bar[0]), 0) : 0u,
                                       TWO ? g2_mapa(smem_u32(&tempty_bar[1]), 0) : 0u};
    for (int tile = work_id; tile < total_tiles; tile += work_stride) {
      int t = tile, mu, nt;
      if (p.nt_fast) { nt = t % p.n_tiles; t /= p.n_tiles; mu = t % m_units; t /= m_units; }
      else           { mu = t % m_units;   t /= m_units;   nt = t % p.n_tiles; t /= p.n_tiles; }
      const int bz = t % p.batch;   t /= p.batch;
      const int sp = t;
      const int mt = mu * MT + crank;          // this CTA's 128-row tile
      const int row = mt * BM + trow;
      const bool row_ok = row < p.M;
      const float rs = (p.row_scale != nullptr && row_ok)
                           ? __ldg(p.row_scale + (long long)bz * p.row_scale_batch_stride + row) * p.alpha
                           : p.alpha;
      const long long obase = (long long)sp * p.out_split_stride + (long long)bz * p.out_batch_stride +
                              (long long)row * p.ldc;
      float s_sum = 0.f, s_sq = 0.f;
      const int cols_here = min(BN, p.N - nt * BN);
      const int slab_cols = (p.tma_store && p.out_f32) ? 32 : 64;   // one 128-byte-row TMA box per slab
      const int n_slabs = (cols_here + slab_cols - 1) / slab_cols;
      const int my_last = ((n_slabs - 1 - grp) >= 0) ? grp + ((n_slabs - 1 - grp) / 2) * 2 : -1;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (my_last < 0) {                      // this group has no slab in a narrow tail tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (TWO) g2_arrive_cluster(tempty_leader[acc]); else mbar_arrive(&tempty_bar[acc]); }
      }
#pragma unroll 1
      for (int sl = grp; sl < n_slabs; sl += 2) {
        const int col0 = nt * BN + sl * slab_cols;
        const bool second = (slab_cols == 64) && (col0 + 32 < p.N);   // second 32-column chunk has valid columns
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + sl * slab_cols);
        tmem_ld32(taddr, r0);
        if (second) tmem_ld32(taddr + 32, r1);
        tmem_ld_wait();
        if (sl == my_last) {
          // accumulator fully read by this warp: hand the TMEM buffer back before finishing the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (TWO) g2_arrive_cluster(tempty_leader[acc]); else mbar_arrive(&tempty_bar[acc]); }
        }
        if (p.tma_store) {
          if (leader) bulk_wait_read<0>();                   // this group's previous slab has left smem
          named_barrier(bar_a, 128);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1 && !second) break;
          float v[32];
          const int c0 = col0 + 32 * h;
          epi_chunk(h == 0 ? r0 : r1, v, p, rs, c0, row, row_ok, bz);
          if (p.stat_sum != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < p.N) { s_sum += v[i]; s_sq += v[i] * v[i]; }
          }
          if (p.tma_store) {
            if (p.out_f32) {                                  // 32 fp32 columns = one 128-byte row
#pragma unroll
              for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(slab + trow * 128 + ((i ^ (trow & 7)) << 4)) =
                    make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 w;
                w.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
                w.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
                w.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
                w.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
                *reinterpret_cast<uint4*>(slab + trow * 128 + (((h * 4 + i) ^ (trow & 7)) << 4)) = w;
              }
            }
          } else if (p.out != nullptr && row_ok) {
            epi_store_direct(v, p, obase, c0);
          }
        }
        if (p.tma_store) {
          fence_proxy_async_smem();
          named_barrier(bar_b, 128);
          if (leader) {
            tma_store_3d(&tmap_c, slab, col0, mt * BM, sp * p.batch + bz);
            bulk_commit();
          }
        }
      }
      if (p.stat_sum != nullptr && row_ok) {
        const long long si = ((long long)(sp * p.batch + bz) * (p.n_tiles * 2) + nt * 2 + grp) * p.M + row;
        p.stat_sum[si] = s_sum;
        p.stat_sq[si] = s_sq;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_store && leader) bulk_wait<0>();   // all output slabs written before exit
    if (p.gated) gating_tail(p, (int)threadIdx.x - 64, smem);   // K3: reduction + context gating, all CTAs of the launch
  }

  tc_fence_before();
  __syncthreads();
  if (TWO) g2_cluster_sync();                  // no CTA of the pair leaves (or frees TMEM) while the other still works
  if (warp == 1) {
    tc_fence_after();
    if (TWO) g2_tmem_dealloc<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------------
// host launcher
// ----------------------------------------------------------------------------------------------
template <int BN, int STAGES, int A_MN, int B_MN, int TWO>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmKernelParams& p,
                       cudaStream_t st) {
  using L = GemmSmem<TWO ? BN / 2 : BN, STAGES>;
  auto kern = gemm_f16_kernel<BN, STAGES, A_MN, B_MN, TWO>;
  static bool attr_set = false;
  if (!attr_set) {
    LPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  const int m_units = TWO ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int total = m_units * p.n_tiles * p.batch * p.splits;
  // Programmatic dependent launch (LPM_PDL=1).  Measured at config 1 (gpurun r2ag): eager inference forward 0.94 -> 0.91 ms,
  // graph-replayed inference unchanged (0.897 ms), graph-replayed training step 3.51 -> 3.54 ms -- inside a graph the
  // programmatic edge buys nothing (the 148-CTA kernels cannot overlap anyway: one CTA per SM by shared memory) and costs a
  // little.  Off by default because the training step is the headline.
  static const bool pdl = getenv("LPM_PDL") != nullptr && getenv("LPM_PDL")[0] == '1';
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl) {   // the kernel waits (pdl_wait) after its block-local set-up
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  // Smallest grid that needs the same number of rounds (LPM_GEMM_MIN_GRID=0: always every SM): 320 pair tiles take five
  // rounds on 74 pairs and on 64, and the SMs left free go to whatever runs next to the product (the forked optimiser
  // branch, the column sums of the bias gradients).  Same tiles, same results.
  static const bool min_grid = !(getenv("LPM_GEMM_MIN_GRID") != nullptr && getenv("LPM_GEMM_MIN_GRID")[0] == '0');
  auto fit = [&](int slots) {
    if (total <= slots) return total;
    if (!min_grid || p.gated) return slots;
    const int rounds = (total + slots - 1) / slots;
    return (total + rounds - 1) / rounds;
  };
  if (!TWO) {
    cfg.gridDim = dim3(fit(num_sms()));
  } else {
    const int pairs = fit(num_sms() / 2);
    cfg.gridDim = dim3(2 * pairs);
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  LPM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, p));
  return LPM_OK;
}

template <int BN, int STAGES, int TWO>
static int dispatch_major(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                          const GemmKernelParams& p, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_gemm<BN, STAGES, 0, 0, TWO>(ta, tb, tc, p, st);
  if (!a_mn && b_mn) return launch_gemm<BN, STAGES, 0, 1, TWO>(ta, tb, tc, p, st);
  if (a_mn && !b_mn) return launch_gemm<BN, STAGES, 1, 0, TWO>(ta, tb, tc, p, st);
  return launch_gemm<BN, STAGES, 1, 1, TWO>(ta, tb, tc, p, st);
}

static int g_gemm_pair_mode = 1;   // 0: never use CTA pairs, 1: automatic
void gemm_set_pair_mode(int mode) { g_gemm_pair_mode = mode; }

int gemm_pick_bn(int N) {
  if (N >= 256 || N > 192) return 256;
  if (N > 64) return 128;
  return 64;
}

int gemm_f16(const GemmArgs& g, cudaStream_t st, const lpm_gating_tail* tail) {
  LPM_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.batch > 0, "gemm: bad dims M=%d N=%d K=%d batch=%d", g.M, g.N, g.K, g.batch);
  LPM_REQUIRE(g.lda % 8 == 0 && g.ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 (TMA 16B strides), got %lld %lld", g.lda, g.ldb);
  LPM_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0, "gemm: A/B must be 16B aligned");
  LPM_REQUIRE(g.a_batch_stride % 8 == 0 && g.b_batch_stride % 8 == 0, "gemm: batch strides must be multiples of 8");
  const int BN = g.force_bn ? g.force_bn : gemm_pick_bn(g.N);
  LPM_REQUIRE(BN == 256 || BN == 128 || BN == 64, "gemm: unsupported BN %d", BN);

  GemmKernelParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K; p.batch = g.batch;
  p.m_tiles = (g.M + BM - 1) / BM;
  p.n_tiles = (g.N + BN - 1) / BN;
  p.kb_total = (g.K + BK - 1) / BK;
  int splits = g.splits > 0 ? g.splits : 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  LPM_REQUIRE(p.splits == 1 || (g.out_f32 && g.out_split_stride > 0 && !g.bias && !g.relu && !g.stat_sum),
              "gemm: split-K needs fp32 partial output with a split stride and a plain epilogue");
  p.a_batched = g.a_batch_stride != 0; p.b_batched = g.b_batch_stride != 0;
  p.out = g.out; p.out_f32 = g.out_f32; p.ldc = g.ldc; p.out_batch_stride = g.out_batch_stride;
  p.out_split_stride = g.out_split_stride;
  p.bias = g.bias; p.row_scale = g.row_scale; p.row_scale_batch_stride = g.row_scale_batch_stride;
  p.relu = g.relu; p.accumulate = g.accumulate; p.alpha = g.alpha;
  p.stat_sum = g.stat_sum; p.stat_sq = g.stat_sq;
  p.mask = reinterpret_cast<const __half*>(g.mask); p.ld_mask = g.ld_mask;
  p.add1 = reinterpret_cast<const __half*>(g.add1); p.add2 = reinterpret_cast<const __half*>(g.add2); p.ld_add = g.ld_add;
  LPM_REQUIRE(!(g.add2 && !g.add1), "gemm: add2 requires add1");
  LPM_REQUIRE(!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
  LPM_REQUIRE(!g.add1 || ((reinterpret_cast<uintptr_t>(g.add1) & 15) == 0 && (!g.add2 || (reinterpret_cast<uintptr_t>(g.add2) & 15) == 0)), "gemm: addends must be 16-byte aligned");
  LPM_REQUIRE(!g.mask || (reinterpret_cast<uintptr_t>(g.mask) & 15) == 0, "gemm: mask must be 16-byte aligned");

  CUtensorMap ta, tb;
  int rc;
  // A: K-major = memory [M][K]; MN-major = memory [K][M]
  if (!g.a_mn) rc = make_tmap_3d(&ta, g.A, 2, g.K, g.M, p.a_batched ? g.batch : 1, g.lda, g.a_batch_stride, BK, BM);
  else         rc = make_tmap_3d(&ta, g.A, 2, g.M, g.K, p.a_batched ? g.batch : 1, g.lda, g.a_batch_stride, 64, BK);
  if (rc) return rc;
  // 2-CTA pairs for products with at least one full wave of 256 x 256 tiles (each CTA then stages half of the B tile)
  const bool two = g_gemm_pair_mode != 0 && BN == 256 &&
                   (long long)((p.m_tiles + 1) / 2) * p.n_tiles * p.batch * p.splits >= num_sms() / 2 && p.m_tiles >= 2;
  if (!g.b_mn) rc = make_tmap_3d(&tb, g.B, 2, g.K, g.N, p.b_batched ? g.batch : 1, g.ldb, g.b_batch_stride, BK, two ? BN / 2 : BN);
  else         rc = make_tmap_3d(&tb, g.B, 2, g.N, g.K, p.b_batched ? g.batch : 1, g.ldb, g.b_batch_stride, 64, BK);
  if (rc) return rc;

  // TMA-store epilogue when the output is TMA-addressable: 16-byte aligned base and row pitch, no read-modify-write
  CUtensorMap tc;
  memset(&tc, 0, sizeof(tc));
  const int es = g.out_f32 ? 4 : 2;
  long long zstride = 0;
  bool z_ok = true;
  if (p.splits > 1 && p.batch > 1) { zstride = g.out_batch_stride; z_ok = (g.out_split_stride == g.out_batch_stride * p.batch); }
  else if (p.splits > 1) zstride = g.out_split_stride;
  else if (p.batch > 1) zstride = g.out_batch_stride;
  p.tma_store = (g.out != nullptr && !g.accumulate && z_ok && (reinterpret_cast<uintptr_t>(g.out) & 15) == 0 &&
                 (g.ldc * es) % 16 == 0 && (zstride * es) % 16 == 0 && !g.no_tma_store) ? 1 : 0;
  if (p.tma_store) {
    rc = make_tmap_3d(&tc, g.out, es, g.N, g.M, (uint64_t)p.splits * p.batch, g.ldc, zstride, g.out_f32 ? 32 : 64, BM);
    if (rc) return rc;
  }
  // Tile order.  With m-tiles fastest the tiles in flight share one B tile and stream distinct A rows, so A is read
  // n_tiles times unless it stays in L2 (FFN2 / the dX of FFN1 at config 1: A = 168 MB against 126 MB of L2, four
  // n-tiles -> 671 MB of DRAM reads, 105 us of a 120 us product); n-tiles fastest reads A once and re-reads the
  // (L2-resident) B instead.
  {
    const double a_bytes = 2.0 * g.M * g.K * (p.a_batched ? 1 : 1), b_bytes = 2.0 * g.N * g.K;
    static const int force = getenv("LPM_GEMM_NT_FAST") ? atoi(getenv("LPM_GEMM_NT_FAST")) : -1;
    p.nt_fast = force >= 0 ? force : ((a_bytes > 64e6 && b_bytes <= 48e6 && p.n_tiles > 1) ? 1 : 0);
  }
  if (tail != nullptr) {
    const long long tiles = (long long)p.m_tiles * p.n_tiles * p.batch * p.splits;
    LPM_REQUIRE(!two && p.m_tiles == 1 && p.batch == 1 && g.out_f32 && g.out_split_stride > 0 && g.ldc == g.N &&
                g.N % 4 == 0 && tiles <= num_sms() && g.alpha == 1.f,
                "gated GEMM: needs a split-K product of <= 128 rows with contiguous fp32 partials [splits][M][N] and at most "
                "one tile per SM (M=%d N=%d tiles=%lld)", g.M, g.N, tiles);
    LPM_REQUIRE(tail->counters && tail->act32 && tail->wg && tail->gamma && tail->beta && tail->moving_mean && tail->moving_var &&
                tail->out32 && tail->ldwg % 4 == 0 && (reinterpret_cast<uintptr_t>(tail->wg) & 15) == 0 &&
                (tail->part2 == nullptr || (reinterpret_cast<uintptr_t>(tail->part2) & 15) == 0),
                "gated GEMM: bad tail arguments");
    p.gated = 1;
    p.tail = *tail;
  }
  if (two) return dispatch_major<256, 6, 1>(g.a_mn, g.b_mn, ta, tb, tc, p, st);
  if (BN == 256) return dispatch_major<256, 4, 0>(g.a_mn, g.b_mn, ta, tb, tc, p, st);
  if (BN == 128) return dispatch_major<128, 6, 0>(g.a_mn, g.b_mn, ta, tb, tc, p, st);
  return dispatch_major<64, 8, 0>(g.a_mn, g.b_mn, ta, tb, tc, p, st);
}

int gemm_effective_splits(int K, int splits) {
  const int kb_total = (K + BK - 1) / BK;
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  const int per = (kb_total + splits - 1) / splits;
  return (kb_total + per - 1) / per;
}

}  // namespace lpm
