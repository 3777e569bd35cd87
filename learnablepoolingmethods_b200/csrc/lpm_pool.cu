// Fused NetVLAD pooling forward (K1) for sm_100a -- frame_level_models.py:2775-2822.
//
// One CTA per video.  The B x T x K assignment tensor lives only in TMEM / registers / shared memory
// (it is written to HBM only when the caller asks for it, i.e. for the training backward).
//
//   phase 1  S[t,k]   = sum_d X[t,d] Wc[d,k]          tcgen05.mma, X and Wc streamed by TMA in 64-wide
//                                                      d-chunks, both 128-frame tiles accumulate in TMEM
//   softmax  A[t,k]   = softmax_k(S*scale_k + shift_k) one thread per frame row, ONE pass over TMEM with a
//                                                      running max (64 logits per tcgen05.wait::ld); exp values
//                                                      are parked as fp16 in the swizzled smem operand tile
//                                                      (MN-major UMMA operand) and rescaled in place; masked
//                                                      for t >= T / invalid frames
//   phase 2  V^T[k,d] = sum_t A[t,k] X[t,d]            tcgen05.mma, A operand = P^T from smem, X streamed
//                                                      again by TMA (L2-resident); N = 128 d-columns per MMA
//                                                      (a pair of 64-column slabs): phase 2 is bound by shared-
//                                                      memory bandwidth (operand re-reads + epilogue staging),
//                                                      so wider N = fewer P^T re-reads per output column
//   epilogue V^T[k,d] -= a_sum[k] C^T[k,d]: the fp16 C^T half-slab (32 columns) is TMA-loaded INTO the output
//            staging buffer and updated in place (thread = cluster row), row norms accumulate, the buffer
//            leaves by TMA store; two 64B-swizzled buffers per cluster tile alternate.  Un-normalised fp16
//            V^T + the combined intra-/global-L2 row scale are emitted:
//              vlad[b,k,:] = z[b,k,:] * rscale[b,k]   (consumers apply rscale in their epilogues)
//
// KSPLIT = 2 (cluster sizes 257..512, the wide config): a 2-CTA thread-block cluster owns one video, each CTA
// half of the clusters.  The softmax row statistics (max, sum) and the global-norm partial are exchanged
// through distributed shared memory (st.shared::cluster + remote mbarrier arrive).
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-9 softmax
// and epilogue (two groups of four warps: one per 128-row tile / TMEM lane quarter), warps 10-11 epilogue
// TMA agents (one per group): each output slab is handled as two 32-column halves in two 64B-swizzled
// buffers, so the centre load of the next half and the store of the previous one overlap the arithmetic.
#include <cstdlib>

#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

constexpr int TP = 256;  // padded frames per video handled by one CTA (two 128-row MMA tiles)

template <int KCT>
struct PoolCfg {
  static constexpr int KCP = KCT < 128 ? 128 : KCT;            // phase-2 M extent (padded clusters)
  static constexpr int NCH = KCT / 64;                         // 64-logit softmax chunks per row
  static constexpr int XS_BYTES = TP * 128;                    // one 64-column slab of X: 256 rows x 128 B
  static constexpr int WS_BYTES = KCT * 128;                   // one 64-row slab of Wc: 64 x KCT fp16
  static constexpr int ST1_BYTES = XS_BYTES + WS_BYTES;
  static constexpr int P_BYTES = (KCP / 64) * TP * 128;
  static constexpr int STG_BYTES = (KCP / 128) * 128 * 128;    // one 128-row x 128-byte output slab per cluster tile
  static constexpr int BUDGET = 224 * 1024;
  static constexpr int NS1 = (BUDGET / ST1_BYTES) > 4 ? 4 : (BUDGET / ST1_BYTES);
  static constexpr int XT_BYTES = 2 * 128 * 128;               // phase-2 stage: one 128-frame tile of a 128-column slab pair
  static constexpr int NS2 = ((BUDGET - P_BYTES - STG_BYTES) / XT_BYTES) > 8 ? 8 : ((BUDGET - P_BYTES - STG_BYTES) / XT_BYTES);
  static constexpr int STG_OFF = BUDGET - STG_BYTES;
  static constexpr int BAR_OFF = BUDGET;
  static constexpr int AFF_OFF = BAR_OFF + 384;                // float2 (scale, shift) per cluster
  static constexpr int RED_OFF = AFF_OFF + KCT * 8;            // 8 floats for the block reduction (+1 from the peer CTA)
  static constexpr int TOTAL = RED_OFF + 64;
  // logits: 2 frame tiles x KCT columns; aggregation: 2 buffers x (KCP/128) cluster tiles x 128 columns (reused)
  static constexpr uint32_t TMEM_COLS = (2 * KCT) < 256 ? 256 : 2 * KCT;
  static_assert(2 * (KCP / 128) * 128 <= TMEM_COLS, "aggregation accumulators must fit the allocation");
  static_assert(NS1 >= 2 && NS2 >= 2, "not enough shared memory for the pipeline");
  static_assert(NS1 * ST1_BYTES <= STG_OFF, "phase-1 ring must not reach the staging slabs (row-statistics mailbox)");
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

struct PoolParams {
  int B, T, D, K;                 // K = real cluster count (<= KSPLIT * KCT)
  const float* logit_scale;       // [K]  cluster_bn folded scale (1 for the bias branch)
  const float* logit_shift;       // [K]  cluster_bn folded shift (or cluster_biases)
  const int* valid_frames;        // [B] or null: frames t >= valid_frames[b] get zero assignment
  float* rscale;                  // [B][K]
  float* a_sum;                   // [B][K] or null
  int save_assign;                // write the assignment tile through tmap_a (training)
  const __half* assign_in;        // [B][T][K] or null: externally supplied assignments (NetVladV2), phase 1 skipped
  long long* debug_clock;         // optional [B][8] clock64 phase stamps (profiling aid)
};

__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {      // 2^x, x <= 0 here; -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ---- distributed shared memory (2-CTA cluster) ----
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float a) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
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
      printf("lpm: cluster mbarrier timeout block=%d thread=%d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int KCT, int KSPLIT>
__global__ void __launch_bounds__(384, 1)
netvlad_pool_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                        const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_c,
                        const __grid_constant__ CUtensorMap tmap_a, const PoolParams p) {
  using Cfg = PoolCfg<KCT>;
  // No static shared memory in this kernel: the dynamic window starts 1024-byte aligned (checked).
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sP = smem;                          // phase 2: P^T operand
  uint8_t* sRing2 = smem + Cfg::P_BYTES;       // phase 2: X slabs
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* full1 = bars;                      // [NS1]
  uint64_t* empty1 = full1 + 4;                // [NS1]
  uint64_t* full2 = empty1 + 4;                // [NS2]
  uint64_t* empty2 = full2 + 8;                // [NS2]
  uint64_t* s_full = empty2 + 8;               // logits complete
  uint64_t* p_ready = s_full + 1;              // assignment tile in smem (8 warps arrive)
  uint64_t* acc_full = p_ready + 1;            // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* c_full = acc_empty + 2;            // [2 groups][2 halves] centre half-slab landed in its staging buffer
  uint64_t* slab_ready = c_full + 4;           // [2][2] half-slab finished in smem (4 warps arrive): the agent may store it
  uint64_t* stat_ready = slab_ready + 4;       // peer CTA delivered its softmax row statistics (256 arrivals)
  uint64_t* norm_ready = stat_ready + 1;       // peer CTA delivered its global-norm partial
  float2* sAff = reinterpret_cast<float2*>(smem + Cfg::AFF_OFF);
  float* sRed = reinterpret_cast<float*>(smem + Cfg::RED_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sRed + 12);
  float2* sStat = reinterpret_cast<float2*>(smem + Cfg::STG_OFF);   // mailbox for the peer's (max, sum); dead before the epilogue

  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  const int crank = KSPLIT > 1 ? (int)(blockIdx.x % KSPLIT) : 0;    // == %cluster_ctarank for (KSPLIT,1,1) clusters
  const int b = blockIdx.x / KSPLIT;
  const int k0 = crank * KCT;                  // first cluster owned by this CTA
  const int n_ft = (p.T + 127) / 128;          // 1 or 2 frame tiles
  const int n_dc = p.D / 64;
  const int n_dc1 = p.assign_in ? 0 : n_dc;    // phase 1 (logits) is skipped when assignments are supplied

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_z);
    tma_prefetch_desc(&tmap_c);
    for (int i = 0; i < 4; ++i) { mbar_init(&full1[i], 1); mbar_init(&empty1[i], 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(&full2[i], 1); mbar_init(&empty2[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * (Cfg::KCP / 128));
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&c_full[i], 1); mbar_init(&slab_ready[i], 4); }
    mbar_init(stat_ready, 256);
    mbar_init(norm_ready, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  // folded logit affine in the log2 domain; padded clusters get -inf
  for (int k = threadIdx.x; k < KCT; k += blockDim.x) {
    float2 a;
    if (k0 + k < p.K && p.assign_in == nullptr) {
      a.x = p.logit_scale[k0 + k] * 1.4426950408889634f;
      a.y = p.logit_shift[k0 + k] * 1.4426950408889634f;
    } else {
      a.x = 0.f; a.y = -INFINITY;
    }
    sAff[k] = a;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (KSPLIT > 1) cluster_sync_all();          // the peer's mbarriers exist before anything is sent to them
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int dc = 0; dc < n_dc1; ++dc) {
        mbar_wait(&empty1[stage], phase ^ 1);
        uint8_t* sx = smem + stage * Cfg::ST1_BYTES;
        uint8_t* sw = sx + Cfg::XS_BYTES;
        mbar_expect_tx(&full1[stage], n_ft * 16384 + Cfg::WS_BYTES);
        for (int ft = 0; ft < n_ft; ++ft) tma_load_3d(sx + ft * 16384, &tmap_x, &full1[stage], dc * 64, ft * 128, b);
#pragma unroll
        for (int j = 0; j < KCT / 64; ++j) tma_load_3d(sw + j * 8192, &tmap_w, &full1[stage], k0 + j * 64, dc * 64, 0);
        if (++stage == Cfg::NS1) { stage = 0; phase ^= 1; }
      }
      // phase 2 reuses the shared memory of phase 1: wait until every phase-1 MMA has retired
      if (!p.assign_in) mbar_wait(s_full, 0);
      stage = 0; phase = 0;
      // stage = one 128-frame tile of a pair of 64-column slabs (N = 128 per MMA halves the P^T operand re-reads)
      for (int dp = 0; dp < (n_dc + 1) / 2; ++dp) {
        const int nsb = min(2, n_dc - 2 * dp);
        for (int ft = 0; ft < n_ft; ++ft) {
          mbar_wait(&empty2[stage], phase ^ 1);
          mbar_expect_tx(&full2[stage], nsb * 16384);
          for (int sb = 0; sb < nsb; ++sb)
            tma_load_3d(sRing2 + stage * Cfg::XT_BYTES + sb * 16384, &tmap_x, &full2[stage], (2 * dp + sb) * 64, ft * 128, b);
          if (++stage == Cfg::NS2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    // The whole warp runs the loops on warp-uniform values; an elected lane issues the instructions of a stage
    // back to back (see umma_f16_w in lpm_common.cuh: an `if (lane == 0)` region costs ~80 clk of issue per MMA,
    // more than the 64 clk an N = 128 aggregation MMA runs).
    {
      constexpr uint32_t idesc1 = umma_idesc_f16(128, KCT, 0, 1);   // A = X (K-major), B = Wc (MN-major)
      constexpr uint32_t idesc2 = umma_idesc_f16(128, 128, 1, 1);   // A = P^T (MN-major), B = X (MN-major), slab pair
      constexpr uint32_t idesc2t = umma_idesc_f16(128, 64, 1, 1);   // odd tail slab
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0; uint32_t phase = 0;
      for (int dc = 0; dc < n_dc1; ++dc) {
        mbar_wait(&full1[stage], phase);
        tc_fence_after();
        const uint32_t sx = smem_base + stage * Cfg::ST1_BYTES;
        const uint64_t xd = umma_smem_desc(sx, 16, 1024), wd = umma_smem_desc(sx + Cfg::XS_BYTES, 8192, 1024);
        if (elect_one()) {
          for (int ft = 0; ft < n_ft; ++ft) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16(tmem_base + ft * KCT, xd + ((ft * 16384 + ks * 32) >> 4), wd + ((ks * 2048) >> 4), idesc1,
                       (dc > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(&empty1[stage]);
        }
        __syncwarp();
        if (++stage == Cfg::NS1) { stage = 0; phase ^= 1; }
      }
      if (!p.assign_in) umma_commit_w(s_full);

      mbar_wait(p_ready, 0);
      tc_fence_after();
      stage = 0; phase = 0;
      int buf = 0; uint32_t bphase = 0;
      const uint64_t pd = umma_smem_desc(smem_u32(sP), TP * 128, 1024);
      for (int dp = 0; dp < (n_dc + 1) / 2; ++dp) {
        const uint32_t idesc = (n_dc - 2 * dp >= 2) ? idesc2 : idesc2t;
        mbar_wait(&acc_empty[buf], bphase ^ 1);
        for (int ft = 0; ft < n_ft; ++ft) {
          mbar_wait(&full2[stage], phase);
          tc_fence_after();
          const uint64_t xd = umma_smem_desc(smem_u32(sRing2) + stage * Cfg::XT_BYTES, 16384, 1024);
          if (elect_one()) {
#pragma unroll
            for (int mt = 0; mt < Cfg::KCP / 128; ++mt) {
              const uint32_t d_tmem = tmem_base + buf * (Cfg::KCP / 128) * 128 + mt * 128;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                umma_f16(d_tmem, pd + ((mt * 2 * (TP * 128) + (ft * 8 + ks) * 2048) >> 4), xd + ((ks * 2048) >> 4), idesc,
                         (ft > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(&empty2[stage]);
          }
          __syncwarp();
          if (++stage == Cfg::NS2) { stage = 0; phase ^= 1; }
        }
        umma_commit_w(&acc_full[buf]);
        if (++buf == 2) { buf = 0; bphase ^= 1; }
      }
    }
  } else if (warp >= 10) {
    // ====================== epilogue TMA agents (one per cluster tile) ======================
    const int g = warp - 10;
    if (lane == 0 && g < Cfg::KCP / 128) {
      mbar_wait(p_ready, 0);                    // assignment tile final; the statistics mailbox is dead
      if (g == 0 && p.save_assign) {
        // training: the assignment tile leaves for the backward straight from the operand tile (fp16 [B][T][K])
#pragma unroll
        for (int kb = 0; kb < KCT / 64; ++kb)
          if (k0 + kb * 64 < p.K) tma_store_3d(&tmap_a, sP + kb * (TP * 128), k0 + kb * 64, 0, b);
        bulk_commit();
      }
      uint8_t* stg = smem + Cfg::STG_OFF + g * (128 * 128);
      const int krow = k0 + g * 128;
      for (int h = 0; h < 2; ++h) {
        mbar_expect_tx(&c_full[g * 2 + h], 128 * 64);
        tma_load_3d(stg + h * 8192, &tmap_c, &c_full[g * 2 + h], h * 32, krow, 0);
      }
      uint32_t ph = 0;
      for (int db = 0; db < n_dc; ++db) {
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&slab_ready[g * 2 + h], ph);
          tma_store_3d(&tmap_z, stg + h * 8192, db * 64 + h * 32, krow, b);   // rows k >= K are clipped by the tensor map
          bulk_commit();
          if (db + 1 < n_dc) {
            bulk_wait_read<0>();                // the half-slab has left shared memory: refill it with C^T
            mbar_expect_tx(&c_full[g * 2 + h], 128 * 64);
            tma_load_3d(stg + h * 8192, &tmap_c, &c_full[g * 2 + h], (db + 1) * 64 + h * 32, krow, 0);
          }
        }
        ph ^= 1;
      }
      bulk_wait<0>();
    }
  } else {
    // ====================== softmax + epilogue warps (2..9) ======================
    const int grp = (warp - 2) >> 2;           // frame tile (softmax) / cluster tile (epilogue)
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;       // row within the 128-row tile
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;

    long long* dbg = (p.debug_clock != nullptr && warp == 2 && lane == 0 && crank == 0) ? p.debug_clock + (size_t)b * 8 : nullptr;
    if (dbg) dbg[0] = clock64();
    // ---------------- softmax over clusters, one thread per frame ----------------
    if (!p.assign_in) {
      mbar_wait(s_full, 0);
      tc_fence_after();
    }
    if (dbg) dbg[1] = clock64();     // logits complete (phase 1 done)
    {
      int tv = p.T;
      if (p.valid_frames) tv = min(tv, p.valid_frames[b]);
      const bool active = grp < n_ft;
      const int t_row = grp * 128 + row;
      // this thread's row of the P^T operand: 16-byte chunk c (8 clusters) lives at
      //   sP + (c>>3)*(TP*128) + t_row*128 + (((c&7) ^ (t_row&7)) << 4)          (128B swizzle)
      uint8_t* prow = sP + t_row * 128;
      const int sw = t_row & 7;
      float fac[Cfg::NCH];                      // per 64-cluster chunk: factor that normalises the parked values
#pragma unroll
      for (int c = 0; c < Cfg::NCH; ++c) fac[c] = 0.f;
      if (active && p.assign_in) {
        // NetVladV2: the assignment row comes from the encoder (video_pooling_modules.py:1628-1638)
        const float on = (t_row < tv) ? 1.f : 0.f;
#pragma unroll
        for (int c = 0; c < Cfg::NCH; ++c) fac[c] = on;
        const __half* arow = p.assign_in + ((size_t)b * p.T + min(t_row, p.T - 1)) * p.K + k0;
#pragma unroll 4
        for (int ch = 0; ch < KCT / 8; ++ch) {
          uint4 v = make_uint4(0, 0, 0, 0);
          if (k0 + ch * 8 < p.K && t_row < tv) v = __ldg(reinterpret_cast<const uint4*>(arow + ch * 8));
          *reinterpret_cast<uint4*>(prow + (ch >> 3) * (TP * 128) + (((ch & 7) ^ sw) << 4)) = v;
        }
      } else if (!p.assign_in) {
        float m = -INFINITY, sum = 0.f;
        float mref[Cfg::NCH];
        if (active) {
          const uint32_t s_addr = tmem_base + lane_addr + grp * KCT;
#pragma unroll
          for (int c = 0; c < Cfg::NCH; ++c) {
            uint32_t r[64];
            tmem_ld32(s_addr + c * 64, r);
            tmem_ld32(s_addr + c * 64 + 32, r + 32);
            tmem_ld_wait();
            float cm = -INFINITY;
#pragma unroll
            for (int i = 0; i < 64; i += 2) {
              const float4 a = *reinterpret_cast<const float4*>(&sAff[c * 64 + i]);
              const float t0 = fmaf(__uint_as_float(r[i]), a.x, a.y);
              const float t1 = fmaf(__uint_as_float(r[i + 1]), a.z, a.w);
              r[i] = __float_as_uint(t0);
              r[i + 1] = __float_as_uint(t1);
              cm = fmaxf(cm, fmaxf(t0, t1));
            }
            // chunk 0 always holds a real cluster (K >= 8), so the running max is finite from here on
            const float m_new = fmaxf(m, cm);
            sum *= ex2_approx(m - m_new);
            m = m_new;
            mref[c] = m_new;
            // un-normalised exp values (<= 1, relative to the running max) parked in the operand tile
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float e0 = ex2_approx(__uint_as_float(r[q * 8 + 2 * j]) - m_new);
                const float e1 = ex2_approx(__uint_as_float(r[q * 8 + 2 * j + 1]) - m_new);
                sum += e0 + e1;
                pk[j] = pack_half2(e0, e1);
              }
              const int ch = c * 8 + q;
              *reinterpret_cast<uint4*>(prow + (ch >> 3) * (TP * 128) + (((ch & 7) ^ sw) << 4)) =
                  make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
        float base = 0.f;                       // 2^(m - M) / total over all clusters of the row
        if (KSPLIT > 1) {
          // exchange (running max, sum) with the CTA that owns the other half of the clusters
          const uint32_t peer = (uint32_t)(crank ^ 1);
          if (active) st_cluster_f32x2(mapa_shared(smem_u32(&sStat[t_row]), peer), m, sum);
          mbar_arrive_remote(mapa_shared(smem_u32(stat_ready), peer));
          if (active) {
            mbar_wait_cluster(stat_ready, 0);
            const float2 ps = sStat[t_row];
            const float M = fmaxf(m, ps.x);
            const float w_own = ex2_approx(m - M);
            base = w_own / (sum * w_own + ps.y * ex2_approx(ps.x - M));
          }
        } else if (active) {
          base = 1.f / sum;
        }
        if (active && t_row < tv) {
#pragma unroll
          for (int c = 0; c < Cfg::NCH; ++c) fac[c] = base * ex2_approx(mref[c] - m);
        }
      }
      tc_fence_before();
      // normalise in place (fp32 multiply, single extra rounding); every row of the padded tile is
      // written: zeros for masked frames, inactive tiles and padded clusters
#pragma unroll
      for (int c = 0; c < Cfg::KCP / 64; ++c) {
        const float f = c < Cfg::NCH ? fac[c < Cfg::NCH ? c : 0] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4* slot = reinterpret_cast<uint4*>(prow + c * (TP * 128) + ((q ^ sw) << 4));
          uint4 v = make_uint4(0, 0, 0, 0);
          if (f != 0.f) {
            v = *slot;
            if (!p.assign_in) {
              __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 g = __half22float2(h[j]);
                h[j] = __floats2half2_rn(g.x * f, g.y * f);
              }
            }
          }
          *slot = v;
        }
      }
    }
    fence_async_smem();                         // generic-proxy smem writes -> visible to tcgen05.mma / TMA
    __syncwarp();
    if (lane == 0) mbar_arrive(p_ready);
    named_bar_sync(1, 256);                     // all assignment rows are in smem
    if (dbg) dbg[2] = clock64();     // softmax done
    // ---------------- a_sum[k] = sum_t A[t,k] from the fp16 tile (consistent with the MMA) ---------
    const int k_own = grp * 128 + row;          // cluster row owned in the epilogue (local to this CTA)
    float a_sum = 0.f;
    if (k_own < Cfg::KCP) {
      const int kb = k_own >> 6, cc = (k_own & 63) >> 3, e = k_own & 7;
      const uint8_t* col = sP + kb * (TP * 128) + e * 2;
      const int rows = n_ft * 128;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
      for (int t = 0; t < rows; t += 4) {
        s0 += __half2float(*reinterpret_cast<const __half*>(col + (t + 0) * 128 + ((cc ^ ((t + 0) & 7)) << 4)));
        s1 += __half2float(*reinterpret_cast<const __half*>(col + (t + 1) * 128 + ((cc ^ ((t + 1) & 7)) << 4)));
        s2 += __half2float(*reinterpret_cast<const __half*>(col + (t + 2) * 128 + ((cc ^ ((t + 2) & 7)) << 4)));
        s3 += __half2float(*reinterpret_cast<const __half*>(col + (t + 3) * 128 + ((cc ^ ((t + 3) & 7)) << 4)));
      }
      a_sum = (s0 + s1) + (s2 + s3);
    }
    const bool k_ok = k_own < Cfg::KCP && k0 + k_own < p.K;
    if (p.a_sum != nullptr && k_ok) p.a_sum[(size_t)b * p.K + k0 + k_own] = a_sum;

    if (dbg) dbg[3] = clock64();     // a_sum done
    // ---------------- phase-2 epilogue: residual, row norm, fp16 slab -> TMA store ----------------
    float sumsq = 0.f;
    if (grp < Cfg::KCP / 128) {
      int buf = 0; uint32_t bphase = 0, cphase = 0;
      uint8_t* stg = smem + Cfg::STG_OFF + grp * (128 * 128);
      const int sw64 = (row >> 1) & 3;                          // 64B swizzle: 16-byte chunk j of row r sits at j ^ ((r>>1)&3)
      for (int dp = 0; dp < (n_dc + 1) / 2; ++dp) {
        const int nsb = min(2, n_dc - 2 * dp);
        mbar_wait(&acc_full[buf], bphase);
        tc_fence_after();
        for (int sb = 0; sb < nsb; ++sb) {
          uint32_t r[64];
          const uint32_t taddr = tmem_base + lane_addr + buf * (Cfg::KCP / 128) * 128 + grp * 128 + sb * 64;
          tmem_ld32(taddr, r);
          tmem_ld32(taddr + 32, r + 32);
          tmem_ld_wait();
          if (sb == nsb - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);        // accumulator drained: MMA may reuse it
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&c_full[grp * 2 + h], cphase);            // C^T[k0+grp*128.., 32 columns] is in the buffer
            uint8_t* rowp = stg + h * 8192 + row * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4* slot = reinterpret_cast<uint4*>(rowp + ((j ^ sw64) << 4));
              uint4 w = *slot;
              __half2* hh = reinterpret_cast<__half2*>(&w);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 c = __half22float2(hh[i]);
                const float v0 = fmaf(-a_sum, c.x, __uint_as_float(r[h * 32 + 8 * j + 2 * i]));
                const float v1 = fmaf(-a_sum, c.y, __uint_as_float(r[h * 32 + 8 * j + 2 * i + 1]));
                sumsq = fmaf(v0, v0, sumsq);
                sumsq = fmaf(v1, v1, sumsq);
                hh[i] = __floats2half2_rn(v0, v1);
              }
              *slot = w;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&slab_ready[grp * 2 + h]);
          }
          cphase ^= 1;
        }
        if (++buf == 2) { buf = 0; bphase ^= 1; }
      }
      if (!k_ok) sumsq = 0.f;
    }
    if (dbg) dbg[4] = clock64();     // phase 2 + epilogue done
    // intra-norm (per cluster row) and global norm (per video): frame_level_models.py:2819-2822
    const float r_intra = rsqrtf(fmaxf(sumsq, 1e-12f));
    float contrib = k_ok ? sumsq * r_intra * r_intra : 0.f;
    contrib = warp_sum(contrib);
    if (lane == 0) sRed[warp - 2] = contrib;
    named_bar_sync(1, 256);
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += sRed[i];
    if (KSPLIT > 1) {
      if (warp == 2 && lane == 0) {
        const uint32_t peer = (uint32_t)(crank ^ 1);
        st_cluster_f32(mapa_shared(smem_u32(&sRed[8]), peer), tot);
        mbar_arrive_remote(mapa_shared(smem_u32(norm_ready), peer));
      }
      mbar_wait_cluster(norm_ready, 0);
      tot += sRed[8];
    }
    const float r_glob = rsqrtf(fmaxf(tot, 1e-12f));
    if (k_ok) p.rscale[(size_t)b * p.K + k0 + k_own] = r_intra * r_glob;
  }

  tc_fence_before();
  __syncthreads();
  if (KSPLIT > 1) cluster_sync_all();           // no CTA leaves while its peer may still write into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int KCT, int KSPLIT>
static int launch_pool(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tz, const CUtensorMap& tc,
                       const CUtensorMap& ta, const PoolParams& p, cudaStream_t st) {
  using Cfg = PoolCfg<KCT>;
  auto kern = netvlad_pool_fwd_kernel<KCT, KSPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    LPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::TOTAL));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.B * KSPLIT);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = Cfg::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = KSPLIT; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = KSPLIT > 1 ? 1 : 0;
  LPM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tx, tw, tz, tc, ta, p));
  return LPM_OK;
}

int netvlad_pool_fwd(const __half* x, long long ldx, long long x_batch_stride, const __half* wc, long long ldw,
                     const float* logit_scale, const float* logit_shift, const __half* centers_t16,
                     const int* valid_frames, int B, int T, int D, int K, __half* z, float* rscale, float* a_sum,
                     __half* assign, const __half* assign_in, long long* debug_clock, cudaStream_t st) {
  LPM_REQUIRE(B > 0 && T > 0 && T <= TP, "netvlad_pool_fwd: frames per video must be in [1,%d] (got %d)", TP, T);
  LPM_REQUIRE(D % 64 == 0 && D >= 64, "netvlad_pool_fwd: feature size must be a multiple of 64 (got %d)", D);
  LPM_REQUIRE(K % 8 == 0 && K >= 8 && K <= 512, "netvlad_pool_fwd: cluster size must be a multiple of 8 in [8,512] (got %d)", K);
  LPM_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0 && x_batch_stride % 8 == 0, "netvlad_pool_fwd: strides must be multiples of 8");
  LPM_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(centers_t16) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(assign) & 15) == 0,
              "netvlad_pool_fwd: z, centers_t16 and assign must be 16-byte aligned");
  PoolParams p{};
  p.B = B; p.T = T; p.D = D; p.K = K;
  p.logit_scale = logit_scale; p.logit_shift = logit_shift;
  p.valid_frames = valid_frames; p.rscale = rscale; p.a_sum = a_sum; p.save_assign = assign != nullptr;
  p.assign_in = assign_in; p.debug_clock = debug_clock;
  CUtensorMap tx, tw, tz, tc, ta;
  if (int rc = make_tmap_3d(&tx, x, 2, D, T, B, ldx, x_batch_stride, 64, 128)) return rc;
  if (assign_in != nullptr) tw = tx;   // unused in this mode
  else if (int rc = make_tmap_3d(&tw, wc, 2, K, D, 1, ldw, 0, 64, 64)) return rc;
  if (int rc = make_tmap_3d(&tz, z, 2, D, K, B, D, (uint64_t)K * D, 32, 128, 64)) return rc;      // 64B-swizzled half-slabs
  if (int rc = make_tmap_3d(&tc, centers_t16, 2, D, K, 1, D, 0, 32, 128, 64)) return rc;
  if (assign == nullptr) ta = tz;      // unused
  else if (int rc = make_tmap_3d(&ta, assign, 2, K, T, B, K, (uint64_t)T * K, 64, 256)) return rc;
  // Experiment switch: LPM_POOL_SPLIT=1 runs 129..256 clusters as a 2-CTA cluster per video (128 clusters per CTA, half the
  // assignment tile per CTA -> a 4-stage X ring in phase 2 instead of 2).  Measured at config 1 (gpurun r2t, cycles per CTA):
  // logits 15.6 k | softmax 7.7 k | a_sum 3.6 k | aggregation 27.1 k = 54.0 k, i.e. 108 k SM-cycles per video against 70.7 k
  // for one CTA per video (304 vs 419 TFLOP/s at B = 80, 516 vs 710 at full waves): half the clusters do not halve the
  // logits phase (X is streamed by both CTAs) and the aggregation stays paced by the centre-load -> update -> store chain
  // of its two staging half-buffers (3.4 k cycles per slab pair whatever the ring depth).  Off by default.
  static const bool split256 = getenv("LPM_POOL_SPLIT") != nullptr && getenv("LPM_POOL_SPLIT")[0] == '1';
  if (K <= 64) return launch_pool<64, 1>(tx, tw, tz, tc, ta, p, st);
  if (K <= 128) return launch_pool<128, 1>(tx, tw, tz, tc, ta, p, st);
  if (K <= 256 && split256 && assign_in == nullptr) return launch_pool<128, 2>(tx, tw, tz, tc, ta, p, st);
  if (K <= 256) return launch_pool<256, 1>(tx, tw, tz, tc, ta, p, st);
  return launch_pool<256, 2>(tx, tw, tz, tc, ta, p, st);
}

}  // namespace lpm
