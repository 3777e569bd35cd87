// Fused NetVLAD pooling forward (K1) for sm_100a -- frame_level_models.py:2775-2822.
//
// One CTA per video.  The B x T x K assignment tensor lives only in TMEM / registers / shared memory
// (it is written to HBM only when the caller asks for it, i.e. for the training backward).
//
//   phase 1  S[t,k]   = sum_d X[t,d] Wc[d,k]          tcgen05.mma, X and Wc streamed by TMA in 64-wide
//                                                      d-chunks, both 128-frame tiles accumulate in TMEM
//   softmax  A[t,k]   = softmax_k(S*scale_k + shift_k) one thread per frame row, TMEM -> registers,
//                                                      result written as fp16 to swizzled smem (MN-major
//                                                      UMMA operand), masked for t >= T / invalid frames
//   phase 2  V^T[k,d] = sum_t A[t,k] X[t,d]            tcgen05.mma, A operand = P^T from smem, X streamed
//                                                      again by TMA (L2-resident), 64 d-columns per stage
//   epilogue V^T[k,d] -= a_sum[k] C[d,k] (coalesced loads of C[d][.]);  row norms;  fp16 slab in swizzled
//            smem -> TMA store;  un-normalised fp16 V^T + the combined
//                                                      intra-/global-L2 row scale are emitted:
//              vlad[b,k,:] = z[b,k,:] * rscale[b,k]   (consumers apply rscale in their epilogues)
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-9 softmax
// and epilogue (two groups of four warps: one per 128-row tile / TMEM lane quarter).
#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

constexpr int TP = 256;  // padded frames per video handled by one CTA (two 128-row MMA tiles)

template <int KCT>
struct PoolCfg {
  static constexpr int KCP = KCT < 128 ? 128 : KCT;            // phase-2 M extent (padded clusters)
  static constexpr int XS_BYTES = TP * 128;                    // one 64-column slab of X: 256 rows x 128 B
  static constexpr int WS_BYTES = KCT * 128;                   // one 64-row slab of Wc: 64 x KCT fp16
  static constexpr int ST1_BYTES = XS_BYTES + WS_BYTES;
  static constexpr int P_BYTES = (KCP / 64) * TP * 128;
  static constexpr int STG_BYTES = (KCP / 128) * 128 * 128;    // one 128-row x 128-byte output slab per cluster tile
  static constexpr int BUDGET = 224 * 1024;
  static constexpr int NS1 = (BUDGET / ST1_BYTES) > 4 ? 4 : (BUDGET / ST1_BYTES);
  static constexpr int NS2 = ((BUDGET - P_BYTES - STG_BYTES) / XS_BYTES) > 4 ? 4 : ((BUDGET - P_BYTES - STG_BYTES) / XS_BYTES);
  static constexpr int STG_OFF = BUDGET - STG_BYTES;
  static constexpr int BAR_OFF = BUDGET;
  static constexpr int AFF_OFF = BAR_OFF + 256;                // float2 (scale, shift) per cluster
  static constexpr int RED_OFF = AFF_OFF + KCT * 8;            // 8 floats for the block reduction
  static constexpr int TOTAL = RED_OFF + 64;
  static constexpr uint32_t TMEM_COLS = (2 * KCT) < 256 ? 256 : 2 * KCT;
  static_assert(NS1 >= 2 && NS2 >= 2, "not enough shared memory for the pipeline");
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

struct PoolParams {
  int B, T, D, K;                 // K = real cluster count (<= KCT)
  const float* logit_scale;       // [K]  cluster_bn folded scale (1 for the bias branch)
  const float* logit_shift;       // [K]  cluster_bn folded shift (or cluster_biases)
  const float* centers;           // [D][K] fp32: cluster_weights2[0] / cluster_centers (native layout)
  const int* valid_frames;        // [B] or null: frames t >= valid_frames[b] get zero assignment
  __half* z;                      // [B][K][D]  un-normalised V^T
  float* rscale;                  // [B][K]
  float* a_sum;                   // [B][K] or null
  __half* assign;                 // [B][T][K] or null (saved for backward)
  const __half* assign_in;        // [B][T][K] or null: externally supplied assignments (NetVladV2), phase 1 skipped
  long long* debug_clock;         // optional [B][8] clock64 phase stamps (profiling aid)
};

__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int KCT>
__global__ void __launch_bounds__(320, 1)
netvlad_pool_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                        const __grid_constant__ CUtensorMap tmap_z, const PoolParams p) {
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
  uint64_t* empty2 = full2 + 4;                // [NS2]
  uint64_t* s_full = empty2 + 4;               // logits complete
  uint64_t* p_ready = s_full + 1;              // assignment tile in smem (8 warps arrive)
  uint64_t* acc_full = p_ready + 1;            // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float2* sAff = reinterpret_cast<float2*>(smem + Cfg::AFF_OFF);
  float* sRed = reinterpret_cast<float*>(smem + Cfg::RED_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int n_ft = (p.T + 127) / 128;          // 1 or 2 frame tiles
  const int n_dc = p.D / 64;
  const int n_dc1 = p.assign_in ? 0 : n_dc;    // phase 1 (logits) is skipped when assignments are supplied

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_z);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full1[i], 1); mbar_init(&empty1[i], 1);
      mbar_init(&full2[i], 1); mbar_init(&empty2[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 8);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * (Cfg::KCP / 128)); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  // folded logit affine in the log2 domain; padded clusters get -inf
  for (int k = threadIdx.x; k < KCT; k += blockDim.x) {
    float2 a;
    if (k < p.K && p.assign_in == nullptr) {
      a.x = p.logit_scale[k] * 1.4426950408889634f;
      a.y = p.logit_shift[k] * 1.4426950408889634f;
    } else {
      a.x = 0.f; a.y = -INFINITY;
    }
    sAff[k] = a;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
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
        for (int j = 0; j < KCT / 64; ++j) tma_load_3d(sw + j * 8192, &tmap_w, &full1[stage], j * 64, dc * 64, 0);
        if (++stage == Cfg::NS1) { stage = 0; phase ^= 1; }
      }
      // phase 2 reuses the shared memory of phase 1: wait until every phase-1 MMA has retired
      if (!p.assign_in) mbar_wait(s_full, 0);
      stage = 0; phase = 0;
      for (int db = 0; db < n_dc; ++db) {
        mbar_wait(&empty2[stage], phase ^ 1);
        uint8_t* sx = sRing2 + stage * Cfg::XS_BYTES;
        mbar_expect_tx(&full2[stage], n_ft * 16384);
        for (int ft = 0; ft < n_ft; ++ft) tma_load_3d(sx + ft * 16384, &tmap_x, &full2[stage], db * 64, ft * 128, b);
        if (++stage == Cfg::NS2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_f16(128, KCT, 0, 1);   // A = X (K-major), B = Wc (MN-major)
      constexpr uint32_t idesc2 = umma_idesc_f16(128, 64, 1, 1);    // A = P^T (MN-major), B = X (MN-major)
      int stage = 0; uint32_t phase = 0;
      for (int dc = 0; dc < n_dc1; ++dc) {
        mbar_wait(&full1[stage], phase);
        tc_fence_after();
        const uint32_t sx = smem_u32(smem + stage * Cfg::ST1_BYTES);
        const uint32_t sw = sx + Cfg::XS_BYTES;
        for (int ft = 0; ft < n_ft; ++ft) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = umma_smem_desc(sx + ft * 16384 + ks * 32, 16, 1024);
            const uint64_t bd = umma_smem_desc(sw + ks * 2048, 8192, 1024);
            umma_f16(tmem_base + ft * KCT, ad, bd, idesc1, (dc > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty1[stage]);
        if (++stage == Cfg::NS1) { stage = 0; phase ^= 1; }
      }
      if (!p.assign_in) umma_commit(s_full);

      mbar_wait(p_ready, 0);
      tc_fence_after();
      stage = 0; phase = 0;
      int buf = 0; uint32_t bphase = 0;
      const uint32_t sp = smem_u32(sP);
      const int n_ks = n_ft * 8;
      for (int db = 0; db < n_dc; ++db) {
        mbar_wait(&full2[stage], phase);
        mbar_wait(&acc_empty[buf], bphase ^ 1);
        tc_fence_after();
        const uint32_t sx = smem_u32(sRing2 + stage * Cfg::XS_BYTES);
#pragma unroll
        for (int mt = 0; mt < Cfg::KCP / 128; ++mt) {
          const uint32_t d_tmem = tmem_base + buf * 128 + mt * 64;
          for (int ks = 0; ks < n_ks; ++ks) {
            const uint64_t ad = umma_smem_desc(sp + mt * 2 * (TP * 128) + ks * 2048, TP * 128, 1024);
            const uint64_t bd = umma_smem_desc(sx + ks * 2048, 8192, 1024);
            umma_f16(d_tmem, ad, bd, idesc2, ks > 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty2[stage]);
        umma_commit(&acc_full[buf]);
        if (++stage == Cfg::NS2) { stage = 0; phase ^= 1; }
        if (++buf == 2) { buf = 0; bphase ^= 1; }
      }
    }
  } else {
    // ====================== softmax + epilogue warps (2..9) ======================
    const int grp = (warp - 2) >> 2;           // frame tile (softmax) / cluster tile (epilogue)
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;       // row within the 128-row tile
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;

    long long* dbg = (p.debug_clock != nullptr && warp == 2 && lane == 0) ? p.debug_clock + (size_t)b * 8 : nullptr;
    if (dbg) dbg[0] = clock64();
    // ---------------- softmax over clusters, one thread per frame ----------------
    if (!p.assign_in) {
      mbar_wait(s_full, 0);
      tc_fence_after();
    }
    if (dbg) dbg[1] = clock64();     // logits complete (phase 1 done)
    {
      const int t = grp * 128 + row;
      int tv = p.T;
      if (p.valid_frames) tv = min(tv, p.valid_frames[b]);
      const bool active = grp < n_ft;
      const int t_row = grp * 128 + row;
      // this thread's row of the P^T operand: 16-byte chunk c (8 clusters) lives at
      //   sP + (c>>3)*(TP*128) + t_row*128 + (((c&7) ^ (t_row&7)) << 4)          (128B swizzle)
      uint8_t* prow = sP + t_row * 128;
      const int sw = t_row & 7;
      float inv = 0.f;
      if (active && p.assign_in) {
        // NetVladV2: the assignment row comes from the encoder (video_pooling_modules.py:1628-1638)
        inv = (t_row < tv) ? 1.f : 0.f;
        const __half* arow = p.assign_in + ((size_t)b * p.T + min(t_row, p.T - 1)) * p.K;
#pragma unroll 4
        for (int ch = 0; ch < KCT / 8; ++ch) {
          uint4 v = make_uint4(0, 0, 0, 0);
          if (ch * 8 < p.K && t_row < tv) v = __ldg(reinterpret_cast<const uint4*>(arow + ch * 8));
          *reinterpret_cast<uint4*>(prow + (ch >> 3) * (TP * 128) + (((ch & 7) ^ sw) << 4)) = v;
        }
      } else if (active) {
        const uint32_t s_addr = tmem_base + lane_addr + grp * KCT;
        float m = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < KCT / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(s_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float4 a = *reinterpret_cast<const float4*>(&sAff[c * 32 + i]);
            m = fmaxf(m, fmaxf(fmaf(__uint_as_float(r[i]), a.x, a.y), fmaf(__uint_as_float(r[i + 1]), a.z, a.w)));
          }
        }
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < KCT / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(s_addr + c * 32, r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float4 a = *reinterpret_cast<const float4*>(&sAff[c * 32 + i]);
            const float e0 = exp2f(fmaf(__uint_as_float(r[i]), a.x, a.y) - m);
            const float e1 = exp2f(fmaf(__uint_as_float(r[i + 1]), a.z, a.w) - m);
            sum += e0 + e1;
            pk[i / 2] = pack_half2(e0, e1);
          }
          // un-normalised exp values (<= 1) parked in the operand tile; rescaled in place below
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ch = c * 4 + q;
            *reinterpret_cast<uint4*>(prow + (ch >> 3) * (TP * 128) + (((ch & 7) ^ sw) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
        inv = (t_row < tv) ? 1.f / sum : 0.f;
      }
      tc_fence_before();
      // normalise in place (fp32 multiply, single extra rounding); every row of the padded tile is
      // written: zeros for masked frames, inactive tiles and padded clusters
      __half* a_out = (p.assign != nullptr && t_row < p.T) ? p.assign + ((size_t)b * p.T + t_row) * p.K : nullptr;
#pragma unroll 4
      for (int ch = 0; ch < Cfg::KCP / 8; ++ch) {
        uint4* slot = reinterpret_cast<uint4*>(prow + (ch >> 3) * (TP * 128) + (((ch & 7) ^ sw) << 4));
        uint4 v = make_uint4(0, 0, 0, 0);
        if (active && ch < KCT / 8 && inv != 0.f) {
          v = *slot;
          __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __half22float2(h[q]);
            h[q] = __floats2half2_rn(f.x * inv, f.y * inv);
          }
        }
        *slot = v;
        if (a_out != nullptr && ch * 8 < p.K) *reinterpret_cast<uint4*>(a_out + ch * 8) = v;
      }
    }
    fence_async_smem();                         // generic-proxy smem writes -> visible to tcgen05.mma
    __syncwarp();
    if (lane == 0) mbar_arrive(p_ready);
    named_bar_sync(1, 256);                     // all assignment rows are in smem
    if (dbg) dbg[2] = clock64();     // softmax done

    // ---------------- a_sum[k] = sum_t A[t,k] from the fp16 tile (consistent with the MMA) ---------
    const int k_own = grp * 128 + row;          // cluster row owned in the epilogue
    float a_sum = 0.f;
    if (k_own < Cfg::KCP) {
      const int kb = k_own >> 6, cc = (k_own & 63) >> 3, e = k_own & 7;
      const uint8_t* col = sP + kb * (TP * 128) + e * 2;
      const int rows = n_ft * 128;
#pragma unroll 8
      for (int t = 0; t < rows; ++t)
        a_sum += __half2float(*reinterpret_cast<const __half*>(col + t * 128 + ((cc ^ (t & 7)) << 4)));
    }
    const bool k_ok = k_own < p.K;
    if (p.a_sum != nullptr && k_ok) p.a_sum[(size_t)b * p.K + k_own] = a_sum;

    if (dbg) dbg[3] = clock64();     // a_sum done
    // ---------------- phase-2 epilogue: residual, row norm, fp16 slab -> TMA store ----------------
    float sumsq = 0.f;
    if (grp < Cfg::KCP / 128) {
      int buf = 0; uint32_t bphase = 0;
      uint8_t* slab = smem + Cfg::STG_OFF + grp * (128 * 128);
      const bool leader = quarter == 0 && lane == 0;          // one thread per group issues the TMA stores
      const int bar_a = 2 + 2 * grp, bar_b = 3 + 2 * grp;
      const float* ccol = p.centers + (k_ok ? k_own : 0);       // C[d][k]: lanes read consecutive k (coalesced)
      for (int db = 0; db < n_dc; ++db) {
        mbar_wait(&acc_full[buf], bphase);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_base + lane_addr + buf * 128 + grp * 64;
        tmem_ld32(taddr, r0);
        tmem_ld32(taddr + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);            // accumulator drained: MMA may reuse it
        if (leader) bulk_wait_read<0>();                        // previous slab of this group has left smem
        named_bar_sync(bar_a, 128);
        const int d0 = db * 64;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t* r = hh == 0 ? r0 : r1;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float c = __ldg(ccol + (size_t)(d0 + hh * 32 + i) * p.K);
            v[i] = __uint_as_float(r[i]) - a_sum * c;
            sumsq += v[i] * v[i];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 w;
            w.x = pack_half2(v[8 * i + 0], v[8 * i + 1]);
            w.y = pack_half2(v[8 * i + 2], v[8 * i + 3]);
            w.z = pack_half2(v[8 * i + 4], v[8 * i + 5]);
            w.w = pack_half2(v[8 * i + 6], v[8 * i + 7]);
            *reinterpret_cast<uint4*>(slab + row * 128 + (((hh * 4 + i) ^ (row & 7)) << 4)) = w;
          }
        }
        fence_async_smem();
        named_bar_sync(bar_b, 128);
        if (leader) {
          tma_store_3d(&tmap_z, slab, d0, grp * 128, b);        // rows k >= K are clipped by the tensor map
          bulk_commit();
        }
        if (++buf == 2) { buf = 0; bphase ^= 1; }
      }
      if (leader) bulk_wait<0>();
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
    const float r_glob = rsqrtf(fmaxf(tot, 1e-12f));
    if (k_ok && grp < Cfg::KCP / 128) p.rscale[(size_t)b * p.K + k_own] = r_intra * r_glob;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int KCT>
static int launch_pool(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tz, const PoolParams& p,
                       cudaStream_t st) {
  using Cfg = PoolCfg<KCT>;
  auto kern = netvlad_pool_fwd_kernel<KCT>;
  static bool attr_set = false;
  if (!attr_set) {
    LPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::TOTAL));
    attr_set = true;
  }
  kern<<<p.B, 320, Cfg::TOTAL, st>>>(tx, tw, tz, p);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int netvlad_pool_fwd(const __half* x, long long ldx, long long x_batch_stride, const __half* wc, long long ldw,
                     const float* logit_scale, const float* logit_shift, const float* centers,
                     const int* valid_frames, int B, int T, int D, int K, __half* z, float* rscale, float* a_sum,
                     __half* assign, const __half* assign_in, long long* debug_clock, cudaStream_t st) {
  LPM_REQUIRE(B > 0 && T > 0 && T <= TP, "netvlad_pool_fwd: frames per video must be in [1,%d] (got %d)", TP, T);
  LPM_REQUIRE(D % 64 == 0 && D >= 64, "netvlad_pool_fwd: feature size must be a multiple of 64 (got %d)", D);
  LPM_REQUIRE(K % 8 == 0 && K >= 8 && K <= 256, "netvlad_pool_fwd: cluster size must be a multiple of 8 in [8,256] (got %d)", K);
  LPM_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0 && x_batch_stride % 8 == 0, "netvlad_pool_fwd: strides must be multiples of 8");
  PoolParams p{};
  p.B = B; p.T = T; p.D = D; p.K = K;
  p.logit_scale = logit_scale; p.logit_shift = logit_shift; p.centers = centers;
  p.valid_frames = valid_frames; p.z = z; p.rscale = rscale; p.a_sum = a_sum; p.assign = assign;
  p.assign_in = assign_in; p.debug_clock = debug_clock;
  CUtensorMap tx, tw;
  if (int rc = make_tmap_3d(&tx, x, 2, D, T, B, ldx, x_batch_stride, 64, 128)) return rc;
  if (assign_in != nullptr) tw = tx;   // unused in this mode
  else if (int rc = make_tmap_3d(&tw, wc, 2, K, D, 1, ldw, 0, 64, 64)) return rc;
  CUtensorMap tz;
  LPM_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0, "netvlad_pool_fwd: z must be 16-byte aligned");
  if (int rc = make_tmap_3d(&tz, z, 2, D, K, B, D, (uint64_t)K * D, 64, 128)) return rc;
  if (K <= 64) return launch_pool<64>(tx, tw, tz, p, st);
  if (K <= 128) return launch_pool<128>(tx, tw, tz, p, st);
  return launch_pool<256>(tx, tw, tz, p, st);
}

}  // namespace lpm
