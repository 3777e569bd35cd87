// Attention core on the 5th-generation tensor cores for the cluster attention of NetVladV1 at 256 clusters
// (transformer_utils.py:563-581: softmax(q * depth^-0.5 k^T) v per head, depth 16, length 256) and its backward.
//
// Layout trick: one CTA owns FOUR adjacent heads of one sample.  Their q (k, v, dO) columns are 64 contiguous fp16 =
// one 128-byte TMA row, so a [256 x 64] box pair lands in shared memory as a 128B-swizzled K-major UMMA operand and
// head h' of the group is simply "k-step h'" of that operand (+32 bytes on the descriptor start address):
//   S  = Q K^T      one tcgen05.mma per 128 queries (M = 128, N = keys, K = 16)
//   the row softmax runs with one thread per query row straight out of TMEM (no shuffles), P goes to shared memory
//   as an fp16 operand tile and O = P V (M = 128, N = 16, K = 256) accumulates in TMEM.
// Backward: S and dP = dO V^T per (128 queries x 128 keys) unit, P = exp2(S c - lse), dS = P o (dP - delta) in
// registers, both parked as fp16 operand tiles; dV += P^T dO, dK += dS^T Q, dQ += dS K are N = 16 MMAs whose
// accumulators stay in TMEM for the whole head.  Results are written in place over the dead q / k / v columns of the
// staged tiles and leave by TMA.  Nothing but q, k, v, o, dO, lse is read from and dq, dk, dv (o, lse) written to HBM.
//
// Warp roles: the last warp(s) = TMA + MMA issuers (whole warp, elected lane per instruction); the others are softmax /
// gradient warps (TMEM lane quarter = warp % 4).  mbarrier waits are bounded (lpm_common.cuh): a protocol bug traps instead of hanging the GPU.
#include <cstdlib>

#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

namespace {

constexpr int AL = 256;                 // sequence length handled by these kernels
constexpr int TILE_BYTES = AL * 128;    // [256 rows][64 fp16] 128B-swizzled tile of four heads
constexpr int HALF_TILE = 128 * 128;    // one 128-row TMA box

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// byte offset of 16-byte chunk c (8 fp16 columns) of row r inside a 128B-swizzled [rows][64] tile
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

// L2 prefetch of one TMA box (no shared memory involved): the CTAs of a wave run in lock step -- all load, all compute,
// all store -- so each CTA asks for the tiles of the CTA that will follow it on the machine while it computes.
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// 16-byte store to a shared-memory address (explicit state space: a pointer selected at run time among several
// buffers otherwise compiles to generic ST.E stores, which go through the local/global queue)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
struct FwdSmem {
  static constexpr int Q_OFF = 0, K_OFF = TILE_BYTES, V_OFF = 2 * TILE_BYTES;
  static constexpr int P_OFF = 3 * TILE_BYTES;             // 2 buffers x [4 key blocks][128 query rows][128 B]
  static constexpr int P_BYTES = 4 * HALF_TILE;
  static constexpr int BAR_OFF = P_OFF + 2 * P_BYTES;      // 229376
  static constexpr int TOTAL = BAR_OFF + 128 + 1024;       // + alignment slack
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

constexpr int O_COL = 0;   // P V partial accumulators inside the logit region of the unit

__global__ void __launch_bounds__(288, 1)
mha_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                  int H, int Dm, float scale_log2, float* __restrict__ lse, int nch) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sQ = smem + FwdSmem::Q_OFF;
  uint8_t* sK = smem + FwdSmem::K_OFF;
  uint8_t* sV = smem + FwdSmem::V_OFF;
  uint8_t* sP = smem + FwdSmem::P_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR_OFF);
  uint64_t* tile_full = bars;          // [1]
  uint64_t* s_full = bars + 1;         // [2] logits of a unit are in TMEM region r
  uint64_t* s_empty = bars + 3;        // [2] region r drained (4 warps arrive)
  uint64_t* p_ready = bars + 5;        // [2] probabilities of a unit are in shared memory (4 warps arrive)
  uint64_t* o_full = bars + 7;         // [2] P V of a unit retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  const int groups = H >> 2;
  const int b = blockIdx.x / groups, h0 = (blockIdx.x % groups) * 4;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv);
      tma_prefetch_desc(&tmap_out);
      mbar_init(tile_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); mbar_init(&p_ready[i], 4); mbar_init(&o_full[i], 1);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // whole warp, warp-uniform values; single instructions are issued by an elected lane (umma_f16_w)
    if (lane == 0) {
      mbar_expect_tx(tile_full, 3 * TILE_BYTES);
      for (int ft = 0; ft < 2; ++ft) {
        tma_load_3d(sQ + ft * HALF_TILE, &tmap_qkv, tile_full, h0 * 16, ft * 128, b);
        tma_load_3d(sK + ft * HALF_TILE, &tmap_qkv, tile_full, Dm + h0 * 16, ft * 128, b);
        tma_load_3d(sV + ft * HALF_TILE, &tmap_qkv, tile_full, 2 * Dm + h0 * 16, ft * 128, b);
      }
    }
    __syncwarp();
    mbar_wait(tile_full, 0);
    tc_fence_after();
    constexpr uint32_t idesc_s = umma_idesc_f16(128, 256, 0, 0);   // A = Q (K-major), B = K (K-major), N = 256 keys
    constexpr uint32_t idesc_o = umma_idesc_f16(128, 16, 0, 1);    // A = P (K-major), B = V (MN-major), N = depth
    const uint64_t q_d = umma_smem_desc(smem_u32(sQ), 16, 1024), k_d = umma_smem_desc(smem_u32(sK), 16, 1024);
    const uint64_t v_d = umma_smem_desc(smem_u32(sV), HALF_TILE, 1024), p_d = umma_smem_desc(smem_u32(sP), 16, 1024);
    auto issue_s = [&](int u) {
      const int hp = u >> 1, ft = u & 1, r = u & 1;
      mbar_wait(&s_empty[r], ((u >> 1) & 1) ^ 1);
      tc_fence_after();
      umma_f16_w(tmem_base + r * 256, q_d + ((ft * HALF_TILE + hp * 32) >> 4), k_d + ((hp * 32) >> 4), idesc_s, 0u);
      umma_commit_w(&s_full[r]);
    };
    issue_s(0);
    issue_s(1);
#pragma unroll 1
    for (int u = 0; u < 8; ++u) {
      const int hp = u >> 1, r = u & 1;
      mbar_wait(&p_ready[r], (u >> 1) & 1);
      tc_fence_after();
      const uint64_t pa = p_d + ((r * FwdSmem::P_BYTES) >> 4), vb = v_d + ((hp * 32) >> 4);
      const uint32_t o_acc = tmem_base + r * 256 + O_COL;
      // four interleaved accumulation chains (keys ks = 0,4,8,12 | 1,5,9,13 | ...): consecutive MMAs never wait on
      // each other's 16-column accumulator; the epilogue adds the four partial sums
#pragma unroll
      for (int ks = 0; ks < 16; ++ks)
        umma_f16_w(o_acc + (ks & 3) * 16, pa + (((ks >> 2) * HALF_TILE + (ks & 3) * 32) >> 4), vb + ((ks * 2048) >> 4),
                   idesc_o, ks >= 4 ? 1u : 0u);
      umma_commit_w(&o_full[r]);
      if (u + 2 < 8) issue_s(u + 2);
    }
  } else {
    // ------------------------ softmax warps: warpgroup w takes the units u = w, w + 2, ... ------------------------
    const int w = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    for (int u = w; u < 8; u += 2) {
      const int hp = u >> 1, ft = u & 1, r = w;
      const uint32_t s_addr = tmem_base + lane_addr + r * 256;
      const uint32_t pbuf = smem_u32(sP) + r * FwdSmem::P_BYTES;
      mbar_wait(&s_full[r], (u >> 1) & 1);
      tc_fence_after();
      // pass 1: row maximum of the raw logits (the next 32 columns are in flight while these are reduced)
      float m = -INFINITY;
      {
        uint32_t va[32], vb[32];
        auto rmax = [&](const uint32_t* v) {
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]);
#pragma unroll
          for (int i = 2; i < 32; i += 2) { m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1])); }
          m = fmaxf(m, fmaxf(m0, m1));
        };
        tmem_ld32(s_addr, va);
#pragma unroll 1
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait();
          tmem_ld32(s_addr + (c + 1) * 32, vb);
          rmax(va);
          tmem_ld_wait();
          if (c + 2 < nch) tmem_ld32(s_addr + (c + 2) * 32, va);
          rmax(vb);
        }
      }
      // pass 2: p = 2^((s - m) c) -> fp16 operand tile; row sum in fp32.  The tile of unit u - 2 was released by the
      // o_full wait of this warpgroup's previous unit.
      const float nm = -m * scale_log2;
      float sum0 = 0.f, sum1 = 0.f;
      {
        uint32_t va[32], vb[32];
        auto emit = [&](const uint32_t* v, int c) {
          const uint32_t prow = pbuf + (c >> 1) * HALF_TILE;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float e0 = fast_exp2(fmaf(__uint_as_float(v[j * 8 + 2 * i]), scale_log2, nm));
              const float e1 = fast_exp2(fmaf(__uint_as_float(v[j * 8 + 2 * i + 1]), scale_log2, nm));
              sum0 += e0; sum1 += e1;
              pk[i] = pack_half2(e0, e1);
            }
            sts128(prow + sw128(row, (c & 1) * 4 + j), pk[0], pk[1], pk[2], pk[3]);
          }
        };
        tmem_ld32(s_addr, va);
#pragma unroll 1
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait();
          tmem_ld32(s_addr + (c + 1) * 32, vb);
          emit(va, c);
          tmem_ld_wait();
          if (c + 2 < nch) tmem_ld32(s_addr + (c + 2) * 32, va);
          emit(vb, c + 1);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[r]);
      const float l = sum0 + sum1;
      const float inv = 1.f / l;
      const int q = ft * 128 + row;
      if (lse != nullptr)
        lse[((long long)b * H + h0 + hp) * AL + q] = (m * scale_log2 + log2f(l)) * 0.6931471805599453f;
      // O of this unit: four partial sums in 64 columns of the (dead) logit region
      mbar_wait(&o_full[r], (u >> 1) & 1);
      tc_fence_after();
      uint32_t ov[32], ow[32];
      tmem_ld32(s_addr + O_COL, ov);
      tmem_ld32(s_addr + O_COL + 32, ow);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[r]);
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x0 = (__uint_as_float(ov[2 * i]) + __uint_as_float(ov[16 + 2 * i])) +
                         (__uint_as_float(ow[2 * i]) + __uint_as_float(ow[16 + 2 * i]));
        const float x1 = (__uint_as_float(ov[2 * i + 1]) + __uint_as_float(ov[16 + 2 * i + 1])) +
                         (__uint_as_float(ow[2 * i + 1]) + __uint_as_float(ow[16 + 2 * i + 1]));
        pk[i] = pack_half2(x0 * inv, x1 * inv);
      }
      // in place over this head's (dead) q columns of the staged tile
      *reinterpret_cast<uint4*>(sQ + sw128(q, 2 * hp)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(sQ + sw128(q, 2 * hp + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    fence_proxy_async_smem();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    if (lane == 0) {
      for (int ft = 0; ft < 2; ++ft) tma_store_3d(&tmap_out, sQ + ft * HALF_TILE, h0 * 16, ft * 128, b);
      bulk_commit();
      bulk_wait<0>();
    }
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// forward, version 2: P stays in TMEM.  One warpgroup + one issuing warp per CTA, 256 TMEM columns and 96 KB of shared
// memory (the three staged tiles), so TWO CTAs share an SM: one CTA's exponentials run while the other waits for its
// logits / P V, and the load / store phases of the CTAs no longer run in lock step.  Per unit (head, 128 queries):
//   S [128 x 256] = Q K^T in columns 0..255; pass 1 row maximum; pass 2 p = 2^((s - m) c) packed as fp16 into columns
//   0..127 IN PLACE (chunk c of 32 logits becomes 16 columns at 16 c: always behind the columns still to be read);
//   O = P V with P as the TMEM A operand, four interleaved accumulation chains in columns 128..191.
// ------------------------------------------------------------------------------------------------------------------
struct Fwd2Smem {
  static constexpr int Q_OFF = 0, K_OFF = TILE_BYTES, V_OFF = 2 * TILE_BYTES;
  static constexpr int BAR_OFF = 3 * TILE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 64 + 1024;      // + alignment slack
  static_assert(2 * (TOTAL + 1024) <= 233472, "two CTAs per SM must fit");
};

__global__ void __launch_bounds__(160, 2)
mha_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                   int H, int Dm, float scale_log2, float* __restrict__ lse, int nch) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sQ = smem + Fwd2Smem::Q_OFF;
  uint8_t* sK = smem + Fwd2Smem::K_OFF;
  uint8_t* sV = smem + Fwd2Smem::V_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Fwd2Smem::BAR_OFF);
  uint64_t* tile_full = bars;          // the three tiles have landed
  uint64_t* s_full = bars + 1;         // logits of the unit are in TMEM
  uint64_t* p_ready = bars + 2;        // probabilities are in TMEM (4 warps arrive)
  uint64_t* o_full = bars + 3;         // P V retired
  uint64_t* o_empty = bars + 4;        // the accumulators have been read (4 warps arrive): the next logits may land
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  constexpr int OCOL = 128;            // P V partial accumulators (4 chains x 16 columns)

  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  const int groups = H >> 2;
  const int b = blockIdx.x / groups, h0 = (blockIdx.x % groups) * 4;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv);
      tma_prefetch_desc(&tmap_out);
      mbar_init(tile_full, 1);
      mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(o_full, 1); mbar_init(o_empty, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(tile_full, 3 * TILE_BYTES);
      for (int ft = 0; ft < 2; ++ft) {
        tma_load_3d(sQ + ft * HALF_TILE, &tmap_qkv, tile_full, h0 * 16, ft * 128, b);
        tma_load_3d(sK + ft * HALF_TILE, &tmap_qkv, tile_full, Dm + h0 * 16, ft * 128, b);
        tma_load_3d(sV + ft * HALF_TILE, &tmap_qkv, tile_full, 2 * Dm + h0 * 16, ft * 128, b);
      }
    }
    __syncwarp();
    mbar_wait(tile_full, 0);
    tc_fence_after();
    constexpr uint32_t idesc_s = umma_idesc_f16(128, 256, 0, 0);   // A = Q (K-major), B = K (K-major), N = 256 keys
    constexpr uint32_t idesc_o = umma_idesc_f16(128, 16, 0, 1);    // A = P (TMEM), B = V (MN-major), N = depth
    const uint64_t q_d = umma_smem_desc(smem_u32(sQ), 16, 1024), k_d = umma_smem_desc(smem_u32(sK), 16, 1024);
    const uint64_t v_d = umma_smem_desc(smem_u32(sV), HALF_TILE, 1024);
#pragma unroll 1
    for (int u = 0; u < 8; ++u) {
      const int hp = u >> 1, ft = u & 1;
      mbar_wait(o_empty, (u & 1) ^ 1);         // unit u - 1 drained (passes immediately for u = 0)
      tc_fence_after();
      umma_f16_w(tmem_base, q_d + ((ft * HALF_TILE + hp * 32) >> 4), k_d + ((hp * 32) >> 4), idesc_s, 0u);
      umma_commit_w(s_full);
      mbar_wait(p_ready, u & 1);
      tc_fence_after();
      const uint64_t vb = v_d + ((hp * 32) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 16; ++ks)
          umma_f16_ts(tmem_base + OCOL + (ks & 3) * 16, tmem_base + ks * 8, vb + ((ks * 2048) >> 4), idesc_o, ks >= 4 ? 1u : 0u);
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t s_addr = tmem_base + (uint32_t(warp * 32) << 16);
#pragma unroll 1
    for (int u = 0; u < 8; ++u) {
      const int hp = u >> 1, ft = u & 1;
      mbar_wait(s_full, u & 1);
      tc_fence_after();
      float m = -INFINITY;
      {
        uint32_t va[32], vb[32];
        auto rmax = [&](const uint32_t* v) {
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]);
#pragma unroll
          for (int i = 2; i < 32; i += 2) { m0 = fmaxf(m0, __uint_as_float(v[i])); m1 = fmaxf(m1, __uint_as_float(v[i + 1])); }
          m = fmaxf(m, fmaxf(m0, m1));
        };
        tmem_ld32(s_addr, va);
#pragma unroll 1
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait();
          tmem_ld32(s_addr + (c + 1) * 32, vb);
          rmax(va);
          tmem_ld_wait();
          if (c + 2 < nch) tmem_ld32(s_addr + (c + 2) * 32, va);
          rmax(vb);
        }
      }
      const float nm = -m * scale_log2;
      float sum0 = 0.f, sum1 = 0.f;
      {
        uint32_t va[32], vb[32];
        auto emit = [&](const uint32_t* v, int c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float e0 = fast_exp2(fmaf(__uint_as_float(v[2 * i]), scale_log2, nm));
            const float e1 = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2, nm));
            sum0 += e0; sum1 += e1;
            pk[i] = pack_half2(e0, e1);
          }
          tmem_st16(s_addr + c * 16, pk);      // in place: columns 16 c .. 16 c + 15 were read at least one chunk ago
        };
        tmem_ld32(s_addr, va);
#pragma unroll 1
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait();
          tmem_ld32(s_addr + (c + 1) * 32, vb);
          emit(va, c);
          tmem_ld_wait();
          if (c + 2 < nch) tmem_ld32(s_addr + (c + 2) * 32, va);
          emit(vb, c + 1);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      const float l = sum0 + sum1;
      const float inv = 1.f / l;
      const int q = ft * 128 + row;
      if (lse != nullptr)
        lse[((long long)b * H + h0 + hp) * AL + q] = (m * scale_log2 + log2f(l)) * 0.6931471805599453f;
      mbar_wait(o_full, u & 1);
      tc_fence_after();
      uint32_t ov[32], ow[32];
      tmem_ld32(s_addr + OCOL, ov);
      tmem_ld32(s_addr + OCOL + 32, ow);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x0 = (__uint_as_float(ov[2 * i]) + __uint_as_float(ov[16 + 2 * i])) +
                         (__uint_as_float(ow[2 * i]) + __uint_as_float(ow[16 + 2 * i]));
        const float x1 = (__uint_as_float(ov[2 * i + 1]) + __uint_as_float(ov[16 + 2 * i + 1])) +
                         (__uint_as_float(ow[2 * i + 1]) + __uint_as_float(ow[16 + 2 * i + 1]));
        pk[i] = pack_half2(x0 * inv, x1 * inv);
      }
      const uint32_t qt = smem_u32(sQ);
      sts128(qt + sw128(q, 2 * hp), pk[0], pk[1], pk[2], pk[3]);
      sts128(qt + sw128(q, 2 * hp + 1), pk[4], pk[5], pk[6], pk[7]);
    }
    fence_proxy_async_smem();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    if (lane == 0) {
      for (int ft = 0; ft < 2; ++ft) tma_store_3d(&tmap_out, sQ + ft * HALF_TILE, h0 * 16, ft * 128, b);
      bulk_commit();
      bulk_wait_read<0>();
    }
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
struct BwdSmem {
  static constexpr int Q_OFF = 0, K_OFF = TILE_BYTES, V_OFF = 2 * TILE_BYTES, DO_OFF = 3 * TILE_BYTES;
  static constexpr int SLOT_OFF = 4 * TILE_BYTES;          // 3 operand slots x [2 key blocks][128 query rows][128 B]
  static constexpr int SLOT_BYTES = 2 * HALF_TILE;
  static constexpr int BAR_OFF = SLOT_OFF + 3 * SLOT_BYTES;   // 229376
  static constexpr int TOTAL = BAR_OFF + 192 + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

// TMEM columns: [0,256) two halves of (S 64 | dP 64) for the unit in flight; [256 + 96 a, +96) accumulators of head
// parity a: dQ (2 query tiles x 16) | dK (2 key tiles x 16) | dV (2 key tiles x 16).
constexpr int NEXT_WAVE = 148;   // one CTA per SM: CTA i + 148 follows CTA i (L2 prefetch distance)
constexpr int ACC_COL = 256;
constexpr int ACC_COLS = 96;

__global__ void __launch_bounds__(608, 1)
mha_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                  const __grid_constant__ CUtensorMap tmap_dqkv, const __half* __restrict__ o, long long ldo,
                  const float* __restrict__ lse, int H, int Dm, float scale, long long* dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sQ = smem + BwdSmem::Q_OFF;
  uint8_t* sK = smem + BwdSmem::K_OFF;
  uint8_t* sV = smem + BwdSmem::V_OFF;
  uint8_t* sdO = smem + BwdSmem::DO_OFF;
  uint8_t* sSlot = smem + BwdSmem::SLOT_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::BAR_OFF);
  uint64_t* tile_full = bars;          // [1]
  uint64_t* sdp_full = bars + 1;       // [2] S | dP of one 64-key half of the unit are in TMEM
  uint64_t* sdp_empty = bars + 3;      // [2] ... and have been read (8 warps arrive)
  uint64_t* slot_full = bars + 13;     // [3] an operand slot has been written (16 warps arrive).  The issuing warps wait on
                                       //     the SLOT barriers: a slot cannot be refilled before its readers' MMAs retired, so a
                                       //     reader is never two phases behind (a per-unit barrier can lap a slow reader)
  uint64_t* slot_free = bars + 6;      // [3] the MMAs reading an operand slot retired
  uint64_t* acc_full = bars + 9;       // [2] gradients of a head complete in TMEM
  uint64_t* acc_empty = bars + 11;     // [2] ... and read (16 warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  const int groups = H >> 2;
  const int b = blockIdx.x / groups, h0 = (blockIdx.x % groups) * 4;
  const float scale_log2 = scale * 1.4426950408889634f;

  if (warp == 16) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv);
      tma_prefetch_desc(&tmap_do);
      tma_prefetch_desc(&tmap_dqkv);
      mbar_init(tile_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&sdp_full[i], 1); mbar_init(&sdp_empty[i], 8);
        mbar_init(&acc_full[i], 3); mbar_init(&acc_empty[i], 16);     // three issuing warps commit a head
      }
      for (int i = 0; i < 3; ++i) mbar_init(&slot_full[i], 16);
      for (int i = 0; i < 3; ++i) mbar_init(&slot_free[i], 2);        // two of the issuing warps read every slot
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 16) {
    // Three issuing warps (one per scheduler), each the whole warp on warp-uniform values with single instructions
    // issued by an elected lane (umma_f16_w): warp 16 = TMA loads, S / dP and the dV chain; warp 17 = dK; warp 18 = dQ.
    // One warp issuing all 28 MMAs of a unit took ~1.1 k cycles per unit (clock stamps), longer than the math.
    const int role = warp - 16;
    if (role == 0 && lane == 0) {
      mbar_expect_tx(tile_full, 4 * TILE_BYTES);
      for (int ft = 0; ft < 2; ++ft) {
        tma_load_3d(sQ + ft * HALF_TILE, &tmap_qkv, tile_full, h0 * 16, ft * 128, b);
        tma_load_3d(sK + ft * HALF_TILE, &tmap_qkv, tile_full, Dm + h0 * 16, ft * 128, b);
        tma_load_3d(sV + ft * HALF_TILE, &tmap_qkv, tile_full, 2 * Dm + h0 * 16, ft * 128, b);
        tma_load_3d(sdO + ft * HALF_TILE, &tmap_do, tile_full, h0 * 16, ft * 128, b);
      }
      const int nxt = blockIdx.x + NEXT_WAVE;            // the group a later CTA of this SM count will load
      if (nxt < (int)gridDim.x) {
        const int nb = nxt / groups, nh = (nxt % groups) * 4;
        for (int ft = 0; ft < 2; ++ft) {
          tma_prefetch_l2_3d(&tmap_qkv, nh * 16, ft * 128, nb);
          tma_prefetch_l2_3d(&tmap_qkv, Dm + nh * 16, ft * 128, nb);
          tma_prefetch_l2_3d(&tmap_qkv, 2 * Dm + nh * 16, ft * 128, nb);
          tma_prefetch_l2_3d(&tmap_do, nh * 16, ft * 128, nb);
        }
      }
    }
    __syncwarp();
    mbar_wait(tile_full, 0);
    tc_fence_after();
    constexpr uint32_t idesc_1 = umma_idesc_f16(128, 64, 0, 0);    // S / dP half: both operands K-major
    constexpr uint32_t idesc_kv = umma_idesc_f16(128, 16, 1, 1);   // dV, dK: A = P^T / dS^T (MN-major), B MN-major
    constexpr uint32_t idesc_q = umma_idesc_f16(128, 16, 0, 1);    // dQ: A = dS (K-major), B = K (MN-major)
    // K-major views (S, dP operands; dS as the A operand of dQ) and MN-major views (B operands of the gradient MMAs,
    // P^T / dS^T as A operands) of the same tiles
    const uint64_t q_k = umma_smem_desc(smem_u32(sQ), 16, 1024), k_k = umma_smem_desc(smem_u32(sK), 16, 1024);
    const uint64_t v_k = umma_smem_desc(smem_u32(sV), 16, 1024), do_k = umma_smem_desc(smem_u32(sdO), 16, 1024);
    const uint64_t q_m = umma_smem_desc(smem_u32(sQ), HALF_TILE, 1024), k_m = umma_smem_desc(smem_u32(sK), HALF_TILE, 1024);
    const uint64_t do_m = umma_smem_desc(smem_u32(sdO), HALF_TILE, 1024);
    const uint64_t slot_k = umma_smem_desc(smem_u32(sSlot), 16, 1024), slot_m = umma_smem_desc(smem_u32(sSlot), HALF_TILE, 1024);
    // unit u = (head hp, query tile ft, key chunk kc) = (u >> 2, (u >> 1) & 1, u & 1)
    auto issue_1 = [&](int u) {
      const int hp = u >> 2, ft = (u >> 1) & 1, kc = u & 1;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        mbar_wait(&sdp_empty[half], (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t qo = (uint32_t)(ft * HALF_TILE + hp * 32) >> 4;
        const uint32_t ko = (uint32_t)((kc * 128 + half * 64) * 128 + hp * 32) >> 4;
        umma_f16_w(tmem_base + half * 128, q_k + qo, k_k + ko, idesc_1, 0u);
        umma_f16_w(tmem_base + half * 128 + 64, do_k + qo, v_k + ko, idesc_1, 0u);
        umma_commit_w(&sdp_full[half]);
      }
    };
    if (role == 0) issue_1(0);
    const bool stamp = dbg != nullptr && blockIdx.x == 0 && lane == 0;
#pragma unroll 1
    for (int u = 0; u < 16; ++u) {
      const int hp = u >> 2, ft = (u >> 1) & 1, kc = u & 1, a = hp & 1;
      if (role == 0 && u + 1 < 16) issue_1(u + 1);
      if (stamp) dbg[256 + role * 64 + u * 4 + 0] = clock64();   // (warp 16) S/dP of unit u + 1 issued
      const int kp = 2 * u, kd = 2 * u + 1;
      if (role < 2) mbar_wait(&slot_full[kp % 3], (kp / 3) & 1);       // dV reads P; dK also arrives on the P slot
      if (role > 0) mbar_wait(&slot_full[kd % 3], (kd / 3) & 1);       // dK, dQ read dS
      tc_fence_after();
      if (stamp) dbg[256 + role * 64 + u * 4 + 1] = clock64();   // P / dS of unit u visible
      if ((u & 3) == 0 && hp >= 2) {
        mbar_wait(&acc_empty[a], ((hp >> 1) & 1) ^ 1);
        tc_fence_after();
      }
      const uint32_t sp = (uint32_t)((kp % 3) * BwdSmem::SLOT_BYTES) >> 4, sd = (uint32_t)((kd % 3) * BwdSmem::SLOT_BYTES) >> 4;
      const uint32_t acc = tmem_base + ACC_COL + a * ACC_COLS;
      const uint32_t bq = (uint32_t)(ft * HALF_TILE + hp * 32) >> 4;     // rows of query tile ft, columns of head hp
      const uint32_t bk = (uint32_t)(kc * HALF_TILE + hp * 32) >> 4;     // rows of key chunk kc
      if (role == 0) {
        // dV[keys of chunk kc] += P^T dO   (contraction over the 128 queries of tile ft)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_f16_w(acc + 64 + kc * 16, slot_m + sp + ((ks * 2048) >> 4), do_m + bq + ((ks * 2048) >> 4), idesc_kv,
                     (ft > 0 || ks > 0) ? 1u : 0u);
        umma_commit_w(&slot_free[kp % 3]);
      } else if (role == 1) {
        // dK[keys of chunk kc] += dS^T Q
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_f16_w(acc + 32 + kc * 16, slot_m + sd + ((ks * 2048) >> 4), q_m + bq + ((ks * 2048) >> 4), idesc_kv,
                     (ft > 0 || ks > 0) ? 1u : 0u);
        umma_commit_w(&slot_free[kp % 3]);     // second arrival on the P slot (a slot barrier always takes two)
        umma_commit_w(&slot_free[kd % 3]);
      } else {
        // dQ[queries of tile ft] += dS K     (contraction over the 128 keys of chunk kc)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_f16_w(acc + ft * 16, slot_k + sd + (((ks >> 2) * HALF_TILE + (ks & 3) * 32) >> 4), k_m + bk + ((ks * 2048) >> 4),
                     idesc_q, (kc > 0 || ks > 0) ? 1u : 0u);
        umma_commit_w(&slot_free[kd % 3]);
      }
      if ((u & 3) == 3) umma_commit_w(&acc_full[a]);
      if (stamp) dbg[256 + role * 64 + u * 4 + 2] = clock64();   // this warp's gradient MMAs of unit u issued
    }
  } else {
    // ---------------- gradient warps: warpgroup g owns keys [32 g, 32 g + 32) of every 128-key chunk ----------------
    const int g = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const int half = g >> 1, cofs = (g & 1) * 32;
    const uint32_t slot_base = smem_u32(sSlot);

    // per-row constants of a (head, query tile): lse in log2 units and delta = rowsum(dO o O); O rows are prefetched
    // from global memory one unit ahead
    uint4 o_pref[2];
    float lse_pref;
    auto prefetch = [&](int combo) {      // combo = u >> 1 = (hp, ft)
      const int hp = combo >> 1, q = (combo & 1) * 128 + row;
      const __half* orow = o + ((long long)b * AL + q) * ldo + (h0 + hp) * 16;
      o_pref[0] = __ldg(reinterpret_cast<const uint4*>(orow));
      o_pref[1] = __ldg(reinterpret_cast<const uint4*>(orow) + 1);
      lse_pref = __ldg(lse + ((long long)b * H + h0 + hp) * AL + q);
    };
    prefetch(0);
    mbar_wait(tile_full, 0);
    float lse2 = 0.f, delta = 0.f;

    auto epilogue = [&](int hp) {
      const int a = hp & 1;
      mbar_wait(&acc_full[a], (hp >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int pc = g + 4 * t;           // piece: 0,1 = dQ tiles; 2,3 = dK tiles; 4,5 = dV tiles
        if (pc < 6) {
          uint32_t v[16];
          tmem_ld16(tmem_base + lane_addr + ACC_COL + a * ACC_COLS + pc * 16, v);
          tmem_ld_wait();
          const float f = pc < 4 ? scale : 1.f;
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = pack_half2(__uint_as_float(v[2 * i]) * f, __uint_as_float(v[2 * i + 1]) * f);
          uint8_t* tile = pc < 2 ? sQ : (pc < 4 ? sK : sV);
          const int r = (pc & 1) * 128 + row;
          *reinterpret_cast<uint4*>(tile + sw128(r, 2 * hp)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(tile + sw128(r, 2 * hp + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
    };

#pragma unroll 1
    for (int u = 0; u < 16; ++u) {
      const int hp = u >> 2, ft = (u >> 1) & 1;
      if ((u & 1) == 0) {
        const int q = ft * 128 + row;
        lse2 = lse_pref * 1.4426950408889634f;
        const uint4 d0 = *reinterpret_cast<const uint4*>(sdO + sw128(q, 2 * hp));
        const uint4 d1 = *reinterpret_cast<const uint4*>(sdO + sw128(q, 2 * hp + 1));
        const __half2* ho = reinterpret_cast<const __half2*>(o_pref);
        const __half2* h0p = reinterpret_cast<const __half2*>(&d0);
        const __half2* h1p = reinterpret_cast<const __half2*>(&d1);
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = __half22float2(ho[i]), y = __half22float2(h0p[i]);
          const float2 z = __half22float2(ho[4 + i]), t = __half22float2(h1p[i]);
          d += x.x * y.x + x.y * y.y + z.x * t.x + z.y * t.y;
        }
        delta = d;
      } else if (u + 1 < 16) {
        prefetch((u + 1) >> 1);
      }
      long long* dg = (dbg != nullptr && blockIdx.x == 0 && (warp & 3) == 0 && lane == 0) ? dbg + (warp >> 2) * 64 + u * 4 : nullptr;
      if (dg) dg[0] = clock64();          // unit begins (row constants done)
      mbar_wait(&sdp_full[half], u & 1);
      tc_fence_after();
      if (dg) dg[1] = clock64();          // S / dP available
      uint32_t s[32], dp[32];
      tmem_ld32(tmem_base + lane_addr + half * 128 + cofs, s);
      tmem_ld32(tmem_base + lane_addr + half * 128 + 64 + cofs, dp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sdp_empty[half]);
      uint32_t pk[16], dk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(s[2 * i]), scale_log2, -lse2));
        const float p1 = fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), scale_log2, -lse2));
        pk[i] = pack_half2(p0, p1);
        dk[i] = pack_half2(p0 * (__uint_as_float(dp[2 * i]) - delta), p1 * (__uint_as_float(dp[2 * i + 1]) - delta));
      }
      const int kp = 2 * u, kd = 2 * u + 1;
      if (dg) dg[2] = clock64();          // math issued
      mbar_wait(&slot_free[kp % 3], ((kp / 3) & 1) ^ 1);
      mbar_wait(&slot_free[kd % 3], ((kd / 3) & 1) ^ 1);
      if (dg) dg[3] = clock64();          // operand slots free
      const uint32_t bp = slot_base + (kp % 3) * BwdSmem::SLOT_BYTES + (g >> 1) * HALF_TILE;
      const uint32_t bd = slot_base + (kd % 3) * BwdSmem::SLOT_BYTES + (g >> 1) * HALF_TILE;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = sw128(row, (g & 1) * 4 + j);
        sts128(bp + off, pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        sts128(bd + off, dk[4 * j], dk[4 * j + 1], dk[4 * j + 2], dk[4 * j + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&slot_full[kp % 3]); mbar_arrive(&slot_full[kd % 3]); }
      if ((u & 3) == 1 && hp >= 1) epilogue(hp - 1);    // one unit late: the head's last MMAs have retired by now
    }
    epilogue(3);
    fence_proxy_async_smem();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    if (lane == 0) {
      for (int ft = 0; ft < 2; ++ft) {
        tma_store_3d(&tmap_dqkv, sQ + ft * HALF_TILE, h0 * 16, ft * 128, b);
        tma_store_3d(&tmap_dqkv, sK + ft * HALF_TILE, Dm + h0 * 16, ft * 128, b);
        tma_store_3d(&tmap_dqkv, sV + ft * HALF_TILE, 2 * Dm + h0 * 16, ft * 128, b);
      }
      bulk_commit();
      bulk_wait_read<0>();     // the tiles have left shared memory; the global writes complete before the grid does
    }
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// backward, version 2: keys on the TMEM lanes.  S^T = K Q^T and dP^T = V dO^T put a KEY on every lane, so P^T and
// dS^T -- the A operands of dV += P^T dO and dK += dS^T Q -- are written back to TMEM (packed fp16, in place over
// the consumed logits) and read by the tensor core from there: only dS^T for dQ += dS K still goes through shared
// memory (32 KB per unit instead of 64 KB written + 96 KB re-read by the MMAs).  A unit is (128 keys x 128 queries);
// its two 64-query halves belong to two warpgroup pairs that run out of phase (own barriers), so one pair's math
// fills the MUFU while the other waits for its next logits.  lse / delta are per COLUMN here: each warpgroup keeps
// the 32 values of its columns in a small shared-memory scratch (one warp computes them, broadcast LDS.128 reads).
// ------------------------------------------------------------------------------------------------------------------
struct Bwd2Smem {
  static constexpr int Q_OFF = 0, K_OFF = TILE_BYTES, V_OFF = 2 * TILE_BYTES, DO_OFF = 3 * TILE_BYTES;
  static constexpr int SLOT_OFF = 4 * TILE_BYTES;          // 3 dS^T slots x [2 query blocks][128 key rows][128 B]
  static constexpr int SLOT_BYTES = 2 * HALF_TILE;
  static constexpr int ROWC_OFF = SLOT_OFF + 3 * SLOT_BYTES;  // [4 warpgroups][2][32] float2 (-lse*log2e, delta)
  static constexpr int BAR_OFF = ROWC_OFF + 4 * 2 * 32 * 8;
  static constexpr int TOTAL = BAR_OFF + 192;
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

__global__ void __launch_bounds__(576, 1)
mha_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                   const __grid_constant__ CUtensorMap tmap_dqkv, const __half* __restrict__ o, long long ldo,
                   const float* __restrict__ lse, int H, int Dm, float scale, long long* dbg, int pair_delay) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem + Bwd2Smem::Q_OFF;
  uint8_t* sK = smem + Bwd2Smem::K_OFF;
  uint8_t* sV = smem + Bwd2Smem::V_OFF;
  uint8_t* sdO = smem + Bwd2Smem::DO_OFF;
  uint8_t* sSlot = smem + Bwd2Smem::SLOT_OFF;
  float2* sRowc = reinterpret_cast<float2*>(smem + Bwd2Smem::ROWC_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Bwd2Smem::BAR_OFF);
  uint64_t* tile_full = bars;          // [1]
  uint64_t* sdp_full = bars + 1;       // [2] S^T | dP^T of one 64-query half are in TMEM
  uint64_t* pds_ready = bars + 3;      // [2] P^T, dS^T of a half are in TMEM / shared memory (8 warps arrive)
  uint64_t* slot_free = bars + 5;      // [3] the dQ MMAs reading a dS^T slot retired
  uint64_t* acc_full = bars + 8;       // [2] gradients of a head complete in TMEM (both issuing warps commit)
  uint64_t* acc_empty = bars + 10;     // [2] ... and read (16 warps arrive)
  uint64_t* slot_full = bars + 16;     // [3] a dS^T slot has been written (16 warps arrive): what the dQ warp waits on -- it may
                                       //     trail the pairs by up to three units, and a per-unit barrier would lap it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = warp_index_uniform(), lane = threadIdx.x & 31;
  const int groups = H >> 2;
  const int b = blockIdx.x / groups, h0 = (blockIdx.x % groups) * 4;
  const float scale_log2 = scale * 1.4426950408889634f;

  if (warp == 16) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_qkv);
      tma_prefetch_desc(&tmap_do);
      tma_prefetch_desc(&tmap_dqkv);
      mbar_init(tile_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&sdp_full[i], 1); mbar_init(&pds_ready[i], 8);
        mbar_init(&acc_full[i], 2); mbar_init(&acc_empty[i], 16);
      }
      for (int i = 0; i < 3; ++i) { mbar_init(&slot_free[i], 1); mbar_init(&slot_full[i], 16); }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 16) {
    // warp 16: TMA loads, S^T / dP^T and the TMEM-operand chains dV, dK (in order on one thread: the next logits of a
    // half overwrite its P^T / dS^T only after the MMAs reading them); warp 17: dQ from the shared-memory dS^T slots
    const int role = warp - 16;
    if (role == 0 && lane == 0) {
      mbar_expect_tx(tile_full, 4 * TILE_BYTES);
      for (int ft = 0; ft < 2; ++ft) {
        tma_load_3d(sQ + ft * HALF_TILE, &tmap_qkv, tile_full, h0 * 16, ft * 128, b);
        tma_load_3d(sK + ft * HALF_TILE, &tmap_qkv, tile_full, Dm + h0 * 16, ft * 128, b);
        tma_load_3d(sV + ft * HALF_TILE, &tmap_qkv, tile_full, 2 * Dm + h0 * 16, ft * 128, b);
        tma_load_3d(sdO + ft * HALF_TILE, &tmap_do, tile_full, h0 * 16, ft * 128, b);
      }
      const int nxt = blockIdx.x + NEXT_WAVE;            // the group a later CTA of this SM count will load
      if (nxt < (int)gridDim.x) {
        const int nb = nxt / groups, nh = (nxt % groups) * 4;
        for (int ft = 0; ft < 2; ++ft) {
          tma_prefetch_l2_3d(&tmap_qkv, nh * 16, ft * 128, nb);
          tma_prefetch_l2_3d(&tmap_qkv, Dm + nh * 16, ft * 128, nb);
          tma_prefetch_l2_3d(&tmap_qkv, 2 * Dm + nh * 16, ft * 128, nb);
          tma_prefetch_l2_3d(&tmap_do, nh * 16, ft * 128, nb);
        }
      }
    }
    __syncwarp();
    mbar_wait(tile_full, 0);
    tc_fence_after();
    constexpr uint32_t idesc_1 = umma_idesc_f16(128, 64, 0, 0);    // S^T / dP^T half: both operands K-major
    constexpr uint32_t idesc_ts = umma_idesc_f16(128, 16, 0, 1);   // dV, dK: A = P^T / dS^T in TMEM, B MN-major
    constexpr uint32_t idesc_q = umma_idesc_f16(128, 16, 1, 1);    // dQ: A = dS (MN-major view of dS^T), B = K (MN-major)
    const uint64_t q_k = umma_smem_desc(smem_u32(sQ), 16, 1024), k_k = umma_smem_desc(smem_u32(sK), 16, 1024);
    const uint64_t v_k = umma_smem_desc(smem_u32(sV), 16, 1024), do_k = umma_smem_desc(smem_u32(sdO), 16, 1024);
    const uint64_t q_m = umma_smem_desc(smem_u32(sQ), HALF_TILE, 1024), k_m = umma_smem_desc(smem_u32(sK), HALF_TILE, 1024);
    const uint64_t do_m = umma_smem_desc(smem_u32(sdO), HALF_TILE, 1024);
    const uint64_t slot_m = umma_smem_desc(smem_u32(sSlot), HALF_TILE, 1024);
    // unit u = (head hp, query tile ft, key chunk kc) = (u >> 2, (u >> 1) & 1, u & 1); half = 64 queries of the tile
    auto issue_1 = [&](int u, int half) {
      const int hp = u >> 2, ft = (u >> 1) & 1, kc = u & 1;
      const uint32_t ko = (uint32_t)(kc * HALF_TILE + hp * 32) >> 4;                     // 128 key rows of chunk kc
      const uint32_t qo = (uint32_t)((ft * 128 + half * 64) * 128 + hp * 32) >> 4;       // 64 query rows
      umma_f16_w(tmem_base + half * 128, k_k + ko, q_k + qo, idesc_1, 0u);
      umma_f16_w(tmem_base + half * 128 + 64, v_k + ko, do_k + qo, idesc_1, 0u);
      umma_commit_w(&sdp_full[half]);
    };
    if (role == 0) {
      // The second pair starts half a period late: its math then runs while the first pair is storing / waiting for its
      // next logits instead of competing with it for the MUFU (the pairs keep whatever phase they start with).
      issue_1(0, 0);
      const long long t_start = clock64();
      while (clock64() - t_start < pair_delay) { }
      issue_1(0, 1);
    }
#pragma unroll 1
    for (int u = 0; u < 16; ++u) {
      const int hp = u >> 2, ft = (u >> 1) & 1, kc = u & 1, a = hp & 1;
      const uint32_t acc = tmem_base + ACC_COL + a * ACC_COLS;
      if (role == 0) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&pds_ready[half], u & 1);
          tc_fence_after();
          if ((u & 3) == 0 && hp >= 2 && half == 0) {
            mbar_wait(&acc_empty[a], ((hp >> 1) & 1) ^ 1);
            tc_fence_after();
          }
          const uint32_t bq = (uint32_t)((ft * 128 + half * 64) * 128 + hp * 32) >> 4;   // query rows of this half
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // queries 16 ks .. 16 ks + 15 of the half: packed fp16 pairs in 8 TMEM columns of warpgroup (ks >> 1)
            const uint32_t acol = tmem_base + half * 128 + (ks >> 1) * 32 + (ks & 1) * 8;
            const uint32_t first = (ft > 0 || half > 0 || ks > 0) ? 1u : 0u;
            if (elect_one()) {
              umma_f16_ts(acc + 64 + kc * 16, acol, do_m + bq + ((ks * 2048) >> 4), idesc_ts, first);       // dV
              umma_f16_ts(acc + 32 + kc * 16, acol + 64, q_m + bq + ((ks * 2048) >> 4), idesc_ts, first);   // dK
            }
            __syncwarp();
          }
          if (u + 1 < 16) issue_1(u + 1, half);
        }
      } else {
        mbar_wait(&slot_full[u % 3], (u / 3) & 1);
        tc_fence_after();
        if ((u & 3) == 0 && hp >= 2) {
          mbar_wait(&acc_empty[a], ((hp >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t sd = (uint32_t)((u % 3) * Bwd2Smem::SLOT_BYTES) >> 4;
        const uint32_t bk = (uint32_t)(kc * HALF_TILE + hp * 32) >> 4;
        // dQ[queries of tile ft] += dS K   (A = rows of dS^T: 16 keys per k-step, 128 queries = two 64-query blocks)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_f16_w(acc + ft * 16, slot_m + sd + ((ks * 2048) >> 4), k_m + bk + ((ks * 2048) >> 4), idesc_q,
                     (kc > 0 || ks > 0) ? 1u : 0u);
        umma_commit_w(&slot_free[u % 3]);
      }
      if ((u & 3) == 3) umma_commit_w(&acc_full[a]);
    }
  } else {
    // ------------- gradient warps: warpgroup g owns queries [32 g, 32 g + 32) of every 128-query tile -------------
    const int g = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;                 // key row inside the chunk = TMEM lane
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const int half = g >> 1, cofs = (g & 1) * 32;
    const uint32_t slot_base = smem_u32(sSlot);
    uint4 o_pref[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    float lse_pref = 0.f;
    auto prefetch = [&](int combo) {      // combo = (hp, ft): O row and lse of query g * 32 + lane of that tile
      const int hp = combo >> 1, q = (combo & 1) * 128 + g * 32 + lane;
      const __half* orow = o + ((long long)b * AL + q) * ldo + (h0 + hp) * 16;
      o_pref[0] = __ldg(reinterpret_cast<const uint4*>(orow));
      o_pref[1] = __ldg(reinterpret_cast<const uint4*>(orow) + 1);
      lse_pref = __ldg(lse + ((long long)b * H + h0 + hp) * AL + q);
    };
    if (quarter == 0) prefetch(0);
    mbar_wait(tile_full, 0);

    auto epilogue = [&](int hp) {
      const int a = hp & 1;
      mbar_wait(&acc_full[a], (hp >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int pc = g + 4 * t;           // piece: 0,1 = dQ tiles; 2,3 = dK tiles; 4,5 = dV tiles
        if (pc < 6) {
          uint32_t v[16];
          tmem_ld16(tmem_base + lane_addr + ACC_COL + a * ACC_COLS + pc * 16, v);
          tmem_ld_wait();
          const float f = pc < 4 ? scale : 1.f;
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = pack_half2(__uint_as_float(v[2 * i]) * f, __uint_as_float(v[2 * i + 1]) * f);
          const uint32_t tile = smem_u32(pc < 2 ? sQ : (pc < 4 ? sK : sV));
          const int r = (pc & 1) * 128 + row;
          sts128(tile + sw128(r, 2 * hp), pk[0], pk[1], pk[2], pk[3]);
          sts128(tile + sw128(r, 2 * hp + 1), pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
    };

#pragma unroll 1
    for (int u = 0; u < 16; ++u) {
      const int hp = u >> 2, ft = (u >> 1) & 1;
      float2* rowc = sRowc + (g * 2 + ((u >> 1) & 1)) * 32;
      if ((u & 1) == 0) {
        // column constants of this warpgroup's 32 queries for the two units of (hp, ft): one warp computes them from
        // the O row / lse it fetched one unit earlier (no global-memory latency in front of the warpgroup barrier)
        if (quarter == 0) {
          const int q = ft * 128 + g * 32 + lane;
          const uint4 d0 = *reinterpret_cast<const uint4*>(sdO + sw128(q, 2 * hp));
          const uint4 d1 = *reinterpret_cast<const uint4*>(sdO + sw128(q, 2 * hp + 1));
          const __half2* a0 = reinterpret_cast<const __half2*>(&o_pref[0]);
          const __half2* a1 = reinterpret_cast<const __half2*>(&o_pref[1]);
          const __half2* b0 = reinterpret_cast<const __half2*>(&d0);
          const __half2* b1 = reinterpret_cast<const __half2*>(&d1);
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 x = __half22float2(a0[i]), y = __half22float2(b0[i]);
            const float2 z = __half22float2(a1[i]), t = __half22float2(b1[i]);
            d += x.x * y.x + x.y * y.y + z.x * t.x + z.y * t.y;
          }
          rowc[lane] = make_float2(-lse_pref * 1.4426950408889634f, d);
        }
        named_barrier(1 + g, 128);      // the warpgroup's four warps (a one-way mbarrier measured the same: 266 us)
      } else if (quarter == 0 && u + 1 < 16) {
        prefetch((u + 1) >> 1);
      }
      long long* dg = (dbg != nullptr && blockIdx.x == 0 && quarter == 0 && lane == 0) ? dbg + g * 128 + u * 8 : nullptr;
      if (dg) dg[0] = clock64();          // unit begins (column constants done)
      mbar_wait(&sdp_full[half], u & 1);
      tc_fence_after();
      if (dg) dg[1] = clock64();          // S^T / dP^T available
      uint32_t s[32], dp[32];
      tmem_ld32(tmem_base + lane_addr + half * 128 + cofs, s);
      tmem_ld32(tmem_base + lane_addr + half * 128 + 64 + cofs, dp);
      tmem_ld_wait();
      uint32_t pk[16], dk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 c = *reinterpret_cast<const float4*>(&rowc[2 * i]);     // (-lse2, delta) of queries 2i, 2i + 1
        const float p0 = fast_exp2(fmaf(__uint_as_float(s[2 * i]), scale_log2, c.x));
        const float p1 = fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), scale_log2, c.z));
        pk[i] = pack_half2(p0, p1);
        dk[i] = pack_half2(p0 * (__uint_as_float(dp[2 * i]) - c.y), p1 * (__uint_as_float(dp[2 * i + 1]) - c.w));
      }
      if (dg) dg[2] = clock64();          // math issued
      // A operands of dV / dK: back into TMEM, in place over this thread's consumed columns
      tmem_st16(tmem_base + lane_addr + half * 128 + cofs, pk);
      tmem_st16(tmem_base + lane_addr + half * 128 + 64 + cofs, dk);
      // dS^T row of this key for dQ: [query block = half][key row][64 queries], 128B-swizzled
      mbar_wait(&slot_free[u % 3], ((u / 3) & 1) ^ 1);
      if (dg) dg[3] = clock64();          // dS^T slot free
      const uint32_t bd = slot_base + (u % 3) * Bwd2Smem::SLOT_BYTES + half * HALF_TILE;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(bd + sw128(row, (g & 1) * 4 + j), dk[4 * j], dk[4 * j + 1], dk[4 * j + 2], dk[4 * j + 3]);
      if (dg) dg[4] = clock64();          // stores issued
      fence_proxy_async_smem();
      if (dg) dg[5] = clock64();          // proxy fence done
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&slot_full[u % 3]); mbar_arrive(&pds_ready[half]); }
      if (dg) dg[6] = clock64();          // arrived
      if ((u & 3) == 1 && hp >= 1) epilogue(hp - 1);
    }
    epilogue(3);
    fence_proxy_async_smem();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    if (lane == 0) {
      for (int ft = 0; ft < 2; ++ft) {
        tma_store_3d(&tmap_dqkv, sQ + ft * HALF_TILE, h0 * 16, ft * 128, b);
        tma_store_3d(&tmap_dqkv, sK + ft * HALF_TILE, Dm + h0 * 16, ft * 128, b);
        tma_store_3d(&tmap_dqkv, sV + ft * HALF_TILE, 2 * Dm + h0 * 16, ft * 128, b);
      }
      bulk_commit();
      bulk_wait_read<0>();     // the tiles have left shared memory; the global writes complete before the grid does
    }
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

long long* g_mha_dbg = nullptr;
int g_mha_tc_mode = -1;   // -1: read LPM_MHA_TC once; 0: warp-level mma.sync kernels only; 1: tcgen05 backward with operands in shared
                          // memory; 2 (default): tcgen05 backward with the dV / dK operands in TMEM

bool tc_enabled() {
  if (g_mha_tc_mode < 0) {
    const char* e = getenv("LPM_MHA_TC");
    g_mha_tc_mode = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
  }
  return g_mha_tc_mode != 0;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

constexpr int MHA_TC_FWD_DEFAULT = 0;
int g_mha_tc_fwd = -1;    // forward on a tcgen05 kernel: -1 = read LPM_MHA_TC_FWD once; 0 = warp-level kernel, 1 = P through shared
                          // memory, 2 = P in TMEM, two CTAs per SM
void mha_set_tc_mode(int mode) {
  g_mha_tc_fwd = (mode >> 2) & 3;
  mode &= 3;
  g_mha_tc_mode = mode > 2 ? 2 : mode;
}
int mha_tc_backward_mode() { tc_enabled(); return g_mha_tc_mode; }
bool mha_tc_forward_enabled() {
  if (g_mha_tc_fwd < 0) {
    const char* e = getenv("LPM_MHA_TC_FWD");
    g_mha_tc_fwd = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : MHA_TC_FWD_DEFAULT;
  }
  return g_mha_tc_fwd != 0;
}
void mha_set_debug_clock(long long* buf) { g_mha_dbg = buf; }

bool mha_tc_eligible(int L, int Dm, int H, long long ld, long long ldo, const void* p0, const void* p1, const void* p2) {
  return H > 0 && Dm == H * 16 && L == AL && (H & 3) == 0 && ld % 8 == 0 && ldo % 8 == 0 &&
         aligned16(p0) && aligned16(p1) && aligned16(p2);
}

int mha_fwd_tc(const __half* qkv, long long ld, int B, int Dm, int H, float scale, __half* out, long long ldo, float* lse,
               cudaStream_t st) {
  CUtensorMap tq, to;
  if (int rc = make_tmap_3d(&tq, qkv, 2, 3 * (uint64_t)Dm, AL, B, ld, (uint64_t)AL * ld, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&to, out, 2, Dm, AL, B, ldo, (uint64_t)AL * ldo, 64, 128)) return rc;
  static bool set = false;
  if (!set) {
    LPM_CUDA_CHECK(cudaFuncSetAttribute(mha_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
    set = true;
  }
  if (g_mha_tc_fwd == 2) {
    static bool set2 = false;
    if (!set2) {
      LPM_CUDA_CHECK(cudaFuncSetAttribute(mha_fwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2Smem::TOTAL));
      set2 = true;
    }
    mha_fwd_tc2_kernel<<<B * (H / 4), 160, Fwd2Smem::TOTAL, st>>>(tq, to, H, Dm, scale * 1.4426950408889634f, lse, AL / 32);
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  mha_fwd_tc_kernel<<<B * (H / 4), 288, FwdSmem::TOTAL, st>>>(tq, to, H, Dm, scale * 1.4426950408889634f, lse, AL / 32);   // nch: run-time trip count (keeps ptxas from flattening the softmax loops)
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int mha_bwd_tc(const __half* qkv, long long ld, const __half* o, const __half* dout, long long ldo, const float* lse, int B,
               int Dm, int H, float scale, __half* dqkv, long long ldd, cudaStream_t st) {
  CUtensorMap tq, tdo, td;
  if (int rc = make_tmap_3d(&tq, qkv, 2, 3 * (uint64_t)Dm, AL, B, ld, (uint64_t)AL * ld, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tdo, dout, 2, Dm, AL, B, ldo, (uint64_t)AL * ldo, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&td, dqkv, 2, 3 * (uint64_t)Dm, AL, B, ldd, (uint64_t)AL * ldd, 64, 128)) return rc;
  static bool set = false;
  if (!set) {
    LPM_CUDA_CHECK(cudaFuncSetAttribute(mha_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL));
    set = true;
  }
  if (g_mha_tc_mode == 2) {
    static const int pair_delay = getenv("LPM_MHA_PAIR_DELAY") ? atoi(getenv("LPM_MHA_PAIR_DELAY")) : 1400;
    static bool set2 = false;
    if (!set2) {
      LPM_CUDA_CHECK(cudaFuncSetAttribute(mha_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Bwd2Smem::TOTAL));
      set2 = true;
    }
    mha_bwd_tc2_kernel<<<B * (H / 4), 576, Bwd2Smem::TOTAL, st>>>(tq, tdo, td, o, ldo, lse, H, Dm, scale, g_mha_dbg, pair_delay);
    LPM_CUDA_CHECK(cudaGetLastError());
    return LPM_OK;
  }
  mha_bwd_tc_kernel<<<B * (H / 4), 608, BwdSmem::TOTAL, st>>>(tq, tdo, td, o, ldo, lse, H, Dm, scale, g_mha_dbg);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
