// Multi-head attention core for short sequences and tiny heads (length <= 512, depth 8 or 16):
//   O = softmax(scale * Q K^T [* key_scale + key_shift]) V        per (sample, head)
// transformer_utils.py:563-581 (V1, q scaled by depth^-0.5) and :641-664 (V2, batch-norm on the
// logits = per-key affine).  One CTA per (sample, head); Q/K/V of the head staged in shared memory;
// each warp owns 16-query blocks, FA2-style online softmax in registers, mma.sync m16n8k16 fp16
// with fp32 accumulation (depth-16 heads make one k-step per score tile: the op is exp/LSU bound,
// not tensor bound, so the legacy warp-level MMA is the right-sized tool here).
#include <cstdlib>

#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// smem tile [L][16] halves, 32 B per row, the two 16-byte chunks swapped on rows with bit 2 set
// (conflict-free ldmatrix).
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4); }

// Asynchronous (cp.async, 16 bytes per request) so that every chunk of every tile of a head is in flight at once: the
// register-staged version (LDG -> STS per chunk in a run-time loop) serialised the round trips and the prologue was 40 %
// of the forward kernel's warp time (ncu: long-scoreboard stalls).  Callers finish with head_tiles_wait() + barrier.
template <int DH>
__device__ __forceinline__ void load_head_tile(uint8_t* dst, const __half* src, long long ld, int L) {
  // src: first element of this head's columns in row 0; DH halves per row
  constexpr int CH = DH / 8;
  for (int i = threadIdx.x; i < L * 2; i += blockDim.x) {
    const int r = i >> 1, c = i & 1;
    const bool real = c < CH;                       // depth-8 heads: the second 16-byte chunk of a row is zero padding
    const __half* g = src + (long long)r * ld + (real ? c * 8 : 0);
    const int bytes = real ? 16 : 0;                // src-size 0 -> zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst + tile_off(r, c))), "l"(g), "r"(bytes) : "memory");
  }
}
__device__ __forceinline__ void head_tiles_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// mma with a zero C operand (no accumulator initialisation instructions)
__device__ __forceinline__ void mma16816_z(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));     // FMNMX3 (sm_100)
  return d;
}

// Forward.  The kernel is bound by the MUFU (one ex2 per score: 335 M at config 1 = 74 us of XU time on 148 SMs), so
// everything else is kept off the issue slots: logits come out of the MMA with a zero C operand, the softmax scale is
// folded into the exponent FMA, the row maximum uses 3-input FMNMX, the running maximum is only moved (and O / l
// rescaled) when it grows by more than 2^8 (fp16 probabilities hold values up to 256 without loss), and the row
// sums l come from the tensor core -- an extra n-tile of P V against a column of ones -- instead of 32 FADDs.
// FULL: L is a multiple of 64 (every 64-key step is complete: no tail predicates in the loop).
template <int DH, bool AFFINE, bool FULL>
__global__ void __launch_bounds__(128) mha_fwd_kernel(const __half* __restrict__ qkv, long long ld, int L, int Dm,
                                                      int H, float scale_log2, const float* __restrict__ key_scale,
                                                      const float* __restrict__ key_shift, __half* __restrict__ out,
                                                      long long ldo, float* __restrict__ lse) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint8_t* sQ = sm;
  uint8_t* sK = sQ + L * 32;
  uint8_t* sV = sK + L * 32;
  float* sKs = reinterpret_cast<float*>(sV + L * 32);  // per-key affine in log2 domain (optional)
  float* sKb = sKs + L;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const __half* base = qkv + (long long)b * L * ld + h * DH;
  load_head_tile<DH>(sQ, base, ld, L);
  load_head_tile<DH>(sK, base + Dm, ld, L);
  load_head_tile<DH>(sV, base + 2 * Dm, ld, L);
  if (AFFINE)
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
      sKs[i] = key_scale[i] * 1.4426950408889634f;
      sKb[i] = key_shift[i] * 1.4426950408889634f;
    }
  head_tiles_wait();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q_s = smem_u32(sQ), k_s = smem_u32(sK), v_s = smem_u32(sV);
  const int nkb = L / 16;  // 16-key blocks
  constexpr uint32_t ONES = 0x3C003C00u;   // half2(1, 1): B fragment of the row-sum tile
  constexpr float LAZY = 8.f;              // log2 of the growth of the maximum that forces a rescale

  for (int qb = warp; qb < L / 16; qb += 4) {
    uint32_t qa0, qa1, qa2, qa3;
    {
      const int row = qb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      ldsm_x4(q_s + tile_off(row, lane >> 4), qa0, qa1, qa2, qa3);
    }
    // m0 / m1: reference maxima of rows g and g+8, in the log2 domain for AFFINE, raw logits otherwise
    float m0 = -INFINITY, m1 = -INFINITY;
    float o[DH / 8][4], lacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < DH / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;

    for (int kb0 = 0; kb0 < nkb; kb0 += 4) {  // up to 64 keys per step
      const int nb = FULL ? 4 : min(4, nkb - kb0);
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (FULL || j < nb) {
          uint32_t b0, b1, b2, b3;
          const int key = (kb0 + j) * 16 + (lane & 7) + (lane >> 4) * 8;
          ldsm_x4(k_s + tile_off(key, (lane >> 3) & 1), b0, b1, b2, b3);
          mma16816_z(s[2 * j], qa0, qa1, qa2, qa3, b0, b1);
          mma16816_z(s[2 * j + 1], qa0, qa1, qa2, qa3, b2, b3);
        } else {
          s[2 * j][0] = s[2 * j][1] = s[2 * j][2] = s[2 * j][3] = -INFINITY;
          s[2 * j + 1][0] = s[2 * j + 1][1] = s[2 * j + 1][2] = s[2 * j + 1][3] = -INFINITY;
        }
      }
      if (AFFINE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (FULL || j < 2 * nb) {
            const int key = kb0 * 16 + j * 8 + (lane & 3) * 2;
            const float2 ks = *reinterpret_cast<const float2*>(&sKs[key]);
            const float2 kb = *reinterpret_cast<const float2*>(&sKb[key]);
            s[j][0] = fmaf(s[j][0], ks.x, kb.x); s[j][1] = fmaf(s[j][1], ks.y, kb.y);
            s[j][2] = fmaf(s[j][2], ks.x, kb.x); s[j][3] = fmaf(s[j][3], ks.y, kb.y);
          }
        }
      }
      // row maxima of this step (3-input max), then across the quad
      float mx0 = fmax3(s[0][0], s[0][1], s[1][0]), mx1 = fmax3(s[0][2], s[0][3], s[1][2]);
      mx0 = fmax3(mx0, s[1][1], s[2][0]); mx1 = fmax3(mx1, s[1][3], s[2][2]);
      mx0 = fmax3(mx0, s[2][1], s[3][0]); mx1 = fmax3(mx1, s[2][3], s[3][2]);
      mx0 = fmax3(mx0, s[3][1], s[4][0]); mx1 = fmax3(mx1, s[3][3], s[4][2]);
      mx0 = fmax3(mx0, s[4][1], s[5][0]); mx1 = fmax3(mx1, s[4][3], s[5][2]);
      mx0 = fmax3(mx0, s[5][1], s[6][0]); mx1 = fmax3(mx1, s[5][3], s[6][2]);
      mx0 = fmax3(mx0, s[6][1], s[7][0]); mx1 = fmax3(mx1, s[6][3], s[7][2]);
      mx0 = fmaxf(mx0, s[7][1]); mx1 = fmaxf(mx1, s[7][3]);
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      // lazy running maximum: move the reference only when it is exceeded by more than 2^LAZY (warp-uniform branch)
      const float sc = AFFINE ? 1.f : scale_log2;     // raw logits are compared in units of 1/sc
      const bool grow = (mx0 - m0) * sc > LAZY || (mx1 - m1) * sc > LAZY || m0 == -INFINITY;
      if (__any_sync(0xffffffffu, grow)) {
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0 = fast_exp2((m0 - mn0) * sc), c1 = fast_exp2((m1 - mn1) * sc);    // first step: 2^-inf = 0 on zeros
        m0 = mn0; m1 = mn1;
        lacc[0] *= c0; lacc[1] *= c0; lacc[2] *= c1; lacc[3] *= c1;
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }
      }
      const float nm0 = -m0 * sc, nm1 = -m1 * sc;
      // P = 2^(s*sc - m*sc) straight into fp16 A fragments; O += P V; l += P 1
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (FULL || j < nb) {
          const uint32_t a0 = pack_half2(fast_exp2(fmaf(s[2 * j][0], sc, nm0)), fast_exp2(fmaf(s[2 * j][1], sc, nm0)));
          const uint32_t a1 = pack_half2(fast_exp2(fmaf(s[2 * j][2], sc, nm1)), fast_exp2(fmaf(s[2 * j][3], sc, nm1)));
          const uint32_t a2 = pack_half2(fast_exp2(fmaf(s[2 * j + 1][0], sc, nm0)), fast_exp2(fmaf(s[2 * j + 1][1], sc, nm0)));
          const uint32_t a3 = pack_half2(fast_exp2(fmaf(s[2 * j + 1][2], sc, nm1)), fast_exp2(fmaf(s[2 * j + 1][3], sc, nm1)));
          uint32_t b0, b1, b2, b3;
          const int key = (kb0 + j) * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          ldsm_x4_t(v_s + tile_off(key, lane >> 4), b0, b1, b2, b3);
          mma16816(o[0], a0, a1, a2, a3, b0, b1);
          if (DH == 16) mma16816(o[DH / 8 - 1], a0, a1, a2, a3, b2, b3);
          mma16816(lacc, a0, a1, a2, a3, ONES, ONES);
        }
      }
    }
    const float l0 = lacc[0], l1 = lacc[2];       // every column of the ones tile holds the row sum
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = qb * 16 + (lane >> 2), r1 = r0 + 8;
    __half* o0 = out + ((long long)b * L + r0) * ldo + h * DH + (lane & 3) * 2;
    __half* o1 = out + ((long long)b * L + r1) * ldo + h * DH + (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < DH / 8; ++j) {
      *reinterpret_cast<__half2*>(o0 + j * 8) = __floats2half2_rn(o[j][0] * i0, o[j][1] * i0);
      *reinterpret_cast<__half2*>(o1 + j * 8) = __floats2half2_rn(o[j][2] * i1, o[j][3] * i1);
    }
    if (lse != nullptr && (lane & 3) == 0) {
      // natural-log log-sum-exp of the (scaled / affine) logits
      const float sc = AFFINE ? 1.f : scale_log2;
      lse[((long long)b * H + h) * L + r0] = (m0 * sc + log2f(l0)) * 0.6931471805599453f;
      lse[((long long)b * H + h) * L + r1] = (m1 * sc + log2f(l1)) * 0.6931471805599453f;
    }
  }
}

int mha_fwd(const __half* qkv, long long ld, int B, int L, int Dm, int H, float scale, const float* key_scale,
            const float* key_shift, __half* out, long long ldo, float* lse, cudaStream_t st) {
  const int DH = Dm / H;
  LPM_REQUIRE(DH * H == Dm && (DH == 8 || DH == 16), "mha_fwd: head depth must be 8 or 16 (Dm=%d H=%d)", Dm, H);
  LPM_REQUIRE(L % 16 == 0 && L >= 16 && L <= 1024, "mha_fwd: length must be a multiple of 16 in [16,1024] (got %d)", L);
  LPM_REQUIRE(ld % 8 == 0 && ldo % 2 == 0, "mha_fwd: leading dimensions must be multiples of 8");
  LPM_REQUIRE(key_scale != nullptr || scale > 0.f, "mha_fwd: the softmax scale must be positive");
  LPM_REQUIRE((key_scale == nullptr) == (key_shift == nullptr), "mha_fwd: key_scale and key_shift come together");
  // depth-16 heads at 256 positions (the cluster attention of config 1): tcgen05 / TMEM kernel, four heads per CTA
  // (forward: opt-in with LPM_MHA_TC_FWD=1 or lpm_debug_set_mha_tc_mode(4 | m) -- measured 170 us against 121 us for the warp-level kernel below, whose
  // running-maximum softmax needs one pass; the backward is the default: 285 us against 366 us)
  if (mha_tc_forward_enabled() && key_scale == nullptr && mha_tc_eligible(L, Dm, H, ld, ldo, qkv, out, out))
    return mha_fwd_tc(qkv, ld, B, Dm, H, scale, out, ldo, lse, st);
  const size_t smem = (size_t)L * 96 + (size_t)L * 8;
  const float scale_log2 = scale * 1.4426950408889634f;
#define LPM_MHA_FWD2(DHV, AFF, FUL)                                                                                      \
  {                                                                                                                      \
    static bool set = false;                                                                                             \
    if (!set) { LPM_CUDA_CHECK(cudaFuncSetAttribute(mha_fwd_kernel<DHV, AFF, FUL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)); set = true; } \
    mha_fwd_kernel<DHV, AFF, FUL><<<B * H, 128, smem, st>>>(qkv, ld, L, Dm, H, scale_log2, key_scale, key_shift, out, ldo, lse); \
  }
#define LPM_MHA_FWD(DHV, AFF) { if (L % 64 == 0) LPM_MHA_FWD2(DHV, AFF, true) else LPM_MHA_FWD2(DHV, AFF, false) }
  if (DH == 16) { if (key_scale) LPM_MHA_FWD(16, true) else LPM_MHA_FWD(16, false) }
  else { if (key_scale) LPM_MHA_FWD(8, true) else LPM_MHA_FWD(8, false) }
#undef LPM_MHA_FWD2
#undef LPM_MHA_FWD
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of the attention core (V1 scaling mode).  One CTA per (sample, head); L <= 512.  Lengths above 256
// are processed in query chunks of QC = 128 so that the parked dS^T ([L keys][QC queries]) still fits in
// shared memory; dK / dV of later chunks are accumulated onto the fp16 results of the earlier ones.
//   pass A (warp owns a 16-key block j, loops over query blocks i):
//       S^T = K_j Q_i^T ; P^T = exp2(S^T*scale - lse_i) ; dP^T = V_j dO_i^T ; dS^T = P^T o (dP^T - delta_i)
//       dV_j += P^T dO_i ; dK_j += dS^T Q_i ; dS^T parked in shared memory (fp16, [key][query])
//   pass B (warp owns a 16-query block i): dQ_i = sum_j dS_ij K_j with dS read back transposed (ldmatrix.trans)
// Gradients arrive / leave scaled by the caller's loss scale (the kernel is linear in dO).
// ------------------------------------------------------------------------------------------------
// halves per dS row = QC + 8 padding (stride = 4 banks mod 32 -> conflict-free stores and ldmatrix)

// MODE 0: logits = scale * q.k (V1).  MODE 1/2: batch-normed logits l' = l*ks_j + kb_j (V2): MODE 1 only accumulates the
// per-key sums of dl' and dl'*lhat (lhat = (l - mu_j)*rstd_j) needed by the batch-norm backward; MODE 2 applies
//   dl = ks_j * (dl' - m1_j - lhat * m2_j)   and produces dQ, dK, dV.
struct MhaBnArgs {
  const float* key_scale; const float* key_shift;   // folded BN affine per key (natural-log domain)
  const float* key_mean; const float* key_rstd;     // batch statistics of the raw logits per key
  const float* m1; const float* m2;                 // mean(dl'), mean(dl'*lhat) per key (MODE 2)
  float* stat_partial;                              // [B*H][2][L] (MODE 1)
};

// NT = threads per CTA: 512 (one CTA per SM, the whole dS^T of a 256-query chunk parked) or 256 with query chunks of 128:
// two CTAs per SM, so that the load prologue and the pass barriers of one head hide under the other head's math.
// Registers: 96 per thread for the 512-thread layout (92 used, no spills) = 48 K of the SM's 64 K registers, which leaves
// room for one 128-thread CTA of the tiled factored Adam (rank_adam_tile_kernel, 16 K registers, 5 KB of shared memory)
// next to it: this kernel is issue-bound and moves < 1 TB/s, so the HBM-bound hidden1_weights update hides under it
// (trainer._fork_hidden_update).  __maxnreg__ and __launch_bounds__ are mutually exclusive.
template <int DH, int MODE, int NT>
__global__ void __maxnreg__(NT == 512 ? 96 : 128) mha_bwd_kernel(const __half* __restrict__ qkv, long long ld,
                                                         const __half* __restrict__ o, const __half* __restrict__ dout,
                                                         long long ldo, const float* __restrict__ lse, int L, int Dm,
                                                         int H, float scale, __half* __restrict__ dqkv, long long ldd, const MhaBnArgs bn,
                                                         int QC) {
  const int DS_STRIDE = QC + 8;
  // 16 warps; one CTA per (sample, head).  The kernel is latency- rather than throughput-bound (one CTA per SM
  // because dS^T is parked in shared memory), so every warp keeps two independent query blocks in flight.
  extern __shared__ __align__(16) uint8_t sm[];
  uint8_t* sQ = sm;
  uint8_t* sK = sQ + L * 32;
  uint8_t* sV = sK + L * 32;
  uint8_t* sdO = sV + L * 32;
  float* sLse = reinterpret_cast<float*>(sdO + L * 32);   // lse in log2 units
  float* sDel = sLse + L;
  float* sBn = sDel + L;                                   // MODE>0: ks2 | kb2 | mu | rstd | m1 | m2  (6 x L)
  __half* sdS = reinterpret_cast<__half*>(sBn + (MODE > 0 ? 6 * L : 0));   // [L keys][DS_STRIDE]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const __half* base = qkv + (long long)b * L * ld + h * DH;
  load_head_tile<DH>(sQ, base, ld, L);
  load_head_tile<DH>(sK, base + Dm, ld, L);
  load_head_tile<DH>(sV, base + 2 * Dm, ld, L);
  load_head_tile<DH>(sdO, dout + (long long)b * L * ldo + h * DH, ldo, L);
  for (int r = threadIdx.x; r < L; r += blockDim.x) {
    const __half* orow = o + ((long long)b * L + r) * ldo + h * DH;
    const __half* drow = dout + ((long long)b * L + r) * ldo + h * DH;
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < DH; c += 8) {
      const uint4 vo = __ldg(reinterpret_cast<const uint4*>(orow + c));
      const uint4 vd = __ldg(reinterpret_cast<const uint4*>(drow + c));
      const __half2* ho = reinterpret_cast<const __half2*>(&vo);
      const __half2* hd = reinterpret_cast<const __half2*>(&vd);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 a = __half22float2(ho[j]), g = __half22float2(hd[j]); d += a.x * g.x + a.y * g.y; }
    }
    sDel[r] = d;
    sLse[r] = lse[((long long)b * H + h) * L + r] * 1.4426950408889634f;
    if (MODE > 0) {
      sBn[r] = bn.key_scale[r] * 1.4426950408889634f;
      sBn[L + r] = bn.key_shift[r] * 1.4426950408889634f;
      sBn[2 * L + r] = bn.key_mean[r];
      sBn[3 * L + r] = bn.key_rstd[r];
      sBn[4 * L + r] = MODE == 2 ? bn.m1[r] : 0.f;
      sBn[5 * L + r] = MODE == 2 ? bn.m2[r] : 0.f;
    }
  }
  head_tiles_wait();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const uint32_t q_s = smem_u32(sQ), k_s = smem_u32(sK), v_s = smem_u32(sV), do_s = smem_u32(sdO);
  const float scale_log2 = scale * 1.4426950408889634f;
  const int nb = L / 16;

  for (int q0 = 0; q0 < L; q0 += QC) {
  const int ib0 = q0 / 16, ib1 = min(nb, (q0 + QC) / 16);
  // ------------------------------- pass A: dK, dV, dS -------------------------------
  for (int j = warp; j < nb; j += nwarps) {
    uint32_t ka0, ka1, ka2, ka3, va0, va1, va2, va3;
    {
      const int row = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      ldsm_x4(k_s + tile_off(row, lane >> 4), ka0, ka1, ka2, ka3);
      ldsm_x4(v_s + tile_off(row, lane >> 4), va0, va1, va2, va3);
    }
    float dk[2][4], dv[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n) dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
    // per-key (row) constants of the batch-normed logits and the statistics accumulators (MODE 1)
    const int kr0 = j * 16 + (lane >> 2), kr1 = kr0 + 8;
    float ks0 = 0.f, ks1 = 0.f, kb0 = 0.f, kb1 = 0.f, mu0 = 0.f, mu1 = 0.f, rs0 = 0.f, rs1 = 0.f;
    float m10 = 0.f, m11 = 0.f, m20 = 0.f, m21 = 0.f, a1r0 = 0.f, a1r1 = 0.f, a2r0 = 0.f, a2r1 = 0.f;
    if (MODE > 0) {
      ks0 = sBn[kr0]; ks1 = sBn[kr1]; kb0 = sBn[L + kr0]; kb1 = sBn[L + kr1];
      mu0 = sBn[2 * L + kr0]; mu1 = sBn[2 * L + kr1]; rs0 = sBn[3 * L + kr0]; rs1 = sBn[3 * L + kr1];
      m10 = sBn[4 * L + kr0]; m11 = sBn[4 * L + kr1]; m20 = sBn[5 * L + kr0]; m21 = sBn[5 * L + kr1];
    }
    for (int i = ib0; i < ib1; i += 2) {
      // two query blocks per iteration (the second is a masked repeat of the first on an odd tail)
      const bool two = i + 1 < ib1;
      const int ib[2] = {i, two ? i + 1 : i};
      float st[2][2][4], dp[2][2][4];
      uint32_t qf[2][4], of[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int qrow = ib[u] * 16 + (lane & 7) + (lane >> 4) * 8;
        ldsm_x4(q_s + tile_off(qrow, (lane >> 3) & 1), qf[u][0], qf[u][1], qf[u][2], qf[u][3]);
        ldsm_x4(do_s + tile_off(qrow, (lane >> 3) & 1), of[u][0], of[u][1], of[u][2], of[u][3]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        mma16816_z(st[u][0], ka0, ka1, ka2, ka3, qf[u][0], qf[u][1]);
        mma16816_z(st[u][1], ka0, ka1, ka2, ka3, qf[u][2], qf[u][3]);
        mma16816_z(dp[u][0], va0, va1, va2, va3, of[u][0], of[u][1]);
        mma16816_z(dp[u][1], va0, va1, va2, va3, of[u][2], of[u][3]);
      }
      // columns of the transposed tiles are queries: n-tile n -> queries i*16 + n*8 + (lane&3)*2 + {0,1}
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const bool dead = !(u == 0 || two);     // masked repeat on an odd tail (never taken when the chunk is a multiple of 32)
        const float live = dead ? 0.f : 1.f;
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int qc = ib[u] * 16 + n * 8 + (lane & 3) * 2;
          const float2 l01 = *reinterpret_cast<const float2*>(&sLse[qc]);
          const float2 d01 = *reinterpret_cast<const float2*>(&sDel[qc]);
          const float l0 = l01.x, l1 = l01.y, d0 = d01.x, d1 = d01.y;
          float p0, p1, p2, p3;
          if (MODE == 0) {
            p0 = fast_exp2(st[u][n][0] * scale_log2 - l0); p1 = fast_exp2(st[u][n][1] * scale_log2 - l1);
            p2 = fast_exp2(st[u][n][2] * scale_log2 - l0); p3 = fast_exp2(st[u][n][3] * scale_log2 - l1);
          } else {
            p0 = fast_exp2(st[u][n][0] * ks0 + kb0 - l0); p1 = fast_exp2(st[u][n][1] * ks0 + kb0 - l1);
            p2 = fast_exp2(st[u][n][2] * ks1 + kb1 - l0); p3 = fast_exp2(st[u][n][3] * ks1 + kb1 - l1);
          }
          if (dead) { p0 = p1 = p2 = p3 = 0.f; }
          float g0 = p0 * (dp[u][n][0] - d0), g1 = p1 * (dp[u][n][1] - d1), g2 = p2 * (dp[u][n][2] - d0), g3 = p3 * (dp[u][n][3] - d1);
          if (MODE > 0) {
            const float h0 = (st[u][n][0] - mu0) * rs0, h1 = (st[u][n][1] - mu0) * rs0;
            const float h2 = (st[u][n][2] - mu1) * rs1, h3 = (st[u][n][3] - mu1) * rs1;
            if (MODE == 1) {
              a1r0 += g0 + g1; a2r0 += g0 * h0 + g1 * h1;
              a1r1 += g2 + g3; a2r1 += g2 * h2 + g3 * h3;
            } else {
              const float c0 = ks0 * 0.6931471805599453f * live, c1 = ks1 * 0.6931471805599453f * live;   // natural-log scale
              g0 = c0 * (g0 - m10 - h0 * m20); g1 = c0 * (g1 - m10 - h1 * m20);
              g2 = c1 * (g2 - m11 - h2 * m21); g3 = c1 * (g3 - m11 - h3 * m21);
            }
          }
          st[u][n][0] = p0; st[u][n][1] = p1; st[u][n][2] = p2; st[u][n][3] = p3;
          dp[u][n][0] = g0; dp[u][n][1] = g1; dp[u][n][2] = g2; dp[u][n][3] = g3;
        }
      }
      if (MODE == 1) continue;                 // statistics pass: no gradients yet
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        // A fragments (rows = keys, k = queries) from the accumulator layout
        const uint32_t pa0 = pack_half2(st[u][0][0], st[u][0][1]), pa1 = pack_half2(st[u][0][2], st[u][0][3]);
        const uint32_t pa2 = pack_half2(st[u][1][0], st[u][1][1]), pa3 = pack_half2(st[u][1][2], st[u][1][3]);
        const uint32_t sa0 = pack_half2(dp[u][0][0], dp[u][0][1]), sa1 = pack_half2(dp[u][0][2], dp[u][0][3]);
        const uint32_t sa2 = pack_half2(dp[u][1][0], dp[u][1][1]), sa3 = pack_half2(dp[u][1][2], dp[u][1][3]);
        uint32_t b0, b1, b2, b3, c0, c1, c2, c3;
        const int qrow = ib[u] * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4_t(do_s + tile_off(qrow, lane >> 4), b0, b1, b2, b3);
        ldsm_x4_t(q_s + tile_off(qrow, lane >> 4), c0, c1, c2, c3);
        mma16816(dv[0], pa0, pa1, pa2, pa3, b0, b1);
        if (DH == 16) mma16816(dv[1], pa0, pa1, pa2, pa3, b2, b3);
        mma16816(dk[0], sa0, sa1, sa2, sa3, c0, c1);
        if (DH == 16) mma16816(dk[1], sa0, sa1, sa2, sa3, c2, c3);
        // park dS^T: element (key = j*16 + lane/4 (+8), query = i*16 + n*8 + (lane&3)*2)
        if (u == 0 || two) {
          __half* r0 = sdS + (size_t)(j * 16 + (lane >> 2)) * DS_STRIDE + (ib[u] - ib0) * 16 + (lane & 3) * 2;
          __half* r1 = r0 + 8 * DS_STRIDE;
          *reinterpret_cast<uint32_t*>(r0) = sa0;
          *reinterpret_cast<uint32_t*>(r1) = sa1;
          *reinterpret_cast<uint32_t*>(r0 + 8) = sa2;
          *reinterpret_cast<uint32_t*>(r1 + 8) = sa3;
        }
      }
    }
    if (MODE == 1) {
      a1r0 += __shfl_xor_sync(0xffffffffu, a1r0, 1); a1r0 += __shfl_xor_sync(0xffffffffu, a1r0, 2);
      a2r0 += __shfl_xor_sync(0xffffffffu, a2r0, 1); a2r0 += __shfl_xor_sync(0xffffffffu, a2r0, 2);
      a1r1 += __shfl_xor_sync(0xffffffffu, a1r1, 1); a1r1 += __shfl_xor_sync(0xffffffffu, a1r1, 2);
      a2r1 += __shfl_xor_sync(0xffffffffu, a2r1, 1); a2r1 += __shfl_xor_sync(0xffffffffu, a2r1, 2);
      if ((lane & 3) == 0) {
        float* sp = bn.stat_partial + (size_t)blockIdx.x * 2 * L;
        if (q0 > 0) { a1r0 += sp[kr0]; a1r1 += sp[kr1]; a2r0 += sp[L + kr0]; a2r1 += sp[L + kr1]; }
        sp[kr0] = a1r0; sp[kr1] = a1r1; sp[L + kr0] = a2r0; sp[L + kr1] = a2r1;
      }
      continue;
    }
    const int r0 = j * 16 + (lane >> 2), r1 = r0 + 8;
    __half* dk0 = dqkv + ((long long)b * L + r0) * ldd + Dm + h * DH + (lane & 3) * 2;
    __half* dk1 = dqkv + ((long long)b * L + r1) * ldd + Dm + h * DH + (lane & 3) * 2;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      float2 k0v = make_float2(dk[n][0] * scale, dk[n][1] * scale), k1v = make_float2(dk[n][2] * scale, dk[n][3] * scale);
      float2 v0v = make_float2(dv[n][0], dv[n][1]), v1v = make_float2(dv[n][2], dv[n][3]);
      if (q0 > 0) {   // later query chunk: this warp owns these rows, add onto its own earlier result
        const float2 a = __half22float2(*reinterpret_cast<__half2*>(dk0 + n * 8)), c = __half22float2(*reinterpret_cast<__half2*>(dk1 + n * 8));
        const float2 e = __half22float2(*reinterpret_cast<__half2*>(dk0 + Dm + n * 8)), f = __half22float2(*reinterpret_cast<__half2*>(dk1 + Dm + n * 8));
        k0v.x += a.x; k0v.y += a.y; k1v.x += c.x; k1v.y += c.y; v0v.x += e.x; v0v.y += e.y; v1v.x += f.x; v1v.y += f.y;
      }
      *reinterpret_cast<__half2*>(dk0 + n * 8) = __floats2half2_rn(k0v.x, k0v.y);
      *reinterpret_cast<__half2*>(dk1 + n * 8) = __floats2half2_rn(k1v.x, k1v.y);
      *reinterpret_cast<__half2*>(dk0 + Dm + n * 8) = __floats2half2_rn(v0v.x, v0v.y);
      *reinterpret_cast<__half2*>(dk1 + Dm + n * 8) = __floats2half2_rn(v1v.x, v1v.y);
    }
  }
  if (MODE == 1) continue;
  __syncthreads();

  // ------------------------------- pass B: dQ -------------------------------
  // two independent accumulator sets (even / odd key blocks) break the MMA dependency chain
  const uint32_t ds_s = smem_u32(sdS);
  for (int i = ib0 + warp; i < ib1; i += nwarps) {
    float dq[2][2][4];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int n = 0; n < 2; ++n) dq[e][n][0] = dq[e][n][1] = dq[e][n][2] = dq[e][n][3] = 0.f;
    for (int j = 0; j < nb; j += 2) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (j + e < nb) {
          uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
          const int key = (j + e) * 16 + (lane & 7) + (lane >> 4) * 8;
          const int qcol = (i - ib0) * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x4_t(ds_s + (uint32_t)(key * DS_STRIDE + qcol) * 2, a0, a1, a2, a3);
          const int krow = (j + e) * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          ldsm_x4_t(k_s + tile_off(krow, lane >> 4), b0, b1, b2, b3);
          mma16816(dq[e][0], a0, a1, a2, a3, b0, b1);
          if (DH == 16) mma16816(dq[e][1], a0, a1, a2, a3, b2, b3);
        }
      }
    }
    const int r0 = i * 16 + (lane >> 2), r1 = r0 + 8;
    __half* q0 = dqkv + ((long long)b * L + r0) * ldd + h * DH + (lane & 3) * 2;
    __half* q1 = dqkv + ((long long)b * L + r1) * ldd + h * DH + (lane & 3) * 2;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      *reinterpret_cast<__half2*>(q0 + n * 8) = __floats2half2_rn((dq[0][n][0] + dq[1][n][0]) * scale, (dq[0][n][1] + dq[1][n][1]) * scale);
      *reinterpret_cast<__half2*>(q1 + n * 8) = __floats2half2_rn((dq[0][n][2] + dq[1][n][2]) * scale, (dq[0][n][3] + dq[1][n][3]) * scale);
    }
  }
  __syncthreads();   // the parked dS^T is rewritten by the next query chunk
  }
}

static int mha_bwd_launch(int mode, const __half* qkv, long long ld, const __half* o, const __half* dout, long long ldo,
                          const float* lse, int B, int L, int Dm, int H, float scale, __half* dqkv, long long ldd,
                          const MhaBnArgs& bn, cudaStream_t st) {
  const int DH = Dm / H;
  LPM_REQUIRE(DH * H == Dm && (DH == 8 || DH == 16), "mha_bwd: head depth must be 8 or 16 (Dm=%d H=%d)", Dm, H);
  LPM_REQUIRE(L % 16 == 0 && L >= 16 && L <= 512, "mha_bwd: length must be a multiple of 16 in [16,512] (got %d)", L);
  LPM_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && ldd % 8 == 0, "mha_bwd: leading dimensions must be multiples of 8");
  LPM_REQUIRE(mode == 0 || DH == 16, "mha_bwd: batch-normed logits need head depth 16");
  // query chunk whose dS^T is parked at a time.  L = 256 (the V1 cluster attention at config 1): the whole dS^T (135 KB,
  // one 512-thread CTA per SM).  Measured alternative (LPM_MHA_BWD_WIDE=0): chunks of 128 in 256-thread CTAs, two CTAs
  // per SM -- 576 us instead of 403 us at [80, 64, 256, 16]: the second pass over the keys and the fp16 read-modify-write
  // of dK / dV cost more than the overlapped prologue saves.
  static const bool wide = !(getenv("LPM_MHA_BWD_WIDE") != nullptr && getenv("LPM_MHA_BWD_WIDE")[0] == '0');
  const int QC = L < 256 ? L : (L == 256 && wide ? 256 : 128);
  const int nt = (L > 256 || (L == 256 && wide)) ? 512 : 256;
  const size_t smem = (size_t)L * 128 + (size_t)L * 8 + (mode ? (size_t)L * 24 : 0) + (size_t)L * (QC + 8) * 2;
  LPM_REQUIRE(smem <= 227 * 1024, "mha_bwd: shared memory budget exceeded (L=%d)", L);
#define LPM_MHA_BWD2(DHV, MODEV, NTV)                                                                                    \
  {                                                                                                                      \
    static bool set = false;                                                                                             \
    if (!set) { LPM_CUDA_CHECK(cudaFuncSetAttribute(mha_bwd_kernel<DHV, MODEV, NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); set = true; } \
    mha_bwd_kernel<DHV, MODEV, NTV><<<B * H, NTV, smem, st>>>(qkv, ld, o, dout, ldo, lse, L, Dm, H, scale, dqkv, ldd, bn, QC);    \
  }
#define LPM_MHA_BWD(DHV, MODEV) { if (nt == 512) LPM_MHA_BWD2(DHV, MODEV, 512) else LPM_MHA_BWD2(DHV, MODEV, 256) }
  if (mode == 0) { if (DH == 16) LPM_MHA_BWD(16, 0) else LPM_MHA_BWD(8, 0) }
  else if (mode == 1) LPM_MHA_BWD(16, 1)
  else LPM_MHA_BWD(16, 2)
#undef LPM_MHA_BWD
#undef LPM_MHA_BWD2
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int mha_bwd(const __half* qkv, long long ld, const __half* o, const __half* dout, long long ldo, const float* lse,
            int B, int L, int Dm, int H, float scale, __half* dqkv, long long ldd, cudaStream_t st) {
  if (mha_tc_backward_mode() != 0 && ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0 && mha_tc_eligible(L, Dm, H, ld, ldo, qkv, dout, dqkv)) {
    LPM_REQUIRE(scale > 0.f, "mha_bwd: the softmax scale must be positive");
    return mha_bwd_tc(qkv, ld, o, dout, ldo, lse, B, Dm, H, scale, dqkv, ldd, st);
  }
  MhaBnArgs bn{};
  return mha_bwd_launch(0, qkv, ld, o, dout, ldo, lse, B, L, Dm, H, scale, dqkv, ldd, bn, st);
}

// Batch-normed-logits attention backward (NetVladV2): mode 1 = per-key statistics partials, mode 2 = gradients.
int mha_bwd_bn(int mode, const __half* qkv, long long ld, const __half* o, const __half* dout, long long ldo,
               const float* lse, int B, int L, int Dm, int H, const float* key_scale, const float* key_shift,
               const float* key_mean, const float* key_rstd, const float* m1, const float* m2, float* stat_partial,
               __half* dqkv, long long ldd, cudaStream_t st) {
  LPM_REQUIRE(mode == 1 || mode == 2, "mha_bwd_bn: mode must be 1 or 2");
  MhaBnArgs bn{key_scale, key_shift, key_mean, key_rstd, m1, m2, stat_partial};
  return mha_bwd_launch(mode, qkv, ld, o, dout, ldo, lse, B, L, Dm, H, 1.0f, dqkv, ldd, bn, st);
}

}  // namespace lpm
