// On-GPU evaluation metrics of the training / eval loops: per-video top-k, hit@1, precision at
// equal recall rate and the global average precision of a batch -- eval_util.py:27-135 and
// average_precision_calculator.py:203-262 (called every logged step from train.py:448-449 on a host copy of the
// [B, 3862] predictions; here only three floats leave the device).
//
//   eval_rows_kernel : one CTA per video.  The row lives in registers (V <= 8192); max(k, num_labels) rounds of block
//                      arg-max (ties -> lowest class index) emit the top-k (value, class, label) triplets
//                      (eval_util.top_k_triplets), hit@1 (eval_util.py:27-42) and PERR (eval_util.py:45-70).
//   eval_gap_kernel  : one CTA: the B*k triplets are sorted by score (bitonic, shared memory) and the
//                      non-interpolated average precision is accumulated against the number of positives of the whole
//                      batch (eval_util.calculate_gap -> ap_at_n with total_num_positives).
#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

constexpr int EV_THREADS = 256;
constexpr int EV_PER = 32;          // values per thread: V <= 8192

__device__ __forceinline__ unsigned long long ev_pack(float v, int idx) {
  // order-preserving key: larger value first, then LOWER index first when compared with '>'
  unsigned int b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned int)(0x7fffffff - idx);
}

__global__ void __launch_bounds__(EV_THREADS) eval_rows_kernel(const float* __restrict__ pred, long long ld,
                                                               const unsigned char* __restrict__ labels, long long ldl,
                                                               int V, int k, float* __restrict__ top_val,
                                                               int* __restrict__ top_idx, unsigned char* __restrict__ top_lab,
                                                               float* __restrict__ row_stats) {
  __shared__ unsigned long long red[EV_THREADS / 32];
  __shared__ int s_cnt[EV_THREADS / 32];
  __shared__ unsigned long long s_best;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = pred + (long long)b * ld;
  const unsigned char* lab = labels + (long long)b * ldl;
  unsigned long long key[EV_PER];
  int n_lab = 0;
#pragma unroll
  for (int j = 0; j < EV_PER; ++j) {
    const int c = tid + j * EV_THREADS;
    key[j] = 0ull;
    if (c < V) {
      key[j] = ev_pack(__ldg(row + c), c);
      n_lab += lab[c] != 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_lab += __shfl_xor_sync(0xffffffffu, n_lab, o);
  if (lane == 0) s_cnt[warp] = n_lab;
  __syncthreads();
  n_lab = 0;
#pragma unroll
  for (int w = 0; w < EV_THREADS / 32; ++w) n_lab += s_cnt[w];
  const int kk = min(k, V);
  const int rounds = min(V, max(kk, n_lab));
  float hit = 0.f, perr_hits = 0.f;
  for (int r = 0; r < rounds; ++r) {
    unsigned long long best = 0ull;
#pragma unroll
    for (int j = 0; j < EV_PER; ++j) best = key[j] > best ? key[j] : best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (tid == 0) {
      unsigned long long m = red[0];
#pragma unroll
      for (int w = 1; w < EV_THREADS / 32; ++w) m = red[w] > m ? red[w] : m;
      s_best = m;
    }
    __syncthreads();
    const unsigned long long m = s_best;
    const int idx = 0x7fffffff - (int)(unsigned int)(m & 0xffffffffull);
#pragma unroll
    for (int j = 0; j < EV_PER; ++j)
      if (key[j] == m) key[j] = 0ull;                       // taken (keys are unique: the class index is part of them)
    if (tid == 0) {
      const float v = row[idx];
      const unsigned char l = lab[idx];
      if (r == 0) hit = l != 0 ? 1.f : 0.f;                 // eval_util.py:39-42
      if (r < n_lab && v > 0.f) perr_hits += l != 0;        // eval_util.py:62-67
      if (r < kk) {
        top_val[(long long)b * k + r] = v;
        top_idx[(long long)b * k + r] = idx;
        top_lab[(long long)b * k + r] = l;
      }
    }
  }
  if (tid == 0) {
    for (int r = kk; r < k; ++r) { top_val[(long long)b * k + r] = -INFINITY; top_idx[(long long)b * k + r] = -1; top_lab[(long long)b * k + r] = 0; }
    row_stats[b * 3 + 0] = hit;
    // num_labels == 0: numpy's [-0:] slice takes the whole row, whose precision is 0 (eval_util.py:60-68)
    row_stats[b * 3 + 1] = n_lab > 0 ? perr_hits / (float)n_lab : 0.f;
    row_stats[b * 3 + 2] = (float)n_lab;
  }
}

// metrics[0] = hit@1, [1] = PERR, [2] = GAP over the n = B*k triplets
__global__ void __launch_bounds__(1024) eval_gap_kernel(const float* __restrict__ top_val, const unsigned char* __restrict__ top_lab,
                                                        int n, int npow2, const float* __restrict__ row_stats, int B,
                                                        float* __restrict__ metrics) {
  extern __shared__ unsigned long long sk[];              // [npow2] (score key << 32 | label bit | slot)
  __shared__ double sred[32];
  __shared__ int sbase[33];
  const int tid = threadIdx.x;
  for (int i = tid; i < npow2; i += 1024) {
    unsigned long long kq = 0ull;
    if (i < n && top_val[i] > -INFINITY) {
      unsigned int bts = __float_as_uint(top_val[i]);
      bts = (bts & 0x80000000u) ? ~bts : (bts | 0x80000000u);
      kq = ((unsigned long long)bts << 32) | ((unsigned long long)(0x7fffffffu - (unsigned)i) << 1) | (top_lab[i] != 0 ? 1ull : 0ull);
    }
    sk[i] = kq;
  }
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < npow2 / 2; i += 1024) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = sk[lo], c = sk[hi];
        if ((a < c) == desc) { sk[lo] = c; sk[hi] = a; }
      }
      __syncthreads();
    }
  }
  // batch statistics
  double hit = 0.0, perr = 0.0, npos = 0.0;
  for (int i = tid; i < B; i += 1024) { hit += row_stats[i * 3]; perr += row_stats[i * 3 + 1]; npos += row_stats[i * 3 + 2]; }
  const int lane = tid & 31, warp = tid >> 5;
  double tot[3] = {hit, perr, npos};
  for (int q = 0; q < 3; ++q) {
    double v = tot[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    v = 0.0;
    for (int w = 0; w < 32; ++w) v += sred[w];
    tot[q] = v;
    __syncthreads();
  }
  // average precision: sum over positives of poscount / rank (average_precision_calculator.py:244-262)
  const int per = (npow2 + 1023) / 1024;
  const int i0 = tid * per, i1 = min(n, i0 + per);
  int cnt = 0;
  for (int i = i0; i < i1; ++i) cnt += (int)(sk[i] & 1ull);
  // exclusive scan of the per-thread positive counts
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) sbase[warp + 1] = incl;
  if (tid == 0) sbase[0] = 0;
  __syncthreads();
  if (tid == 0) for (int w = 1; w <= 32; ++w) sbase[w] += sbase[w - 1];
  __syncthreads();
  int pos = sbase[warp] + incl - cnt;
  double ap = 0.0;
  for (int i = i0; i < i1; ++i)
    if (sk[i] & 1ull) { pos += 1; ap += (double)pos / (double)(i + 1); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ap += __shfl_xor_sync(0xffffffffu, ap, o);
  if (lane == 0) sred[warp] = ap;
  __syncthreads();
  if (tid == 0) {
    double a = 0.0;
    for (int w = 0; w < 32; ++w) a += sred[w];
    metrics[0] = B > 0 ? (float)(tot[0] / B) : 0.f;
    metrics[1] = B > 0 ? (float)(tot[1] / B) : 0.f;
    metrics[2] = tot[2] > 0.0 ? (float)(a / tot[2]) : 0.f;
  }
}

int eval_topk(const float* pred, long long ld, const unsigned char* labels, long long ldl, int B, int V, int k,
              float* top_val, int* top_idx, unsigned char* top_lab, float* row_stats, cudaStream_t st) {
  LPM_REQUIRE(B > 0 && V > 0 && V <= EV_THREADS * EV_PER, "eval_topk: vocabulary must be in [1,%d] (got %d)", EV_THREADS * EV_PER, V);
  LPM_REQUIRE(k > 0 && k <= 1024, "eval_topk: k must be in [1,1024] (got %d)", k);
  eval_rows_kernel<<<B, EV_THREADS, 0, st>>>(pred, ld, labels, ldl, V, k, top_val, top_idx, top_lab, row_stats);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

int eval_metrics(const float* top_val, const unsigned char* top_lab, int B, int k, const float* row_stats, float* metrics,
                 cudaStream_t st) {
  const long long n = (long long)B * k;
  LPM_REQUIRE(n > 0 && n <= 16384, "eval_metrics: B*k must be in [1,16384] (got %lld)", n);
  int npow2 = 2;
  while (npow2 < n) npow2 <<= 1;
  const size_t smem = (size_t)npow2 * 8;
  static bool attr = false;
  if (!attr) {
    LPM_CUDA_CHECK(cudaFuncSetAttribute(eval_gap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    attr = true;
  }
  eval_gap_kernel<<<1, 1024, smem, st>>>(top_val, top_lab, (int)n, npow2, row_stats, B, metrics);
  LPM_CUDA_CHECK(cudaGetLastError());
  return LPM_OK;
}

}  // namespace lpm
