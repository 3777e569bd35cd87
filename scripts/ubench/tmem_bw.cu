// Micro-benchmark: tcgen05.ld throughput per SM (alone and under concurrent tcgen05.mma), sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I learnablepoolingmethods_b200/csrc -o scripts/ubench/tmem_bw scripts/ubench/tmem_bw.cu
#include "lpm_common.cuh"
#include <vector>
using namespace lpm;

__global__ void __launch_bounds__(320, 1) k_tmem(long long* out, int iters, int n_ld_warps, int do_mma, int mma_n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) tmem_alloc<512>(&slot);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    if (lane == 0 && do_mma) {
      const uint32_t idesc = umma_idesc_f16(128, mma_n, 0, 1);
      const uint32_t sa = smem_u32(smem), sb = sa + 32768;
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = umma_smem_desc(sa + ks * 32, 16, 1024);
          const uint64_t bd = umma_smem_desc(sb + ks * 2048, 8192, 1024);
          umma_f16(tb + 256, ad, bd, idesc, 1u);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      t1 = clock64();
      out[blockIdx.x * 16 + 8] = t1 - t0;
    }
  } else if (warp >= 2 && warp < 2 + n_ld_warps) {
    const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t r0[32], r1[32];
      tmem_ld32(tb + lane_addr + ((i * 64) & 255), r0);
      tmem_ld32(tb + lane_addr + ((i * 64 + 32) & 255), r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += r0[j] ^ r1[j];
    }
    t1 = clock64();
    if (lane == 0) out[blockIdx.x * 16 + (warp - 2)] = t1 - t0;
    if (acc == 0x12345) out[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 16 * 8);
  cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int grid : {1, 148}) for (int nw : {4, 8}) for (int mma : {0, 1, 2}) {
    int mma_n = mma == 2 ? 64 : 256;
    cudaMemset(d, 0, 148 * 16 * 8);
    k_tmem<<<grid, 320, 200 * 1024>>>(d, iters, nw, mma ? 1 : 0, mma_n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(148 * 16); cudaMemcpy(h.data(), d, 148 * 16 * 8, cudaMemcpyDeviceToHost);
    double ldmax = 0; for (int w = 0; w < nw; ++w) ldmax = std::max<double>(ldmax, h[w]);
    double bytes = double(nw) * iters * 2 * 32 * 32 * 4;
    printf("grid %3d ld_warps %d mma %d(N=%3d): ld %.0f clk -> %.1f B/clk/SM ; mma %lld clk for %d MMAs (%.1f clk each)\n",
           grid, nw, mma, mma_n, ldmax, bytes / ldmax, h[8], iters * 4, double(h[8]) / (iters * 4));
  }
  return 0;
}
