// Internal (C++) interfaces between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lpm_b200.h"

namespace lpm {

using GemmArgs = lpm_gemm_desc;

int gemm_f16(const GemmArgs& g, cudaStream_t st);
int gemm_pick_bn(int N);
int gemm_effective_splits(int K, int splits);

// lpm_elementwise.cu
int sample_stats_blocks();
int sample_stats(const float* x, const int* nf, int B, int max_frames, int F, int T, float* partial, cudaStream_t st);
int sample_apply(const float* x, const int* nf, int B, int max_frames, int F, int T, const float* scale,
                 const float* shift, __half* y, cudaStream_t st);
int bn_finalize(const float* psum, const float* psq, int P, long long pstride, int C, double count,
                const float* gamma, const float* beta, float* mm, float* mv, float decay, float eps, int bessel,
                int training, float* scale, float* shift, float* save_mean, float* save_rstd, cudaStream_t st);
int layernorm_joint(__half* a, const __half* b, const float* b_row_scale, int B, int rows, int D,
                    long long a_stride, long long b_stride, const float* gamma, const float* beta, float eps,
                    __half* y, long long y_stride, float* partial, float* save_mean_rstd, cudaStream_t st);
int splitk_reduce(const float* part, int splits, long long split_stride, long long n, int cols, const float* bias,
                  int relu, float alpha, int accumulate, float* out32, __half* out16, cudaStream_t st);
int gating_fwd(const float* act, const float* g, int B, int H, const float* wg_diag, const float* gamma,
               const float* beta, float* mm, float* mv, float decay, float eps, int training, float* out32,
               __half* out16, float* save_mean, float* save_rstd, cudaStream_t st);
int moe_mix(const float* logits, long long ld, int B, int V, int M, float* pred, cudaStream_t st);
int xent_loss(const float* pred, const uint8_t* labels, int B, int V, float* row_loss, float* loss, cudaStream_t st);
int cast_2d(const float* src, long long ld_src, int rows, int cols, __half* dst, long long ld_dst, int cols_dst,
            cudaStream_t st);
int transpose_2d(const float* src, int rows, int cols, float* dst, cudaStream_t st);
int vlad_finalize(const __half* z, const float* rscale, int B, int K, int D, int d_major, float* out, cudaStream_t st);

// lpm_attn.cu
int mha_fwd(const __half* qkv, long long ld, int B, int L, int Dm, int H, float scale, const float* key_scale,
            const float* key_shift, __half* out, long long ldo, float* lse, cudaStream_t st);

// lpm_pool.cu
int netvlad_pool_fwd(const __half* x, long long ldx, long long x_batch_stride, const __half* wc, long long ldw,
                     const float* logit_scale, const float* logit_shift, const float* centers_t,
                     const int* valid_frames, int B, int T, int D, int K, __half* z, float* rscale, float* a_sum,
                     __half* assign, cudaStream_t st);

}  // namespace lpm
