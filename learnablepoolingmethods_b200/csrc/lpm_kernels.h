// Internal (C++) interfaces between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lpm_b200.h"

namespace lpm {

using GemmArgs = lpm_gemm_desc;

int gemm_f16(const GemmArgs& g, cudaStream_t st, const lpm_gating_tail* tail = nullptr);
int gemm_pick_bn(int N);
int gemm_effective_splits(int K, int splits);
void gemm_set_pair_mode(int mode);

// lpm_elementwise.cu
int sample_stats_blocks();
// frame_index: null = SampleUniformFrames rule from nf; else explicit int32 [B*T] gather indices (nf unused)
int sample_stats(const void* x, int codes, float qmax, float qmin, const int* nf, const int* frame_index, int B,
                 int max_frames, int F, int T, float* partial, cudaStream_t st);
int sample_apply(const void* x, int codes, float qmax, float qmin, const int* nf, const int* frame_index, int B,
                 int max_frames, int F, int T, const float* scale, const float* shift, __half* y, int split_col,
                 __half* y2, cudaStream_t st);
int bn_finalize(const float* psum, const float* psq, int P, long long pstride, int C, double count,
                const float* gamma, const float* beta, float* mm, float* mv, float decay, float eps, int bessel,
                int training, float* scale, float* shift, float* save_mean, float* save_rstd, cudaStream_t st);
int scale_rows(const __half* x, const float* rs, long long rows, int D, __half* y, cudaStream_t st);
int layernorm_joint(__half* a, const __half* b, const float* b_row_scale, __half* u_out, int B, int rows, int D,
                    long long a_stride, long long b_stride, const float* gamma, const float* beta, float eps,
                    __half* y, long long y_stride, float* partial, float* save_mean_rstd, cudaStream_t st);
int splitk_reduce(const float* part, int splits, long long split_stride, const float* part2, int splits2, long long split_stride2,
                  long long n, int cols, const float* bias, int relu, float alpha, int accumulate, float* out32, __half* out16,
                  int split3, cudaStream_t st);
int gating_fwd(const float* act, const float* g, int g_splits, long long g_split_stride, float* g_sum, int split3, int B, int H,
               const float* wg_diag, const float* gamma, const float* beta, float* mm, float* mv, float decay, float eps,
               int training, float* out32, __half* out16, float* save_mean, float* save_rstd, cudaStream_t st);
int moe_mix(const float* logits, long long ld, int B, int V, int M, int expert_off, float* pred, cudaStream_t st);
int xent_loss(const float* pred, const uint8_t* labels, int B, int V, float* row_loss, float* loss, cudaStream_t st);
int cast_2d(const float* src, long long ld_src, int rows, int cols, __half* dst, long long ld_dst, int cols_dst,
            cudaStream_t st);
int transpose_2d(const float* src, int rows, int cols, float* dst, __half* dst16, cudaStream_t st);
int split_hi_lo(const float* src, long long ld_src, int rows, int cols, __half* dst, long long ld_dst, int mode, cudaStream_t st);
int vlad_finalize(const __half* z, const float* rscale, int B, int K, int D, int d_major, float* out, cudaStream_t st);

// lpm_attn.cu
int mha_fwd(const __half* qkv, long long ld, int B, int L, int Dm, int H, float scale, const float* key_scale,
            const float* key_shift, __half* out, long long ldo, float* lse, cudaStream_t st);

int mha_bwd(const __half* qkv, long long ld, const __half* o, const __half* dout, long long ldo, const float* lse,
            int B, int L, int Dm, int H, float scale, __half* dqkv, long long ldd, cudaStream_t st);

int mha_bwd_bn(int mode, const __half* qkv, long long ld, const __half* o, const __half* dout, long long ldo,
               const float* lse, int B, int L, int Dm, int H, const float* key_scale, const float* key_shift,
               const float* key_mean, const float* key_rstd, const float* m1, const float* m2, float* stat_partial,
               __half* dqkv, long long ldd, cudaStream_t st);

// lpm_attn_tc.cu: tcgen05 / TMEM attention core for depth-16 heads at length 256 (four heads per CTA)
void mha_set_tc_mode(int mode);
bool mha_tc_forward_enabled();
int mha_tc_backward_mode();
void mha_set_debug_clock(long long* buf);
bool mha_tc_eligible(int L, int Dm, int H, long long ld, long long ldo, const void* p0, const void* p1, const void* p2);
int mha_fwd_tc(const __half* qkv, long long ld, int B, int Dm, int H, float scale, __half* out, long long ldo, float* lse,
               cudaStream_t st);
int mha_bwd_tc(const __half* qkv, long long ld, const __half* o, const __half* dout, long long ldo, const float* lse, int B,
               int Dm, int H, float scale, __half* dqkv, long long ldd, cudaStream_t st);

// lpm_backward.cu
int xent_bwd(const float* pred, const uint8_t* labels, long long n, float gscale, const float* upstream, float* dpred,
             cudaStream_t st);
int moe_mix_bwd(const float* logits, long long ld, int B, int V, int M, int expert_off, const float* dpred,
                float loss_scale, __half* dl, long long ldo, int ncols, cudaStream_t st);
int colsum_chunks(long long rows);
int colsum(const void* x, int is_f32, long long ld, long long rows, int cols, float alpha, int accumulate,
           float* partial, float* out, cudaStream_t st);
int colsum_final(const float* partial, int chunks, long long pstride, int cols, float alpha, int accumulate,
                 float* out, cudaStream_t st);
int gating_bwd(const float* act, const float* g, int B, int H, const float* gamma, const float* beta,
               const float* mean, const float* rstd, const float* dout, float inv_scale, float* dact, __half* dg,
               float* dgamma, float* dbeta, const float* wg_diag, float* ddiag, cudaStream_t st);
int ln_bwd_chunks();
int layernorm_joint_bwd(const __half* u, const __half* dy, long long dy_stride, int B, int rows, int D,
                        const float* mean_rstd, const float* gamma, const __half* mask, __half* du,
                        __half* du_masked, float* part_sample, float* part_cols, float* part_cols_du,
                        cudaStream_t st);
int vlad_norm_bwd(const __half* z, const float* rs, const __half* dvh, long long rows, int K, int D,
                  const float* centers_t, __half* dz, float* q, cudaStream_t st);
int assign_bwd_blocks();
int assign_bwd1(const float* G, const __half* A, const float* q, const __half* S, const float* mean,
                const float* rstd, long long rows, int T, int K, __half* dsh, float* partial, cudaStream_t st);
int assign_bwd2(__half* dsh, const __half* S, const float* mean, const float* rstd, const float* gamma,
                const float* csum, long long rows, int K, cudaStream_t st);
int center_bwd(const __half* dV, const __half* Z, const float* a_sum, int B, int K, int D, const float* centers_t,
               const float* beta_in, float inv_scale, float* dCt, float* E, cudaStream_t st);
int input_bn_grad(const float* Wc, const float* dWc, const float* dCt, const float* E, int D, int K,
                  const float* gamma_in, float* dgamma_in, float* dbeta_in, cudaStream_t st);
int cast_scaled(const float* x, long long n, float alpha, __half* y, cudaStream_t st);

// lpm_optim.cu
int adam_clip_step(float* p, const float* g, float* m, float* v, const int* table, int n_chunks,
                   const int* chunk_begin, int n_tensors, const float* wd, const unsigned long long* sh_ptr,
                   const int* sh_cols, const long long* sh_ld, float clip, float lr_t, const float* lr_dev, float b1, float b2,
                   float eps, float* partial, float* factor, float* norms, int* flag, cudaStream_t st, int chunk0 = 0, int tensor0 = 0);
int step_begin(int* flag, int* skipped, cudaStream_t st);

int rank_grad_clip(const float* gram_a, const float* gram_g, int R, float alpha, float clip, float* factor, float* norm,
                   int* flag, cudaStream_t st);
int rank_adam_step(const __half* a16, long long lda, const __half* g16, long long ldg, int R, long long Kd, int N,
                   float alpha, const float* factor, const int* flag, float* w, float* m, float* v, __half* w16,
                   long long ldw16, float lr_t, const float* lr_dev, int tiled, float b1, float b2, float eps, void* workspace,
                   size_t workspace_bytes, cudaStream_t st);
size_t rank_adam_workspace_bytes(int R, int N);
int shard_sqnorm(const float* g, const float* p, const int* table, int n_chunks, const float* wd1, float* partial,
                 float* sumsq, cudaStream_t st);
int shard_adam(float* p, const float* g, float* m, float* v, const int* table, int n_chunks, const float* wd1,
               const float* sumsq, float clip, float* factor, float* norm, int* flag, const unsigned long long* sh_ptr,
               const int* sh_cols, const long long* sh_ld, float lr_t, float b1, float b2, float eps, cudaStream_t st);

// lpm_layernorm.cu
int layernorm_chain_supported(int rows, int D);
int layernorm_chain_fwd(const __half* a, long long a_stride, const __half* b, long long b_stride, const float* b_row_scale,
                        int B, int rows, int D, float eps, const float* gamma1, const float* beta1, __half* u1_out,
                        long long u1_stride, float* stats1, const float* gamma2, const float* beta2, __half* u2_out,
                        long long u2_stride, float* stats2, __half* y, long long y_stride, __half* y_lo, cudaStream_t st);

// lpm_eval.cu
int eval_topk(const float* pred, long long ld, const unsigned char* labels, long long ldl, int B, int V, int k,
              float* top_val, int* top_idx, unsigned char* top_lab, float* row_stats, cudaStream_t st);
int eval_metrics(const float* top_val, const unsigned char* top_lab, int B, int k, const float* row_stats, float* metrics,
                 cudaStream_t st);

// lpm_pool.cu
int netvlad_pool_fwd(const __half* x, long long ldx, long long x_batch_stride, const __half* wc, long long ldw,
                     const float* logit_scale, const float* logit_shift, const __half* centers_t16,
                     const int* valid_frames, int B, int T, int D, int K, __half* z, float* rscale, float* a_sum,
                     __half* assign, const __half* assign_in, long long* debug_clock, cudaStream_t st);

// lpm_v2.cu
int mha_logit_stats(const __half* qkv, long long ld, int B, int L, int Dm, int H, float* partial, cudaStream_t st);
int colstats_chunks(long long rows, int C);
int colstats(const __half* x, long long ld, long long rows, int C, float* partial, cudaStream_t st);
int affine_cols(const __half* x, __half* y, long long rows, int C, const float* scale, const float* shift,
                cudaStream_t st);
int bn_bwd_stats(const void* dy, int dy_f32, long long ld_dy, const float* q, int T, const __half* x, long long ld_x,
                 long long rows, int C, const float* p0, const float* p1, int mode, float* partial, cudaStream_t st);
int bn_bwd_apply(const void* dy, int dy_f32, const float* q, int T, __half* dx, const __half* x, long long rows, int C,
                 const float* mean, const float* rstd, const float* gamma, const float* csum, int relu, cudaStream_t st);
int sub_q_cast(const float* G, const float* q, long long rows, int T, int K, __half* out, cudaStream_t st);
int dmajor_to_kmajor_f16(const __half* in, long long in_stride, int B, int K, int D, __half* out, cudaStream_t st);
int dropout_f16(__half* x, __half* out, long long n, const __half* mask_in, __half* mask_out, unsigned long long seed,
                const unsigned long long* seed_dev, float rate, cudaStream_t st);
int vlad_dmajor_f16(const __half* z, const float* rscale, int B, int K, int D, __half* out, long long out_stride,
                    cudaStream_t st);

// lpm_willow.cu
int random_frame_index(const int* nf, const float* uniform, unsigned long long seed, int B, int T, int max_frames,
                       int mode, int* idx, cudaStream_t st);
unsigned long long ortho_reg_workspace_bytes(int D, int K);
int ortho_reg(const float* w, int D, int K, float scale, float grad_scale, int accumulate, float* value, float* dw,
              float* ws, unsigned long long ws_bytes, cudaStream_t st);

// lpm_head.cu
int hidden_bn_relu6_fwd(const float* x, int B, int H, const float* gamma, const float* beta, float* mm, float* mv,
                        float decay, float eps, int training, int relu6, float* out32, __half* out16, float* save_mean,
                        float* save_rstd, cudaStream_t st);
int hidden_bn_relu6_bwd(const float* x, const float* y, const float* dy, int B, int H, const float* gamma, const float* mean,
                        const float* rstd, int relu6, float inv_scale, float* dx, float* dgamma, float* dbeta, cudaStream_t st);
int add_diag(float* m, int n, long long ld, const float* d, float alpha, cudaStream_t st);
int l2_normalize_rows(const float* x, long long rows, int F, float* y, cudaStream_t st);

}  // namespace lpm
