/* lpm_b200.h -- C ABI of liblpm_b200.so: B200 (sm_100a) kernels for the NetVLAD learnable-pooling
 * hot path of pomonam/LearnablePoolingMethods (NetVladV1 / NetVladV2).
 *
 * The reference has NO FFI (it is TensorFlow-1.x graph code); each entry point below names the
 * reference lines it replaces.  Conventions:
 *   - every function returns int: 0 = ok, negative = error (lpm_last_error() gives the message,
 *     thread-local);
 *   - raw device pointers + explicit shapes/strides + caller-owned workspaces + a cudaStream_t
 *     (passed as void*); the library never allocates device memory, never synchronises and never
 *     owns tensors;
 *   - fp16 ("f16") operands, fp32 accumulation, fp32 parameters/statistics;
 *   - all kernels are sm_100a only; there is no fallback path.
 */
#ifndef LPM_B200_H_
#define LPM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPM_OK 0
#define LPM_ERR_ARG (-1)
#define LPM_ERR_CUDA (-2)
#define LPM_ERR_WORKSPACE (-3)
#define LPM_ERR_DEVICE (-4)

typedef void* lpm_stream_t; /* cudaStream_t */

/* Library version (major*10000 + minor*100 + patch). */
int lpm_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* lpm_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Generic dense product  D[b] = epilogue(alpha * rowscale * A[b] x B[b])   (fp16 in, fp32 accumulate)
 * Replaces tf.matmul / tf.layers.dense / slim.fully_connected on the hot path and their autodiff
 * transposes: frame_level_models.py:2319,2347,2781,2815; transformer_utils.py:559-561,583-585,
 * 701-711,641-643,673,742,754; video_level_models.py:86-114.
 *   A: a_mn=0 -> memory [M][K] (K contiguous, row stride lda);  a_mn=1 -> memory [K][M].
 *   B: b_mn=0 -> memory [N][K] (K contiguous, row stride ldb);  b_mn=1 -> memory [K][N].
 *   batch strides of 0 share the operand across the batch.  lda/ldb/batch strides: multiples of 8.
 *   splits>1: split-K; out must be an fp32 workspace [splits][batch][M][ldc] (out_split_stride
 *   elements apart) reduced afterwards with lpm_splitk_reduce.
 *   mask / add1 / add2 (optional, fp16, batch stride = out_batch_stride): fused ReLU-backward mask and
 *   residual-gradient sums, applied after bias/ReLU in the order "+= add1 + add2, then mask".
 *   stat_sum/stat_sq (optional): per-row sum / sum of squares of the epilogue values over the N
 *   columns of each N-tile half-set, written to [batch][2*n_tiles][M] (deterministic partials, one row per
 *   N-tile and epilogue warp group; rows of groups without columns stay untouched: zero them first).
 * ------------------------------------------------------------------------------------------- */
typedef struct lpm_gemm_desc {
  const void* A; int a_mn; long long lda; long long a_batch_stride;
  const void* B; int b_mn; long long ldb; long long b_batch_stride;
  int M, N, K, batch, splits, force_bn;
  void* out; int out_f32; long long ldc; long long out_batch_stride; long long out_split_stride;
  const float* bias;       /* [N] or NULL */
  const float* row_scale;  /* [batch][M] or NULL */
  long long row_scale_batch_stride;
  int relu; int accumulate; float alpha;
  float* stat_sum; float* stat_sq;
  const void* mask; long long ld_mask;                     /* fp16 [M][ld_mask]: out = mask>0 ? out : 0 */
  const void* add1; const void* add2; long long ld_add;    /* fp16 addends [M][ld_add]: out += add1 (+ add2) */
  int no_tma_store;                                        /* debug: force the direct-store epilogue */
} lpm_gemm_desc;

int lpm_gemm_f16(const lpm_gemm_desc* desc, lpm_stream_t stream);
/* ---------------------------------------------------------------------------------------------
 * K3: the hidden projection with the context-gating epilogue fused INTO the split-K GEMM kernel
 * (frame_level_models.py:2314-2368: hidden1_weights product + biases, gating_weights product, gating_bn,
 * sigmoid, element-wise product).  `desc` must describe a split-K product with fp32 partial output
 * [splits][B][H] (B <= 128 rows, N = H); all CTAs of the launch are resident (grid <= SM count), so after its last
 * tile every CTA
 *   1. waits on a grid-wide barrier (tail->counters, four ints, zero before the first call; the kernel leaves them zero),
 *   2. sums a slice of H/grid-rounded-to-4 columns of all partials in a fixed order (plus the partials `part2` of an
 *      earlier launch: the other pass of the split-precision product), adds the bias -> act32 / act16,
 *   3. waits on a second barrier, forms the gate pre-activations of ITS columns in fp32 from act32 and the fp32 gating
 *      weights (exact: no fp16 operands), batch-norms them over the B rows (training: batch statistics + moving-average
 *      update; else moving statistics), and writes act * sigmoid(.) -> out32 / out16.
 * act16 / out16: fp16 [B][H], or with *_split3 != 0 the split-precision operand [B][3H] = [hi | lo | hi].
 * Replaces lpm_gemm_f16 + lpm_splitk_reduce_ex + lpm_gemm_f16 (gate) + lpm_gating_fwd_ex: one launch instead of four.
 * ------------------------------------------------------------------------------------------- */
typedef struct lpm_gating_tail {
  int* counters;                                   /* [4] grid-barrier scratch */
  const float* part2; int splits2; long long split_stride2;   /* optional second partial set [splits2][B][H] */
  const float* bias;                               /* hidden1_biases [H] or NULL */
  float* act32; void* act16; int act_split3;       /* hidden activation */
  const float* wg; long long ldwg;                 /* gating_weights fp32 [H][ldwg] */
  const float* wg_diag;                            /* diag(gating_weights) for --gating_remove_diag, or NULL */
  const float* gamma; const float* beta; float* moving_mean; float* moving_var;
  float decay, eps; int training;
  float* g_sum;                                    /* [B][H] gate pre-activations (for the backward) */
  float* out32; void* out16; int out_split3;       /* gated activation */
  float* save_mean; float* save_rstd;              /* [H] or NULL */
} lpm_gating_tail;
int lpm_gemm_splitk_gated_fwd(const lpm_gemm_desc* desc, const lpm_gating_tail* tail, lpm_stream_t stream);
/* N-tile width the kernel will use for a given N (for sizing stat partial buffers). */
int lpm_gemm_tile_n(int N);
/* Number of non-empty K splits the kernel will use. */
int lpm_gemm_splits(int K, int requested_splits);

/* ---------------------------------------------------------------------------------------------
 * Split-K reduction:  out[r][c] = act(alpha * sum_s part[s][r][c] + bias[c]) -> fp32 and/or fp16.
 * Completes the hidden projection (frame_level_models.py:2319,2329-2334) and split-K weight
 * gradients.  n = rows*cols elements per split, splits are split_stride elements apart.
 * ------------------------------------------------------------------------------------------- */
int lpm_splitk_reduce(const float* part, int splits, long long split_stride, long long n, int cols,
                      const float* bias, int relu, float alpha, int accumulate, float* out_f32,
                      void* out_f16, lpm_stream_t stream);
/* Same with two options used by the hidden projection (frame_level_models.py:2319-2334): part2 (may be NULL) is a second
 * set of partials summed in -- the low-order pass of the split-precision product -- and split3 != 0 makes out_f16 the
 * split-precision operand [n / cols][3 * cols] = [ hi | lo | hi ] of the result (see lpm_split_hi_lo_f16). */
int lpm_splitk_reduce_ex(const float* part, int splits, long long split_stride, const float* part2, int splits2,
                         long long split_stride2, long long n, int cols, const float* bias, int relu, float alpha,
                         int accumulate, float* out_f32, void* out_f16, int split3, lpm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Frame sampling + input batch norm.
 *   lpm_sample_bn_stats : per-feature sum / sum-of-squares partials of the uniformly sampled frames
 *                         (model_utils.py:101-122 gather; frame_level_models.py:2265-2271 statistics).
 *                         partial must hold lpm_sample_stats_blocks()*2*F floats.
 *   lpm_sample_bn_apply : y[b*T+i][:] = fp16(x[b, idx(b,i), :] * scale + shift), idx = int32(fl32(i/T)*nf).
 * x: fp32 [B][max_frames][F] (already L2-normalised by the caller, train.py:264); num_frames int32 [B].
 * ------------------------------------------------------------------------------------------- */
int lpm_sample_stats_blocks(void);
int lpm_sample_bn_stats(const float* x, const int* num_frames, int B, int max_frames, int F, int T,
                        float* partial, lpm_stream_t stream);
/* y2_f16 == NULL: one [B*T][F] matrix; else columns [0,split_col) -> y_f16 [B*T][split_col], rest -> y2_f16. */
int lpm_sample_bn_apply(const float* x, const int* num_frames, int B, int max_frames, int F, int T,
                        const float* scale, const float* shift, void* y_f16, int split_col, void* y2_f16,
                        lpm_stream_t stream);
/* Ingest side of the boundary: the same two passes fed with the YT8M uint8 codes
 * [B][max_frames][F] as the reader decodes them (readers.py:185-193).  Each gathered frame is dequantised
 * (utils.py:28-43: q*range/255 + range/512 + min, range = max - min) and L2-normalised over all F features
 * (train.py:264: x*rsqrt(max(sum x^2, 1e-12))) on the fly, so the fp32 [B][300][1152] tensor never exists
 * and the host->device copy shrinks 4x.  Results equal the fp32 entry points on dequantised+normalised input. */
int lpm_sample_bn_stats_u8(const unsigned char* codes, float max_quantized_value, float min_quantized_value,
                           const int* num_frames, int B, int max_frames, int F, int T, float* partial,
                           lpm_stream_t stream);
int lpm_sample_bn_apply_u8(const unsigned char* codes, float max_quantized_value, float min_quantized_value,
                           const int* num_frames, int B, int max_frames, int F, int T, const float* scale,
                           const float* shift, void* y_f16, int split_col, void* y2_f16, lpm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * slim.batch_norm finalisation (eps 1e-3, decay 0.999 passed by the caller): reduces P partial
 * (sum, sumsq) rows (pstride floats apart) over `count` samples into the folded affine
 * y = x*scale + shift, updates the moving statistics (Bessel-corrected variance when bessel!=0),
 * or (training==0) folds the moving statistics.  frame_level_models.py:2266,2784; slim semantics.
 * ------------------------------------------------------------------------------------------- */
int lpm_batchnorm_finalize(const float* psum, const float* psq, int P, long long pstride, int C, double count,
                           const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                           float decay, float eps, int bessel, int training, float* scale, float* shift,
                           float* save_mean, float* save_rstd, lpm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused NetVLAD pooling forward (K1): soft-assignment + residual aggregation + norms.
 * frame_level_models.py:2775-2822 (NetVLAD.forward) for one modality.
 *   x            fp16 [B][T][D] (row stride ldx, video stride x_batch_stride), the batch-normed frames
 *   wc           fp16 [D][K] cluster_weights (row stride ldw)
 *   logit_scale/shift  fp32 [K]: cluster_bn folded affine, or (1, cluster_biases)
 *   centers_t16  fp16 [K][D]: cluster_weights2[0] (V1) / cluster_centers (V2) transposed to cluster-major and
 *                rounded to fp16 (lpm_transpose_f32_dual); TMA-loaded into the output staging slab
 *   valid_frames int32 [B] or NULL: frames t >= valid_frames[b] get zero assignment (masked mode)
 *   z            fp16 [B][K][D]  un-normalised cluster-major descriptor V^T
 *   rscale       fp32 [B][K]     vlad[b,k,:] = z[b,k,:]*rscale[b,k]  (intra-norm x global norm)
 *   a_sum        fp32 [B][K] or NULL;  assign fp16 [B][T][K] or NULL (saved for the backward)
 *   assign_in    fp16 [B][T][K] or NULL: externally supplied cluster similarities (NetVladAttenCluster,
 *                video_pooling_modules.py:1628-1652): the logits/softmax phase is skipped, wc/logit_* unused
 * Limits: T <= 256, D % 64 == 0, K % 8 == 0, K <= 512 (K > 256 runs as a 2-CTA cluster per video, each CTA
 * owning 256 clusters; softmax row statistics and the global norm are exchanged through distributed shared memory).
 * ------------------------------------------------------------------------------------------- */
int lpm_netvlad_pool_fwd(const void* x, long long ldx, long long x_batch_stride, const void* wc, long long ldw,
                         const float* logit_scale, const float* logit_shift, const void* centers_t16,
                         const int* valid_frames, int B, int T, int D, int K, void* z, float* rscale,
                         float* a_sum, void* assign, const void* assign_in, lpm_stream_t stream);
/* Profiling aid: when non-NULL, lpm_netvlad_pool_fwd writes 8 clock64 phase stamps per video ([B][8] int64:
 * start, logits done, softmax done, a_sum done, aggregation done). */
void lpm_debug_set_pool_clock(long long* buf);
/* Measurement aid: 0 = lpm_gemm_f16 never uses 2-CTA (cta_group::2) tiles, 1 = automatic (default). */
void lpm_debug_set_gemm_pair_mode(int mode);
/* Measurement aid.  Low two bits select the backward of lpm_mha_core_bwd when the shape is eligible (depth 16, length 256,
 * heads a multiple of 4): 0 = warp-level (mma.sync) kernel, 1 = tcgen05 kernel with the P / dS operands in shared memory,
 * 2 = tcgen05 kernel with the dV / dK operands in TMEM (default; LPM_MHA_TC=0/1/2).  Bits 2-3 select the forward of
 * lpm_mha_core_fwd: 0 = warp-level kernel (default), 4 = tcgen05 with P through shared memory, 8 = tcgen05 with P in TMEM and
 * two CTAs per SM (LPM_MHA_TC_FWD=0/1/2). */
void lpm_debug_set_mha_tc_mode(int mode);
/* Profiling aid: when non-NULL, CTA 0 of the tcgen05 attention backward writes clock64 stamps per unit:
 * [4 warpgroups][16 units][4] (unit begins, S/dP available, math issued, operand slots free) then [16 units][4] of the
 * MMA-issuing warp at offset 256 (next S/dP issued, P/dS visible, gradient MMAs issued). */
void lpm_debug_set_mha_clock(long long* buf);
/* vlad = z * rscale as fp32: d_major!=0 -> [B][D*K] (reference flatten, :2821), else [B][K][D]. */
int lpm_netvlad_finalize(const void* z, const float* rscale, int B, int K, int D, int d_major, float* out,
                         lpm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-head attention core: out = softmax(scale*QK^T [*key_scale + key_shift]) V per (sample, head).
 * transformer_utils.py:563-581 (V1) and :641-664 (V2, logits batch norm = per-key affine).
 *   qkv fp16 [B*L][ld] with q | k | v blocks of Dm columns; head h = columns [h*depth, (h+1)*depth).
 *   depth = Dm/H in {8,16}; L % 16 == 0.  lse fp32 [B][H][L] or NULL.
 * ------------------------------------------------------------------------------------------- */
int lpm_mha_core_fwd(const void* qkv, long long ld, int B, int L, int Dm, int H, float scale,
                     const float* key_scale, const float* key_shift, void* out, long long ldo, float* lse,
                     lpm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Joint-axis layer norm with fused residual (tf.contrib.layers.layer_norm, begin_norm_axis=1):
 *   u = a + b*b_row_scale (fp16, written to u_out, or over a when u_out is NULL);
 *   y = (u-mean_b)*rstd_b*gamma[d] + beta[d].
 * transformer_utils.py:406-411,712-713.  partial: B*64 floats; save_mean_rstd: [B][2] or NULL.
 * ------------------------------------------------------------------------------------------- */
int lpm_layernorm_joint_fwd(void* a, const void* b, const float* b_row_scale, void* u_out, int B, int rows, int D,
                            long long a_stride, long long b_stride, const float* gamma, const float* beta,
                            float eps, void* y, long long y_stride, float* partial, float* save_mean_rstd,
                            lpm_stream_t stream);

/* One-pass variant for one or two chained joint-axis layer norms sharing the residual b (the tail of the encoder
 * block, transformer_utils.py:712-713 then :410-411):
 *   u1 = a + b*b_row_scale ; y1 = LN(u1; gamma1, beta1) ; [u2 = y1 + b ; y2 = LN(u2; gamma2, beta2)]
 * A thread-block cluster owns a sample, keeps u in shared memory and exchanges the moments over DSMEM, so a and b
 * are read once and y (= y2 when gamma2 is given, else y1) written once.  u1_out / u2_out (fp16, may be NULL; u1_out
 * may alias a) and stats1 / stats2 ([B][2] = mean, rstd; may be NULL) are what the backward needs.
 * lpm_layernorm_chain_supported: 1 when a sample of rows x D fits the cluster's shared memory. */
int lpm_layernorm_chain_supported(int rows, int D);
int lpm_layernorm_chain_fwd(const void* a, long long a_stride, const void* b, long long b_stride, const float* b_row_scale,
                            int B, int rows, int D, float eps, const float* gamma1, const float* beta1, void* u1_out,
                            long long u1_stride, float* stats1, const float* gamma2, const float* beta2, void* u2_out,
                            long long u2_stride, float* stats2, void* y, long long y_stride, lpm_stream_t stream);
/* Same, additionally writing y_lo = fp16(y - fp16(y)) with y's layout: (y, y_lo) is the split-precision left operand of
 * the hidden projection (frame_level_models.py:2319), see lpm_split_hi_lo_f16. */
int lpm_layernorm_chain_fwd_split(const void* a, long long a_stride, const void* b, long long b_stride, const float* b_row_scale,
                                  int B, int rows, int D, float eps, const float* gamma1, const float* beta1, void* u1_out,
                                  long long u1_stride, float* stats1, const float* gamma2, const float* beta2, void* u2_out,
                                  long long u2_stride, float* stats2, void* y, long long y_stride, void* y_lo,
                                  lpm_stream_t stream);

/* Context gating (frame_level_models.py:2342-2368): act * sigmoid(BN_batch(g - diag*act)). */
int lpm_gating_fwd(const float* act, const float* g, int B, int H, const float* wg_diag, const float* gamma,
                   const float* beta, float* moving_mean, float* moving_var, float decay, float eps,
                   int training, float* out_f32, void* out_f16, float* save_mean, float* save_rstd,
                   lpm_stream_t stream);
/* Same, fed straight from the gate product's split-K partials: g = [g_splits][B][H] (g_split_stride elements apart) is
 * summed in a fixed order into g_sum [B][H] (kept for the backward) before the batch norm; split3 != 0 makes out_f16
 * the split-precision operand [B][3H] = [ hi | lo | hi ] of the gated activation for the MoE product. */
int lpm_gating_fwd_ex(const float* act, const float* g, int g_splits, long long g_split_stride, float* g_sum, int B, int H,
                      const float* wg_diag, const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                      float decay, float eps, int training, float* out_f32, void* out_f16, int split3, float* save_mean,
                      float* save_rstd, lpm_stream_t stream);

/* MoE mixing (video_level_models.py:116-126): logits fp32 [B][ld]; gates V*(M+1) at column 0, experts V*M at
 * column expert_off (lets the caller pad the gate block to a 16-byte boundary). */
int lpm_moe_mix_fwd(const float* logits, long long ld, int B, int V, int M, int expert_off, float* pred,
                    lpm_stream_t stream);

/* CrossEntropyLoss (losses.py:44-51): labels uint8 [B][V]; row_loss [B]; loss scalar = mean_b. */
int lpm_xent_fwd(const float* pred, const uint8_t* labels, int B, int V, float* row_loss, float* loss,
                 lpm_stream_t stream);

/* y[r][:] = fp16(x[r][:] * row_scale[r]): materialises vlad = z*rscale for the training path. */
int lpm_scale_rows_f16(const void* x, const float* row_scale, long long rows, int D, void* y, lpm_stream_t stream);

/* fp32 -> fp16 2-D copy with zero column padding (parameter shadows) and fp32 transpose. */
int lpm_cast_f32_to_f16(const float* src, long long ld_src, int rows, int cols, void* dst, long long ld_dst,
                        int cols_dst, lpm_stream_t stream);
int lpm_transpose_f32(const float* src, int rows, int cols, float* dst, lpm_stream_t stream);
/* dst32 (fp32) and/or dst16 (fp16) [cols][rows] = src^T; either destination may be NULL. */
int lpm_transpose_f32_dual(const float* src, int rows, int cols, float* dst32, void* dst16, lpm_stream_t stream);

/* =============================================================================================
 * Backward entry points (autodiff of the reference lines cited by the matching forward call).
 * Activation gradients are fp16 scaled by the caller's loss scale; parameter gradients are fp32 and
 * unscaled (inv_scale = 1/loss_scale).
 * ============================================================================================= */
/* losses.py:44-51: dpred = -(y/(p+1e-5) - (1-y)/(1-p+1e-5)) * gscale  (gscale = upstream/B). */
int lpm_xent_bwd(const float* pred, const uint8_t* labels, long long n, float gscale, float* dpred,
                 lpm_stream_t stream);
/* Same with the upstream gradient of the loss read from device memory: gscale_effective = gscale * upstream_dev[0]
 * (the autograd edge `loss.backward()` hands dLoss over as a device scalar; no host round trip). */
int lpm_xent_bwd_dev(const float* pred, const uint8_t* labels, long long n, float gscale, const float* upstream_dev,
                     float* dpred, lpm_stream_t stream);
/* video_level_models.py:116-126: dlogits fp16 [B][ldo] (x loss_scale), padding columns zeroed. */
int lpm_moe_mix_bwd(const float* logits, long long ld, int B, int V, int M, int expert_off, const float* dpred,
                    float loss_scale, void* dlogits_f16, long long ldo, int ncols, lpm_stream_t stream);
/* out[c] (+)= alpha*sum_r x[r][c]; partial: lpm_colsum_chunks(rows)*cols floats. */
int lpm_colsum_chunks(long long rows);
int lpm_colsum(const void* x, int is_f32, long long ld, long long rows, int cols, float alpha, int accumulate,
               float* partial, float* out, lpm_stream_t stream);
int lpm_colsum_final(const float* partial, int chunks, long long pstride, int cols, float alpha, int accumulate,
                     float* out, lpm_stream_t stream);
/* frame_level_models.py:2342-2368 backward. */
/* wg_diag / ddiag (both or neither; --gating_remove_diag, :2349-2352): the batch norm saw g - diag*act; dact then also
 * carries -diag*dv and ddiag[c] = -sum_b dv*act (unscaled) is the extra gradient of gating_weights_2's diagonal. */
int lpm_gating_bwd(const float* act, const float* g, int B, int H, const float* gamma, const float* beta,
                   const float* mean, const float* rstd, const float* dout, float inv_scale, float* dact,
                   void* dg_f16, float* dgamma, float* dbeta, const float* wg_diag, float* ddiag, lpm_stream_t stream);
/* m[i][i] += alpha*d[i], i < n (fp32, row stride ld). */
int lpm_add_diag(float* m, int n, long long ld, const float* d, float alpha, lpm_stream_t stream);
/* --netvlad_relu head (frame_level_models.py:2321-2327, 2339-2340): y = relu6(slim.batch_norm(x)) over the batch rows of
 * x fp32 [B][H] (relu6 = 0: batch norm only).  Training updates the moving statistics (decay, Bessel-corrected variance)
 * and saves (mean, rstd); backward: dy at y (fp32, loss-scaled) -> dx (may alias dy), dgamma / dbeta (x inv_scale). */
int lpm_hidden_bn_relu6_fwd(const float* x, int B, int H, const float* gamma, const float* beta, float* moving_mean,
                            float* moving_var, float decay, float eps, int training, int relu6, float* out_f32,
                            void* out_f16, float* save_mean, float* save_rstd, lpm_stream_t stream);
int lpm_hidden_bn_relu6_bwd(const float* x, const float* y, const float* dy, int B, int H, const float* gamma,
                            const float* mean, const float* rstd, int relu6, float inv_scale, float* dx, float* dgamma,
                            float* dbeta, lpm_stream_t stream);
/* Joint layer norm backward: du = rstd*(dy*gamma - c1/N - xhat*c2/N); with `mask`, du_masked = du o (mask>0)
 * (ReLU backward of the pre-residual branch).  part_sample: B*chunks*2, part_cols: B*chunks*2*D (sum dy*xhat |
 * sum dy per column), part_cols_du (optional): B*chunks*D column sums of du_masked (du without mask);
 * chunks = lpm_layernorm_bwd_chunks(). */
int lpm_layernorm_bwd_chunks(void);
int lpm_layernorm_joint_bwd(const void* u, const void* dy, long long dy_stride, int B, int rows, int D,
                            const float* mean_rstd, const float* gamma, const void* mask, void* du,
                            void* du_masked, float* part_sample, float* part_cols, float* part_cols_du,
                            lpm_stream_t stream);
/* frame_level_models.py:2819-2822 backward: dz fp16 [rows][D], q[rows] = dz . C[:,k]. */
int lpm_netvlad_norm_bwd(const void* z, const float* rscale, const void* dvhat, long long rows, int K, int D,
                         const float* centers_t, void* dz, float* q, lpm_stream_t stream);
/* Soft-assignment (softmax + cluster_bn) backward, two passes; partial: lpm_assign_bwd_blocks()*2*K. */
int lpm_assign_bwd_blocks(void);
int lpm_assign_bwd1(const float* G, const void* assign, const float* q, const void* S, const float* mean,
                    const float* rstd, long long rows, int T, int K, void* dshat, float* partial,
                    lpm_stream_t stream);
int lpm_assign_bwd2(void* dshat, const void* S, const float* mean, const float* rstd, const float* gamma,
                    const float* csum, long long rows, int K, lpm_stream_t stream);
/* cluster_weights2 / input_bn parameter gradients (see lpm_backward.cu for the algebra). */
int lpm_center_bwd(const void* dV, const void* Z, const float* a_sum, int B, int K, int D, const float* centers_t,
                   const float* beta_in, float inv_scale, float* dCt, float* E, lpm_stream_t stream);
int lpm_input_bn_grad(const float* Wc, const float* dWc, const float* dCt, const float* E, int D, int K,
                      const float* gamma_in, float* dgamma_in, float* dbeta_in, lpm_stream_t stream);
int lpm_cast_scaled_f16(const float* x, long long n, float alpha, void* y, lpm_stream_t stream);
/* Split-precision operands for the head's products (context gating frame_level_models.py:2347, MoE
 * video_level_models.py:86-114): x = hi + lo, hi = fp16(x), lo = fp16(x - hi).  A W ~= A_hi W_hi + A_lo W_hi + A_hi W_lo is
 * then one lpm_gemm_f16 over a reduction of 3K:
 *   along_rows = 0 (activations): src fp32 [rows][cols] -> dst fp16 [rows][3*cols] = [ hi | lo | hi ]
 *   along_rows = 1 (weights [K][N]): dst rows [0,K) = hi, [K,2K) = hi, [2K,3K) = lo  (dst holds 3*rows rows)
 *   along_rows = 2: dst fp16 [rows][cols] = lo only (hidden1_weights: the hi part is the ordinary fp16 operand) */
int lpm_split_hi_lo_f16(const float* src, long long ld_src, int rows, int cols, void* dst_f16, long long ld_dst,
                        int along_rows, lpm_stream_t stream);
/* transformer_utils.py:563-581 backward: dqkv fp16 [B*L][ldd] in the qkv layout; L <= 256. */
int lpm_mha_core_bwd(const void* qkv, long long ld, const void* o, const void* dout, long long ldo, const float* lse,
                     int B, int L, int Dm, int H, float scale, void* dqkv, long long ldd, lpm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Optimiser step on a flat fp32 buffer (train.py:321-336, utils.py:170-189, tf.train.AdamOptimizer):
 * per-tensor L2-regulariser term (wd[t]*p), per-tensor clip_by_norm(clip), Adam with the TF bias-corrected
 * step lr_t = lr*sqrt(1-b2^t)/(1-b1^t).  table: int32 [n_chunks][4] = {tensor id, start/32, length, element
 * offset inside the tensor / 32}; chunk_begin: int32 [n_tensors+1].  Optional fused refresh of the fp16 operand
 * shadows: sh_ptr[t] (device address or 0), sh_cols[t] (inner dimension), sh_ld[t] (shadow row stride).
 * Scratch: partial [n_chunks], factor/norms [n_tensors], flag [1] (set to 1 and the update skipped when a
 * gradient norm is non-finite).
 * ------------------------------------------------------------------------------------------- */
int lpm_adam_clip_step(float* p, const float* g, float* m, float* v, const int* table, int n_chunks,
                       const int* chunk_begin, int n_tensors, const float* wd, const unsigned long long* sh_ptr,
                       const int* sh_cols, const long long* sh_ld, float clip, float lr_t, float b1, float b2,
                       float eps, float* partial, float* factor, float* norms, int* flag, lpm_stream_t stream);
/* Same step with the bias-corrected step size read from device memory (lr_t_dev[0]) at execution time: the launch can
 * then be captured in a CUDA graph and replayed while the learning-rate schedule (train.py:244-249) advances. */
int lpm_adam_clip_step_dev(float* p, const float* g, float* m, float* v, const int* table, int n_chunks,
                           const int* chunk_begin, int n_tensors, const float* wd, const unsigned long long* sh_ptr,
                           const int* sh_cols, const long long* sh_ld, float clip, const float* lr_t_dev, float b1, float b2,
                           float eps, float* partial, float* factor, float* norms, int* flag, lpm_stream_t stream);
/* The same step restricted to tensors tensor0 .. tensor0 + n_tensors - 1, whose chunks are chunk0 .. chunk0 + n_chunks - 1 of
 * `table` (all arrays stay indexed by absolute chunk / tensor ids).  Clipping is per tensor (utils.py:181-188), so a group of
 * variables can be updated as soon as ITS gradients are final: the trainer updates the MoE / gating variables (43 % of the
 * non-factored parameters) underneath the modalities' backward.  lr_t_dev non-NULL overrides lr_t. */
int lpm_adam_clip_step_range(float* p, const float* g, float* m, float* v, const int* table, int chunk0, int n_chunks,
                             const int* chunk_begin, int tensor0, int n_tensors, const float* wd,
                             const unsigned long long* sh_ptr, const int* sh_cols, const long long* sh_ld, float clip, float lr_t,
                             const float* lr_t_dev, float b1, float b2, float eps, float* partial, float* factor, float* norms,
                             int* flag, lpm_stream_t stream);
/* Start-of-step latch of the overflow flag shared by the optimiser entry points: if *flag is set, *skipped += 1 and
 * *flag = 0, so that one non-finite gradient norm skips exactly one update (the reference has no skip: TF would
 * write NaNs into the variables, utils.py:181-188). */
int lpm_step_begin(int* flag, int* skipped, lpm_stream_t stream);

/* Split-phase variant for ONE tensor whose rows are sharded over data-parallel ranks (hidden1_weights under DP:
 * every rank owns rows [r*Kd/W, (r+1)*Kd/W), updates them, and the fp16 operand shards are all-gathered).
 * p/g/m/v point at the shard; table = chunk table of the shard ({0, start/32, len, start/32}); wd1 = [1] regulariser.
 *   lpm_shard_sqnorm: sumsq[0] = sum (g + wd p)^2 over the shard   -> the caller all-reduces sumsq over the ranks
 *   lpm_shard_adam  : factor = clip / max(sqrt(sumsq), clip) (tf.clip_by_norm over the whole tensor), then Adam and
 *                     the fp16 shadow of the shard (sh_* as in lpm_adam_clip_step, one entry). */
int lpm_shard_sqnorm(const float* g, const float* p, const int* table, int n_chunks, const float* wd1, float* partial,
                     float* sumsq, lpm_stream_t stream);
int lpm_shard_adam(float* p, const float* g, float* m, float* v, const int* table, int n_chunks, const float* wd1,
                   const float* sumsq, float clip, float* factor, float* norm, int* flag,
                   const unsigned long long* sh_ptr, const int* sh_cols, const long long* sh_ld, float lr_t, float b1,
                   float b2, float eps, lpm_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Evaluation metrics on the device (the reference computes them with numpy on a host copy of the
 * predictions every logged step, train.py:448-449).
 *   lpm_eval_topk    : per video the k largest predictions, descending (ties -> lower class index):
 *                      eval_util.top_k_triplets (eval_util.py:128-135) as top_val/top_idx/top_lab [B][k]
 *                      (unused slots when V < k: -inf / -1 / 0), and row_stats [B][3] =
 *                      { hit@1 (eval_util.py:27-42), PERR (eval_util.py:45-70; 0 for a video without labels),
 *                        number of labels }.  pred fp32 [B][V] (row stride ld), labels uint8 [B][V] (stride ldl).
 *   lpm_eval_metrics : metrics[3] = { mean hit@1, mean PERR, GAP } with GAP = eval_util.calculate_gap
 *                      (eval_util.py:73-91 -> average_precision_calculator.py:203-262: all B*k triplets ranked by
 *                      score, average precision against the number of positives of the whole batch).
 * Limits: V <= 8192, B*k <= 16384.
 * ------------------------------------------------------------------------------------------- */
int lpm_eval_topk(const float* pred, long long ld, const unsigned char* labels, long long ldl, int B, int V, int k,
                  float* top_val, int* top_idx, unsigned char* top_lab, float* row_stats, lpm_stream_t stream);
int lpm_eval_metrics(const float* top_val, const unsigned char* top_lab, int B, int k, const float* row_stats,
                     float* metrics, lpm_stream_t stream);

/* Factored optimiser step for a dense layer whose weight gradient is the rank-R product dW = alpha * A^T G
 * (hidden1_weights, frame_level_models.py:2314-2319: A = VLAD descriptor fp16 [R][Kd], G = output gradient fp16
 * [R][N], R = tower batch <= 128).  dW is never materialised: lpm_rank_grad_clip derives the tf.clip_by_norm factor
 * (utils.py:181-188) from the Gram matrices gram_a = A A^T, gram_g = G G^T (fp32 [R][R]):
 * ||A^T G||_F^2 = sum_ij gram_a_ij gram_g_ij; lpm_rank_adam_step recomputes 128-row panels of dW on the tensor
 * cores and applies factor[0], Adam (same formulas as lpm_adam_clip_step, no regulariser) and the fp16 shadow
 * refresh (w16, row stride ldw16, may be null) in registers (persistent CTAs, cp.async-prefetched panels).  flag: the step is skipped when *flag != 0;
 * lpm_rank_grad_clip sets it on a non-finite norm. */
int lpm_rank_grad_clip(const float* gram_a, const float* gram_g, int R, float alpha, float clip, float* factor,
                       float* norm, int* flag, lpm_stream_t stream);
int lpm_rank_adam_step(const void* a16, long long lda, const void* g16, long long ldg, int R, long long Kd, int N,
                       float alpha, const float* factor, const int* flag, float* w, float* m, float* v, void* w16,
                       long long ldw16, float lr_t, float b1, float b2, float eps, void* workspace,
                       unsigned long long workspace_bytes, lpm_stream_t stream);
/* lpm_rank_adam_step with two options: lr_t_dev (non-null: the step size is read from device memory, see
 * lpm_adam_clip_step_dev) and tiled (non-zero: the same arithmetic, bit-identical, as 16-row CTAs of 128 threads and
 * 4 KB of shared memory instead of persistent 512-thread CTAs -- meant for a side stream underneath the
 * backward, where the small CTAs co-reside with the GEMM CTAs and fill idle SMs; tiled > 1 additionally splits the
 * N columns over that many CTAs per row block, i.e. shorter-lived CTAs). */
int lpm_rank_adam_step_ex(const void* a16, long long lda, const void* g16, long long ldg, int R, long long Kd, int N,
                          float alpha, const float* factor, const int* flag, float* w, float* m, float* v, void* w16,
                          long long ldw16, float lr_t, const float* lr_t_dev, int tiled, float b1, float b2, float eps,
                          void* workspace, unsigned long long workspace_bytes, lpm_stream_t stream);
/* bytes of caller-owned device workspace lpm_rank_adam_step needs (the column-permuted copy of G); 0 when the
 * (R, N) pair is not supported (G must fit in shared memory): callers then keep the dense-gradient path */
unsigned long long lpm_rank_adam_workspace_bytes(int R, int N);

/* =============================================================================================
 * NetVladV2 (attention-based cluster similarities) helpers
 * ============================================================================================= */
/* Batch-norm statistics of the attention logits q.k^T per key channel without materialising them
 * (transformer_utils.py:646-654): partial [B*H][2][L] = (sum_i q_i.k_j | sum_i (q_i.k_j)^2); head depth 16. */
int lpm_mha_logit_stats(const void* qkv, long long ld, int B, int L, int Dm, int H, float* partial, lpm_stream_t stream);
/* Per-column (sum | sum of squares) partials of an fp16 matrix: partial [lpm_colstats_chunks(rows, C)][2][C]
 * (attention_bn, filter_bn, feed_output_bn: transformer_utils.py:666,747,760). */
int lpm_colstats_chunks(long long rows, int C);
int lpm_colstats_f16(const void* x, long long ld, long long rows, int C, float* partial, lpm_stream_t stream);
/* y[r][c] = x[r][c]*scale[c] + shift[c] (applies a folded batch norm; y may alias x). */
int lpm_affine_cols_f16(const void* x, void* y, long long rows, int C, const float* scale, const float* shift,
                        lpm_stream_t stream);
/* slim.batch_norm backward over the rows of an fp16 matrix (training statistics).  stats: partial
 * [lpm_colstats_chunks(rows)][2][C] = (sum dy | sum dy*xhat), xhat = (x-p0)*p1 (mode 0: x = BN input, p0 = mean,
 * p1 = rstd) or (x-p0)/p1 (mode 1: x = BN output, p0 = beta, p1 = gamma).  apply: dx = gamma*rstd*(dy - c1/N -
 * xhat*c2/N) with csum = [2][C] totals, optionally masked by (x > 0) (ReLU in front of the batch norm).
 * dy is fp16 or fp32 (dy_f32); q (optional fp32 [rows/T][C]) is subtracted on the fly: dy_eff = dy - q[r/T][c]. */
int lpm_batchnorm_bwd_stats(const void* dy, int dy_f32, long long ld_dy, const float* q, int T, const void* x,
                            long long ld_x, long long rows, int C, const float* p0, const float* p1, int mode,
                            float* partial, lpm_stream_t stream);
int lpm_batchnorm_bwd_apply(const void* dy, int dy_f32, const float* q, int T, void* dx, const void* x, long long rows,
                            int C, const float* mean, const float* rstd, const float* gamma, const float* csum,
                            int relu, lpm_stream_t stream);
/* out[r][k] = fp16(G[r][k] - q[r/T][k]): gradient of the residual aggregation wrt externally supplied assignments. */
int lpm_sub_q_cast_f16(const float* G, const float* q, long long rows, int T, int K, void* out, lpm_stream_t stream);
/* out[b][k][d] = in[b*in_stride + d*K + k]: gradient of the d-major flatten back to the cluster-major layout. */
int lpm_dmajor_to_kmajor_f16(const void* in, long long in_stride, int B, int K, int D, void* out, lpm_stream_t stream);
/* Attention backward with batch-normed logits (transformer_utils.py:646-664): mode 1 writes per-key partial sums
 * [B*H][2][L] of (dl' | dl'*lhat); mode 2 takes their means m1/m2 [L] and writes dqkv. */
int lpm_mha_core_bwd_bn(int mode, const void* qkv, long long ld, const void* o, const void* dout, long long ldo,
                        const float* lse, int B, int L, int Dm, int H, const float* key_scale, const float* key_shift,
                        const float* key_mean, const float* key_rstd, const float* m1, const float* m2,
                        float* stat_partial, void* dqkv, long long ldd, lpm_stream_t stream);
/* tf.layers.dropout (transformer_utils.py:450): x *= keep/(1-rate); keep from mask_in (fp16 0/1) or a hash of
 * (seed + *seed_dev, index) -- seed_dev (device scalar, may be NULL) lets a captured CUDA graph draw a fresh mask on every
 * replay; the mask used is written to mask_out when given; out_f16 (may be NULL = in place) receives the result. */
int lpm_dropout_f16(void* x, long long n, const void* mask_in, void* mask_out, unsigned long long seed,
                    const unsigned long long* seed_dev, float rate, void* out_f16, lpm_stream_t stream);
/* out[b][d*K + k] = fp16(z[b][k][d]*rscale[b][k]): the reference's d-major flatten as an fp16 GEMM operand. */
int lpm_netvlad_finalize_f16(const void* z, const float* rscale, int B, int K, int D, void* out, long long out_stride,
                             lpm_stream_t stream);

/* =============================================================================================
 * Baseline NetVLAD (WillowModelReg / NetVladOrthoReg / LightVLAD)
 * ============================================================================================= */
/* Random frame sampling indices, int32 [B][T] (consumed by lpm_gather_bn_*):
 *   mode 0  SampleRandomFrames   (model_utils.py:54-73): idx = int32(u[b][i] * fl32(num_frames[b]))
 *   mode 1  SampleRandomSequence (model_utils.py:26-51): start = int32(u[b] * fl32(max(nf-T,0)+1)), idx = min(start+i, nf-1)
 * uniform: device fp32 [B][T] (mode 0) / [B] (mode 1) in [0,1), or NULL: a counter-based generator keyed by seed. */
int lpm_random_frame_index(const int* num_frames, const float* uniform, unsigned long long seed, int B, int T,
                           int max_frames, int mode, int* frame_index, lpm_stream_t stream);
/* lpm_sample_bn_stats / lpm_sample_bn_apply with explicit gather indices frame_index int32 [B*T] (tf.gather_nd,
 * model_utils.py:50-51,72-73).  x: fp32 frames (is_codes = 0) or the reader's uint8 codes (is_codes = 1:
 * dequantised with the given range and L2-normalised per frame, as in the *_u8 entry points). */
int lpm_gather_bn_stats(const void* x, int is_codes, float max_quantized_value, float min_quantized_value,
                        const int* frame_index, int B, int max_frames, int F, int T, float* partial,
                        lpm_stream_t stream);
int lpm_gather_bn_apply(const void* x, int is_codes, float max_quantized_value, float min_quantized_value,
                        const int* frame_index, int B, int max_frames, int F, int T, const float* scale,
                        const float* shift, void* y_f16, int split_col, void* y2_f16, lpm_stream_t stream);
/* Orthogonal regulariser on the cluster centres (module_utils.py:55-90; video_pooling_modules.py:1560-1568):
 *   value[0] = scale * sum_ij |(N^T N - I)_ij|,  N = l2_normalize(w [D][K], axis=1)   (fp32 throughout)
 *   dw [D][K] (optional) = (accumulate ? dw : 0) + grad_scale * d value / d w         (sign(0) = 0, as tf.abs)
 * workspace: lpm_ortho_reg_workspace_bytes(D, K) bytes of device memory. */
unsigned long long lpm_ortho_reg_workspace_bytes(int D, int K);
int lpm_ortho_reg(const float* w, int D, int K, float scale, float grad_scale, int accumulate, float* value,
                  float* dw, void* workspace, unsigned long long workspace_bytes, lpm_stream_t stream);

/* Host utility (no GPU work): CRC-32C (Castagnoli, reflected, init/xorout 0xFFFFFFFF) continued from `crc` (0 to start).
 * TensorFlow's tensor-bundle checkpoints store it (masked) per tensor and per index block; used by the checkpoint
 * reader/writer of learnablepoolingmethods_b200/checkpoint.py (train.py:415-423 saver, :390-411). */
unsigned int lpm_crc32c(unsigned int crc, const void* data, unsigned long long n);

/* Caller prelude of the model (train.py:262-264, eval.py:140-143, export_model.py:91-92): tf.nn.l2_normalize(model_input,
 * 2) on fp32 frames, y[r][:] = x[r][:] * rsqrt(max(sum x^2, 1e-12)); rows = B*max_frames; y may alias x.  (With uint8
 * input the *_u8 / lpm_gather_bn_* entry points do this on the fly.) */
int lpm_l2_normalize_rows(const float* x, long long rows, int F, float* y, lpm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LPM_B200_H_ */
