// C-ABI layer of liblpm_b200.so: argument checks, thread-local error messages, TMA descriptor
// creation (driver entry point resolved at run time, so the library links without libcuda).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "lpm_common.cuh"
#include "lpm_kernels.h"

namespace lpm {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_3d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                 uint64_t batch, uint64_t row_stride_elems, uint64_t batch_stride_elems,
                 uint32_t box_cols, uint32_t box_rows, uint32_t swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(LPM_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if ((swizzle_bytes != 128 && swizzle_bytes != 64) || box_cols * elem_bytes != swizzle_bytes)
    return fail(LPM_ERR_ARG, "tmap: box inner extent must equal the swizzle span (64 or 128 bytes)");
  if (batch_stride_elems == 0) batch_stride_elems = rows * row_stride_elems;  // unused when batch == 1
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstride[2] = {row_stride_elems * (uint64_t)elem_bytes, batch_stride_elems * (uint64_t)elem_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(map, dt, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(LPM_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): base=%p cols=%llu rows=%llu batch=%llu rs=%llu bs=%llu box=%ux%u",
                (int)r, base, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)batch,
                (unsigned long long)row_stride_elems, (unsigned long long)batch_stride_elems, box_cols, box_rows);
  return LPM_OK;
}

int check_device() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(LPM_ERR_DEVICE, "no CUDA device");
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    ok = (major == 10) ? 1 : 0;
  }
  if (!ok) return fail(LPM_ERR_DEVICE, "liblpm_b200 requires an sm_100 (B200) device");
  return LPM_OK;
}

}  // namespace lpm

using namespace lpm;

static long long* g_pool_debug = nullptr;

extern "C" {

int lpm_version(void) { return 100; }

const char* lpm_last_error(void) { return g_last_error.c_str(); }

int lpm_gemm_f16(const lpm_gemm_desc* desc, lpm_stream_t stream) {
  if (!desc) return fail(LPM_ERR_ARG, "lpm_gemm_f16: null desc");
  if (int rc = check_device()) return rc;
  return gemm_f16(*desc, static_cast<cudaStream_t>(stream));
}

int lpm_gemm_splitk_gated_fwd(const lpm_gemm_desc* desc, const lpm_gating_tail* tail, lpm_stream_t stream) {
  if (!desc || !tail) return fail(LPM_ERR_ARG, "lpm_gemm_splitk_gated_fwd: null argument");
  if (int rc = check_device()) return rc;
  return gemm_f16(*desc, static_cast<cudaStream_t>(stream), tail);
}

int lpm_gemm_tile_n(int N) { return gemm_pick_bn(N); }

int lpm_gemm_splits(int K, int requested_splits) { return gemm_effective_splits(K, requested_splits); }
void lpm_debug_set_gemm_pair_mode(int mode) { gemm_set_pair_mode(mode); }
void lpm_debug_set_mha_tc_mode(int mode) { mha_set_tc_mode(mode); }
void lpm_debug_set_mha_clock(long long* buf) { mha_set_debug_clock(buf); }

#define ST(s) static_cast<cudaStream_t>(s)
#define H16(p) reinterpret_cast<__half*>(p)
#define CH16(p) reinterpret_cast<const __half*>(p)
#define DEVCHK() do { if (int rc_ = check_device()) return rc_; } while (0)

int lpm_splitk_reduce(const float* part, int splits, long long split_stride, long long n, int cols,
                      const float* bias, int relu, float alpha, int accumulate, float* out_f32,
                      void* out_f16, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(part && splits > 0 && n > 0 && cols > 0, "lpm_splitk_reduce: bad arguments");
  return splitk_reduce(part, splits, split_stride, nullptr, 0, 0, n, cols, bias, relu, alpha, accumulate, out_f32, H16(out_f16), 0,
                       ST(stream));
}

int lpm_splitk_reduce_ex(const float* part, int splits, long long split_stride, const float* part2, int splits2,
                         long long split_stride2, long long n, int cols, const float* bias, int relu, float alpha,
                         int accumulate, float* out_f32, void* out_f16, int split3, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(part && splits > 0 && n > 0 && cols > 0 && (part2 == nullptr || splits2 > 0), "lpm_splitk_reduce_ex: bad arguments");
  return splitk_reduce(part, splits, split_stride, part2, splits2, split_stride2, n, cols, bias, relu, alpha, accumulate, out_f32,
                       H16(out_f16), split3, ST(stream));
}

int lpm_sample_stats_blocks(void) { return sample_stats_blocks(); }

int lpm_sample_bn_stats(const float* x, const int* num_frames, int B, int max_frames, int F, int T,
                        float* partial, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && num_frames && partial && B > 0 && T > 0 && max_frames > 0, "lpm_sample_bn_stats: bad arguments");
  return sample_stats(x, 0, 0.f, 0.f, num_frames, nullptr, B, max_frames, F, T, partial, ST(stream));
}

int lpm_sample_bn_stats_u8(const unsigned char* codes, float max_quantized_value, float min_quantized_value,
                           const int* num_frames, int B, int max_frames, int F, int T, float* partial,
                           lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(codes && num_frames && partial && B > 0 && T > 0 && max_frames > 0, "lpm_sample_bn_stats_u8: bad arguments");
  return sample_stats(codes, 1, max_quantized_value, min_quantized_value, num_frames, nullptr, B, max_frames, F, T, partial, ST(stream));
}

int lpm_sample_bn_apply(const float* x, const int* num_frames, int B, int max_frames, int F, int T,
                        const float* scale, const float* shift, void* y_f16, int split_col, void* y2_f16,
                        lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && num_frames && scale && shift && y_f16 && B > 0 && T > 0, "lpm_sample_bn_apply: bad arguments");
  return sample_apply(x, 0, 0.f, 0.f, num_frames, nullptr, B, max_frames, F, T, scale, shift, H16(y_f16), split_col, H16(y2_f16), ST(stream));
}

int lpm_sample_bn_apply_u8(const unsigned char* codes, float max_quantized_value, float min_quantized_value,
                           const int* num_frames, int B, int max_frames, int F, int T, const float* scale,
                           const float* shift, void* y_f16, int split_col, void* y2_f16, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(codes && num_frames && scale && shift && y_f16 && B > 0 && T > 0, "lpm_sample_bn_apply_u8: bad arguments");
  return sample_apply(codes, 1, max_quantized_value, min_quantized_value, num_frames, nullptr, B, max_frames, F, T, scale, shift,
                      H16(y_f16), split_col, H16(y2_f16), ST(stream));
}

int lpm_batchnorm_finalize(const float* psum, const float* psq, int P, long long pstride, int C, double count,
                           const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                           float decay, float eps, int bessel, int training, float* scale, float* shift,
                           float* save_mean, float* save_rstd, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(scale && shift && C > 0, "lpm_batchnorm_finalize: bad arguments");
  LPM_REQUIRE(training ? (psum && psq && P > 0 && count > 0) : (moving_mean && moving_var),
              "lpm_batchnorm_finalize: training needs partial sums, inference needs moving statistics");
  return bn_finalize(psum, psq, P, pstride, C, count, gamma, beta, moving_mean, moving_var, decay, eps, bessel,
                     training, scale, shift, save_mean, save_rstd, ST(stream));
}

int lpm_netvlad_pool_fwd(const void* x, long long ldx, long long x_batch_stride, const void* wc, long long ldw,
                         const float* logit_scale, const float* logit_shift, const void* centers_t16,
                         const int* valid_frames, int B, int T, int D, int K, void* z, float* rscale,
                         float* a_sum, void* assign, const void* assign_in, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && centers_t16 && z && rscale, "lpm_netvlad_pool_fwd: null pointer");
  LPM_REQUIRE(assign_in || (wc && logit_scale && logit_shift), "lpm_netvlad_pool_fwd: need cluster weights or assign_in");
  return netvlad_pool_fwd(CH16(x), ldx, x_batch_stride, CH16(wc), ldw, logit_scale, logit_shift, CH16(centers_t16),
                          valid_frames, B, T, D, K, H16(z), rscale, a_sum, H16(assign), CH16(assign_in), g_pool_debug, ST(stream));
}

/* profiling aid: when set, lpm_netvlad_pool_fwd writes 8 clock64 phase stamps per video into this buffer */
void lpm_debug_set_pool_clock(long long* buf) { g_pool_debug = buf; }

int lpm_netvlad_finalize(const void* z, const float* rscale, int B, int K, int D, int d_major, float* out,
                         lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(z && rscale && out, "lpm_netvlad_finalize: null pointer");
  return vlad_finalize(CH16(z), rscale, B, K, D, d_major, out, ST(stream));
}

int lpm_mha_core_fwd(const void* qkv, long long ld, int B, int L, int Dm, int H, float scale,
                     const float* key_scale, const float* key_shift, void* out, long long ldo, float* lse,
                     lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(qkv && out && B > 0 && H > 0, "lpm_mha_core_fwd: bad arguments");
  return mha_fwd(CH16(qkv), ld, B, L, Dm, H, scale, key_scale, key_shift, H16(out), ldo, lse, ST(stream));
}

int lpm_scale_rows_f16(const void* x, const float* row_scale, long long rows, int D, void* y, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && row_scale && y && rows > 0, "lpm_scale_rows_f16: bad arguments");
  return scale_rows(CH16(x), row_scale, rows, D, H16(y), ST(stream));
}

int lpm_layernorm_joint_fwd(void* a, const void* b, const float* b_row_scale, void* u_out, int B, int rows, int D,
                            long long a_stride, long long b_stride, const float* gamma, const float* beta,
                            float eps, void* y, long long y_stride, float* partial, float* save_mean_rstd,
                            lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(a && gamma && beta && y && partial && B > 0 && rows > 0, "lpm_layernorm_joint_fwd: bad arguments");
  return layernorm_joint(H16(a), CH16(b), b_row_scale, H16(u_out), B, rows, D, a_stride, b_stride, gamma, beta, eps, H16(y),
                         y_stride, partial, save_mean_rstd, ST(stream));
}

int lpm_gating_fwd(const float* act, const float* g, int B, int H, const float* wg_diag, const float* gamma,
                   const float* beta, float* moving_mean, float* moving_var, float decay, float eps,
                   int training, float* out_f32, void* out_f16, float* save_mean, float* save_rstd,
                   lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(act && g && gamma && beta && moving_mean && moving_var && out_f32, "lpm_gating_fwd: null pointer");
  return gating_fwd(act, g, 1, 0, nullptr, 0, B, H, wg_diag, gamma, beta, moving_mean, moving_var, decay, eps, training, out_f32,
                    H16(out_f16), save_mean, save_rstd, ST(stream));
}

int lpm_gating_fwd_ex(const float* act, const float* g, int g_splits, long long g_split_stride, float* g_sum, int B, int H,
                      const float* wg_diag, const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                      float decay, float eps, int training, float* out_f32, void* out_f16, int split3, float* save_mean,
                      float* save_rstd, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(act && g && gamma && beta && moving_mean && moving_var && out_f32, "lpm_gating_fwd_ex: null pointer");
  LPM_REQUIRE(g_splits >= 1 && (g_splits == 1 || g_sum != nullptr), "lpm_gating_fwd_ex: split-K partials need g_sum");
  return gating_fwd(act, g, g_splits, g_split_stride, g_sum, split3, B, H, wg_diag, gamma, beta, moving_mean, moving_var, decay, eps,
                    training, out_f32, H16(out_f16), save_mean, save_rstd, ST(stream));
}

int lpm_moe_mix_fwd(const float* logits, long long ld, int B, int V, int M, int expert_off, float* pred,
                    lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(logits && pred && B > 0 && V > 0 && M > 0, "lpm_moe_mix_fwd: bad arguments");
  return moe_mix(logits, ld, B, V, M, expert_off, pred, ST(stream));
}

int lpm_xent_fwd(const float* pred, const uint8_t* labels, int B, int V, float* row_loss, float* loss,
                 lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(pred && labels && row_loss && loss, "lpm_xent_fwd: null pointer");
  return xent_loss(pred, labels, B, V, row_loss, loss, ST(stream));
}

int lpm_cast_f32_to_f16(const float* src, long long ld_src, int rows, int cols, void* dst, long long ld_dst,
                        int cols_dst, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(src && dst && rows > 0 && cols > 0 && cols_dst >= cols, "lpm_cast_f32_to_f16: bad arguments");
  return cast_2d(src, ld_src, rows, cols, H16(dst), ld_dst, cols_dst, ST(stream));
}

int lpm_transpose_f32(const float* src, int rows, int cols, float* dst, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(src && dst && rows > 0 && cols > 0, "lpm_transpose_f32: bad arguments");
  return transpose_2d(src, rows, cols, dst, nullptr, ST(stream));
}

int lpm_transpose_f32_dual(const float* src, int rows, int cols, float* dst32, void* dst16, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(src && (dst32 || dst16) && rows > 0 && cols > 0, "lpm_transpose_f32_dual: bad arguments");
  return transpose_2d(src, rows, cols, dst32, H16(dst16), ST(stream));
}

int lpm_xent_bwd(const float* pred, const uint8_t* labels, long long n, float gscale, float* dpred,
                 lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(pred && labels && dpred && n > 0, "lpm_xent_bwd: bad arguments");
  return xent_bwd(pred, labels, n, gscale, nullptr, dpred, ST(stream));
}
int lpm_xent_bwd_dev(const float* pred, const uint8_t* labels, long long n, float gscale, const float* upstream_dev,
                     float* dpred, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(pred && labels && dpred && upstream_dev && n > 0, "lpm_xent_bwd_dev: bad arguments");
  return xent_bwd(pred, labels, n, gscale, upstream_dev, dpred, ST(stream));
}
int lpm_moe_mix_bwd(const float* logits, long long ld, int B, int V, int M, int expert_off, const float* dpred,
                    float loss_scale, void* dlogits_f16, long long ldo, int ncols, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(logits && dpred && dlogits_f16, "lpm_moe_mix_bwd: null pointer");
  return moe_mix_bwd(logits, ld, B, V, M, expert_off, dpred, loss_scale, H16(dlogits_f16), ldo, ncols, ST(stream));
}
int lpm_colsum_chunks(long long rows) { return colsum_chunks(rows); }
int lpm_colsum(const void* x, int is_f32, long long ld, long long rows, int cols, float alpha, int accumulate,
               float* partial, float* out, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && partial && out && rows > 0 && cols > 0, "lpm_colsum: bad arguments");
  return colsum(x, is_f32, ld, rows, cols, alpha, accumulate, partial, out, ST(stream));
}
int lpm_colsum_final(const float* partial, int chunks, long long pstride, int cols, float alpha, int accumulate,
                     float* out, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(partial && out && chunks > 0 && cols > 0, "lpm_colsum_final: bad arguments");
  return colsum_final(partial, chunks, pstride, cols, alpha, accumulate, out, ST(stream));
}
int lpm_gating_bwd(const float* act, const float* g, int B, int H, const float* gamma, const float* beta,
                   const float* mean, const float* rstd, const float* dout, float inv_scale, float* dact,
                   void* dg_f16, float* dgamma, float* dbeta, const float* wg_diag, float* ddiag, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(act && g && gamma && beta && mean && rstd && dout && dact && dg_f16 && dgamma && dbeta, "lpm_gating_bwd: null pointer");
  return gating_bwd(act, g, B, H, gamma, beta, mean, rstd, dout, inv_scale, dact, H16(dg_f16), dgamma, dbeta, wg_diag, ddiag,
                    ST(stream));
}
int lpm_layernorm_bwd_chunks(void) { return ln_bwd_chunks(); }
int lpm_layernorm_joint_bwd(const void* u, const void* dy, long long dy_stride, int B, int rows, int D,
                            const float* mean_rstd, const float* gamma, const void* mask, void* du,
                            void* du_masked, float* part_sample, float* part_cols, float* part_cols_du,
                            lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(u && dy && mean_rstd && gamma && du && part_sample && part_cols, "lpm_layernorm_joint_bwd: null pointer");
  return layernorm_joint_bwd(CH16(u), CH16(dy), dy_stride, B, rows, D, mean_rstd, gamma, CH16(mask), H16(du),
                             H16(du_masked), part_sample, part_cols, part_cols_du, ST(stream));
}
int lpm_netvlad_norm_bwd(const void* z, const float* rscale, const void* dvhat, long long rows, int K, int D,
                         const float* centers_t, void* dz, float* q, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(z && rscale && dvhat && centers_t && dz && q, "lpm_netvlad_norm_bwd: null pointer");
  return vlad_norm_bwd(CH16(z), rscale, CH16(dvhat), rows, K, D, centers_t, H16(dz), q, ST(stream));
}
int lpm_assign_bwd_blocks(void) { return assign_bwd_blocks(); }
int lpm_assign_bwd1(const float* G, const void* assign, const float* q, const void* S, const float* mean,
                    const float* rstd, long long rows, int T, int K, void* dshat, float* partial,
                    lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(G && assign && q && S && mean && rstd && dshat && partial, "lpm_assign_bwd1: null pointer");
  return assign_bwd1(G, CH16(assign), q, CH16(S), mean, rstd, rows, T, K, H16(dshat), partial, ST(stream));
}
int lpm_assign_bwd2(void* dshat, const void* S, const float* mean, const float* rstd, const float* gamma,
                    const float* csum, long long rows, int K, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(dshat && S && mean && rstd && gamma && csum, "lpm_assign_bwd2: null pointer");
  return assign_bwd2(H16(dshat), CH16(S), mean, rstd, gamma, csum, rows, K, ST(stream));
}
int lpm_center_bwd(const void* dV, const void* Z, const float* a_sum, int B, int K, int D, const float* centers_t,
                   const float* beta_in, float inv_scale, float* dCt, float* E, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(dV && Z && a_sum && centers_t && beta_in && dCt && E, "lpm_center_bwd: null pointer");
  return center_bwd(CH16(dV), CH16(Z), a_sum, B, K, D, centers_t, beta_in, inv_scale, dCt, E, ST(stream));
}
int lpm_input_bn_grad(const float* Wc, const float* dWc, const float* dCt, const float* E, int D, int K,
                      const float* gamma_in, float* dgamma_in, float* dbeta_in, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(Wc && dWc && dCt && E && gamma_in && dgamma_in && dbeta_in, "lpm_input_bn_grad: null pointer");
  return input_bn_grad(Wc, dWc, dCt, E, D, K, gamma_in, dgamma_in, dbeta_in, ST(stream));
}
int lpm_split_hi_lo_f16(const float* src, long long ld_src, int rows, int cols, void* dst, long long ld_dst, int along_rows,
                        lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_src >= cols, "lpm_split_hi_lo_f16: bad arguments");
  LPM_REQUIRE(along_rows >= 0 && along_rows <= 2, "lpm_split_hi_lo_f16: along_rows must be 0, 1 or 2");
  LPM_REQUIRE(along_rows ? ld_dst >= cols : ld_dst >= 3ll * cols, "lpm_split_hi_lo_f16: destination row stride too small");
  return split_hi_lo(src, ld_src, rows, cols, H16(dst), ld_dst, along_rows, ST(stream));
}

int lpm_cast_scaled_f16(const float* x, long long n, float alpha, void* y, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && y && n > 0, "lpm_cast_scaled_f16: bad arguments");
  return cast_scaled(x, n, alpha, H16(y), ST(stream));
}
int lpm_mha_core_bwd(const void* qkv, long long ld, const void* o, const void* dout, long long ldo, const float* lse,
                     int B, int L, int Dm, int H, float scale, void* dqkv, long long ldd, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(qkv && o && dout && lse && dqkv, "lpm_mha_core_bwd: null pointer");
  return mha_bwd(CH16(qkv), ld, CH16(o), CH16(dout), ldo, lse, B, L, Dm, H, scale, H16(dqkv), ldd, ST(stream));
}

int lpm_adam_clip_step(float* p, const float* g, float* m, float* v, const int* table, int n_chunks,
                       const int* chunk_begin, int n_tensors, const float* wd, const unsigned long long* sh_ptr,
                       const int* sh_cols, const long long* sh_ld, float clip, float lr_t, float b1, float b2,
                       float eps, float* partial, float* factor, float* norms, int* flag, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(p && g && m && v && table && chunk_begin && wd && partial && factor && norms && flag && n_chunks > 0 && n_tensors > 0,
              "lpm_adam_clip_step: bad arguments");
  return adam_clip_step(p, g, m, v, table, n_chunks, chunk_begin, n_tensors, wd, sh_ptr, sh_cols, sh_ld, clip, lr_t, nullptr, b1, b2, eps,
                        partial, factor, norms, flag, ST(stream));
}

int lpm_adam_clip_step_dev(float* p, const float* g, float* m, float* v, const int* table, int n_chunks,
                           const int* chunk_begin, int n_tensors, const float* wd, const unsigned long long* sh_ptr,
                           const int* sh_cols, const long long* sh_ld, float clip, const float* lr_t_dev, float b1, float b2,
                           float eps, float* partial, float* factor, float* norms, int* flag, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(p && g && m && v && table && chunk_begin && wd && partial && factor && norms && flag && lr_t_dev && n_chunks > 0 &&
              n_tensors > 0, "lpm_adam_clip_step_dev: bad arguments");
  return adam_clip_step(p, g, m, v, table, n_chunks, chunk_begin, n_tensors, wd, sh_ptr, sh_cols, sh_ld, clip, 0.f, lr_t_dev, b1, b2, eps,
                        partial, factor, norms, flag, ST(stream));
}

int lpm_adam_clip_step_range(float* p, const float* g, float* m, float* v, const int* table, int chunk0, int n_chunks,
                             const int* chunk_begin, int tensor0, int n_tensors, const float* wd,
                             const unsigned long long* sh_ptr, const int* sh_cols, const long long* sh_ld, float clip, float lr_t,
                             const float* lr_t_dev, float b1, float b2, float eps, float* partial, float* factor, float* norms,
                             int* flag, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(p && g && m && v && table && chunk_begin && wd && partial && factor && norms && flag && n_chunks > 0 &&
              n_tensors > 0 && chunk0 >= 0 && tensor0 >= 0, "lpm_adam_clip_step_range: bad arguments");
  return adam_clip_step(p, g, m, v, table, n_chunks, chunk_begin, n_tensors, wd, sh_ptr, sh_cols, sh_ld, clip, lr_t, lr_t_dev, b1, b2,
                        eps, partial, factor, norms, flag, ST(stream), chunk0, tensor0);
}

int lpm_step_begin(int* flag, int* skipped, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(flag && skipped, "lpm_step_begin: null pointer");
  return step_begin(flag, skipped, ST(stream));
}

int lpm_shard_sqnorm(const float* g, const float* p, const int* table, int n_chunks, const float* wd1, float* partial,
                     float* sumsq, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(g && p && table && wd1 && partial && sumsq && n_chunks > 0, "lpm_shard_sqnorm: bad arguments");
  return shard_sqnorm(g, p, table, n_chunks, wd1, partial, sumsq, ST(stream));
}

int lpm_shard_adam(float* p, const float* g, float* m, float* v, const int* table, int n_chunks, const float* wd1,
                   const float* sumsq, float clip, float* factor, float* norm, int* flag,
                   const unsigned long long* sh_ptr, const int* sh_cols, const long long* sh_ld, float lr_t, float b1,
                   float b2, float eps, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(p && g && m && v && table && wd1 && sumsq && factor && norm && flag && n_chunks > 0, "lpm_shard_adam: bad arguments");
  return shard_adam(p, g, m, v, table, n_chunks, wd1, sumsq, clip, factor, norm, flag, sh_ptr, sh_cols, sh_ld, lr_t, b1, b2,
                    eps, ST(stream));
}

int lpm_eval_topk(const float* pred, long long ld, const unsigned char* labels, long long ldl, int B, int V, int k,
                  float* top_val, int* top_idx, unsigned char* top_lab, float* row_stats, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(pred && labels && top_val && top_idx && top_lab && row_stats, "lpm_eval_topk: null pointer");
  return eval_topk(pred, ld, labels, ldl, B, V, k, top_val, top_idx, top_lab, row_stats, ST(stream));
}

int lpm_eval_metrics(const float* top_val, const unsigned char* top_lab, int B, int k, const float* row_stats,
                     float* metrics, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(top_val && top_lab && row_stats && metrics, "lpm_eval_metrics: null pointer");
  return eval_metrics(top_val, top_lab, B, k, row_stats, metrics, ST(stream));
}

int lpm_layernorm_chain_supported(int rows, int D) { return layernorm_chain_supported(rows, D); }

int lpm_layernorm_chain_fwd(const void* a, long long a_stride, const void* b, long long b_stride, const float* b_row_scale,
                            int B, int rows, int D, float eps, const float* gamma1, const float* beta1, void* u1_out,
                            long long u1_stride, float* stats1, const float* gamma2, const float* beta2, void* u2_out,
                            long long u2_stride, float* stats2, void* y, long long y_stride, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(a && b && y && gamma1 && beta1 && B > 0, "lpm_layernorm_chain_fwd: bad arguments");
  return layernorm_chain_fwd(CH16(a), a_stride, CH16(b), b_stride, b_row_scale, B, rows, D, eps, gamma1, beta1, H16(u1_out),
                             u1_stride, stats1, gamma2, beta2, H16(u2_out), u2_stride, stats2, H16(y), y_stride, nullptr, ST(stream));
}

int lpm_layernorm_chain_fwd_split(const void* a, long long a_stride, const void* b, long long b_stride, const float* b_row_scale,
                                  int B, int rows, int D, float eps, const float* gamma1, const float* beta1, void* u1_out,
                                  long long u1_stride, float* stats1, const float* gamma2, const float* beta2, void* u2_out,
                                  long long u2_stride, float* stats2, void* y, long long y_stride, void* y_lo,
                                  lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(a && b && y && y_lo && gamma1 && beta1 && B > 0, "lpm_layernorm_chain_fwd_split: bad arguments");
  return layernorm_chain_fwd(CH16(a), a_stride, CH16(b), b_stride, b_row_scale, B, rows, D, eps, gamma1, beta1, H16(u1_out),
                             u1_stride, stats1, gamma2, beta2, H16(u2_out), u2_stride, stats2, H16(y), y_stride, H16(y_lo), ST(stream));
}

int lpm_rank_grad_clip(const float* gram_a, const float* gram_g, int R, float alpha, float clip, float* factor,
                       float* norm, int* flag, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(gram_a && gram_g && factor && norm && flag && R > 0, "lpm_rank_grad_clip: bad arguments");
  return rank_grad_clip(gram_a, gram_g, R, alpha, clip, factor, norm, flag, ST(stream));
}

int lpm_rank_adam_step(const void* a16, long long lda, const void* g16, long long ldg, int R, long long Kd, int N,
                       float alpha, const float* factor, const int* flag, float* w, float* m, float* v, void* w16,
                       long long ldw16, float lr_t, float b1, float b2, float eps, void* workspace,
                       unsigned long long workspace_bytes, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(a16 && g16 && factor && flag && w && m && v && Kd > 0, "lpm_rank_adam_step: bad arguments");
  return rank_adam_step(CH16(a16), lda, CH16(g16), ldg, R, Kd, N, alpha, factor, flag, w, m, v, H16(w16), ldw16, lr_t, nullptr, 0,
                        b1, b2, eps, workspace, (size_t)workspace_bytes, ST(stream));
}

int lpm_rank_adam_step_ex(const void* a16, long long lda, const void* g16, long long ldg, int R, long long Kd, int N,
                          float alpha, const float* factor, const int* flag, float* w, float* m, float* v, void* w16,
                          long long ldw16, float lr_t, const float* lr_t_dev, int tiled, float b1, float b2, float eps,
                          void* workspace, unsigned long long workspace_bytes, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(a16 && g16 && factor && flag && w && m && v && Kd > 0, "lpm_rank_adam_step_ex: bad arguments");
  return rank_adam_step(CH16(a16), lda, CH16(g16), ldg, R, Kd, N, alpha, factor, flag, w, m, v, H16(w16), ldw16, lr_t, lr_t_dev,
                        tiled, b1, b2, eps, workspace, (size_t)workspace_bytes, ST(stream));
}

unsigned long long lpm_rank_adam_workspace_bytes(int R, int N) { return rank_adam_workspace_bytes(R, N); }

int lpm_mha_logit_stats(const void* qkv, long long ld, int B, int L, int Dm, int H, float* partial, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(qkv && partial && B > 0 && L > 0, "lpm_mha_logit_stats: bad arguments");
  return mha_logit_stats(CH16(qkv), ld, B, L, Dm, H, partial, ST(stream));
}
int lpm_colstats_chunks(long long rows, int C) { return colstats_chunks(rows, C); }
int lpm_colstats_f16(const void* x, long long ld, long long rows, int C, float* partial, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && partial && rows > 0 && C > 0, "lpm_colstats_f16: bad arguments");
  return colstats(CH16(x), ld, rows, C, partial, ST(stream));
}
int lpm_affine_cols_f16(const void* x, void* y, long long rows, int C, const float* scale, const float* shift,
                        lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && y && scale && shift && rows > 0, "lpm_affine_cols_f16: bad arguments");
  return affine_cols(CH16(x), H16(y), rows, C, scale, shift, ST(stream));
}
int lpm_batchnorm_bwd_stats(const void* dy, int dy_f32, long long ld_dy, const float* q, int T, const void* x,
                            long long ld_x, long long rows, int C, const float* p0, const float* p1, int mode,
                            float* partial, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(dy && x && p0 && p1 && partial && rows > 0, "lpm_batchnorm_bwd_stats: bad arguments");
  return bn_bwd_stats(dy, dy_f32, ld_dy, q, T, CH16(x), ld_x, rows, C, p0, p1, mode, partial, ST(stream));
}
int lpm_batchnorm_bwd_apply(const void* dy, int dy_f32, const float* q, int T, void* dx, const void* x, long long rows,
                            int C, const float* mean, const float* rstd, const float* gamma, const float* csum,
                            int relu, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(dy && dx && x && mean && rstd && gamma && csum, "lpm_batchnorm_bwd_apply: null pointer");
  return bn_bwd_apply(dy, dy_f32, q, T, H16(dx), CH16(x), rows, C, mean, rstd, gamma, csum, relu, ST(stream));
}
int lpm_sub_q_cast_f16(const float* G, const float* q, long long rows, int T, int K, void* out, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(G && q && out && rows > 0, "lpm_sub_q_cast_f16: bad arguments");
  return sub_q_cast(G, q, rows, T, K, H16(out), ST(stream));
}
int lpm_dmajor_to_kmajor_f16(const void* in, long long in_stride, int B, int K, int D, void* out, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(in && out, "lpm_dmajor_to_kmajor_f16: null pointer");
  return dmajor_to_kmajor_f16(CH16(in), in_stride, B, K, D, H16(out), ST(stream));
}
int lpm_mha_core_bwd_bn(int mode, const void* qkv, long long ld, const void* o, const void* dout, long long ldo,
                        const float* lse, int B, int L, int Dm, int H, const float* key_scale, const float* key_shift,
                        const float* key_mean, const float* key_rstd, const float* m1, const float* m2,
                        float* stat_partial, void* dqkv, long long ldd, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(qkv && o && dout && lse && key_scale && key_shift && key_mean && key_rstd, "lpm_mha_core_bwd_bn: null pointer");
  LPM_REQUIRE(mode == 1 ? stat_partial != nullptr : (m1 && m2 && dqkv), "lpm_mha_core_bwd_bn: mode 1 needs stat_partial, mode 2 needs m1/m2/dqkv");
  return mha_bwd_bn(mode, CH16(qkv), ld, CH16(o), CH16(dout), ldo, lse, B, L, Dm, H, key_scale, key_shift, key_mean,
                    key_rstd, m1, m2, stat_partial, H16(dqkv), ldd, ST(stream));
}
int lpm_dropout_f16(void* x, long long n, const void* mask_in, void* mask_out, unsigned long long seed,
                    const unsigned long long* seed_dev, float rate, void* out_f16, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && n > 0, "lpm_dropout_f16: bad arguments");
  return dropout_f16(H16(x), H16(out_f16), n, CH16(mask_in), H16(mask_out), seed, seed_dev, rate, ST(stream));
}
int lpm_netvlad_finalize_f16(const void* z, const float* rscale, int B, int K, int D, void* out, long long out_stride,
                             lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(z && rscale && out, "lpm_netvlad_finalize_f16: null pointer");
  return vlad_dmajor_f16(CH16(z), rscale, B, K, D, H16(out), out_stride, ST(stream));
}

int lpm_random_frame_index(const int* num_frames, const float* uniform, unsigned long long seed, int B, int T,
                           int max_frames, int mode, int* frame_index, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(num_frames && frame_index && B > 0 && T > 0 && max_frames > 0, "lpm_random_frame_index: bad arguments");
  return random_frame_index(num_frames, uniform, seed, B, T, max_frames, mode, frame_index, ST(stream));
}
int lpm_gather_bn_stats(const void* x, int is_codes, float max_quantized_value, float min_quantized_value,
                        const int* frame_index, int B, int max_frames, int F, int T, float* partial,
                        lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && frame_index && partial && B > 0 && T > 0 && max_frames > 0, "lpm_gather_bn_stats: bad arguments");
  return sample_stats(x, is_codes != 0, max_quantized_value, min_quantized_value, nullptr, frame_index, B, max_frames, F, T,
                      partial, ST(stream));
}
int lpm_gather_bn_apply(const void* x, int is_codes, float max_quantized_value, float min_quantized_value,
                        const int* frame_index, int B, int max_frames, int F, int T, const float* scale,
                        const float* shift, void* y_f16, int split_col, void* y2_f16, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && frame_index && scale && shift && y_f16 && B > 0 && T > 0, "lpm_gather_bn_apply: bad arguments");
  return sample_apply(x, is_codes != 0, max_quantized_value, min_quantized_value, nullptr, frame_index, B, max_frames, F, T,
                      scale, shift, H16(y_f16), split_col, H16(y2_f16), ST(stream));
}
unsigned long long lpm_ortho_reg_workspace_bytes(int D, int K) { return ortho_reg_workspace_bytes(D, K); }
int lpm_ortho_reg(const float* w, int D, int K, float scale, float grad_scale, int accumulate, float* value,
                  float* dw, void* workspace, unsigned long long workspace_bytes, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(w && workspace && (value || dw), "lpm_ortho_reg: null pointer");
  return ortho_reg(w, D, K, scale, grad_scale, accumulate, value, dw, static_cast<float*>(workspace), workspace_bytes,
                   ST(stream));
}

// ---- host utility for checkpoint interchange: CRC-32C (Castagnoli), slicing-by-8 ----
static uint32_t g_crc_tab[8][256];
static std::once_flag g_crc_once;
static void crc_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xff];
}
unsigned int lpm_crc32c(unsigned int crc, const void* data, unsigned long long n) {
  std::call_once(g_crc_once, crc_init);
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8); --n; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = g_crc_tab[7][w & 0xff] ^ g_crc_tab[6][(w >> 8) & 0xff] ^ g_crc_tab[5][(w >> 16) & 0xff] ^ g_crc_tab[4][(w >> 24) & 0xff] ^
        g_crc_tab[3][(w >> 32) & 0xff] ^ g_crc_tab[2][(w >> 40) & 0xff] ^ g_crc_tab[1][(w >> 48) & 0xff] ^ g_crc_tab[0][w >> 56];
    p += 8; n -= 8;
  }
  while (n--) c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return ~c;
}

int lpm_hidden_bn_relu6_fwd(const float* x, int B, int H, const float* gamma, const float* beta, float* moving_mean,
                            float* moving_var, float decay, float eps, int training, int relu6, float* out_f32,
                            void* out_f16, float* save_mean, float* save_rstd, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && gamma && beta && moving_mean && moving_var && out_f32 && B > 0 && H > 0, "lpm_hidden_bn_relu6_fwd: bad arguments");
  return hidden_bn_relu6_fwd(x, B, H, gamma, beta, moving_mean, moving_var, decay, eps, training, relu6, out_f32, H16(out_f16),
                             save_mean, save_rstd, ST(stream));
}
int lpm_hidden_bn_relu6_bwd(const float* x, const float* y, const float* dy, int B, int H, const float* gamma,
                            const float* mean, const float* rstd, int relu6, float inv_scale, float* dx, float* dgamma,
                            float* dbeta, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && y && dy && gamma && mean && rstd && dx && dgamma && dbeta, "lpm_hidden_bn_relu6_bwd: null pointer");
  return hidden_bn_relu6_bwd(x, y, dy, B, H, gamma, mean, rstd, relu6, inv_scale, dx, dgamma, dbeta, ST(stream));
}
int lpm_add_diag(float* m, int n, long long ld, const float* d, float alpha, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(m && d && n > 0 && ld >= n, "lpm_add_diag: bad arguments");
  return add_diag(m, n, ld, d, alpha, ST(stream));
}

int lpm_l2_normalize_rows(const float* x, long long rows, int F, float* y, lpm_stream_t stream) {
  DEVCHK();
  LPM_REQUIRE(x && y && rows > 0, "lpm_l2_normalize_rows: bad arguments");
  return l2_normalize_rows(x, rows, F, y, ST(stream));
}

}  // extern "C"
