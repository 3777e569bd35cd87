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
 *   stat_sum/stat_sq (optional): per-row sum / sum of squares of the epilogue values over the N
 *   columns of each N-tile, written to [batch][n_tiles][M] (deterministic partials).
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
} lpm_gemm_desc;

int lpm_gemm_f16(const lpm_gemm_desc* desc, lpm_stream_t stream);
/* N-tile width the kernel will use for a given N (for sizing stat partial buffers). */
int lpm_gemm_tile_n(int N);
/* Number of non-empty K splits the kernel will use. */
int lpm_gemm_splits(int K, int requested_splits);

#ifdef __cplusplus
}
#endif
#endif /* LPM_B200_H_ */
