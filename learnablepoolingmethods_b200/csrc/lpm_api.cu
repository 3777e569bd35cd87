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
                 uint32_t box_cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(LPM_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if (box_cols * elem_bytes != 128) return fail(LPM_ERR_ARG, "tmap: box inner extent must be 128 bytes");
  if (batch_stride_elems == 0) batch_stride_elems = rows * row_stride_elems;  // unused when batch == 1
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstride[2] = {row_stride_elems * (uint64_t)elem_bytes, batch_stride_elems * (uint64_t)elem_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(map, dt, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

extern "C" {

int lpm_version(void) { return 100; }

const char* lpm_last_error(void) { return g_last_error.c_str(); }

int lpm_gemm_f16(const lpm_gemm_desc* desc, lpm_stream_t stream) {
  if (!desc) return fail(LPM_ERR_ARG, "lpm_gemm_f16: null desc");
  if (int rc = check_device()) return rc;
  return gemm_f16(*desc, static_cast<cudaStream_t>(stream));
}

int lpm_gemm_tile_n(int N) { return gemm_pick_bn(N); }

int lpm_gemm_splits(int K, int requested_splits) { return gemm_effective_splits(K, requested_splits); }

}  // extern "C"
