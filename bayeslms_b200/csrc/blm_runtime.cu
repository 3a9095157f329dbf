// Library runtime: error string, device check, TMA tensor-map encoding.
#include <stdarg.h>
#include <string.h>

#include "blm_host.h"

namespace blm {

static thread_local char g_err[512] = "";
static int g_num_sms = 0;
static int g_device = -1;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() { return g_num_sms; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return BLM_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
    return BLM_ERR_CUDA;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return BLM_OK;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                     int box_rows) {
  int rc = load_encode();
  if (rc != BLM_OK) return rc;
  BLM_REQUIRE(aligned16(base), BLM_ERR_ALIGN, "tensor base %p is not 16-byte aligned", base);
  BLM_REQUIRE((ld % 8) == 0 && ld >= cols, BLM_ERR_ALIGN,
              "leading dimension %lld must be a multiple of 8 and >= cols %lld", (long long)ld,
              (long long)cols);
  BLM_REQUIRE(rows > 0 && cols > 0 && box_rows > 0 && box_rows <= 256, BLM_ERR_SHAPE,
              "bad tensor-map shape rows=%lld cols=%lld box_rows=%d", (long long)rows,
              (long long)cols, box_rows);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2u};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld box=%d)",
              (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows);
    return BLM_ERR_CUDA;
  }
  return BLM_OK;
}

// fp32 row-major [rows, cols] tensor (leading dimension ld elements), [box_rows x 32] box = 128-byte inner
// extent with 128-byte swizzle: the residual-in / fp32-out tiles of the LayerNorm-fused GEMM epilogue.
int encode_tmap_f32(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  int rc = load_encode();
  if (rc != BLM_OK) return rc;
  BLM_REQUIRE(aligned16(base), BLM_ERR_ALIGN, "tensor base %p is not 16-byte aligned", base);
  BLM_REQUIRE((ld % 4) == 0 && ld >= cols, BLM_ERR_ALIGN,
              "leading dimension %lld must be a multiple of 4 and >= cols %lld", (long long)ld, (long long)cols);
  BLM_REQUIRE(rows > 0 && cols > 0 && box_rows > 0 && box_rows <= 256, BLM_ERR_SHAPE,
              "bad tensor-map shape rows=%lld cols=%lld box_rows=%d", (long long)rows, (long long)cols, box_rows);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4u};
  cuuint32_t box[2] = {32u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d (rows=%lld cols=%lld ld=%lld box=%d)", (int)r,
              (long long)rows, (long long)cols, (long long)ld, box_rows);
    return BLM_ERR_CUDA;
  }
  return BLM_OK;
}

// bf16 row-major [rows, cols] output tensor, [32 rows x 32 cols] box = 64-byte inner extent with 64-byte swizzle:
// the 2 KB per-warp staging tiles of the 16-epilogue-warp GEMM variant.
int encode_tmap_bf16_box32(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  int rc = load_encode();
  if (rc != BLM_OK) return rc;
  BLM_REQUIRE(aligned16(base) && (ld % 8) == 0 && ld >= cols && rows > 0 && cols > 0, BLM_ERR_ALIGN,
              "bad bf16 tensor for a 32 x 32 box map (ld=%lld cols=%lld)", (long long)ld, (long long)cols);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2u};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (32 x 32 bf16 box) failed with CUresult %d", (int)r);
    return BLM_ERR_CUDA;
  }
  return BLM_OK;
}

int gemm_init();  // blm_gemm.cu: raise dynamic shared-memory limits
int lstm_init();  // blm_lstm.cu
int gemm_sampled_init();  // blm_gemm_sampled.cu
int gemm2_init();         // blm_gemm2.cu
int gemm_ln_init();       // blm_gemm_ln.cu

}  // namespace blm

extern "C" {

int blm_version(void) { return 100; }

const char* blm_last_error(void) { return blm::g_err; }

int blm_num_sms(void) { return blm::g_num_sms; }

int blm_init(int device) {
  using namespace blm;
  int count = 0;
  BLM_CHECK_CUDA(cudaGetDeviceCount(&count));
  BLM_REQUIRE(device >= 0 && device < count, BLM_ERR_ARG, "device %d out of range (%d visible)",
              device, count);
  cudaDeviceProp prop;
  BLM_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  BLM_REQUIRE(prop.major == 10 && prop.minor == 0, BLM_ERR_ARCH,
              "bayeslm_b200 only runs on sm_100 (B200); device %d is sm_%d%d (%s)", device,
              prop.major, prop.minor, prop.name);
  BLM_CHECK_CUDA(cudaSetDevice(device));
  g_num_sms = prop.multiProcessorCount;
  g_device = device;
  int rc = load_encode();
  if (rc != BLM_OK) return rc;
  rc = gemm_init();
  if (rc != BLM_OK) return rc;
  rc = lstm_init();
  if (rc != BLM_OK) return rc;
  rc = gemm_sampled_init();
  if (rc != BLM_OK) return rc;
  rc = gemm_ln_init();
  if (rc != BLM_OK) return rc;
  rc = gemm2_init();
  if (rc != BLM_OK) return rc;
  return BLM_OK;
}

}  // extern "C"
