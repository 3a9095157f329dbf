// Host-side helpers shared by the C-ABI translation units: error reporting,
// argument checks and the TMA tensor-map encoder (driver entry point fetched at
// run time so the library links against cudart only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bayeslm_b200.h"

namespace blm {

void set_error(const char* fmt, ...);
int num_sms();

#define BLM_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::blm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                     \
      return BLM_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define BLM_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::blm::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Encode a 2-D bf16 row-major [rows, cols] tensor (leading dimension ld elements)
// with a [box_rows x 64] box and 128-byte swizzle.  Out-of-bounds elements read
// as zero, which is what makes ragged M / N / K edges correct.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                     int box_rows);

int encode_tmap_bf16_box32(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld);
int encode_tmap_f32(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

// generate-once sampled weights inside a blm_gemm launch (blm_gemm.cu: generate_weights)
struct GemmGen {
  const void* mu;      // bf16 [N, K], leading dimension ldmu
  int64_t ldmu;
  const void* sigma;   // bf16 [N, K] dense
  const float* eps;    // fp32 [N, K] dense, or null for Philox(seed, stream_id)
  const float* mu32;   // optional fp32 mean [N, K] (ld ldmu32) + dense fp32 lgstd: replaces (mu, sigma)
  int64_t ldmu32;
  const float* lgstd32;
  uint64_t seed, stream_id;
  void* wt;            // bf16 [N, K] dense scratch = B operand of the GEMM
  unsigned int* sync;  // two zeroed counters
};
int gemm_impl(const blm_gemm_desc* d, const GemmGen* gen, blm_stream stream);

inline cudaStream_t as_stream(blm_stream s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace blm
