// Philox4x32-10 + Box-Muller, shared by the materialising reparam kernel and the tile-fused
// sampled GEMM so both see bit-identical noise for the same (seed, stream, element).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace blm {

// ------------------------------------------------------------------ Philox
// Philox4x32-10 (Salmon et al., SC'11).  counter = (idx_lo, idx_hi, stream_lo,
// stream_hi), key = (seed_lo, seed_hi).  One call yields four uniform words ->
// four N(0,1) values by two Box-Muller pairs, so element i of a tensor uses
// counter i/4, lane i%4: the noise is a pure function of (seed, stream, i) and
// therefore identical on every rank and for every launch geometry.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kM0, ctr.x), lo0 = kM0 * ctr.x;
    const uint32_t hi1 = __umulhi(kM1, ctr.z), lo1 = kM1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += kW0;
    key.y += kW1;
  }
  return ctr;
}

__device__ __forceinline__ float u32_to_unit_open(uint32_t u) {
  // (0, 1]: never 0 so the log below is finite
  return (static_cast<float>(u >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Four N(0,1) values from one Philox call: two Box-Muller pairs on MUFU intrinsics (lg2, sqrt, sin,
// cos = 2 special-function ops per normal -- the budget that bounds the tile-fused sampled GEMM).
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t stream, uint64_t idx4) {
  const uint4 r = philox4x32_10(
      make_uint4(static_cast<uint32_t>(idx4), static_cast<uint32_t>(idx4 >> 32),
                 static_cast<uint32_t>(stream), static_cast<uint32_t>(stream >> 32)),
      make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  constexpr float kTwoPi = 6.283185307179586f;
  constexpr float kNeg2Ln2 = -1.3862943611198906f;
  const float r0 = sqrt_approx(kNeg2Ln2 * __log2f(u32_to_unit_open(r.x)));
  const float r1 = sqrt_approx(kNeg2Ln2 * __log2f(u32_to_unit_open(r.z)));
  float s0, c0, s1, c1;
  __sincosf(kTwoPi * u32_to_unit_open(r.y), &s0, &c0);
  __sincosf(kTwoPi * u32_to_unit_open(r.w), &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// w = mu + sigma * eps with sigma = exp(lgstd): one definition so the fused and the materialising
// kernels round identically.
__device__ __forceinline__ float reparam_value(float mu, float lgstd, float eps) {
  return fmaf(__expf(lgstd), eps, mu);
}

}  // namespace blm
