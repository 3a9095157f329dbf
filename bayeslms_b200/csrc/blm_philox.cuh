// Philox4x32-10 + Box-Muller, shared by the materialising reparam kernel and the tile-fused
// sampled GEMM so both see bit-identical noise for the same (seed, stream, element).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace blm {

// ------------------------------------------------------------------ Philox
// Philox4x32-10 (Salmon et al., SC'11).  counter = (idx_lo, idx_hi, stream_lo,
// stream_hi), key = (seed_lo, seed_hi).  One call yields four uniform words ->
// eight N(0,1) values (below), so element i of a tensor uses counter i/8, lane
// i%8: the noise is a pure function of (seed, stream, i) and therefore identical
// on every rank and for every launch geometry.
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

// Eight N(0,1) values from ONE Philox call: each 32-bit word is split into two 16-bit uniforms that
// drive one Box-Muller pair -- radius from u1 = (k + 0.5) / 65536 (|z| <= 4.66, P(|z| > 4.66) = 3e-6),
// angle from (j + 0.5) / 65536 of a turn.  Four MUFU ops (lg2, sqrt, sin, cos) and ~12 ALU ops per pair;
// halving the Philox rounds per normal is what lets the tile-fused sampled GEMM keep its generator
// warps below the tensor pipe's time per K block.  Element i of a tensor uses counter i/8, lane i%8.
struct Normal8 {
  float v[8];
};

__device__ __forceinline__ void box_muller16(uint32_t w, float& z0, float& z1) {
  constexpr float kInv = 1.0f / 65536.0f;
  const float u1 = (static_cast<float>(w & 0xffffu) + 0.5f) * kInv;   // (0, 1)
  const float turn = (static_cast<float>(w >> 16) + 0.5f) * kInv;     // (0, 1) of a full turn
  float r, sn, cs;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r * -1.3862943611198906f));  // sqrt(-2 ln u1)
  const float ang = turn * 6.283185307179586f;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
  z0 = r * cs;
  z1 = r * sn;
}

__device__ __forceinline__ Normal8 philox_normal8(uint64_t seed, uint64_t stream, uint64_t idx8) {
  const uint4 r = philox4x32_10(
      make_uint4(static_cast<uint32_t>(idx8), static_cast<uint32_t>(idx8 >> 32),
                 static_cast<uint32_t>(stream), static_cast<uint32_t>(stream >> 32)),
      make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  Normal8 o;
  box_muller16(r.x, o.v[0], o.v[1]);
  box_muller16(r.y, o.v[2], o.v[3]);
  box_muller16(r.z, o.v[4], o.v[5]);
  box_muller16(r.w, o.v[6], o.v[7]);
  return o;
}

// Four consecutive normals of the stream: elements [4 idx4, 4 idx4 + 4) = half of counter idx4 / 2.
// (The elementwise kernels walk tensors four elements at a time.)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t stream, uint64_t idx4) {
  const uint4 r = philox4x32_10(
      make_uint4(static_cast<uint32_t>(idx4 >> 1), static_cast<uint32_t>(idx4 >> 33),
                 static_cast<uint32_t>(stream), static_cast<uint32_t>(stream >> 32)),
      make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  float4 o;
  if (idx4 & 1u) {
    box_muller16(r.z, o.x, o.y);
    box_muller16(r.w, o.z, o.w);
  } else {
    box_muller16(r.x, o.x, o.y);
    box_muller16(r.y, o.z, o.w);
  }
  return o;
}

// ------------------------------------------------------------------ dropout multipliers
// Element i: word (i & 3) of Philox counter (i >> 2); kept (multiplier 1/(1-p)) iff word >= thresh = floor(p 2^32).
struct DropParams {
  const float* mask;          // explicit multipliers (parity), or null
  const uint64_t* seed_dev;   // device word added to the key at kernel entry (CUDA-graph replays), or null
  uint64_t seed, stream;
  uint32_t thresh;
  float scale;                // 1 / (1 - p)
};

// host: C-ABI descriptor -> kernel parameters
inline DropParams make_drop_params(const float* mask, float p, uint64_t seed, const uint64_t* seed_dev, uint64_t stream) {
  DropParams d;
  d.mask = mask;
  d.seed_dev = seed_dev;
  d.seed = seed;
  d.stream = stream;
  const double t = static_cast<double>(p) * 4294967296.0;
  d.thresh = t >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(t);
  d.scale = 1.0f / (1.0f - p);
  return d;
}

__device__ __forceinline__ DropParams drop_resolve(DropParams d) {
  if (d.seed_dev) d.seed += *d.seed_dev;
  return d;
}

__device__ __forceinline__ uint4 philox_words4(uint64_t seed, uint64_t stream, uint64_t idx4) {
  return philox4x32_10(make_uint4(static_cast<uint32_t>(idx4), static_cast<uint32_t>(idx4 >> 32),
                                  static_cast<uint32_t>(stream), static_cast<uint32_t>(stream >> 32)),
                       make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
}

// multipliers of elements idx, idx + 1 (idx even, both valid) / of element idx
__device__ __forceinline__ float2 drop_mult2(const DropParams& d, long long idx) {
  if (d.mask) return make_float2(__ldg(d.mask + idx), __ldg(d.mask + idx + 1));
  const uint4 w = philox_words4(d.seed, d.stream, static_cast<uint64_t>(idx) >> 2);
  const uint32_t a = (idx & 2) ? w.z : w.x, b = (idx & 2) ? w.w : w.y;
  return make_float2(a >= d.thresh ? d.scale : 0.0f, b >= d.thresh ? d.scale : 0.0f);
}
__device__ __forceinline__ float drop_mult1(const DropParams& d, long long idx) {
  if (d.mask) return __ldg(d.mask + idx);
  const uint4 w = philox_words4(d.seed, d.stream, static_cast<uint64_t>(idx) >> 2);
  const uint32_t q = static_cast<uint32_t>(idx) & 3u;
  const uint32_t a = q == 0 ? w.x : q == 1 ? w.y : q == 2 ? w.z : w.w;
  return a >= d.thresh ? d.scale : 0.0f;
}

// Keep bits of elements [base, base + n) of a dropout site (base % 4 == 0, n % 4 == 0) into shared memory, one bit per
// element, by all `nthreads` threads of the CTA (ends with __syncthreads).  With an explicit mask *kept receives the bits
// of the multiplier of a kept element (every kept element of a mask carries the same value 1 / (1 - p)); with Philox the
// multiplier is d.scale.  The attention kernels read their multipliers from these bits: regenerating Philox words inside
// the MMA loops cost 60 registers and half the occupancy of the backward kernel (151 -> 351 us per step).
__device__ __forceinline__ void drop_keep_bits(const DropParams& d, long long base, int n, uint32_t* bits, unsigned* kept,
                                               int tid, int nthreads) {
  if (d.mask) {
    if (tid == 0) *kept = 0u;
    __syncthreads();
  }
  unsigned vmax = 0u;
  for (int w = tid; w * 32 < n; w += nthreads) {
    uint32_t word = 0u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = w * 32 + q * 4;
      if (e < n) {
        uint32_t b;
        if (d.mask) {
          const float2 m0 = __ldg(reinterpret_cast<const float2*>(d.mask + base + e));
          const float2 m1 = __ldg(reinterpret_cast<const float2*>(d.mask + base + e + 2));
          b = (m0.x != 0.0f ? 1u : 0u) | (m0.y != 0.0f ? 2u : 0u) | (m1.x != 0.0f ? 4u : 0u) | (m1.y != 0.0f ? 8u : 0u);
          vmax = max(vmax, __float_as_uint(fmaxf(fmaxf(m0.x, m0.y), fmaxf(m1.x, m1.y))));
        } else {
          const uint4 r = philox_words4(d.seed, d.stream, static_cast<uint64_t>(base + e) >> 2);
          b = (r.x >= d.thresh ? 1u : 0u) | (r.y >= d.thresh ? 2u : 0u) | (r.z >= d.thresh ? 4u : 0u) |
              (r.w >= d.thresh ? 8u : 0u);
        }
        word |= b << (q * 4);
      }
    }
    bits[w] = word;
  }
  if (d.mask && vmax) atomicMax(kept, vmax);
  __syncthreads();
}
// multipliers of elements idx, idx + 1 (idx even) / of element idx, relative to `base` above
__device__ __forceinline__ float2 keep_mult2(const uint32_t* bits, float scale, int idx) {
  const uint32_t b = bits[idx >> 5] >> (idx & 31);
  return make_float2((b & 1u) ? scale : 0.0f, (b & 2u) ? scale : 0.0f);
}
__device__ __forceinline__ float keep_mult1(const uint32_t* bits, float scale, int idx) {
  return ((bits[idx >> 5] >> (idx & 31)) & 1u) ? scale : 0.0f;
}

// w = mu + sigma * eps with sigma = exp(lgstd): one definition so the fused and the materialising
// kernels round identically.
__device__ __forceinline__ float reparam_value(float mu, float lgstd, float eps) {
  return fmaf(__expf(lgstd), eps, mu);
}

}  // namespace blm
