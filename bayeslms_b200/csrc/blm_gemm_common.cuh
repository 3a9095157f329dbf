// Pieces shared by the tcgen05 GEMM kernels: tile constants, kernel parameters, the fused
// epilogue math (bias / q-scale / GELU / GP mixture / residual / hi-lo split) and its helpers.
#pragma once
#include <cuda_fp16.h>

#include "blm_host.h"
#include "blm_philox.cuh"
#include "blm_ptx.cuh"

namespace blm {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kEpiWarp0 = 4;  // warps 0-3: TMA producer, MMA issuer, TMEM allocator, spare

enum { EPI_STORE = 0, EPI_NLL = 1 };

struct GemmParams {
  CUtensorMap tmA[BLM_MAX_SEG];
  CUtensorMap tmB[BLM_MAX_SEG];
  CUtensorMap tmC;      // STG == 2: bf16 output [M, N], box 64 columns x 32 rows, 128B swizzle (TMA store)
  int kblocks[BLM_MAX_SEG];
  int nseg;
  int M, N;
  int m_tiles, n_tiles;
  // work decomposition: work w -> (m_tile = w / n_groups, group = w % n_groups),
  // n tiles [group * tiles_per_group, min(n_tiles, (group+1) * tiles_per_group))
  int n_groups, tiles_per_group, num_works;
  // EPI_STORE
  const float* bias;
  const float* coef;
  float col_scale;
  int col_scale_cols;
  const float* resid;
  long long ldr;
  float* out_f32;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  long long ldc;
  int a_mn, b_mn;       // 1: the operand is MN-major: A[s] is [K_s, M] / B[s] is [K_s, N] row-major (wgrad / dgrad
                        // straight from the activations / weights, no transposed copies); boxes of 64 x 64
  int fast_act;         // 1: packed-fp16 evaluation of the activation derivative (GELU_GRAD on the transposed path)
  int x3;               // chunked kernels: segments come in (A.hi, B.lo), (A.lo, B.hi), (A.hi, B.hi) triples of one K range;
                        // each ring stage then carries the FOUR distinct tiles once and feeds all three products
  int a_f16;            // 1: A operands are fp16 (instruction descriptor A format F16, B stays BF16)
  int use_stg;          // 1: transpose finished chunks through shared memory for row-coalesced stores
  int f32_rows32;       // 1: out_f32 is stored in 32-row blocks (blm_gemm_desc.f32_rows32): row-per-thread stores are
                        // then 512 contiguous bytes per warp instruction, and so are the loads of a consumer that
                        // owns one row per thread (the LSTM gate math on tensor-memory lanes)
  int chunk_kb;         // CHUNK kernels: K blocks per TMEM accumulation chunk
  float* out_pre;       // fp32 copy of the value BEFORE the activation (training: saved pre-activation)
  const float* aux;     // [M, N] fp32, leading dimension ldaux: the pre-activation the *_GRAD epilogues differentiate at
  long long ldaux;
  // BLM_ACT_SOFTMAX_GRAD: dZ = (softmax - onehot) * grad_scale, softmax = exp(z - lse[m])
  const float* lse;
  float grad_scale;
  // EPI_NLL (and the softmax-grad epilogue)
  const int* targets;
  int seg_tiles;    // vocabulary sweep: tiles per canonical segment (the online LSE restarts at every segment boundary)
  float* part_max;  // [n_segments * column groups, M]  (part_tgt: [n_groups * column groups, M])
  float* part_sum;
  float* part_tgt;
  // generate-once mode of blm_gemm_sampled (gen_wt != null): before the first B tile is loaded every CTA builds
  // its share of W~ = bf16(mu + sigma eps) into gen_wt -- the dense [N, gen_K] bf16 tensor tmB[0] points at, 4 MB
  // for the FFN weight, L2 resident -- and the TMA producers wait on a grid-wide arrival counter.
  const __nv_bfloat16* gen_mu;     // [N, gen_K] bf16, leading dimension gen_ldmu
  long long gen_ldmu;
  const __nv_bfloat16* gen_sigma;  // [N, gen_K] bf16, dense
  const float* gen_eps;            // [N, gen_K] fp32 dense (BLM_EPS_PTR) or null (Philox)
  const float* gen_mu32;           // optional fp32 sources [N, gen_K] (ld gen_ldmu32) / lgstd dense: W~ is then
  long long gen_ldmu32;            // bf16(mu + exp(lgstd) eps) from fp32, bit-identical to blm_reparam's bf16 output
  const float* gen_lgstd32;
  unsigned long long gen_seed, gen_stream;
  int gen_K;
  __nv_bfloat16* gen_wt;
  unsigned int* gen_sync;          // [0] arrivals, [1] departures; zero between launches
  // dropout fused into the storing epilogue (blm_gemm_desc.drop): the multiplier of element (m, col) is element
  // m * N + col of the site's mask / Philox stream, applied to the activated value before the residual is added
  // (forward) or to the incoming gradient before the activation derivative and out_pre (the *_GRAD epilogues)
  int drop_on;
  DropParams drop;
};

// ARES > 0: the A operand of a work item (ARES K blocks of 128 x 64) stays resident in shared
// memory for the whole sweep over its N tiles and only B streams through the ring.
template <int BN, int STAGES, int ARES = 0>
struct SmemLayout {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kResBytes = ARES * kABytes;
  static constexpr int kStageBytes = (ARES ? 0 : kABytes) + kBBytes;
  static constexpr int kBiasOffset = kResBytes + STAGES * kStageBytes;  // fp32 bias[2][BN], one per accumulator stage
  // 4 KB per epilogue warp (8 warps) for the store transpose; the A-resident kernels (NLL) store nothing
  static constexpr int kStgOffset = kBiasOffset + 2 * BN * 4;
  // (the A-resident storing variant with TWO ring stages keeps 32 KB of TMA-store staging; the three-stage NLL variant stores nothing)
  static constexpr int kStgBytes = (ARES && STAGES > 2) ? 0 : 8 * 4096;
  static constexpr int kBarOffset = kStgOffset + kStgBytes;
  // full[STAGES] empty[STAGES] tmem_full[2] tmem_empty[2] a_full a_empty + tmem ptr
  static constexpr int kBytes = kBarOffset + (2 * STAGES + 6) * 8 + 16;
  // the dynamic shared-memory window of a kernel without static shared memory starts 1024-B aligned
  // (checked at kernel entry), so no alignment slack is reserved: the A-resident NLL variant needs
  // 224 KB of tiles and would not fit with it
  static constexpr int kDynBytes = kBytes;
};

// barrier among the epilogue warps only (named barrier 1; barrier 0 is __syncthreads)
__device__ __forceinline__ void epi_bar_sync(int threads) {
  asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exact-erf GELU to 4e-7 absolute: erf(t) = 1 - 2^(-t q(t)) with q a degree-6 fit of -log2(erfc(t))/t
// on [0, 4] (erfc(4) = 1.5e-8, so t is clamped there); one MUFU.EX2 and nine FMAs per element
// instead of libdevice erff -- the FFN1 epilogue has ~16 issue slots per element before it, not
// the tensor pipe, becomes the bound at K = 512.
__device__ __forceinline__ float gelu_fast(float x) {
  // GELU(x) = relu(x) - |x|/2 * erfc(|x|/sqrt 2): one formula for both signs, no select
  const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float q = fmaf(t, -1.002195230e-04f, 4.615629764e-04f);
  q = fmaf(q, t, 2.302262028e-03f);
  q = fmaf(q, t, -2.945254180e-02f);
  q = fmaf(q, t, 1.489636837e-01f);
  q = fmaf(q, t, 9.183286407e-01f);
  q = fmaf(q, t, 1.627913732e+00f);
  const float e = ex2_approx(-(q * t));  // erfc(t)
  return fmaf(fabsf(x) * -0.5f, e, fmaxf(x, 0.0f));
}

// The same formula on two elements at a time in packed fp16 (HFMA2 / HMNMX2 / ex2.approx.f16x2): ~8 issue
// slots per element instead of ~15.  fp16's 11-bit mantissa costs <= 5e-4 relative on the result -- below the
// bf16 rounding (4e-3) of the only output this variant is used for (BLM_ACT_GELU_FAST: bf16-hi output of the
// fast mode), where the fp32 GELU made the FFN1 epilogue, not the tensor pipe, the bound.
__device__ __forceinline__ __half2 gelu_fast_core_h2(__half2 x, __half2 relu_x) {
  // GELU(x) = max(x, 0) - |x| 2^(-(p(s) s) - 1), s = min(|x|, 4 sqrt 2), p(s) = q(s / sqrt 2) / sqrt 2 with q the
  // degree-3 fit of -log2(erfc(t)) / t (weighted by the sensitivity |x|/2 erfc(t) ln2 t of the result: 1.2e-5
  // absolute on GELU in exact arithmetic, far below the fp16 evaluation itself).  The 1/sqrt 2 scaling and the
  // factor 1/2 live in the coefficients / the exponent: 9 packed instructions per pair + 2 MUFU.
  const __half2 s = __hmin2(__habs2(x), __float2half2_rn(5.65685f));
  __half2 p = __hfma2(s, __float2half2_rn(-4.0813875e-03f), __float2half2_rn(4.5319763e-02f));
  p = __hfma2(p, s, __float2half2_rn(4.6557564e-01f));
  p = __hfma2(p, s, __float2half2_rn(1.1492719e+00f));
  const __half2 ex = __hfma2(__hneg2(p), s, __float2half2_rn(-1.0f));   // -(p s) - 1
  // ex2.approx.f16x2 straight from PTX: two MUFU.EX2.F16 + one PRMT.  The h2exp2() intrinsic expands to seven
  // instructions (half -> float, MUFU.EX2, an FFMA rounding fix-up, float -> half) -- measured in the SASS
  uint32_t e_bits;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(e_bits) : "r"(*reinterpret_cast<const uint32_t*>(&ex)));
  const __half2 e = *reinterpret_cast<const __half2*>(&e_bits);          // erfc(|x| / sqrt 2) / 2
  return __hfma2(__hneg2(__habs2(x)), e, relu_x);
}

__device__ __forceinline__ void gelu_fast_h2(float& a, float& b) {
  const __half2 x = __floats2half2_rn(a, b);
  const float2 f = __half22float2(gelu_fast_core_h2(x, __hmax2(x, __float2half2_rn(0.0f))));
  a = f.x;
  b = f.y;
}

__device__ __forceinline__ __half2 tanh_h2(__half2 x) {
  uint32_t r;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(*reinterpret_cast<const uint32_t*>(&x)));   // 2 MUFU.TANH.F16 + PRMT
  return *reinterpret_cast<const __half2*>(&r);
}

// GP activation mixture sum_i coef[i, n] act_i(z), acts tanh, sigmoid, relu, gelu (model.py:1893-1899, 2263), on
// the column pair (n, n + 1) in packed fp16: sigmoid(z) = 0.5 tanh(z / 2) + 0.5, both tanh through MUFU.TANH.F16.
// The fp32 evaluation costs ~45 issue slots per element (706 us for the GP layer's FFN1 vs 273 us with GELU).
__device__ __forceinline__ void gpmix_fast_h2(float& a, float& b, const float* __restrict__ coef, int N, int n) {
  n = min(n, N - 2);   // the TMA-store path evaluates (and then clips) columns past the ragged edge; N is even
  const __half2 x = __floats2half2_rn(a, b);
  const __half2 half = __float2half2_rn(0.5f);
  const __half2 th = tanh_h2(x);
  const __half2 sg = __hfma2(tanh_h2(__hmul2(x, half)), half, half);
  const __half2 rl = __hmax2(x, __float2half2_rn(0.0f));
  const __half2 gl = gelu_fast_core_h2(x, rl);
  const float2 c0 = __ldg(reinterpret_cast<const float2*>(coef + n));
  const float2 c1 = __ldg(reinterpret_cast<const float2*>(coef + N + n));
  const float2 c2 = __ldg(reinterpret_cast<const float2*>(coef + 2 * N + n));
  const float2 c3 = __ldg(reinterpret_cast<const float2*>(coef + 3 * N + n));
  __half2 r = __hmul2(__floats2half2_rn(c0.x, c0.y), th);
  r = __hfma2(__floats2half2_rn(c1.x, c1.y), sg, r);
  r = __hfma2(__floats2half2_rn(c2.x, c2.y), rl, r);
  r = __hfma2(__floats2half2_rn(c3.x, c3.y), gl, r);
  const float2 f = __half22float2(r);
  a = f.x;
  b = f.y;
}

// d/dz of the exact-erf GELU: Phi(z) + z phi(z), with erfc from the same fit as gelu_fast
__device__ __forceinline__ float gelu_grad(float z) {
  const float t = fminf(fabsf(z) * 0.70710678118654752440f, 4.0f);
  float q = fmaf(t, -1.002195230e-04f, 4.615629764e-04f);
  q = fmaf(q, t, 2.302262028e-03f);
  q = fmaf(q, t, -2.945254180e-02f);
  q = fmaf(q, t, 1.489636837e-01f);
  q = fmaf(q, t, 9.183286407e-01f);
  q = fmaf(q, t, 1.627913732e+00f);
  const float half_erfc = 0.5f * ex2_approx(-(q * t));
  const float Phi = z >= 0.0f ? 1.0f - half_erfc : half_erfc;
  const float phi = 0.3989422804014327f * ex2_approx(z * z * -0.7213475204444817f);
  return fmaf(z, phi, Phi);
}

// The same derivative for two elements in packed fp16 (fast mode: <= 1e-3 absolute): E = erfc(|z| / sqrt 2) / 2 from
// the exponent trick of gelu_fast_core_h2, Phi = 1/2 + copysign(1/2 - E, z), phi through a second ex2.approx.f16x2.
__device__ __forceinline__ void gelu_grad_h2(float z0, float z1, float& g0, float& g1) {
  const __half2 x = __floats2half2_rn(z0, z1);
  const __half2 s = __hmin2(__habs2(x), __float2half2_rn(5.65685f));
  __half2 p = __hfma2(s, __float2half2_rn(-4.0813875e-03f), __float2half2_rn(4.5319763e-02f));
  p = __hfma2(p, s, __float2half2_rn(4.6557564e-01f));
  p = __hfma2(p, s, __float2half2_rn(1.1492719e+00f));
  const __half2 ex = __hfma2(__hneg2(p), s, __float2half2_rn(-1.0f));
  const __half2 ex2 = __hmul2(__hmul2(x, x), __float2half2_rn(-0.7213475f));        // -z^2 / 2 in log2 units
  uint32_t e_bits, f_bits;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(e_bits) : "r"(*reinterpret_cast<const uint32_t*>(&ex)));
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(f_bits) : "r"(*reinterpret_cast<const uint32_t*>(&ex2)));
  const __half2 E = *reinterpret_cast<const __half2*>(&e_bits);
  const __half2 half = __float2half2_rn(0.5f);
  __half2 dlt = __hsub2(half, E);                                                   // >= 0
  uint32_t d_bits = *reinterpret_cast<const uint32_t*>(&dlt) | (*reinterpret_cast<const uint32_t*>(&x) & 0x80008000u);
  const __half2 Phi = __hadd2(half, *reinterpret_cast<const __half2*>(&d_bits));
  const __half2 g = __hfma2(__hmul2(x, __float2half2_rn(0.3989423f)), *reinterpret_cast<const __half2*>(&f_bits), Phi);
  const float2 f = __half22float2(g);
  g0 = f.x;
  g1 = f.y;
}

// d/dz of sum_i coef[i, n] act_i(z), acts tanh, sigmoid, relu, gelu (model.py:1893-1899)
__device__ __forceinline__ float gpmix_grad(float z, const float* __restrict__ coef, int N, int n) {
  const float c0 = __ldg(coef + n), c1 = __ldg(coef + N + n), c2 = __ldg(coef + 2 * N + n), c3 = __ldg(coef + 3 * N + n);
  const float th = tanhf(z), sg = 1.0f / (1.0f + expf(-z));
  return c0 * (1.0f - th * th) + c1 * sg * (1.0f - sg) + (z > 0.0f ? c2 : 0.0f) + c3 * gelu_grad(z);
}

template <int ACT>
__device__ __forceinline__ float apply_act(float z, const float* __restrict__ coef, int N, int n) {
  if constexpr (ACT == BLM_ACT_GPMIX_FAST) {
    return z;   // (never reached: the packed variant only runs on the TMA-store path, which takes no ragged-edge branch)
  } else if constexpr (ACT == BLM_ACT_GELU || ACT == BLM_ACT_GELU_FAST) {
    return gelu_fast(z);   // (the packed-fp16 variant is applied pairwise in store_chunk; this is its ragged-edge path)
  } else if constexpr (ACT == BLM_ACT_GPMIX) {
    n = min(n, N - 1);  // the TMA-store path evaluates (and then clips) columns past the ragged edge
    const float c0 = __ldg(coef + n), c1 = __ldg(coef + N + n), c2 = __ldg(coef + 2 * N + n),
                c3 = __ldg(coef + 3 * N + n);
    // sigmoid and tanh from ONE exponential: e = exp(-z) (z clamped to +-15, where both have saturated to
    // 3e-7), sigmoid = 1/(1+e), tanh = (1-e^2)/(1+e^2); ex2 / rcp are MUFU approximations (<= 2 ulp)
    const float zc = fminf(fmaxf(z, -15.0f), 15.0f);
    const float e1 = ex2_approx(zc * -1.4426950408889634f);
    const float e2 = e1 * e1;
    const float sg = rcp_approx(1.0f + e1);
    const float th = (1.0f - e2) * rcp_approx(1.0f + e2);
    return fmaf(c0, th, fmaf(c1, sg, fmaf(c2, fmaxf(z, 0.0f), c3 * gelu_fast(z))));
  } else {
    return z;
  }
}

// ---- vocabulary NLL epilogue state (shared by the 1-CTA and 2-CTA kernels) -------------------------------
struct NllState {
  float run_max;  // log2 domain: fl(max logit * log2 e)
  float run_sum, tgt_logit;
  int tgt;
};

// online log-sum-exp over one 32-column chunk of logits (natural-log units; exponentials via ex2)
// sb: the chunk's 32 bias values in shared memory (zero where there is no bias / past column N)
__device__ __forceinline__ void nll_chunk(const GemmParams& p, float (&v)[32], int col0, NllState& st,
                                          const float* sb) {
  constexpr float kLog2e = 1.4426950408889634f;
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bb = *reinterpret_cast<const float4*>(sb + j);
      v[j] += bb.x;
      v[j + 1] += bb.y;
      v[j + 2] += bb.z;
      v[j + 3] += bb.w;
    }
  }
  if (col0 + 32 > p.N) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j >= p.N) v[j] = -INFINITY;
  }
  const unsigned int rel = static_cast<unsigned int>(st.tgt - col0);
  if (rel < 32u) {  // the target column lives in this chunk: once per row per sweep
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (rel == static_cast<unsigned int>(j)) st.tgt_logit = v[j];
  }
  float m0 = fmaxf(v[0], v[1]), m1 = fmaxf(v[2], v[3]), m2 = fmaxf(v[4], v[5]), m3 = fmaxf(v[6], v[7]);
#pragma unroll
  for (int j = 8; j < 32; j += 8) {
    m0 = fmaxf(m0, fmaxf(v[j], v[j + 1]));
    m1 = fmaxf(m1, fmaxf(v[j + 2], v[j + 3]));
    m2 = fmaxf(m2, fmaxf(v[j + 4], v[j + 5]));
    m3 = fmaxf(m3, fmaxf(v[j + 6], v[j + 7]));
  }
  // running maximum kept in the log2 domain as the ROUNDED product max * log2(e): every term and
  // every rescale is then measured against exactly the same power of two (a rescale by the
  // unchanged maximum is exactly 1, so nothing compounds over the ~1000 chunks of a row)
  const float new_m2 = fmaxf(st.run_max, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * kLog2e);
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    s0 += ex2_approx(fmaf(v[j], kLog2e, -new_m2));
    s1 += ex2_approx(fmaf(v[j + 1], kLog2e, -new_m2));
    s2 += ex2_approx(fmaf(v[j + 2], kLog2e, -new_m2));
    s3 += ex2_approx(fmaf(v[j + 3], kLog2e, -new_m2));
  }
  st.run_sum = st.run_sum * ex2_approx(st.run_max - new_m2) + ((s0 + s1) + (s2 + s3));
  st.run_max = new_m2;
}

// ---- epilogue building blocks: one accumulator row per thread, 32 columns per chunk ----------

// Dropout multipliers of the 32 consecutive elements (m, col0 ..) of a dense [M, N] tensor, N % 32 == 0: eight Philox
// counters (or eight float4 loads of an explicit mask).
__device__ __forceinline__ void drop_chunk(const GemmParams& p, float (&v)[32], int m, int col0) {
  const long long base = static_cast<long long>(m) * p.N + col0;
  if (p.drop.mask) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 k = __ldg(reinterpret_cast<const float4*>(p.drop.mask + base + j));
      v[j] *= k.x;
      v[j + 1] *= k.y;
      v[j + 2] *= k.z;
      v[j + 3] *= k.w;
    }
    return;
  }
  const uint64_t seed = p.drop.seed + (p.drop.seed_dev ? __ldg(reinterpret_cast<const unsigned long long*>(p.drop.seed_dev)) : 0ull);
  const uint32_t th = p.drop.thresh;
  const float sc = p.drop.scale;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 w = philox_words4(seed, p.drop.stream, static_cast<uint64_t>(base >> 2) + q);
    v[4 * q] *= w.x >= th ? sc : 0.0f;
    v[4 * q + 1] *= w.y >= th ? sc : 0.0f;
    v[4 * q + 2] *= w.z >= th ? sc : 0.0f;
    v[4 * q + 3] *= w.w >= th ? sc : 0.0f;
  }
}

// Element offset of (row m, column col) of the fp32 output.  rows32 layout: [ceil(M / 32)][N / 4][32 rows][4 floats].
__device__ __forceinline__ long long f32_off(const GemmParams& p, int m, int col) {
  if (!p.f32_rows32) return static_cast<long long>(m) * p.ldc + col;
  return ((static_cast<long long>(m >> 5) * (p.ldc >> 2) + (col >> 2)) * 32 + (m & 31)) * 4 + (col & 3);
}

// bias / q-scale / activation / residual / (hi, lo) split / stores for one 32-column chunk.
// Called by all 32 lanes of an epilogue warp; thread `lane` holds row m = m0 + lane (row_ok = m < M).
//   sb : this chunk's 32 bias values staged in shared memory (null: read p.bias from global)
//   stg: 4 KB of shared memory private to the warp.  The tcgen05.ld layout gives every thread one
//        ROW, so direct stores touch 32 different 128-byte lines per instruction (16 bytes each) and
//        the LSU / L2 request rate, not bandwidth, bounds the epilogue.  With stg the finished chunk
//        is transposed through shared memory (16-byte slots XOR-swizzled by row, conflict free both
//        ways) so that 8 lanes cover one row: every global access -- residual load, fp32 / bf16 hi /
//        bf16 lo stores -- is a full 128-byte (64-byte for bf16) row segment, 4 rows per instruction.
// STG selects the transposed-store path at compile time (one path per kernel keeps the epilogue
// under the register cap: with both inlined every storing kernel spilled its loop state).
template <int ACT, int STG = 0>
__device__ __forceinline__ void store_chunk(const GemmParams& p, float (&v)[32], int m, bool row_ok, int lane,
                                            int col0, const float* sb = nullptr, float4* stg = nullptr) {
  // STG == 2: arithmetic only -- the caller packs the chunk to bf16 and hands it to a TMA store, which clips
  // the tensor edges itself, so every chunk takes the vector path (staged bias is zero past column N)
  const bool full = (STG == 2) || (col0 + 32 <= p.N);
  if (full) {
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bb = sb ? *reinterpret_cast<const float4*>(sb + j)
                             : __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        v[j] += bb.x;
        v[j + 1] += bb.y;
        v[j + 2] += bb.z;
        v[j + 3] += bb.w;
      }
    }
    if (col0 + 32 <= p.col_scale_cols) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.col_scale;
    } else if (col0 < p.col_scale_cols) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.col_scale_cols) v[j] *= p.col_scale;
    }
    constexpr bool kGradAct = ACT == BLM_ACT_GELU_GRAD || ACT == BLM_ACT_GPMIX_GRAD;
    if constexpr (kGradAct && STG != 2) {
      if (p.drop_on && row_ok) drop_chunk(p, v, m, col0);   // h = mask . act(z): the mask multiplies dL/dh first
    }
    if (p.out_pre && row_ok) {
      float* o = p.out_pre + static_cast<long long>(m) * p.ldc + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if constexpr (ACT == BLM_ACT_SOFTMAX_GRAD) {
      const float nl = row_ok ? -__ldg(p.lse + m) * 1.4426950408889634f : 0.0f;
      const int rel = row_ok ? __ldg(p.targets + m) - col0 : -1;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = (ex2_approx(fmaf(v[j], 1.4426950408889634f, nl)) - (j == rel ? 1.0f : 0.0f)) * p.grad_scale;
    } else if constexpr (ACT == BLM_ACT_GELU_GRAD && STG == 1) {
      // applied after the transpose below, where the saved pre-activation is read with row-coalesced loads
    } else if constexpr (ACT == BLM_ACT_GELU_GRAD || ACT == BLM_ACT_GPMIX_GRAD) {
      if (row_ok) {
        const float* a = p.aux + static_cast<long long>(m) * p.ldaux + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 z = __ldg(reinterpret_cast<const float4*>(a + j));
          const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            v[j + q] *= (ACT == BLM_ACT_GELU_GRAD) ? gelu_grad(zz[q]) : gpmix_grad(zz[q], p.coef, p.N, col0 + j + q);
        }
      }
    } else if constexpr (ACT == BLM_ACT_GELU_FAST) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) gelu_fast_h2(v[j], v[j + 1]);
    } else if constexpr (ACT == BLM_ACT_GPMIX_FAST) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) gpmix_fast_h2(v[j], v[j + 1], p.coef, p.N, col0 + j);
    } else if constexpr (ACT != BLM_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act<ACT>(v[j], p.coef, p.N, col0 + j);
    }
    if constexpr (!kGradAct && ACT != BLM_ACT_SOFTMAX_GRAD && STG != 2) {   // (the TMA-store kernels are inference-only)
      if (p.drop_on && row_ok) drop_chunk(p, v, m, col0);   // dropout(act(z)), then the residual
    }
    if constexpr (STG == 2) return;
    if constexpr (STG == 1) {
      // ---- transpose through shared memory, then row-coalesced residual + stores
#pragma unroll
      for (int q = 0; q < 8; ++q)
        stg[lane * 8 + (q ^ (lane & 7))] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      __syncwarp();
      const int c4 = lane & 7, rsub = lane >> 3;
      const int m0 = m - lane;
      const int col = col0 + c4 * 4;
      // residual loads are issued four rows ahead of their use (4 x 128 bit in flight per lane)
#pragma unroll
      for (int k0 = 0; k0 < 8; k0 += 4) {
        float4 rr[4];
        if (p.resid) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int mr = m0 + rsub + 4 * (k0 + k);
            rr[k] = mr < p.M ? __ldg(reinterpret_cast<const float4*>(p.resid + static_cast<long long>(mr) * p.ldr + col))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float4 za[4];
        if constexpr (ACT == BLM_ACT_GELU_GRAD) {   // the saved pre-activation of these four rows, coalesced
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int mr = m0 + rsub + 4 * (k0 + k);
            za[k] = mr < p.M ? __ldg(reinterpret_cast<const float4*>(p.aux + static_cast<long long>(mr) * p.ldaux + col))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int row = rsub + 4 * (k0 + k);
          const int mr = m0 + row;
          float4 x = stg[row * 8 + (c4 ^ (row & 7))];
          if constexpr (ACT == BLM_ACT_GELU_GRAD) {
            if (p.fast_act) {
              float g0, g1, g2, g3;
              gelu_grad_h2(za[k].x, za[k].y, g0, g1);
              gelu_grad_h2(za[k].z, za[k].w, g2, g3);
              x.x *= g0;
              x.y *= g1;
              x.z *= g2;
              x.w *= g3;
            } else {
              x.x *= gelu_grad(za[k].x);
              x.y *= gelu_grad(za[k].y);
              x.z *= gelu_grad(za[k].z);
              x.w *= gelu_grad(za[k].w);
            }
          }
          if (p.resid) {
            x.x += rr[k].x;
            x.y += rr[k].y;
            x.z += rr[k].z;
            x.w += rr[k].w;
          }
          if (mr < p.M) {
            const long long off = static_cast<long long>(mr) * p.ldc + col;
            if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + f32_off(p, mr, col)) = x;
            if (p.out_hi) {
              const uint32_t h0 = pack_bf16x2(x.x, x.y), h1 = pack_bf16x2(x.z, x.w);
              *reinterpret_cast<uint2*>(p.out_hi + off) = make_uint2(h0, h1);
              if (p.out_lo) {
                const uint32_t l0 = pack_bf16x2(x.x - __uint_as_float(h0 << 16), x.y - __uint_as_float(h0 & 0xffff0000u));
                const uint32_t l1 = pack_bf16x2(x.z - __uint_as_float(h1 << 16), x.w - __uint_as_float(h1 & 0xffff0000u));
                *reinterpret_cast<uint2*>(p.out_lo + off) = make_uint2(l0, l1);
              }
            }
          }
        }
      }
      __syncwarp();  // the staging buffer is reused by the next chunk
      return;
    }
    if (!row_ok) return;
    if constexpr (STG != 0) return;  // (not reached: the staged paths returned above)
    if (p.resid) {
      const float* r = p.resid + static_cast<long long>(m) * p.ldr + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 rr = __ldg(reinterpret_cast<const float4*>(r + j));
        v[j] += rr.x;
        v[j + 1] += rr.y;
        v[j + 2] += rr.z;
        v[j + 3] += rr.w;
      }
    }
    const long long off = static_cast<long long>(m) * p.ldc + col0;
    if (p.out_f32) {
      float* o = p.out_f32 + f32_off(p, m, col0);
      const int step = p.f32_rows32 ? 128 : 4;   // consecutive float4 columns of a row are 32 rows apart in rows32
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(o + (j >> 2) * step) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if (p.out_hi) {
      __nv_bfloat16* oh = p.out_hi + off;
      __nv_bfloat16* ol = p.out_lo ? p.out_lo + off : nullptr;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t h[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) h[q] = pack_bf16x2(v[j + 2 * q], v[j + 2 * q + 1]);
        *reinterpret_cast<uint4*>(oh + j) = make_uint4(h[0], h[1], h[2], h[3]);
        if (ol) {
          uint32_t l[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float a = v[j + 2 * q] - __uint_as_float(h[q] << 16);
            const float b = v[j + 2 * q + 1] - __uint_as_float(h[q] & 0xffff0000u);
            l[q] = pack_bf16x2(a, b);
          }
          *reinterpret_cast<uint4*>(ol + j) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
    }
  } else {
    // ragged right edge (N not a multiple of 32): scalar path, only the last chunk of a row
    if (!row_ok) return;
    const long long off = static_cast<long long>(m) * p.ldc + col0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {  // static register indices: v must not fall into local memory
      const int col = col0 + j;
      if (col >= p.N) continue;
      float z = v[j];
      if (p.bias) z += sb ? sb[j] : __ldg(p.bias + col);
      if (col < p.col_scale_cols) z *= p.col_scale;
      if (p.out_pre) p.out_pre[off + j] = z;
      if constexpr (ACT == BLM_ACT_SOFTMAX_GRAD) {
        z = (ex2_approx((z - __ldg(p.lse + m)) * 1.4426950408889634f) - (col == __ldg(p.targets + m) ? 1.0f : 0.0f)) *
            p.grad_scale;
      } else if constexpr (ACT == BLM_ACT_GELU_GRAD) {
        z *= gelu_grad(__ldg(p.aux + static_cast<long long>(m) * p.ldaux + col));
      } else if constexpr (ACT == BLM_ACT_GPMIX_GRAD) {
        z *= gpmix_grad(__ldg(p.aux + static_cast<long long>(m) * p.ldaux + col), p.coef, p.N, col);
      } else {
        z = apply_act<ACT>(z, p.coef, p.N, col);
      }
      if (p.resid) z += __ldg(p.resid + static_cast<long long>(m) * p.ldr + col);
      if (p.out_f32) p.out_f32[f32_off(p, m, col)] = z;
      if (p.out_hi) {
        const __nv_bfloat16 hh = __float2bfloat16_rn(z);
        p.out_hi[off + j] = hh;
        if (p.out_lo) p.out_lo[off + j] = __float2bfloat16_rn(z - __bfloat162float(hh));
      }
    }
  }
}

// STG == 2: pack one finished 32-column chunk to bf16 and write it into this lane's 128-byte row of the
// warp's [32 rows x 64 columns] staging tile (half = 0 / 1: columns [0, 32) / [32, 64)), in the 128B-swizzle
// layout the output tensor map expects: 16-byte slot j of row r sits at slot j ^ (r & 7).
__device__ __forceinline__ void stage_chunk_bf16(const float (&v)[32], uint8_t* stg, int lane, int half) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = pack_bf16x2(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
    const int slot = (half * 4 + q) ^ (lane & 7);
    *reinterpret_cast<uint4*>(stg + lane * 128 + slot * 16) = make_uint4(h[0], h[1], h[2], h[3]);
  }
}


// 16-epilogue-warp variant: one 32-column chunk -> this lane's 64-byte row of the warp's [32 rows x 32 columns]
// bf16 staging tile in the 64B-swizzle layout (16-byte slot j of row r sits at slot j ^ ((r >> 1) & 3)).
__device__ __forceinline__ void stage_chunk_bf16_sw64(const float (&v)[32], uint8_t* stg, int lane) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = pack_bf16x2(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
    const int slot = q ^ ((lane >> 1) & 3);
    *reinterpret_cast<uint4*>(stg + lane * 64 + slot * 16) = make_uint4(h[0], h[1], h[2], h[3]);
  }
}

}  // namespace blm
