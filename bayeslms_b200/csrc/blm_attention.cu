// Causal multi-head self-attention for short, packed, variable-length
// hypotheses (rescoring: T ~ 6..26 tokens; fine-tuning: T = 100).
//
// One CTA per (hypothesis, head).  K and V of that head live in shared memory
// (rows padded to head_dim + 1 floats so the stride-head_dim reads of the
// q.k dot products hit 32 different banks); each warp owns query rows
// i = warp, warp + nwarps, ...: lanes split the keys j <= i for the scores,
// a warp-shuffle max / sum gives the softmax, then lanes split the head_dim
// output columns for P.V.  The additive -inf mask of the reference
// (model.py:906-912) is realised by simply not visiting j > i.
#include "blm_host.h"
#include "blm_ptx.cuh"

namespace blm {

constexpr int kAttnThreads = 128;
constexpr int kAttnMaxLen = 128;
constexpr int kAttnMaxHd = 128;

__global__ void __launch_bounds__(kAttnThreads) mha_causal_kernel(
    const float* __restrict__ qkv, const int* __restrict__ seq_offsets, int nhead, int hd, int max_len,
    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  extern __shared__ float sm[];
  const int seq = blockIdx.x / nhead;
  const int head = blockIdx.x - seq * nhead;
  const int row0 = seq_offsets[seq];
  const int T = seq_offsets[seq + 1] - row0;
  if (T > max_len) {  // shared memory was sized for max_len rows
    if (threadIdx.x == 0) printf("blm: sequence %d has %d tokens > max_len %d\n", seq, T, max_len);
    __trap();
  }
  const int d = nhead * hd;
  const int ld = 3 * d;
  const int hp = hd + 1;
  float* sK = sm;                     // [T, hp]
  float* sV = sK + T * hp;            // [T, hp]
  float* sQ = sV + T * hp;            // [nwarps, hd]  current query row of each warp
  float* sP = sQ + (kAttnThreads / 32) * hd;  // [nwarps, T] softmax numerators

  // cooperative load of K and V (coalesced along head_dim)
  for (int idx = threadIdx.x; idx < T * hd; idx += kAttnThreads) {
    const int t = idx / hd, c = idx - t * hd;
    const float* base = qkv + static_cast<long long>(row0 + t) * ld + head * hd + c;
    sK[t * hp + c] = __ldg(base + d);
    sV[t * hp + c] = __ldg(base + 2 * d);
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = sQ + warp * hd;
  float* pr = sP + warp * T;
  for (int i = warp; i < T; i += kAttnThreads / 32) {
    const float* qrow = qkv + static_cast<long long>(row0 + i) * ld + head * hd;
    for (int c = lane; c < hd; c += 32) q[c] = __ldg(qrow + c);
    __syncwarp();
    // scores for keys j = lane, lane + 32, ... <= i
    float mx = -INFINITY;
    for (int j = lane; j <= i; j += 32) {
      const float* k = sK + j * hp;
      float s = 0.0f;
#pragma unroll 8
      for (int c = 0; c < hd; ++c) s = fmaf(q[c], k[c], s);
      pr[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = lane; j <= i; j += 32) {
      const float e = expf(pr[j] - mx);
      pr[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    for (int c = lane; c < hd; c += 32) {
      float o = 0.0f;
      for (int j = 0; j <= i; ++j) o = fmaf(pr[j], sV[j * hp + c], o);
      o *= inv;
      const long long off = static_cast<long long>(row0 + i) * d + head * hd + c;
      if (out_f32) out_f32[off] = o;
      if (out_hi) {
        const __nv_bfloat16 h = __float2bfloat16_rn(o);
        out_hi[off] = h;
        if (out_lo) out_lo[off] = __float2bfloat16_rn(o - __bfloat162float(h));
      }
    }
    __syncwarp();
  }
}

static size_t attn_smem_bytes(int max_len, int hd) {
  return sizeof(float) * (2ull * max_len * (hd + 1) + (kAttnThreads / 32) * hd + (kAttnThreads / 32) * max_len);
}

int attention_init() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(attn_smem_bytes(kAttnMaxLen, kAttnMaxHd))));
  return BLM_OK;
}

}  // namespace blm

extern "C" int blm_mha_causal(const float* qkv, const int32_t* seq_offsets, int64_t nseq, int32_t nhead,
                              int32_t head_dim, int32_t max_len, float* out_f32, blm_bf16* out_hi,
                              blm_bf16* out_lo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(qkv && seq_offsets && nseq > 0 && nhead > 0, BLM_ERR_ARG, "bad attention arguments");
  BLM_REQUIRE(head_dim > 0 && head_dim <= kAttnMaxHd, BLM_ERR_SHAPE, "head_dim %d not in (0, %d]", head_dim,
              kAttnMaxHd);
  BLM_REQUIRE(max_len > 0 && max_len <= kAttnMaxLen, BLM_ERR_SHAPE, "max_len %d not in (0, %d]", max_len,
              kAttnMaxLen);
  BLM_REQUIRE(out_f32 || out_hi, BLM_ERR_ARG, "no output buffer");
  BLM_REQUIRE(nseq * nhead < (1ll << 31), BLM_ERR_SHAPE, "too many (sequence, head) pairs");
  static bool attr_set = false;
  if (!attr_set) {
    int rc = attention_init();
    if (rc != BLM_OK) return rc;
    attr_set = true;
  }
  mha_causal_kernel<<<static_cast<unsigned>(nseq * nhead), kAttnThreads, attn_smem_bytes(max_len, head_dim),
                      as_stream(stream)>>>(qkv, seq_offsets, nhead, head_dim, max_len, out_f32,
                                           reinterpret_cast<__nv_bfloat16*>(out_hi),
                                           reinterpret_cast<__nv_bfloat16*>(out_lo));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}
