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

// ---- short sequences (T <= 32, head_dim 64): one WARP per (hypothesis, head) -------------------
// The rescoring case: T ~ 6..26.  K and V of the head sit in the warp's slice of shared memory;
// the warp is split into 32/W groups of W = 8/16/32 lanes so that R = 32/W query rows are scored
// at once (lane = key index inside the group): q.k dot products with q broadcast from L1, a
// group-wide shuffle softmax, then P.V with lane = output column.  No block-level barrier.
constexpr int kWarpAttnWarps = 4;

template <int HD>
__global__ void __launch_bounds__(kWarpAttnWarps * 32) mha_causal_warp_kernel(
    const float* __restrict__ qkv, const int* __restrict__ seq_offsets, long long n_pairs, int nhead, int max_len,
    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = static_cast<long long>(blockIdx.x) * kWarpAttnWarps + warp;
  if (gw >= n_pairs) return;
  const int seq = static_cast<int>(gw / nhead);
  const int head = static_cast<int>(gw - static_cast<long long>(seq) * nhead);
  const int row0 = seq_offsets[seq];
  const int T = seq_offsets[seq + 1] - row0;
  if (T > max_len || T > 32) {
    if (lane == 0) printf("blm: sequence %d has %d tokens > %d\n", seq, T, max_len < 32 ? max_len : 32);
    __trap();
  }
  constexpr int KP = HD + 1;  // padded K rows: lanes read different rows of the same column
  const int d = nhead * HD, ld = 3 * d;
  // per-warp slice: V [max_len, HD], Q [max_len, HD] (16-byte aligned), P [4, 32], K [max_len, HD + 1]
  const size_t per_warp = (static_cast<size_t>(max_len) * (KP + 2 * HD) + 4 * 32 + 3) & ~static_cast<size_t>(3);
  float* sV = sm + static_cast<size_t>(warp) * per_warp;
  float* sQ = sV + max_len * HD;
  float* sP = sQ + max_len * HD;
  float* sK = sP + 4 * 32;

  // stage q, k, v of this (hypothesis, head): every global load is issued up front, 12 x 128 bit
  // in flight per lane, so the DRAM latency is paid once per warp rather than once per query row
  const int n4 = T * (HD / 4);
  for (int i0 = lane; i0 < n4; i0 += 128) {
    float4 q[4], k[4], v[4];
    int t[4], c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + 32 * u;
      t[u] = i / (HD / 4);
      c[u] = (i - t[u] * (HD / 4)) * 4;
      if (i < n4) {
        const float* base = qkv + static_cast<long long>(row0 + t[u]) * ld + head * HD + c[u];
        q[u] = __ldg(reinterpret_cast<const float4*>(base));
        k[u] = __ldg(reinterpret_cast<const float4*>(base + d));
        v[u] = __ldg(reinterpret_cast<const float4*>(base + 2 * d));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + 32 * u < n4) {
        float* kd = sK + t[u] * KP + c[u];
        kd[0] = k[u].x; kd[1] = k[u].y; kd[2] = k[u].z; kd[3] = k[u].w;
        *reinterpret_cast<float4*>(sV + t[u] * HD + c[u]) = v[u];
        *reinterpret_cast<float4*>(sQ + t[u] * HD + c[u]) = q[u];
      }
    }
  }
  __syncwarp();

  const int W = T <= 8 ? 8 : (T <= 16 ? 16 : 32);
  const int R = 32 / W;
  const int g = lane / W, j = lane - g * W;
  // register blocking: lane (g, j) keeps key row j in registers for all query rows it scores, so the
  // inner product only streams the broadcast q row from shared memory
  float kreg[HD];
  {
    const float* k = sK + min(j, T - 1) * KP;
#pragma unroll
    for (int c = 0; c < HD; ++c) kreg[c] = k[c];
  }
  for (int i0 = 0; i0 < T; i0 += R) {
    const int i = i0 + g;
    const bool valid = (i < T) && (j <= i);
    const int ic = min(i, T - 1);
    const float4* q4 = reinterpret_cast<const float4*>(sQ + ic * HD);  // broadcast inside the group
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
    for (int c4 = 0; c4 < HD / 4; ++c4) {
      const float4 q = q4[c4];
      a0 = fmaf(q.x, kreg[4 * c4], a0);
      a1 = fmaf(q.y, kreg[4 * c4 + 1], a1);
      a0 = fmaf(q.z, kreg[4 * c4 + 2], a0);
      a1 = fmaf(q.w, kreg[4 * c4 + 3], a1);
    }
    const float s = valid ? a0 + a1 : -INFINITY;
    float mx = s;
    for (int o = W >> 1; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e = valid ? expf(s - mx) : 0.0f;
    float sum = e;
    for (int o = W >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sP[g * 32 + j] = (i < T) ? e / sum : 0.0f;
    __syncwarp();
    for (int g2 = 0; g2 < R; ++g2) {
      const int i2 = i0 + g2;
      if (i2 >= T) break;
      const float* pr = sP + g2 * 32;
      float o0 = 0.0f, o1 = 0.0f;
      for (int jj = 0; jj <= i2; ++jj) {
        const float pj = pr[jj];
        o0 = fmaf(pj, sV[jj * HD + lane], o0);
        if (HD > 32) o1 = fmaf(pj, sV[jj * HD + lane + 32], o1);
      }
      const long long off = static_cast<long long>(row0 + i2) * d + head * HD + lane;
      if (out_f32) {
        out_f32[off] = o0;
        if (HD > 32) out_f32[off + 32] = o1;
      }
      if (out_hi) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(o0), h1 = __float2bfloat16_rn(o1);
        out_hi[off] = h0;
        if (HD > 32) out_hi[off + 32] = h1;
        if (out_lo) {
          out_lo[off] = __float2bfloat16_rn(o0 - __bfloat162float(h0));
          if (HD > 32) out_lo[off + 32] = __float2bfloat16_rn(o1 - __bfloat162float(h1));
        }
      }
    }
    __syncwarp();
  }
}

static size_t warp_attn_smem_bytes(int max_len, int hd) {
  const size_t per_warp = (static_cast<size_t>(max_len) * (3 * hd + 1) + 4 * 32 + 3) & ~static_cast<size_t>(3);
  return sizeof(float) * kWarpAttnWarps * per_warp;
}

static size_t attn_smem_bytes(int max_len, int hd) {
  return sizeof(float) * (2ull * max_len * (hd + 1) + (kAttnThreads / 32) * hd + (kAttnThreads / 32) * max_len);
}

int attention_init() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(attn_smem_bytes(kAttnMaxLen, kAttnMaxHd))));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_warp_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(warp_attn_smem_bytes(32, 64))));
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
  if (head_dim == 64 && max_len <= 32) {
    const long long pairs = nseq * nhead;
    const unsigned blocks = static_cast<unsigned>((pairs + kWarpAttnWarps - 1) / kWarpAttnWarps);
    mha_causal_warp_kernel<64><<<blocks, kWarpAttnWarps * 32, warp_attn_smem_bytes(max_len, 64), as_stream(stream)>>>(
        qkv, seq_offsets, pairs, nhead, max_len, out_f32, reinterpret_cast<__nv_bfloat16*>(out_hi),
        reinterpret_cast<__nv_bfloat16*>(out_lo));
    BLM_CHECK_CUDA(cudaGetLastError());
    return BLM_OK;
  }
  mha_causal_kernel<<<static_cast<unsigned>(nseq * nhead), kAttnThreads, attn_smem_bytes(max_len, head_dim),
                      as_stream(stream)>>>(qkv, seq_offsets, nhead, head_dim, max_len, out_f32,
                                           reinterpret_cast<__nv_bfloat16*>(out_hi),
                                           reinterpret_cast<__nv_bfloat16*>(out_lo));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}
