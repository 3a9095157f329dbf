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
#include "blm_philox.cuh"

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

// ---- tensor-core attention over bf16 (hi[, lo]) q/k/v ------------------------------------------
// The QKV projection writes its output as bf16 (hi[, lo]) pairs; this kernel stages the [T x 64]
// q, k, v blocks of each (hypothesis, head) with cp.async (16-byte chunks, XOR-swizzled rows, zero
// fill past the hypothesis end), and runs S = Q K^T and O = P V on mma.sync m16n8k16 (bf16 in, fp32
// accumulate) with the flash-attention register pipeline: the S accumulator fragments are masked,
// soft-maxed (online, log2 domain) and re-used in place as the A fragments of P V; V is read through
// ldmatrix.trans.  PRECISE carries every operand as hi + lo and issues hi*hi + hi*lo + lo*hi.
// Hypotheses are 6..26 tokens (rescoring) or 100 (fine-tuning): a tcgen05 128-row tile would be
// mostly padding and the kernel is bound by the q/k/v read anyway (4 KB per token), so the warp-level
// MMA is the right atom here.
//   KV_ROWS = 32 : one warp per (hypothesis, head), four pairs per CTA          (max_len <= 32)
//   KV_ROWS = 128: one CTA per (hypothesis, head, 128-row query block), warp w owns query rows [32w, 32w + 32) of
//                  the block; the CTA of query block b walks the key / value blocks 0..b (flash-attention style:
//                  K, V re-staged 128 rows at a time, running max / sum / output carried in registers), so a
//                  hypothesis may be any length -- the reference scores whatever fits its 5000-row positional table
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kMmaAttnHd = 64;
constexpr int kMmaAttnRowBytes = kMmaAttnHd * 2;  // one q/k/v row of a head: 128 B = 8 chunks of 16 B

// DROP: dropout on the attention probabilities (training, model.py:912-913): P is multiplied by the keep multipliers
// AFTER the row sum l_i has been taken, so O = (m . P) V with P the full softmax.
template <bool PRECISE, int KV_ROWS, bool DROP = false>
__global__ void __launch_bounds__(128) mha_causal_mma_kernel(
    const __nv_bfloat16* __restrict__ qkv_hi, const __nv_bfloat16* __restrict__ qkv_lo, long long ld,
    const int* __restrict__ seq_offsets, long long n_pairs, int nhead, float* __restrict__ out_f32,
    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, long long ldo,
    DropParams dparams = DropParams{}, int ldm = 0) {
  constexpr int PARTS = PRECISE ? 2 : 1;
  constexpr int WPP = KV_ROWS / 32;  // warps per (hypothesis, head) pair
  constexpr int PPC = 4 / WPP;       // pairs per CTA
  extern __shared__ __align__(128) uint8_t attn_sm[];
  // [matrix q,k,v][part hi,lo][128 rows][128 B]
  const uint32_t sm_base = smem_u32(attn_sm);
  auto tile = [&](int mat, int part) -> uint32_t { return sm_base + static_cast<uint32_t>((mat * PARTS + part) * 128 * 128); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pair = static_cast<long long>(blockIdx.x) * PPC + warp / WPP;
  const int qt = warp % WPP;
  int row0 = 0, T = 0, head = 0;
  if (pair < n_pairs) {
    const int seq = static_cast<int>(pair / nhead);
    head = static_cast<int>(pair - static_cast<long long>(seq) * nhead);
    row0 = __ldg(seq_offsets + seq);
    T = __ldg(seq_offsets + seq + 1) - row0;
    if (KV_ROWS == 32 && T > KV_ROWS) {
      if (lane == 0) printf("blm: sequence %d has %d tokens > %d\n", seq, T, KV_ROWS);
      __trap();
    }
  }
  const int qb = KV_ROWS == 128 ? static_cast<int>(blockIdx.y) : 0;   // 128-row query block of the pair
  const int q0 = qb * 128;                                            // its first token
  if (KV_ROWS == 128 && q0 >= T) return;                              // (whole CTA: one pair per CTA in this variant)
  const int Tall = T;                                                 // tokens of the hypothesis
  T = min(T - q0, 128);                                               // query rows of this block (= T when T <= 128)
  const int d = nhead * kMmaAttnHd;
  const int pb = (warp / WPP) * KV_ROWS;  // first shared-memory row of this pair
  // one pair per CTA (KV_ROWS == 128, T <= 128 when dropping): its keep bits are generated once into shared memory
  __shared__ uint32_t s_keep[(DROP && KV_ROWS == 128) ? 512 : 1];
  __shared__ unsigned s_kept;
  float kscale = 0.0f;
  if constexpr (DROP && KV_ROWS == 128) {
    const DropParams dr = drop_resolve(dparams);
    drop_keep_bits(dr, pair * ldm * ldm, ldm * ldm, s_keep, &s_kept, threadIdx.x, 128);
    kscale = dr.mask ? __uint_as_float(s_kept) : dr.scale;
  }
  // ---- staging: rows [32 qt, 32 qt + 32) of matrices [m0, m1) (0 = q, 1 = k, 2 = v) of this pair, taken from tokens
  // tok0.. of the hypothesis, zero fill past `valid` rows.  Lane l copies 16-byte chunk l % 8 of rows l / 8 + 4 it: one
  // pointer and one swizzled offset per (matrix, part), advanced by constants (the address arithmetic of the first
  // version was 60 % of this kernel's instructions, ncu r01az).  Rows past the last 16-row tile that holds a valid
  // token are never read by an MMA and are skipped.
  auto stage = [&](int m0, int m1, int tok0, int valid) {
    const int r0 = lane >> 3, ch = lane & 7;
    const int rows_used = min(32, ((valid - qt * 32 + 15) & ~15));      // 16 or 32 (<= 0: nothing to stage)
    const long long row_step = 4 * ld;                                   // elements between two iterations
    // swizzled chunk position alternates with (row & 7) = r0 or r0 + 4
    const uint32_t sw0 = static_cast<uint32_t>((ch ^ r0) << 4), sw1 = static_cast<uint32_t>((ch ^ (r0 + 4)) << 4);
    const uint32_t dst_row = static_cast<uint32_t>((pb + qt * 32 + r0) * kMmaAttnRowBytes);
    for (int mat = m0; mat < m1; ++mat) {
#pragma unroll
      for (int part = 0; part < PARTS; ++part) {
        const __nv_bfloat16* src = (part == 0 ? qkv_hi : qkv_lo) + mat * d + head * kMmaAttnHd + ch * 8 +
                                   static_cast<long long>(row0 + tok0 + qt * 32 + r0) * ld;
        const uint32_t dst = tile(mat, part) + dst_row;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if (it * 4 < rows_used) {
            const bool ok = qt * 32 + r0 + it * 4 < valid;
            cp_async16(dst + static_cast<uint32_t>(it * 4 * kMmaAttnRowBytes) + ((it & 1) ? sw1 : sw0),
                       ok ? src + it * row_step : src, ok ? 16u : 0u);
          }
        }
      }
    }
  };
  stage(0, 1, q0, T);                       // the query rows of this block
  const bool active = qt * 32 < T;          // this warp owns query rows (all warps stage and meet at the barriers)

  const int g = lane >> 2, t4 = lane & 3;
  const int i_hi = min(T, qt * 32 + 32) - 1;      // last valid query row of this warp (inside the query block)
  const int nmt = active ? (i_hi - qt * 32) / 16 + 1 : 0;   // m16 tiles in use (1 or 2)
  constexpr float kLog2e = 1.4426950408889634f;

  float o[2][8][4];
  float mrow[2][2], lrow[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    mrow[mt][0] = mrow[mt][1] = -INFINITY;
    lrow[mt][0] = lrow[mt][1] = 0.0f;
#pragma unroll
    for (int ct = 0; ct < 8; ++ct)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][ct][e] = 0.0f;
  }
  // swizzled address of (row, 16-byte chunk) inside a tile
  auto addr = [&](uint32_t base, int row, int chunk) -> uint32_t {
    return base + static_cast<uint32_t>(row * kMmaAttnRowBytes + ((chunk ^ (row & 7)) << 4));
  };

  for (int c = 0; c <= qb; ++c) {           // key / value blocks of 128 tokens, up to the query block's own
  if (c > 0) __syncthreads();               // every warp is done with the previous block's tiles
  stage(1, 3, c * 128, min(Tall - c * 128, 128));
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int joff = (c - qb) * 128;          // key index relative to the query block: <= -128 for earlier blocks
  const int kb_last = !active ? -1 : (c == qb ? qt : 3);
  for (int kb = 0; kb <= kb_last; ++kb) {
    const int jmax = c == qb ? min(i_hi, kb * 32 + 31) : kb * 32 + 31;
    const int nnt = (jmax - kb * 32) / 8 + 1;  // n8 key tiles in use (1..4)
    float s[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.0f;
    // ---- S = Q K^T over the 64 head dimensions
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t aq[2][PARTS][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (mt < nmt) {
          const int row = pb + qt * 32 + mt * 16 + (lane & 15);
#pragma unroll
          for (int part = 0; part < PARTS; ++part) ldsm_x4(addr(tile(0, part), row, ks * 2 + (lane >> 4)), aq[mt][part]);
        }
      }
#pragma unroll
      for (int ntp = 0; ntp < 2; ++ntp) {
        if (ntp * 2 < nnt) {
          uint32_t bk[PARTS][4];
          const int row = pb + kb * 32 + ntp * 16 + (lane & 7) + ((lane >> 4) << 3);
#pragma unroll
          for (int part = 0; part < PARTS; ++part) ldsm_x4(addr(tile(1, part), row, ks * 2 + ((lane >> 3) & 1)), bk[part]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (mt < nmt) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                mma_bf16(s[mt][ntp * 2 + u], aq[mt][0], bk[0][2 * u], bk[0][2 * u + 1]);
                if constexpr (PRECISE) {
                  mma_bf16(s[mt][ntp * 2 + u], aq[mt][0], bk[1][2 * u], bk[1][2 * u + 1]);
                  mma_bf16(s[mt][ntp * 2 + u], aq[mt][1], bk[0][2 * u], bk[0][2 * u + 1]);
                }
              }
            }
          }
        }
      }
    }
    // ---- causal mask + online softmax (log2 domain), P re-packed as the A fragments of P V
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      if (mt < nmt) {
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = qt * 32 + mt * 16 + g + ((e >> 1) << 3);
            const int j = kb * 32 + nt * 8 + 2 * t4 + (e & 1);
            const float v = (j + joff <= i && nt < nnt) ? s[mt][nt][e] * kLog2e : -INFINITY;
            s[mt][nt][e] = v;
            mx[e >> 1] = fmaxf(mx[e >> 1], v);
          }
        float alpha[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
          mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
          const float mnew = fmaxf(mrow[mt][h], mx[h]);  // finite: key kb*32 <= every row of this warp
          alpha[h] = ex2f_(mrow[mt][h] - mnew);
          mrow[mt][h] = mnew;
          lrow[mt][h] *= alpha[h];
        }
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) {
          o[mt][ct][0] *= alpha[0];
          o[mt][ct][1] *= alpha[0];
          o[mt][ct][2] *= alpha[1];
          o[mt][ct][3] *= alpha[1];
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pv = ex2f_(s[mt][nt][e] - mrow[mt][e >> 1]);
            s[mt][nt][e] = pv;
            lrow[mt][e >> 1] += pv;
          }
        if constexpr (DROP) {
          const DropParams dr = drop_resolve(dparams);
          const long long mbase = pair * ldm * ldm;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int i = qt * 32 + mt * 16 + g + 8 * h;
              const int j0 = kb * 32 + nt * 8 + 2 * t4;
              if (nt < nnt && i < T && j0 + joff <= i) {   // (elements with j > i are already zero)
                float2 dm;
                if constexpr (KV_ROWS == 128)
                  dm = keep_mult2(s_keep, kscale, (q0 + i) * ldm + c * 128 + j0);
                else
                  dm = drop_mult2(dr, mbase + static_cast<long long>(q0 + i) * ldm + c * 128 + j0);
                s[mt][nt][2 * h] *= dm.x;
                s[mt][nt][2 * h + 1] *= dm.y;
              }
            }
        }
      }
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      if (kk * 2 < nnt) {
        uint32_t ap[2][PARTS][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (mt < nmt) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float x0 = s[mt][2 * kk + (q >> 1)][(q & 1) * 2], x1 = s[mt][2 * kk + (q >> 1)][(q & 1) * 2 + 1];
              const uint32_t hi = pack_bf16x2(x0, x1);
              ap[mt][0][q] = hi;
              if constexpr (PRECISE)
                ap[mt][1][q] = pack_bf16x2(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
            }
          }
        }
#pragma unroll
        for (int ctp = 0; ctp < 4; ++ctp) {
          uint32_t bv[PARTS][4];
          const int row = pb + kb * 32 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
#pragma unroll
          for (int part = 0; part < PARTS; ++part) ldsm_x4_trans(addr(tile(2, part), row, ctp * 2 + (lane >> 4)), bv[part]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (mt < nmt) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                mma_bf16(o[mt][ctp * 2 + u], ap[mt][0], bv[0][2 * u], bv[0][2 * u + 1]);
                if constexpr (PRECISE) {
                  mma_bf16(o[mt][ctp * 2 + u], ap[mt][0], bv[1][2 * u], bv[1][2 * u + 1]);
                  mma_bf16(o[mt][ctp * 2 + u], ap[mt][1], bv[0][2 * u], bv[0][2 * u + 1]);
                }
              }
            }
          }
        }
      }
    }
  }
  }   // key / value blocks
  if (!active) return;
  // ---- normalise and store: thread holds rows g, g + 8 of each m tile, columns 8 ct + 2 t4 + {0, 1}
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    if (mt < nmt) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float l = lrow[mt][h];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const float inv = 1.0f / l;
        const int i = qt * 32 + mt * 16 + g + 8 * h;
        if (i < T) {
          const long long off = static_cast<long long>(row0 + q0 + i) * ldo + head * kMmaAttnHd + 2 * t4;
#pragma unroll
          for (int ct = 0; ct < 8; ++ct) {
            const float x0 = o[mt][ct][2 * h] * inv, x1 = o[mt][ct][2 * h + 1] * inv;
            if (out_f32) *reinterpret_cast<float2*>(out_f32 + off + ct * 8) = make_float2(x0, x1);
            if (out_hi) {
              const uint32_t hi = pack_bf16x2(x0, x1);
              *reinterpret_cast<uint32_t*>(out_hi + off + ct * 8) = hi;
              if (out_lo)
                *reinterpret_cast<uint32_t*>(out_lo + off + ct * 8) =
                    pack_bf16x2(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ tensor-core attention backward
// dqkv = d(causal MHA)/d(q, k, v) for one (sequence, head) per CTA, T <= 128, head_dim 64, on mma.sync m16n8k16
// bf16 (hi[, lo] parts as in the forward kernel; PRECISE = three products per term).  q, k, v (q already scaled)
// and dO arrive as fp32 and are split into swizzled bf16 tiles while they are staged.
// Two phases, nothing of size T x T is stored and no fragment is transposed:
//   A (warp w = query rows [32w, 32w+32)): pass 1 over the key blocks j <= i rebuilds S = Q K^T and dP = dO V^T
//     and folds them into the row statistics m_i, l_i, D_i = sum_j P_ij dP_ij (online, log2 domain); pass 2
//     rebuilds both again, forms dS = P (dP - D) in registers and accumulates dQ += dS K with dS re-packed from
//     the accumulator layout into A fragments (the forward kernel's P V trick).
//   B (warp w = key rows [32w, 32w+32)): for every 16-row query block i >= j, S^T = K Q^T and dP^T = V dO^T come out
//     with the KEY index as the accumulator row, so P^T and dS^T re-pack into A fragments of
//     dV += P^T dO and dK += dS^T Q; the statistics of phase A are read from shared memory per column.
// Recomputing the two small products three times costs ~0.3 GFLOP per launch; the fp32 SIMT kernel it replaces
// read 5 MB of shared memory per CTA and took 155 us per layer (19 % of the fine-tune step).
// MT = m16 tiles per warp: 1 -> eight warps of 16 rows (bf16 mode: <= 128 registers, two CTAs per SM; the precise
// mode's 129 KB of tiles allow one CTA per SM anyway, so it keeps the full register file): the launch is one wave of
// nseq * nhead latency-bound CTAs, so warps per CTA are what shortens it.  2 -> four warps of 32 rows.
// DROP (attention-probability dropout, P' = m . P): dV = P'^T dO, dP = m . (dO V^T), D_i = sum_j P_ij dP_ij,
// dS = P (dP - D): the multipliers are re-derived from the same mask tensor / Philox stream as in the forward kernel.
template <bool PRECISE, int MT, bool DROP = false>
__global__ void __launch_bounds__(128 / (16 * MT) * 32, (MT == 1 && !PRECISE) ? 2 : 1) mha_causal_bwd_mma_kernel(
    const float* __restrict__ qkv, long long ld, const float* __restrict__ dout, long long ldo,
    const int* __restrict__ seq_offsets, int nhead, float q_scale, float* __restrict__ dqkv, long long ldd,
    DropParams dparams = DropParams{}, int ldm = 0) {
  constexpr int PARTS = PRECISE ? 2 : 1;
  constexpr int RB = 16 * MT;               // rows per warp
  constexpr int NT = 128 / RB * 32;         // threads
  extern __shared__ __align__(128) uint8_t attn_sm[];
  const uint32_t sm_base = smem_u32(attn_sm);
  // [matrix q,k,v,dO][part hi,lo][128 rows][128 B], then the row statistics
  auto tile = [&](int mat, int part) -> uint32_t { return sm_base + static_cast<uint32_t>((mat * PARTS + part) * 128 * 128); };
  float* sM = reinterpret_cast<float*>(attn_sm + 4 * PARTS * 128 * 128);
  float* sL = sM + 128;
  float* sD = sL + 128;
  auto addr = [&](uint32_t base, int row, int chunk) -> uint32_t {
    return base + static_cast<uint32_t>(row * kMmaAttnRowBytes + ((chunk ^ (row & 7)) << 4));
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.x / nhead, head = blockIdx.x - seq * nhead;
  const int row0 = __ldg(seq_offsets + seq);
  const int T = __ldg(seq_offsets + seq + 1) - row0;
  if (T > 128) {
    if (threadIdx.x == 0) printf("blm: sequence %d has %d tokens > 128\n", seq, T);
    __trap();
  }
  const int d = nhead * kMmaAttnHd;
  const int Tpad = (T + 31) & ~31;
  // ---- stage q, k, v, dO: fp32 -> bf16 hi (lo), zero rows up to the next multiple of 32
  for (int idx = threadIdx.x; idx < 4 * Tpad * 8; idx += NT) {
    const int mat = idx / (Tpad * 8), rem = idx - mat * (Tpad * 8);
    const int r = rem >> 3, ch = rem & 7;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (r < T) {
      const float* src = mat < 3 ? qkv + static_cast<long long>(row0 + r) * ld + mat * d + head * kMmaAttnHd + ch * 8
                                 : dout + static_cast<long long>(row0 + r) * ldo + head * kMmaAttnHd + ch * 8;
      a = __ldg(reinterpret_cast<const float4*>(src));
      b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    }
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = pack_bf16x2(x[2 * e], x[2 * e + 1]);
      lo[e] = pack_bf16x2(x[2 * e] - __uint_as_float(hi[e] << 16), x[2 * e + 1] - __uint_as_float(hi[e] & 0xffff0000u));
    }
    const uint32_t off = static_cast<uint32_t>(r * kMmaAttnRowBytes + ((ch ^ (r & 7)) << 4));
    *reinterpret_cast<uint4*>(attn_sm + (mat * PARTS) * 128 * 128 + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if constexpr (PRECISE)
      *reinterpret_cast<uint4*>(attn_sm + (mat * PARTS + 1) * 128 * 128 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  __syncthreads();

  const int g = lane >> 2, t4 = lane & 3;
  constexpr float kLog2e = 1.4426950408889634f;
  // C[16 x 8] += A(rows ra.., mat ma) B(rows rb.., mat mb)^T over the 64 head dimensions, for 2 m tiles x 4 n tiles
  auto rows_product = [&](float (&c)[MT][4][4], int ma, int ra, int mb, int rb, int nmt, int nnt) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[mt][nt][e] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t fa[MT][PARTS][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (mt < nmt) {
          const int row = ra + mt * 16 + (lane & 15);
#pragma unroll
          for (int part = 0; part < PARTS; ++part) ldsm_x4(addr(tile(ma, part), row, ks * 2 + (lane >> 4)), fa[mt][part]);
        }
      }
#pragma unroll
      for (int ntp = 0; ntp < 2; ++ntp) {
        if (ntp * 2 < nnt) {
          uint32_t fb[PARTS][4];
          const int row = rb + ntp * 16 + (lane & 7) + ((lane >> 4) << 3);
#pragma unroll
          for (int part = 0; part < PARTS; ++part) ldsm_x4(addr(tile(mb, part), row, ks * 2 + ((lane >> 3) & 1)), fb[part]);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (mt < nmt) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                mma_bf16(c[mt][ntp * 2 + u], fa[mt][0], fb[0][2 * u], fb[0][2 * u + 1]);
                if constexpr (PRECISE) {
                  mma_bf16(c[mt][ntp * 2 + u], fa[mt][0], fb[1][2 * u], fb[1][2 * u + 1]);
                  mma_bf16(c[mt][ntp * 2 + u], fa[mt][1], fb[0][2 * u], fb[0][2 * u + 1]);
                }
              }
            }
          }
        }
      }
    }
  };
  // acc[2][8][4] += X (accumulator-layout [2 m tiles][4 n tiles], contraction over its 32 columns) . rows rb.. of mat mb
  auto acc_product = [&](float (&acc)[MT][8][4], const float (&x)[MT][4][4], int mb, int rb, int nmt, int nnt) {
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      if (kk * 2 < nnt) {
        uint32_t ap[MT][PARTS][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (mt < nmt) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float x0 = x[mt][2 * kk + (q >> 1)][(q & 1) * 2], x1 = x[mt][2 * kk + (q >> 1)][(q & 1) * 2 + 1];
              const uint32_t hi = pack_bf16x2(x0, x1);
              ap[mt][0][q] = hi;
              if constexpr (PRECISE)
                ap[mt][1][q] = pack_bf16x2(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
            }
          }
        }
#pragma unroll
        for (int ctp = 0; ctp < 4; ++ctp) {
          uint32_t bv[PARTS][4];
          const int row = rb + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
#pragma unroll
          for (int part = 0; part < PARTS; ++part) ldsm_x4_trans(addr(tile(mb, part), row, ctp * 2 + (lane >> 4)), bv[part]);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (mt < nmt) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                mma_bf16(acc[mt][ctp * 2 + u], ap[mt][0], bv[0][2 * u], bv[0][2 * u + 1]);
                if constexpr (PRECISE) {
                  mma_bf16(acc[mt][ctp * 2 + u], ap[mt][0], bv[1][2 * u], bv[1][2 * u + 1]);
                  mma_bf16(acc[mt][ctp * 2 + u], ap[mt][1], bv[0][2 * u], bv[0][2 * u + 1]);
                }
              }
            }
          }
        }
      }
    }
  };

  // keep bits of this (sequence, head) pair's [ldm x ldm] multipliers, generated once
  __shared__ uint32_t s_keep[DROP ? 512 : 1];
  __shared__ unsigned s_kept;
  float kscale = 0.0f;
  if constexpr (DROP) {
    const DropParams dr = drop_resolve(dparams);
    drop_keep_bits(dr, static_cast<long long>(blockIdx.x) * ldm * ldm, ldm * ldm, s_keep, &s_kept, threadIdx.x, NT);
    kscale = dr.mask ? __uint_as_float(s_kept) : dr.scale;
  }
  // dP' -> dP = m . dP' on an accumulator block whose rows are queries r0.. and columns keys c0..
  auto drop_rows = [&](float (&x)[MT][4][4], int r0, int c0, int nmt_, int nnt_) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = r0 + mt * 16 + g + 8 * h;
          const int j0 = c0 + nt * 8 + 2 * t4;
          if (mt < nmt_ && nt < nnt_ && i < T && j0 <= i) {
            const float2 dm = keep_mult2(s_keep, kscale, i * ldm + j0);
            x[mt][nt][2 * h] *= dm.x;
            x[mt][nt][2 * h + 1] *= dm.y;
          }
        }
  };
  // ================================================================= phase A: query rows of this warp
  const bool active = warp * RB < T;
  const int qt = warp;
  const int i_hi = min(T, qt * RB + RB) - 1;
  const int nmt = active ? (i_hi - qt * RB) / 16 + 1 : 0;
  float mrow[MT][2], lrow[MT][2], drow[MT][2];
  float s[MT][4][4], dp[MT][4][4];
  if (active) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mrow[mt][h] = -INFINITY;
        lrow[mt][h] = 0.0f;
        drow[mt][h] = 0.0f;
      }
    // ---- pass 1: statistics
    for (int kb = 0; kb * 32 <= i_hi; ++kb) {
      const int jmax = min(i_hi, kb * 32 + 31);
      const int nnt = (jmax - kb * 32) / 8 + 1;
      rows_product(s, 0, qt * RB, 1, kb * 32, nmt, nnt);
      rows_product(dp, 3, qt * RB, 2, kb * 32, nmt, nnt);
      if constexpr (DROP) drop_rows(dp, qt * RB, kb * 32, nmt, nnt);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (mt < nmt) {
          float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = qt * RB + mt * 16 + g + ((e >> 1) << 3);
              const int j = kb * 32 + nt * 8 + 2 * t4 + (e & 1);
              const float v = (j <= i && nt < nnt) ? s[mt][nt][e] * kLog2e : -INFINITY;
              s[mt][nt][e] = v;
              mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            const float mnew = fmaxf(mrow[mt][h], mx[h]);
            const float alpha = ex2f_(mrow[mt][h] - mnew);
            mrow[mt][h] = mnew;
            lrow[mt][h] *= alpha;
            drow[mt][h] *= alpha;
          }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float pv = ex2f_(s[mt][nt][e] - mrow[mt][e >> 1]);   // 0 where masked
              lrow[mt][e >> 1] += pv;
              drow[mt][e >> 1] = fmaf(pv, (nt < nnt) ? dp[mt][nt][e] : 0.0f, drow[mt][e >> 1]);
            }
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float l = lrow[mt][h], dd = drow[mt][h];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        dd += __shfl_xor_sync(0xffffffffu, dd, 1);
        dd += __shfl_xor_sync(0xffffffffu, dd, 2);
        const float inv = (mt < nmt && l > 0.0f) ? 1.0f / l : 0.0f;
        lrow[mt][h] = inv;
        drow[mt][h] = dd * inv;
        if (t4 == 0) {
          const int i = qt * RB + mt * 16 + g + 8 * h;
          sM[i] = mrow[mt][h];
          sL[i] = inv;
          sD[i] = dd * inv;
        }
      }
  }
  __syncthreads();   // the statistics of every query row are in shared memory: pass 2 and phase B need no further sync
  if (active) {
    // ---- pass 2: dS and dQ
    float dq[MT][8][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int ct = 0; ct < 8; ++ct)
#pragma unroll
        for (int e = 0; e < 4; ++e) dq[mt][ct][e] = 0.0f;
    for (int kb = 0; kb * 32 <= i_hi; ++kb) {
      const int jmax = min(i_hi, kb * 32 + 31);
      const int nnt = (jmax - kb * 32) / 8 + 1;
      rows_product(s, 0, qt * RB, 1, kb * 32, nmt, nnt);
      rows_product(dp, 3, qt * RB, 2, kb * 32, nmt, nnt);
      if constexpr (DROP) drop_rows(dp, qt * RB, kb * 32, nmt, nnt);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = qt * RB + mt * 16 + g + ((e >> 1) << 3);
            const int j = kb * 32 + nt * 8 + 2 * t4 + (e & 1);
            const bool ok = j <= i && nt < nnt && mt < nmt;
            const float pv = ok ? ex2f_(fmaf(s[mt][nt][e], kLog2e, -mrow[mt][e >> 1])) * lrow[mt][e >> 1] : 0.0f;
            s[mt][nt][e] = ok ? pv * (dp[mt][nt][e] - drow[mt][e >> 1]) : 0.0f;   // dS
          }
      acc_product(dq, s, 1, kb * 32, nmt, nnt);
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      if (mt < nmt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = qt * RB + mt * 16 + g + 8 * h;
          if (i < T) {
            float* o = dqkv + static_cast<long long>(row0 + i) * ldd + head * kMmaAttnHd + 2 * t4;
#pragma unroll
            for (int ct = 0; ct < 8; ++ct)
              *reinterpret_cast<float2*>(o + ct * 8) = make_float2(dq[mt][ct][2 * h] * q_scale, dq[mt][ct][2 * h + 1] * q_scale);
          }
        }
      }
    }
  }

  // ================================================================= phase B: key rows of this warp
  if (active) {
    const int jt = warp;
    float dk[MT][8][4], dv[MT][8][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int ct = 0; ct < 8; ++ct)
#pragma unroll
        for (int e = 0; e < 4; ++e) dk[mt][ct][e] = dv[mt][ct][e] = 0.0f;
    float (&st)[MT][4][4] = s;
    float (&dpt)[MT][4][4] = dp;
    const int nblk = (T + 31) / 32;
    for (int ib = (jt * RB) / 32; ib < nblk; ++ib) {
      const int imax = min(T - 1, ib * 32 + 31);
      const int nnt = (imax - ib * 32) / 8 + 1;   // n8 tiles of query columns in use
      rows_product(st, 1, jt * RB, 0, ib * 32, nmt, nnt);    // S^T = K Q^T
      rows_product(dpt, 2, jt * RB, 3, ib * 32, nmt, nnt);   // dP^T = V dO^T
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float cm[2], cl[2], cd[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = ib * 32 + nt * 8 + 2 * t4 + u;   // rows beyond T hold finite leftovers or zeros; masked below
          cm[u] = sM[i & 127];
          cl[u] = sL[i & 127];
          cd[u] = sD[i & 127];
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = jt * RB + mt * 16 + g + ((e >> 1) << 3);
            const int i = ib * 32 + nt * 8 + 2 * t4 + (e & 1);
            const bool ok = j <= i && i < T && nt < nnt && mt < nmt;
            const float pv = ok ? ex2f_(fmaf(st[mt][nt][e], kLog2e, -cm[e & 1])) * cl[e & 1] : 0.0f;
            float dm = 1.0f;
            if constexpr (DROP) dm = ok ? keep_mult1(s_keep, kscale, i * ldm + j) : 0.0f;
            st[mt][nt][e] = pv * dm;                                           // P'^T = (m . P)^T
            dpt[mt][nt][e] = ok ? pv * (dm * dpt[mt][nt][e] - cd[e & 1]) : 0.0f;   // dS^T
          }
      }
      acc_product(dv, st, 3, ib * 32, nmt, nnt);    // dV += P^T dO
      acc_product(dk, dpt, 0, ib * 32, nmt, nnt);   // dK += dS^T Q
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      if (mt < nmt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = jt * RB + mt * 16 + g + 8 * h;
          if (j < T) {
            float* o = dqkv + static_cast<long long>(row0 + j) * ldd + d + head * kMmaAttnHd + 2 * t4;
#pragma unroll
            for (int ct = 0; ct < 8; ++ct) {
              *reinterpret_cast<float2*>(o + ct * 8) = make_float2(dk[mt][ct][2 * h], dk[mt][ct][2 * h + 1]);
              *reinterpret_cast<float2*>(o + d + ct * 8) = make_float2(dv[mt][ct][2 * h], dv[mt][ct][2 * h + 1]);
            }
          }
        }
      }
    }
  }
}

template <bool PRECISE>
static constexpr int mma_attn_bwd_smem_bytes() { return 4 * (PRECISE ? 2 : 1) * 128 * 128 + 3 * 128 * 4; }

template <bool PRECISE>
static constexpr int mma_attn_smem_bytes() { return 3 * (PRECISE ? 2 : 1) * 128 * 128; }

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
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_mma_kernel<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_smem_bytes<false>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_mma_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_smem_bytes<false>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_mma_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_smem_bytes<true>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_mma_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_smem_bytes<true>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_bwd_mma_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_bwd_smem_bytes<false>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_bwd_mma_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_bwd_smem_bytes<true>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_mma_kernel<false, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_smem_bytes<false>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_mma_kernel<false, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_smem_bytes<false>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_mma_kernel<true, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_smem_bytes<true>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_mma_kernel<true, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_smem_bytes<true>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_bwd_mma_kernel<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_bwd_smem_bytes<false>()));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_bwd_mma_kernel<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      mma_attn_bwd_smem_bytes<true>()));
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


extern "C" int blm_mha_causal_bf16(const blm_bf16* qkv_hi, const blm_bf16* qkv_lo, int64_t ld,
                                   const int32_t* seq_offsets, int64_t nseq, int32_t nhead, int32_t head_dim,
                                   int32_t max_len, float* out_f32, blm_bf16* out_hi, blm_bf16* out_lo, int64_t ldo,
                                   blm_stream stream) {
  return blm_mha_causal_bf16_dropout(qkv_hi, qkv_lo, ld, seq_offsets, nseq, nhead, head_dim, max_len, nullptr, out_f32,
                                     out_hi, out_lo, ldo, stream);
}

static bool drop_active(const blm_dropout_desc* d) { return d && (d->mask || d->p > 0.0f); }

extern "C" int blm_mha_causal_bf16_dropout(const blm_bf16* qkv_hi, const blm_bf16* qkv_lo, int64_t ld,
                                           const int32_t* seq_offsets, int64_t nseq, int32_t nhead, int32_t head_dim,
                                           int32_t max_len, const blm_dropout_desc* drop, float* out_f32,
                                           blm_bf16* out_hi, blm_bf16* out_lo, int64_t ldo, blm_stream stream) {
  using namespace blm;
  const bool dropping = drop_active(drop);
  if (dropping) {
    BLM_REQUIRE(drop->p >= 0.0f && drop->p < 1.0f, BLM_ERR_ARG, "dropout probability %g not in [0, 1)", drop->p);
    BLM_REQUIRE((reinterpret_cast<uintptr_t>(drop->mask) & 7u) == 0, BLM_ERR_ALIGN, "attention dropout mask must be 8-byte aligned");
  }
  const DropParams dpar = dropping ? make_drop_params(drop->mask, drop->p, drop->seed, drop->seed_dev, drop->stream_id)
                                   : DropParams{};
  const int ldm = (max_len + 3) & ~3;
  BLM_REQUIRE(qkv_hi && seq_offsets && nseq > 0 && nhead > 0, BLM_ERR_ARG, "bad attention arguments");
  BLM_REQUIRE(head_dim == kMmaAttnHd, BLM_ERR_SHAPE, "the tensor-core attention kernel needs head_dim 64, got %d", head_dim);
  BLM_REQUIRE(max_len > 0 && max_len <= 8192, BLM_ERR_SHAPE, "max_len %d not in (0, 8192]", max_len);
  BLM_REQUIRE(!dropping || max_len <= 128, BLM_ERR_SHAPE, "attention dropout (training) is limited to 128 tokens per "
              "sequence like the backward kernel, got %d", max_len);
  BLM_REQUIRE(out_f32 || out_hi, BLM_ERR_ARG, "no output buffer");
  BLM_REQUIRE(!out_lo || out_hi, BLM_ERR_ARG, "out_lo requires out_hi");
  BLM_REQUIRE((ld % 8) == 0 && ld >= 3ll * nhead * head_dim && (ldo % 2) == 0 && ldo >= (int64_t)nhead * head_dim,
              BLM_ERR_ALIGN, "bad leading dimensions ld=%lld ldo=%lld", (long long)ld, (long long)ldo);
  BLM_REQUIRE(aligned16(qkv_hi) && aligned16(qkv_lo) && aligned16(out_f32) && aligned16(out_hi) && aligned16(out_lo),
              BLM_ERR_ALIGN, "attention pointers must be 16-byte aligned");
  BLM_REQUIRE(nseq * nhead < (1ll << 31), BLM_ERR_SHAPE, "too many (sequence, head) pairs");
  static bool attr_set = false;
  if (!attr_set) {
    int rc = attention_init();
    if (rc != BLM_OK) return rc;
    attr_set = true;
  }
  const long long pairs = nseq * nhead;
  const __nv_bfloat16* qh = reinterpret_cast<const __nv_bfloat16*>(qkv_hi);
  const __nv_bfloat16* ql = reinterpret_cast<const __nv_bfloat16*>(qkv_lo);
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(out_lo);
  cudaStream_t st = as_stream(stream);
  if (dropping) {
    if (max_len <= 32) {
      const unsigned blocks = static_cast<unsigned>((pairs + 3) / 4);
      if (ql)
        mha_causal_mma_kernel<true, 32, true><<<blocks, 128, mma_attn_smem_bytes<true>(), st>>>(
            qh, ql, ld, seq_offsets, pairs, nhead, out_f32, oh, ol, ldo, dpar, ldm);
      else
        mha_causal_mma_kernel<false, 32, true><<<blocks, 128, mma_attn_smem_bytes<false>(), st>>>(
            qh, ql, ld, seq_offsets, pairs, nhead, out_f32, oh, ol, ldo, dpar, ldm);
    } else {
      const unsigned blocks = static_cast<unsigned>(pairs);
      if (ql)
        mha_causal_mma_kernel<true, 128, true><<<blocks, 128, mma_attn_smem_bytes<true>(), st>>>(
            qh, ql, ld, seq_offsets, pairs, nhead, out_f32, oh, ol, ldo, dpar, ldm);
      else
        mha_causal_mma_kernel<false, 128, true><<<blocks, 128, mma_attn_smem_bytes<false>(), st>>>(
            qh, ql, ld, seq_offsets, pairs, nhead, out_f32, oh, ol, ldo, dpar, ldm);
    }
    BLM_CHECK_CUDA(cudaGetLastError());
    return BLM_OK;
  }
  if (max_len <= 32) {
    const unsigned blocks = static_cast<unsigned>((pairs + 3) / 4);
    if (ql)
      mha_causal_mma_kernel<true, 32><<<blocks, 128, mma_attn_smem_bytes<true>(), st>>>(qh, ql, ld, seq_offsets, pairs, nhead,
                                                                                    out_f32, oh, ol, ldo);
    else
      mha_causal_mma_kernel<false, 32><<<blocks, 128, mma_attn_smem_bytes<false>(), st>>>(qh, ql, ld, seq_offsets, pairs,
                                                                                      nhead, out_f32, oh, ol, ldo);
  } else {
    // one CTA per (pair, 128-row query block); blocks past a hypothesis' length exit at once
    const dim3 blocks(static_cast<unsigned>(pairs), static_cast<unsigned>((max_len + 127) / 128));
    if (ql)
      mha_causal_mma_kernel<true, 128><<<blocks, 128, mma_attn_smem_bytes<true>(), st>>>(qh, ql, ld, seq_offsets, pairs, nhead,
                                                                                     out_f32, oh, ol, ldo);
    else
      mha_causal_mma_kernel<false, 128><<<blocks, 128, mma_attn_smem_bytes<false>(), st>>>(qh, ql, ld, seq_offsets, pairs,
                                                                                       nhead, out_f32, oh, ol, ldo);
  }
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

extern "C" int blm_mha_causal_bwd_tc(const float* qkv, int64_t ld, const float* dout, int64_t ldo, const int32_t* seq_offsets,
                                     int64_t nseq, int32_t nhead, int32_t head_dim, int32_t max_len, float q_scale,
                                     int32_t precise, float* dqkv, int64_t ldd, blm_stream stream) {
  return blm_mha_causal_bwd_tc_dropout(qkv, ld, dout, ldo, seq_offsets, nseq, nhead, head_dim, max_len, q_scale, precise,
                                       nullptr, dqkv, ldd, stream);
}

extern "C" int blm_mha_causal_bwd_tc_dropout(const float* qkv, int64_t ld, const float* dout, int64_t ldo,
                                             const int32_t* seq_offsets, int64_t nseq, int32_t nhead, int32_t head_dim,
                                             int32_t max_len, float q_scale, int32_t precise, const blm_dropout_desc* drop,
                                             float* dqkv, int64_t ldd, blm_stream stream) {
  using namespace blm;
  const bool dropping = drop_active(drop);
  if (dropping) {
    BLM_REQUIRE(drop->p >= 0.0f && drop->p < 1.0f, BLM_ERR_ARG, "dropout probability %g not in [0, 1)", drop->p);
    BLM_REQUIRE((reinterpret_cast<uintptr_t>(drop->mask) & 7u) == 0, BLM_ERR_ALIGN, "attention dropout mask must be 8-byte aligned");
  }
  const DropParams dpar = dropping ? make_drop_params(drop->mask, drop->p, drop->seed, drop->seed_dev, drop->stream_id)
                                   : DropParams{};
  const int ldm = (max_len + 3) & ~3;
  BLM_REQUIRE(qkv && dout && seq_offsets && dqkv && nseq > 0 && nhead > 0, BLM_ERR_ARG, "bad attention-backward arguments");
  BLM_REQUIRE(head_dim == kMmaAttnHd, BLM_ERR_SHAPE, "tensor-core attention backward needs head_dim 64, got %d", head_dim);
  BLM_REQUIRE(max_len > 0 && max_len <= 128, BLM_ERR_SHAPE, "max_len %d not in (0, 128]", max_len);
  BLM_REQUIRE((ld % 4) == 0 && (ldo % 4) == 0 && (ldd % 4) == 0 && aligned16(qkv) && aligned16(dout) && aligned16(dqkv),
              BLM_ERR_ALIGN, "attention-backward operands must be 16-byte aligned with leading dimensions %% 4 == 0");
  static bool attr_set = false;
  if (!attr_set) {
    int rc = attention_init();
    if (rc != BLM_OK) return rc;
    attr_set = true;
  }
  const unsigned grid = static_cast<unsigned>(nseq * nhead);
  if (dropping) {
    if (precise)
      mha_causal_bwd_mma_kernel<true, 1, true><<<grid, 256, mma_attn_bwd_smem_bytes<true>(), as_stream(stream)>>>(
          qkv, ld, dout, ldo, seq_offsets, nhead, q_scale, dqkv, ldd, dpar, ldm);
    else
      mha_causal_bwd_mma_kernel<false, 1, true><<<grid, 256, mma_attn_bwd_smem_bytes<false>(), as_stream(stream)>>>(
          qkv, ld, dout, ldo, seq_offsets, nhead, q_scale, dqkv, ldd, dpar, ldm);
    BLM_CHECK_CUDA(cudaGetLastError());
    return BLM_OK;
  }
  if (precise)
    mha_causal_bwd_mma_kernel<true, 1><<<grid, 256, mma_attn_bwd_smem_bytes<true>(), as_stream(stream)>>>(
        qkv, ld, dout, ldo, seq_offsets, nhead, q_scale, dqkv, ldd);
  else
    mha_causal_bwd_mma_kernel<false, 1><<<grid, 256, mma_attn_bwd_smem_bytes<false>(), as_stream(stream)>>>(
        qkv, ld, dout, ldo, seq_offsets, nhead, q_scale, dqkv, ldd);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}
