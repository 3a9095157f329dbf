// Backward twins and the optimiser of the fine-tune step (reference train.py:306-438; SURVEY.md 8
// row a19).  The contractions of the backward pass (dgrad, wgrad) run on the tcgen05 GEMM of
// blm_gemm.cu with transposed bf16 operand copies made here; everything in this file is the
// HBM-bound glue around them: transposes, column sums, LayerNorm / attention / activation
// backward, the variational hidden-noise layer, reparameterisation and KL gradients, the embedding
// scatter, the global gradient norm and the SGD-momentum update.  All reductions have a fixed
// order (deterministic) except the embedding scatter (fp32 atomics).
#include "blm_host.h"
#include "blm_ptx.cuh"
#include "blm_philox.cuh"

namespace blm {

static int tgrid(long long items, int threads, int per_sm) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------ transposes
// out[c, r] = bf16 split of x[r, c]; 64 x 64 tiles through shared memory.  Loads: 128 B per warp row
// (fp32: two 4-byte columns per lane, bf16: one bf16x2 per lane); stores: one bf16x2 per lane = 128 B per
// warp row.  dir_hi / dir_lo optionally receive the untransposed bf16 split of the same elements.
template <typename Src>
__global__ void __launch_bounds__(256) transpose_kernel(const Src* __restrict__ x_hi, const Src* __restrict__ x_lo,
                                                        long long ldx, long long R, long long C,
                                                        __nv_bfloat16* __restrict__ out_hi,
                                                        __nv_bfloat16* __restrict__ out_lo, long long ldo,
                                                        __nv_bfloat16* __restrict__ dir_hi = nullptr,
                                                        __nv_bfloat16* __restrict__ dir_lo = nullptr,
                                                        long long ldd = 0) {
  __shared__ float tile[64][65];
  const long long tiles_c = (C + 63) / 64, tiles_r = (R + 63) / 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (long long t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
    const long long r0 = (t / tiles_c) * 64, c0 = (t % tiles_c) * 64;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int rr = ty + 8 * k;
      const long long r = r0 + rr;
      float v0 = 0.0f, v1 = 0.0f;
      int cc0, cc1;
      if constexpr (sizeof(Src) == 4) {
        cc0 = tx;
        cc1 = tx + 32;
        if (r < R) {
          const float* p = reinterpret_cast<const float*>(x_hi) + r * ldx + c0;
          if (c0 + cc0 < C) v0 = p[cc0];
          if (c0 + cc1 < C) v1 = p[cc1];
        }
      } else {
        cc0 = 2 * tx;
        cc1 = 2 * tx + 1;
        if (r < R) {
          const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(x_hi) + r * ldx + c0 + cc0;
          const __nv_bfloat16* pl = x_lo ? reinterpret_cast<const __nv_bfloat16*>(x_lo) + r * ldx + c0 + cc0 : nullptr;
          if (c0 + cc1 < C && ((reinterpret_cast<uintptr_t>(p) & 3u) == 0)) {
            const uint32_t q = *reinterpret_cast<const uint32_t*>(p);
            v0 = __uint_as_float(q << 16);
            v1 = __uint_as_float(q & 0xffff0000u);
            if (pl) {
              const uint32_t l = *reinterpret_cast<const uint32_t*>(pl);
              v0 += __uint_as_float(l << 16);
              v1 += __uint_as_float(l & 0xffff0000u);
            }
          } else {
            if (c0 + cc0 < C) v0 = __bfloat162float(p[0]) + (pl ? __bfloat162float(pl[0]) : 0.0f);
            if (c0 + cc1 < C) v1 = __bfloat162float(p[1]) + (pl ? __bfloat162float(pl[1]) : 0.0f);
          }
        }
      }
      tile[rr][cc0] = v0;
      tile[rr][cc1] = v1;
      if (dir_hi && r < R) {  // the untransposed bf16 split of the same elements, for free
        if (c0 + cc0 < C) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v0);
          dir_hi[r * ldd + c0 + cc0] = h;
          if (dir_lo) dir_lo[r * ldd + c0 + cc0] = __float2bfloat16_rn(v0 - __bfloat162float(h));
        }
        if (c0 + cc1 < C) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v1);
          dir_hi[r * ldd + c0 + cc1] = h;
          if (dir_lo) dir_lo[r * ldd + c0 + cc1] = __float2bfloat16_rn(v1 - __bfloat162float(h));
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int cc = ty + 8 * k;
      const long long c = c0 + cc, r = r0 + 2 * tx;
      if (c < C && r < R) {
        const float a = tile[2 * tx][cc], bq = tile[2 * tx + 1][cc];
        __nv_bfloat16* oh = out_hi + c * ldo + r;
        __nv_bfloat16* ol = out_lo ? out_lo + c * ldo + r : nullptr;
        if (r + 1 < R && ((reinterpret_cast<uintptr_t>(oh) & 3u) == 0)) {
          const uint32_t h = pack_bf16x2(a, bq);
          *reinterpret_cast<uint32_t*>(oh) = h;
          if (ol) *reinterpret_cast<uint32_t*>(ol) = pack_bf16x2(a - __uint_as_float(h << 16), bq - __uint_as_float(h & 0xffff0000u));
        } else {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(a);
          oh[0] = h0;
          if (ol) ol[0] = __float2bfloat16_rn(a - __bfloat162float(h0));
          if (r + 1 < R) {
            const __nv_bfloat16 h1 = __float2bfloat16_rn(bq);
            oh[1] = h1;
            if (ol) ol[1] = __float2bfloat16_rn(bq - __bfloat162float(h1));
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ column sums
// out[n] (+)= scale * sum_m x[m, n].  Grid = (column blocks of 128) x (row splits): a thread owns 4
// adjacent columns (one 128-bit fp32 / 64-bit bf16 load per row), the 8 warps of a CTA take rows
// round-robin inside the CTA's row range, partial sums go to the workspace [row_splits, N] and the
// last CTA of each column block (atomic ticket) folds them in split order -- deterministic, and
// enough CTAs to fill the chip even for N = 512.
struct ColsumWs {
  unsigned int tickets[1024];  // one per column block
};

template <typename Src>
__global__ void __launch_bounds__(256) colsum_kernel(const Src* __restrict__ x_hi, const Src* __restrict__ x_lo,
                                                     long long ldx, long long M, long long N, float scale,
                                                     int accumulate, float* __restrict__ out,
                                                     float* __restrict__ partial, unsigned int* __restrict__ tickets) {
  __shared__ float4 part[8][32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n = (static_cast<long long>(blockIdx.x) * 32 + lane) * 4;
  const long long rows_per = (M + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
    const bool vec = (n + 4 <= N);
    for (long long m = r0 + warp; m < r1; m += 8) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if constexpr (sizeof(Src) == 4) {
        const float* p = reinterpret_cast<const float*>(x_hi) + m * ldx + n;
        if (vec && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(p));
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
          for (int k = 0; k < 4; ++k)
            if (n + k < N) v[k] = p[k];
        }
      } else {
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(x_hi) + m * ldx + n;
        const __nv_bfloat16* pl = x_lo ? reinterpret_cast<const __nv_bfloat16*>(x_lo) + m * ldx + n : nullptr;
        if (vec && ((reinterpret_cast<uintptr_t>(p) & 7u) == 0)) {
          const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
          v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
          v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
          if (pl) {
            const uint2 l = __ldg(reinterpret_cast<const uint2*>(pl));
            v[0] += __uint_as_float(l.x << 16); v[1] += __uint_as_float(l.x & 0xffff0000u);
            v[2] += __uint_as_float(l.y << 16); v[3] += __uint_as_float(l.y & 0xffff0000u);
          }
        } else {
          for (int k = 0; k < 4; ++k)
            if (n + k < N) v[k] = __bfloat162float(p[k]) + (pl ? __bfloat162float(pl[k]) : 0.0f);
        }
      }
      s.x += v[0]; s.y += v[1]; s.z += v[2]; s.w += v[3];
    }
  }
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0) {
    float4 t = part[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      t.x += part[k][lane].x; t.y += part[k][lane].y; t.z += part[k][lane].z; t.w += part[k][lane].w;
    }
    if (n < N) {
      float* dst = partial + static_cast<long long>(blockIdx.y) * N + n;
      dst[0] = t.x;
      if (n + 1 < N) dst[1] = t.y;
      if (n + 2 < N) dst[2] = t.z;
      if (n + 3 < N) dst[3] = t.w;
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) is_last = (atomicAdd(&tickets[blockIdx.x], 1u) == gridDim.y - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int c = threadIdx.x; c < 128; c += 256) {
      const long long col = static_cast<long long>(blockIdx.x) * 128 + c;
      if (col < N) {
        float t = 0.0f;
        for (unsigned int r = 0; r < gridDim.y; ++r) t += partial[static_cast<long long>(r) * N + col];
        t *= scale;
        out[col] = accumulate ? out[col] + t : t;
      }
    }
    if (threadIdx.x == 0) tickets[blockIdx.x] = 0;  // ready for the next launch
  }
}

// ------------------------------------------------------------------ LayerNorm backward
// One warp per row: recompute mean / rstd of the LayerNorm input x, then
//   dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dgamma += dy*xhat;  dbeta += dy.
// Per-CTA partial (dgamma, dbeta) go to the workspace; the second kernel folds them in block order.
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, float eps, long long M,
                                                            int d, float* __restrict__ dx,
                                                            float* __restrict__ partial /* [grid, NP, d] */,
                                                            __nv_bfloat16* __restrict__ dx_hi = nullptr,
                                                            __nv_bfloat16* __restrict__ dx_lo = nullptr, int NP = 2) {
  // NP == 3: a third column sum, of dx itself -- the bias gradient of the projection that feeds this LayerNorm's
  // residual branch (dbias = colsum(dY) and dY = dx) -- leaves with dgamma / dbeta; dx_hi / dx_lo: the bf16 operand
  // copies of dx for the dgrad / wgrad GEMMs that follow (one split launch and one colsum launch less per LayerNorm)
  extern __shared__ float sred[];  // [8 warps][NP][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long warps = static_cast<long long>(gridDim.x) * 8;
  const float inv_d = 1.0f / static_cast<float>(d);
  float4 ag[MAXV], ab[MAXV], ax[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) ag[i] = ab[i] = ax[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long m = static_cast<long long>(blockIdx.x) * 8 + warp; m < M; m += warps) {
    const float* xr = x + m * d;
    const float* dr = dy + m * d;
    float4 v[MAXV], g[MAXV], w[MAXV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        v[i] = __ldg(reinterpret_cast<const float4*>(xr + c));
        w[i] = __ldg(reinterpret_cast<const float4*>(dr + c));
        g[i] = __ldg(reinterpret_cast<const float4*>(gamma + c));
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
        ag[i].x += w[i].x * v[i].x; ag[i].y += w[i].y * v[i].y; ag[i].z += w[i].z * v[i].z; ag[i].w += w[i].w * v[i].w;
        ab[i].x += w[i].x; ab[i].y += w[i].y; ab[i].z += w[i].z; ab[i].w += w[i].w;
        w[i].x *= g[i].x; w[i].y *= g[i].y; w[i].z *= g[i].z; w[i].w *= g[i].w;  // g * dy
        s1 += (w[i].x + w[i].y) + (w[i].z + w[i].w);
        s2 += (w[i].x * v[i].x + w[i].y * v[i].y) + (w[i].z * v[i].z + w[i].w * v[i].w);
      }
    }
    const float m1 = warp_sum(s1) * inv_d, m2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        float4 o;
        o.x = rstd * (w[i].x - m1 - v[i].x * m2);
        o.y = rstd * (w[i].y - m1 - v[i].y * m2);
        o.z = rstd * (w[i].z - m1 - v[i].z * m2);
        o.w = rstd * (w[i].w - m1 - v[i].w * m2);
        *reinterpret_cast<float4*>(dx + m * d + c) = o;
        ax[i].x += o.x; ax[i].y += o.y; ax[i].z += o.z; ax[i].w += o.w;
        if (dx_hi) {
          const uint32_t h0 = pack_bf16x2(o.x, o.y), h1 = pack_bf16x2(o.z, o.w);
          *reinterpret_cast<uint2*>(dx_hi + m * d + c) = make_uint2(h0, h1);
          if (dx_lo) {
            const uint32_t l0 = pack_bf16x2(o.x - __uint_as_float(h0 << 16), o.y - __uint_as_float(h0 & 0xffff0000u));
            const uint32_t l1 = pack_bf16x2(o.z - __uint_as_float(h1 << 16), o.w - __uint_as_float(h1 & 0xffff0000u));
            *reinterpret_cast<uint2*>(dx_lo + m * d + c) = make_uint2(l0, l1);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < d) {
      *reinterpret_cast<float4*>(sred + (warp * NP + 0) * d + c) = ag[i];
      *reinterpret_cast<float4*>(sred + (warp * NP + 1) * d + c) = ab[i];
      if (NP == 3) *reinterpret_cast<float4*>(sred + (warp * NP + 2) * d + c) = ax[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < NP * d; c += 256) {
    const int which = c / d, col = c - which * d;
    float t = 0.0f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += sred[(wv * NP + which) * d + col];
    partial[(static_cast<long long>(blockIdx.x) * NP + which) * d + col] = t;
  }
}

// Fold of the per-CTA partials: a block owns 32 of the 2 d columns, its eight warps take every eighth partial (fixed
// order: deterministic) and meet in shared memory.  (One thread per column walking all ~300 partials was a chain of
// dependent L2 round trips: 25 us for a 12 KB result, 6.5 % of the fine-tune step, profiles/r02a.)
__global__ void __launch_bounds__(256) layernorm_bwd_fold_kernel(const float* __restrict__ partial, int blocks, int d,
                                                                 int accumulate, float* __restrict__ dgamma,
                                                                 float* __restrict__ dbeta, float* __restrict__ dxsum = nullptr,
                                                                 int dxsum_accumulate = 0) {
  __shared__ float red[8][32];
  const int NP = dxsum ? 3 : 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;      // column of the [NP, d] result
  float t0 = 0.0f, t1 = 0.0f;
  if (c < NP * d) {
    const float* p = partial + c;
    int b = warp;
    for (; b + 8 < blocks; b += 16) {
      t0 += p[static_cast<long long>(b) * NP * d];
      t1 += p[static_cast<long long>(b + 8) * NP * d];
    }
    if (b < blocks) t0 += p[static_cast<long long>(b) * NP * d];
  }
  red[warp][lane] = t0 + t1;
  __syncthreads();
  if (warp == 0 && c < NP * d) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    const int which = c / d, col = c - which * d;
    float* out = which == 0 ? dgamma : which == 1 ? dbeta : dxsum;
    const int acc = which == 2 ? dxsum_accumulate : accumulate;
    out[col] = acc ? out[col] + t : t;
  }
}

// ------------------------------------------------------------------ attention backward
// One CTA per (sequence, head), T <= 128, head_dim 64, fp32.  q, k, v, dO of the head are staged in
// shared memory with a row pitch of 68 floats (16-byte aligned rows; a quarter-warp reading the same
// 16-byte column chunk of eight consecutive rows touches 32 distinct banks).
// Phase 1 (warp per query row i): q_i and dO_i live in registers; lane j reads k_j / v_j with 128-bit
// loads and forms s_ij and dP_ij; softmax statistics (m_i, 1/l_i), D_i = sum_j p_ij dP_ij and
// dS_ij = p_ij (dP_ij - D_i) follow; dq_i = sum_j dS_ij k_j is accumulated by two half-warps (even / odd
// j, four columns per lane) and combined with one shuffle.
// Phase 2 (warp per key row j): k_j and v_j live in registers; lane i recomputes p_ij and dS_ij from the
// stored statistics, then dk_j = sum_i dS_ij q_i and dv_j = sum_i p_ij dO_i the same half-warp way.
// Nothing of size T x T is kept.  dq is multiplied by q_scale (model.py:877 scales q after the projection).
constexpr int kBwdHd = 64;
constexpr int kBwdHP = 68;

__device__ __forceinline__ float dot64(const float4 (&a)[16], const float* __restrict__ row) {
  float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float4 b = *reinterpret_cast<const float4*>(row + 4 * c);
    s0 = fmaf(a[c].x, b.x, s0);
    s1 = fmaf(a[c].y, b.y, s1);
    s0 = fmaf(a[c].z, b.z, s0);
    s1 = fmaf(a[c].w, b.w, s1);
  }
  return s0 + s1;
}

__global__ void __launch_bounds__(256) mha_causal_bwd_kernel(const float* __restrict__ qkv, long long ld,
                                                             const float* __restrict__ dout, long long ldo,
                                                             const int* __restrict__ seq_offsets, int nhead,
                                                             int max_len, float q_scale, float* __restrict__ dqkv,
                                                             long long ldd) {
  extern __shared__ __align__(16) float sm[];
  constexpr int HP = kBwdHP;
  const int seq = blockIdx.x / nhead, head = blockIdx.x - seq * nhead;
  const int row0 = seq_offsets[seq];
  const int T = seq_offsets[seq + 1] - row0;
  if (T > max_len) {
    if (threadIdx.x == 0) printf("blm: sequence %d has %d tokens > max_len %d\n", seq, T, max_len);
    __trap();
  }
  const int d = nhead * kBwdHd;
  float* sQ = sm;
  float* sK = sQ + max_len * HP;
  float* sV = sK + max_len * HP;
  float* sG = sV + max_len * HP;          // dO
  float* sM = sG + max_len * HP;          // row max
  float* sL = sM + max_len;               // 1 / row sum
  float* sD = sL + max_len;               // D_i
  float* sP = sD + max_len;               // [8 warps][2][max_len] scratch
  for (int idx = threadIdx.x; idx < T * (kBwdHd / 4); idx += 256) {
    const int t = idx / (kBwdHd / 4), c = (idx - t * (kBwdHd / 4)) * 4;
    const float* base = qkv + static_cast<long long>(row0 + t) * ld + head * kBwdHd + c;
    *reinterpret_cast<float4*>(sQ + t * HP + c) = __ldg(reinterpret_cast<const float4*>(base));
    *reinterpret_cast<float4*>(sK + t * HP + c) = __ldg(reinterpret_cast<const float4*>(base + d));
    *reinterpret_cast<float4*>(sV + t * HP + c) = __ldg(reinterpret_cast<const float4*>(base + 2 * d));
    *reinterpret_cast<float4*>(sG + t * HP + c) =
        __ldg(reinterpret_cast<const float4*>(dout + static_cast<long long>(row0 + t) * ldo + head * kBwdHd + c));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, cl = lane & 15;
  float* pa = sP + (warp * 2) * max_len;      // per-warp scratch rows
  float* pb = pa + max_len;
  // ---- phase 1: per query row
  for (int i = warp; i < T; i += 8) {
    float4 q4[16], g4[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      q4[c] = *reinterpret_cast<const float4*>(sQ + i * HP + 4 * c);
      g4[c] = *reinterpret_cast<const float4*>(sG + i * HP + 4 * c);
    }
    float mx = -INFINITY;
    for (int j = lane; j <= i; j += 32) {
      const float s = dot64(q4, sK + j * HP);
      pa[j] = s;
      pb[j] = dot64(g4, sV + j * HP);
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = lane; j <= i; j += 32) sum += expf(pa[j] - mx);
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float dsum = 0.0f;
    for (int j = lane; j <= i; j += 32) {
      const float p = expf(pa[j] - mx) * inv;
      pa[j] = p;
      dsum = fmaf(p, pb[j], dsum);
    }
    const float D = warp_sum(dsum);
    for (int j = lane; j <= i; j += 32) pa[j] = pa[j] * (pb[j] - D);  // dS_ij
    if (lane == 0) {
      sM[i] = mx;
      sL[i] = inv;
      sD[i] = D;
    }
    __syncwarp();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = half; j <= i; j += 2) {
      const float ds = pa[j];
      const float4 k4 = *reinterpret_cast<const float4*>(sK + j * HP + 4 * cl);
      acc.x = fmaf(ds, k4.x, acc.x);
      acc.y = fmaf(ds, k4.y, acc.y);
      acc.z = fmaf(ds, k4.z, acc.z);
      acc.w = fmaf(ds, k4.w, acc.w);
    }
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
    if (half == 0) {
      float* o = dqkv + static_cast<long long>(row0 + i) * ldd + head * kBwdHd + 4 * cl;
      *reinterpret_cast<float4*>(o) = make_float4(acc.x * q_scale, acc.y * q_scale, acc.z * q_scale, acc.w * q_scale);
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- phase 2: per key row
  for (int j = warp; j < T; j += 8) {
    float4 k4[16], v4[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      k4[c] = *reinterpret_cast<const float4*>(sK + j * HP + 4 * c);
      v4[c] = *reinterpret_cast<const float4*>(sV + j * HP + 4 * c);
    }
    for (int i = j + lane; i < T; i += 32) {
      const float s = dot64(k4, sQ + i * HP);
      const float dp = dot64(v4, sG + i * HP);
      const float p = expf(s - sM[i]) * sL[i];
      pa[i] = p;
      pb[i] = p * (dp - sD[i]);
    }
    __syncwarp();
    float4 ak = make_float4(0.f, 0.f, 0.f, 0.f), av = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = j + half; i < T; i += 2) {
      const float p = pa[i], ds = pb[i];
      const float4 qq = *reinterpret_cast<const float4*>(sQ + i * HP + 4 * cl);
      const float4 gg = *reinterpret_cast<const float4*>(sG + i * HP + 4 * cl);
      ak.x = fmaf(ds, qq.x, ak.x); ak.y = fmaf(ds, qq.y, ak.y); ak.z = fmaf(ds, qq.z, ak.z); ak.w = fmaf(ds, qq.w, ak.w);
      av.x = fmaf(p, gg.x, av.x); av.y = fmaf(p, gg.y, av.y); av.z = fmaf(p, gg.z, av.z); av.w = fmaf(p, gg.w, av.w);
    }
    ak.x += __shfl_xor_sync(0xffffffffu, ak.x, 16); ak.y += __shfl_xor_sync(0xffffffffu, ak.y, 16);
    ak.z += __shfl_xor_sync(0xffffffffu, ak.z, 16); ak.w += __shfl_xor_sync(0xffffffffu, ak.w, 16);
    av.x += __shfl_xor_sync(0xffffffffu, av.x, 16); av.y += __shfl_xor_sync(0xffffffffu, av.y, 16);
    av.z += __shfl_xor_sync(0xffffffffu, av.z, 16); av.w += __shfl_xor_sync(0xffffffffu, av.w, 16);
    if (half == 0) {
      float* o = dqkv + static_cast<long long>(row0 + j) * ldd + head * kBwdHd + 4 * cl;
      *reinterpret_cast<float4*>(o + d) = ak;
      *reinterpret_cast<float4*>(o + 2 * d) = av;
    }
    __syncwarp();
  }
}

static size_t mha_bwd_smem_bytes(int max_len) {
  return sizeof(float) * (4ull * max_len * kBwdHP + 3ull * max_len + 16ull * max_len);
}

// ------------------------------------------------------------------ activation column gradients
// GP mixture: dcoef[i, n] (+)= sum_m dh[m, n] * act_i(z[m, n])   (acts tanh, sigmoid, relu, gelu)
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(256) gpmix_dcoef_kernel(const float* __restrict__ z, const float* __restrict__ dh,
                                                          long long ld, long long M, long long N, int accumulate,
                                                          float* __restrict__ dcoef) {
  __shared__ float part[8][4][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long n0 = static_cast<long long>(blockIdx.x) * 32; n0 < N; n0 += static_cast<long long>(gridDim.x) * 32) {
    const long long n = n0 + tx;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < N) {
      for (long long m = ty; m < M; m += 8) {
        const float zz = z[m * ld + n], g = dh[m * ld + n];
        s[0] += g * tanhf(zz);
        s[1] += g * (1.0f / (1.0f + expf(-zz)));
        s[2] += g * fmaxf(zz, 0.0f);
        s[3] += g * gelu_exact(zz);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) part[ty][i][tx] = s[i];
    __syncthreads();
    if (ty < 4 && n < N) {
      float t = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += part[k][ty][tx];
      float* o = dcoef + ty * N + n;
      *o = accumulate ? *o + t : t;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ variational hidden noise
// forward:  fp = f + e * exp(f * rho[t]),  e = explicit noise or noise_std * Philox normal; row m = b*T + t
__global__ void vnoise_fwd_kernel(const float* __restrict__ f, const float* __restrict__ rho,
                                  const float* __restrict__ eps, int eps_mode, uint64_t seed, uint64_t stream,
                                  float noise_std, const float* __restrict__ resid, long long M, int T, int d,
                                  float* __restrict__ fp) {
  const long long n4 = M * d / 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long m = (i * 4) / d;
    const int c = static_cast<int>(i * 4 - m * d);
    const int t = static_cast<int>(m % T);
    const float4 x = __ldg(reinterpret_cast<const float4*>(f) + i);
    const float4 r = __ldg(reinterpret_cast<const float4*>(rho + static_cast<long long>(t) * d + c));
    float4 e;
    if (eps_mode == BLM_EPS_PTR) {
      e = __ldg(reinterpret_cast<const float4*>(eps) + i);
    } else {
      e = philox_normal4(seed, stream, static_cast<uint64_t>(i));
      e.x *= noise_std; e.y *= noise_std; e.z *= noise_std; e.w *= noise_std;
    }
    float4 o;
    o.x = fmaf(e.x, expf(x.x * r.x), x.x);
    o.y = fmaf(e.y, expf(x.y * r.y), x.y);
    o.z = fmaf(e.z, expf(x.z * r.z), x.z);
    o.w = fmaf(e.w, expf(x.w * r.w), x.w);
    if (resid) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(resid) + i);
      o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
    }
    *(reinterpret_cast<float4*>(fp) + i) = o;
  }
}

// backward + KL: thread per (t, c), fixed-order loop over the B sequences.
//   KL = 0.5 * mean_{T,B,d}( ((1 - mp) fp)^2 - 2 rho + exp(2 rho) ),  klc = kl_scale / (T*B*d)
//   g  = dfp + klc * (1 - mp)^2 fp            (loss gradient w.r.t. fp, KL term included)
//   df = g * (1 + e rho E),  E = exp(f rho);  drho = sum_b g e f E + klc * B * (exp(2 rho) - 1)
//   dmp = -klc * sum_b (1 - mp) fp^2;         klpart[t, c] = sum_b ((1-mp) fp)^2 + B (exp(2 rho) - 2 rho)
__global__ void vnoise_bwd_kernel(const float* __restrict__ dfp, const float* __restrict__ f,
                                  const float* __restrict__ rho, const float* __restrict__ mean_p,
                                  const float* __restrict__ eps, int eps_mode, uint64_t seed, uint64_t stream,
                                  float noise_std, int B, int T, int d, float klc, float* __restrict__ df,
                                  float* __restrict__ drho, float* __restrict__ dmean_p, float* __restrict__ klpart) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * d) return;
  const int t = idx / d, c = idx - t * d;
  const float r = rho[idx], mp = mean_p[idx];
  const float om = 1.0f - mp;
  float a_rho = 0.0f, a_mp = 0.0f, a_kl = 0.0f;
  for (int b = 0; b < B; ++b) {
    const long long o = (static_cast<long long>(b) * T + t) * d + c;
    const float x = f[o];
    float e;
    if (eps_mode == BLM_EPS_PTR) {
      e = eps[o];
    } else {
      const float4 z = philox_normal4(seed, stream, static_cast<uint64_t>(o >> 2));
      const float zz[4] = {z.x, z.y, z.z, z.w};
      e = zz[o & 3] * noise_std;
    }
    const float E = expf(x * r);
    const float y = fmaf(e, E, x);  // fp
    const float g = (dfp ? dfp[o] : 0.0f) + klc * om * om * y;
    df[o] = g * (1.0f + e * r * E);
    a_rho = fmaf(g * e, x * E, a_rho);
    a_mp = fmaf(om * y, y, a_mp);
    a_kl = fmaf(om * y, om * y, a_kl);
  }
  const float e2 = expf(2.0f * r);
  drho[idx] = a_rho + klc * static_cast<float>(B) * (e2 - 1.0f);
  dmean_p[idx] = -klc * a_mp;
  klpart[idx] = a_kl + static_cast<float>(B) * (e2 - 2.0f * r);
}

// ------------------------------------------------------------------ dropout (model.py:116, 1039-1045, 218-221)
// out = x * m (+ resid); m from an explicit multiplier tensor or Philox (blm_philox.cuh: drop_mult*)
__global__ void dropout_kernel(const float* x, long long n4, DropParams dp, const float* __restrict__ resid,
                               float* out_f32 /* may alias x */, __nv_bfloat16* __restrict__ out_hi,
                               __nv_bfloat16* __restrict__ out_lo) {
  const DropParams d = drop_resolve(dp);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = x ? *(reinterpret_cast<const float4*>(x) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 m;
    if (d.mask) {
      m = __ldg(reinterpret_cast<const float4*>(d.mask) + i);
    } else {
      const uint4 w = philox_words4(d.seed, d.stream, static_cast<uint64_t>(i));
      m = make_float4(w.x >= d.thresh ? d.scale : 0.f, w.y >= d.thresh ? d.scale : 0.f,
                      w.z >= d.thresh ? d.scale : 0.f, w.w >= d.thresh ? d.scale : 0.f);
    }
    v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
    if (resid) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(resid) + i);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (out_f32) *(reinterpret_cast<float4*>(out_f32) + i) = v;
    if (out_hi) {
      const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
      *(reinterpret_cast<uint2*>(out_hi) + i) = make_uint2(h0, h1);
      if (out_lo) {
        const uint32_t l0 = pack_bf16x2(v.x - __uint_as_float(h0 << 16), v.y - __uint_as_float(h0 & 0xffff0000u));
        const uint32_t l1 = pack_bf16x2(v.z - __uint_as_float(h1 << 16), v.w - __uint_as_float(h1 & 0xffff0000u));
        *(reinterpret_cast<uint2*>(out_lo) + i) = make_uint2(l0, l1);
      }
    }
  }
}

// ------------------------------------------------------------------ embedding scatter
__global__ void embed_bwd_kernel(const float* __restrict__ dx, const int* __restrict__ tok, float scale, long long M,
                                 int d, float* __restrict__ dE) {
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (long long m = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
    float* e = dE + static_cast<long long>(__ldg(tok + m)) * d;
    for (int c = lane; c < d; c += 32) atomicAdd(e + c, dx[m * d + c] * scale);
  }
}

// ------------------------------------------------------------------ KL / reparameterisation gradients
// dmu += coef * mu;  dlgstd += coef * (exp(2 lgstd) - 1)
__global__ void kl_bwd_kernel(const float* __restrict__ mu, long long ldmu, const float* __restrict__ lgstd,
                              long long rows, long long cols, float coef, float* __restrict__ dmu, long long lddmu,
                              float* __restrict__ dlgstd) {
  const long long n = rows * cols;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long r = i / cols, c = i - r * cols;
    dmu[r * lddmu + c] += coef * mu[r * ldmu + c];
    dlgstd[i] += coef * (expf(2.0f * lgstd[i]) - 1.0f);
  }
}

// G = dL/dW~ : dmu (+)= G ; dlgstd (+)= G * eps * exp(lgstd)
__global__ void reparam_bwd_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ lgstd,
                                   const float* __restrict__ eps, int eps_mode, uint64_t seed, uint64_t stream,
                                   long long rows, long long cols, int accumulate, float* __restrict__ dmu,
                                   long long lddmu, float* __restrict__ dlgstd) {
  const long long c4 = cols / 4, n4 = rows * c4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long r = i / c4, c = (i - r * c4) * 4;
    const float4 g = *reinterpret_cast<const float4*>(G + r * ldg + c);
    const float4 ls = __ldg(reinterpret_cast<const float4*>(lgstd + r * cols + c));
    float4 e;
    if (eps_mode == BLM_EPS_PTR)
      e = __ldg(reinterpret_cast<const float4*>(eps + r * cols + c));
    else
      e = philox_normal4(seed, stream, static_cast<uint64_t>(i));
    float4* pm = reinterpret_cast<float4*>(dmu + r * lddmu + c);
    float4* ps = reinterpret_cast<float4*>(dlgstd + r * cols + c);
    float4 m = accumulate ? *pm : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 s = accumulate ? *ps : make_float4(0.f, 0.f, 0.f, 0.f);
    if (dmu != G || accumulate) {
      m.x += g.x; m.y += g.y; m.z += g.z; m.w += g.w;
      *pm = m;
    }
    s.x += g.x * e.x * __expf(ls.x);
    s.y += g.y * e.y * __expf(ls.y);
    s.z += g.z * e.z * __expf(ls.z);
    s.w += g.w * e.w * __expf(ls.w);
    *ps = s;
  }
}

// ------------------------------------------------------------------ reductions + optimiser
struct ReduceWs {
  unsigned int counter;
  unsigned int pad;
  double partial[1024];
};

// out[0] (+)= scale * sum x   (mode 0)   or   scale * sum x^2   (mode 1)
// VEC: x is 16-byte aligned -- 128-bit loads, two in flight per thread (the 42 M-element gradient norm of the fine-tune
// step ran at 3.1 TB/s on scalar loads); the n % 4 tail goes to one thread.
template <bool VEC>
__global__ void __launch_bounds__(256) reduce_kernel(const float* __restrict__ x, long long n, int mode, float scale,
                                                     int accumulate, float* __restrict__ out,
                                                     ReduceWs* __restrict__ ws) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if constexpr (VEC) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const long long n4 = n >> 2;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    int cnt = 0;
    long long i = tid;
    for (; i + stride < n4; i += 2 * stride) {
      const float4 u = x4[i], v = x4[i + stride];
      if (mode) {
        a0 = fmaf(u.x, u.x, a0), a1 = fmaf(u.y, u.y, a1), a2 = fmaf(u.z, u.z, a2), a3 = fmaf(u.w, u.w, a3);
        a0 = fmaf(v.x, v.x, a0), a1 = fmaf(v.y, v.y, a1), a2 = fmaf(v.z, v.z, a2), a3 = fmaf(v.w, v.w, a3);
      } else {
        a0 += u.x + v.x, a1 += u.y + v.y, a2 += u.z + v.z, a3 += u.w + v.w;
      }
      if (++cnt == 8) {  // bound the fp32 run length (16 values per partial sum)
        acc += static_cast<double>(a0) + static_cast<double>(a1) + static_cast<double>(a2) + static_cast<double>(a3);
        a0 = a1 = a2 = a3 = 0.0f;
        cnt = 0;
      }
    }
    if (i < n4) {
      const float4 u = x4[i];
      if (mode) a0 = fmaf(u.x, u.x, a0), a1 = fmaf(u.y, u.y, a1), a2 = fmaf(u.z, u.z, a2), a3 = fmaf(u.w, u.w, a3);
      else a0 += u.x, a1 += u.y, a2 += u.z, a3 += u.w;
    }
    if (tid == 0)
      for (long long k = n4 << 2; k < n; ++k) a0 += mode ? x[k] * x[k] : x[k];
    acc += static_cast<double>(a0) + static_cast<double>(a1) + static_cast<double>(a2) + static_cast<double>(a3);
  } else {
    float a = 0.0f;
    int cnt = 0;
    for (long long i = tid; i < n; i += stride) {
      const float v = x[i];
      a += mode ? v * v : v;
      if (++cnt == 64) {  // bound the fp32 run length
        acc += static_cast<double>(a);
        a = 0.0f;
        cnt = 0;
      }
    }
    acc += static_cast<double>(a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double wsum[8];
  __shared__ bool is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) wsum[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int w = 0; w < 8; ++w) b += wsum[w];
    ws->partial[blockIdx.x] = b;
    __threadfence();
    is_last = (atomicAdd(&ws->counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    double t = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32) t += ws->partial[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      const float r = static_cast<float>(t * static_cast<double>(scale));
      out[0] = accumulate ? out[0] + r : r;
      ws->counter = 0;
    }
  }
}

// torch.nn.utils.clip_grad_norm_ + SGD(momentum) of train.py:419,466:
//   c = min(1, max_norm / (sqrt(norm_sq) * grad_scale + 1e-6));  g' = c * grad_scale * g
//   v = momentum * v + g';  p -= lr * v
// out_hi / out_lo (optional): the bf16 (hi, lo) operand copies of the UPDATED parameters, so the next step's GEMMs
// need no per-weight split launches.
template <bool VEC>
__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v,
                                    long long n, float lr, float momentum, const float* __restrict__ norm_sq,
                                    float max_norm, float grad_scale, __nv_bfloat16* __restrict__ out_hi,
                                    __nv_bfloat16* __restrict__ out_lo) {
  float c = grad_scale;
  if (norm_sq && max_norm > 0.0f) {
    const float norm = sqrtf(norm_sq[0]) * grad_scale;
    c *= fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  auto one = [&](long long i) {
    const float nv = fmaf(momentum, v[i], c * g[i]);
    v[i] = nv;
    const float np_ = fmaf(-lr, nv, p[i]);
    p[i] = np_;
    if (out_hi) {
      const __nv_bfloat16 h = __float2bfloat16_rn(np_);
      out_hi[i] = h;
      if (out_lo) out_lo[i] = __float2bfloat16_rn(np_ - __bfloat162float(h));
    }
  };
  if constexpr (VEC) {   // 128-bit accesses on the three fp32 streams, 64-bit on the bf16 copies
    const long long n4 = n >> 2;
    for (long long i = tid; i < n4; i += stride) {
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      float4 pv = reinterpret_cast<float4*>(p)[i];
      vv.x = fmaf(momentum, vv.x, c * gv.x), vv.y = fmaf(momentum, vv.y, c * gv.y);
      vv.z = fmaf(momentum, vv.z, c * gv.z), vv.w = fmaf(momentum, vv.w, c * gv.w);
      pv.x = fmaf(-lr, vv.x, pv.x), pv.y = fmaf(-lr, vv.y, pv.y), pv.z = fmaf(-lr, vv.z, pv.z), pv.w = fmaf(-lr, vv.w, pv.w);
      reinterpret_cast<float4*>(v)[i] = vv;
      reinterpret_cast<float4*>(p)[i] = pv;
      if (out_hi) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(pv.x), h1 = __float2bfloat16_rn(pv.y);
        const __nv_bfloat16 h2 = __float2bfloat16_rn(pv.z), h3 = __float2bfloat16_rn(pv.w);
        auto pk = [](__nv_bfloat16 a, __nv_bfloat16 b) {
          return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
        };
        reinterpret_cast<uint2*>(out_hi)[i] = make_uint2(pk(h0, h1), pk(h2, h3));
        if (out_lo) {
          const __nv_bfloat16 l0 = __float2bfloat16_rn(pv.x - __bfloat162float(h0));
          const __nv_bfloat16 l1 = __float2bfloat16_rn(pv.y - __bfloat162float(h1));
          const __nv_bfloat16 l2 = __float2bfloat16_rn(pv.z - __bfloat162float(h2));
          const __nv_bfloat16 l3 = __float2bfloat16_rn(pv.w - __bfloat162float(h3));
          reinterpret_cast<uint2*>(out_lo)[i] = make_uint2(pk(l0, l1), pk(l2, l3));
        }
      }
    }
    if (tid == 0)
      for (long long k = n4 << 2; k < n; ++k) one(k);
  } else {
    for (long long i = tid; i < n; i += stride) one(i);
  }
}

// ---------------------------------------------------------------- LSTM backward (BPTT)
// The forward recurrence kernel (blm_lstm.cu) keeps no gate values.  For the fine-tune step the
// pre-activations Z = gates_x + H_prev W_hh^T are rebuilt by ONE GEMM over all timesteps (H_prev is the
// saved layer output shifted by a step), then this kernel turns Z into the activated gates in place and
// recomputes the cell states: one thread per (row b, unit u) walks t = 0..T-1.
__global__ void lstm_gates_act_kernel(float* __restrict__ gates, const float* __restrict__ c0, int T, int B, int H,
                                      float* __restrict__ c_all) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * H) return;
  const int b = static_cast<int>(idx / H), u = static_cast<int>(idx % H);
  float c = c0[idx];
  for (int t = 0; t < T; ++t) {
    float* z = gates + (static_cast<long long>(t) * B + b) * 4 * H + u;
    const float i = 1.0f / (1.0f + expf(-z[0]));
    const float f = 1.0f / (1.0f + expf(-z[H]));
    const float g = tanhf(z[2 * H]);
    const float o = 1.0f / (1.0f + expf(-z[3 * H]));
    c = fmaf(f, c, i * g);
    z[0] = i, z[H] = f, z[2 * H] = g, z[3 * H] = o;
    c_all[(static_cast<long long>(t) * B + b) * H + u] = c;
  }
}

// One step of the backward recurrence (gate order i,f,g,o):
//   dh = dout_t + dh_rec;  do = dh tanh(c_t) o(1-o);  dc' = dc + dh o (1 - tanh(c_t)^2)
//   di = dc' g i(1-i);  df = dc' c_{t-1} f(1-f);  dg = dc' i (1-g^2);  dc <- dc' f
// dgates leave as fp32 (for the weight gradients) and as bf16 hi[, lo] (A operand of dh_rec = dgates W_hh).
__global__ void lstm_bwd_step_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                     const float* __restrict__ c_t, const float* __restrict__ dout,
                                     const float* __restrict__ dh_rec, float* __restrict__ dc, int dc_is_zero, int B, int H,
                                     float* __restrict__ dg32, __nv_bfloat16* __restrict__ dg_hi,
                                     __nv_bfloat16* __restrict__ dg_lo) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * H) return;
  const int b = static_cast<int>(idx / H), u = static_cast<int>(idx % H);
  const long long row = static_cast<long long>(b) * 4 * H + u;
  const float i = gates[row], f = gates[row + H], g = gates[row + 2 * H], o = gates[row + 3 * H];
  const float dh = dout[idx] + (dh_rec ? dh_rec[idx] : 0.0f);
  const float tc = tanhf(c_t[idx]);
  const float dcp = (dc_is_zero ? 0.0f : dc[idx]) + dh * o * (1.0f - tc * tc);
  float d[4];
  d[0] = dcp * g * i * (1.0f - i);
  d[1] = dcp * c_prev[idx] * f * (1.0f - f);
  d[2] = dcp * i * (1.0f - g * g);
  d[3] = dh * tc * o * (1.0f - o);
  dc[idx] = dcp * f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    dg32[row + k * H] = d[k];
    const __nv_bfloat16 h = __float2bfloat16_rn(d[k]);
    dg_hi[row + k * H] = h;
    if (dg_lo) dg_lo[row + k * H] = __float2bfloat16_rn(d[k] - __bfloat162float(h));
  }
}

int train_init() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(mha_causal_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(mha_bwd_smem_bytes(128))));
  return BLM_OK;
}

}  // namespace blm

static int colsum_row_splits(int64_t M, int64_t N) {
  const int64_t col_blocks = (N + 127) / 128;
  // wide matrices (the [tokens, V] logit gradient: 235 column blocks) need more CTAs per SM to cover the latency of
  // their one-load-per-row loops: 2 row splits ran at 1.45 TB/s
  const int64_t per_sm = col_blocks >= 64 ? 8 : 2;
  int64_t splits = (per_sm * (blm::num_sms() > 0 ? blm::num_sms() : 148) + col_blocks - 1) / col_blocks;
  const int64_t max_by_rows = (M + 31) / 32;  // at least 32 rows per CTA
  if (splits > max_by_rows) splits = max_by_rows;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  return static_cast<int>(splits);
}

template <typename Src>
static int colsum_launch(const Src* hi, const Src* lo, int64_t ld, int64_t M, int64_t N, float scale, int32_t accumulate,
                         float* out, void* workspace, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(hi && out && workspace && M > 0 && N > 0 && ld >= N, BLM_ERR_ARG, "bad colsum arguments");
  BLM_REQUIRE((N + 127) / 128 <= 1024, BLM_ERR_SHAPE, "colsum: N=%lld too wide", (long long)N);
  ColsumWs* ws = reinterpret_cast<ColsumWs*>(workspace);
  float* partial = reinterpret_cast<float*>(ws + 1);
  const dim3 grid(static_cast<unsigned>((N + 127) / 128), static_cast<unsigned>(colsum_row_splits(M, N)));
  colsum_kernel<Src><<<grid, 256, 0, as_stream(stream)>>>(hi, lo, ld, M, N, scale, accumulate, out, partial, ws->tickets);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

extern "C" {

int blm_transpose_split(const float* x, int64_t ldx, int64_t R, int64_t C, blm_bf16* out_hi, blm_bf16* out_lo,
                        int64_t ldo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(x && out_hi && R > 0 && C > 0 && ldx >= C && ldo >= R, BLM_ERR_ARG, "bad transpose arguments");
  transpose_kernel<float><<<tgrid(((R + 63) / 64) * ((C + 63) / 64), 1, 8), 256, 0, as_stream(stream)>>>(
      x, nullptr, ldx, R, C, reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo), ldo);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_split_transpose(const float* x, int64_t ldx, int64_t R, int64_t C, blm_bf16* hi, blm_bf16* lo, int64_t ld,
                        blm_bf16* t_hi, blm_bf16* t_lo, int64_t ldt, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(x && hi && t_hi && R > 0 && C > 0 && ldx >= C && ld >= C && ldt >= R, BLM_ERR_ARG,
              "bad split_transpose arguments");
  transpose_kernel<float><<<tgrid(((R + 63) / 64) * ((C + 63) / 64), 1, 8), 256, 0, as_stream(stream)>>>(
      x, nullptr, ldx, R, C, reinterpret_cast<__nv_bfloat16*>(t_hi), reinterpret_cast<__nv_bfloat16*>(t_lo), ldt,
      reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), ld);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_transpose_bf16(const blm_bf16* hi, const blm_bf16* lo, int64_t ld, int64_t R, int64_t C, blm_bf16* out_hi,
                       blm_bf16* out_lo, int64_t ldo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(hi && out_hi && R > 0 && C > 0 && ld >= C && ldo >= R, BLM_ERR_ARG, "bad transpose arguments");
  transpose_kernel<__nv_bfloat16><<<tgrid(((R + 63) / 64) * ((C + 63) / 64), 1, 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(hi), reinterpret_cast<const __nv_bfloat16*>(lo), ld, R, C,
      reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo), ldo);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int64_t blm_colsum_workspace_bytes(int64_t M, int64_t N) {
  (void)M;
  return static_cast<int64_t>(sizeof(blm::ColsumWs)) + 64 * N * static_cast<int64_t>(sizeof(float));
}

int blm_colsum(const float* x, int64_t ldx, int64_t M, int64_t N, float scale, int32_t accumulate, float* out,
               void* workspace, blm_stream stream) {
  return colsum_launch<float>(x, nullptr, ldx, M, N, scale, accumulate, out, workspace, stream);
}

int blm_colsum_bf16(const blm_bf16* hi, const blm_bf16* lo, int64_t ld, int64_t M, int64_t N, float scale,
                    int32_t accumulate, float* out, void* workspace, blm_stream stream) {
  return colsum_launch<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(hi), reinterpret_cast<const __nv_bfloat16*>(lo),
                                      ld, M, N, scale, accumulate, out, workspace, stream);
}

static int ln_bwd_blocks(int64_t M) {
  const int cap = (blm::num_sms() > 0 ? blm::num_sms() : 148) * 2;
  const int64_t need = (M + 7) / 8;
  return static_cast<int>(need < cap ? need : cap);
}

int64_t blm_layernorm_bwd_workspace_bytes(int64_t M, int32_t d) {
  (void)M;
  return static_cast<int64_t>(148 * 2 * 2) * d * static_cast<int64_t>(sizeof(float)) * 2;
}

int blm_layernorm_bwd(const float* dy, const float* x, const float* gamma, float eps, int64_t M, int32_t d,
                      float* dx, float* dgamma, float* dbeta, int32_t accumulate, void* workspace,
                      blm_stream stream) {
  return blm_layernorm_bwd_ex(dy, x, gamma, eps, M, d, dx, nullptr, nullptr, dgamma, dbeta, accumulate, nullptr, 0, workspace,
                              stream);
}

int blm_layernorm_bwd_ex(const float* dy, const float* x, const float* gamma, float eps, int64_t M, int32_t d, float* dx,
                         blm_bf16* dx_hi, blm_bf16* dx_lo, float* dgamma, float* dbeta, int32_t accumulate, float* dxsum,
                         int32_t dxsum_accumulate, void* workspace, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(dy && x && gamma && dx && dgamma && dbeta && workspace && M > 0, BLM_ERR_ARG, "bad layernorm_bwd arguments");
  BLM_REQUIRE((d % 4) == 0 && d > 0 && d <= 1024, BLM_ERR_SHAPE, "layernorm_bwd width %d must be a multiple of 4 and <= 1024", d);
  BLM_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(gamma) && aligned16(dx) && aligned16(workspace) &&
                  (reinterpret_cast<uintptr_t>(dx_hi) & 7u) == 0 && (reinterpret_cast<uintptr_t>(dx_lo) & 7u) == 0,
              BLM_ERR_ALIGN, "layernorm_bwd pointers must be 16-byte aligned (bf16 copies: 8-byte)");
  BLM_REQUIRE(!dx_lo || dx_hi, BLM_ERR_ARG, "dx_lo requires dx_hi");
  const int blocks = ln_bwd_blocks(M);
  float* partial = reinterpret_cast<float*>(workspace);
  cudaStream_t st = as_stream(stream);
  const int NP = dxsum ? 3 : 2;
  const size_t smem = sizeof(float) * 8 * NP * d;
  static bool attr = false;
  if (!attr) {
    BLM_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
    BLM_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
    attr = true;
  }
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(dx_hi);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(dx_lo);
  if (d <= 512)
    layernorm_bwd_kernel<4><<<blocks, 256, smem, st>>>(dy, x, gamma, eps, M, d, dx, partial, hi, lo, NP);
  else
    layernorm_bwd_kernel<8><<<blocks, 256, smem, st>>>(dy, x, gamma, eps, M, d, dx, partial, hi, lo, NP);
  BLM_CHECK_CUDA(cudaGetLastError());
  layernorm_bwd_fold_kernel<<<(NP * d + 31) / 32, 256, 0, st>>>(partial, blocks, d, accumulate, dgamma, dbeta, dxsum,
                                                               dxsum_accumulate);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_mha_causal_bwd(const float* qkv, int64_t ld, const float* dout, int64_t ldo, const int32_t* seq_offsets,
                       int64_t nseq, int32_t nhead, int32_t head_dim, int32_t max_len, float q_scale, float* dqkv,
                       int64_t ldd, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(qkv && dout && seq_offsets && dqkv && nseq > 0 && nhead > 0, BLM_ERR_ARG, "bad attention-backward arguments");
  BLM_REQUIRE(head_dim == kBwdHd, BLM_ERR_SHAPE, "attention backward needs head_dim 64, got %d", head_dim);
  BLM_REQUIRE(max_len > 0 && max_len <= 128, BLM_ERR_SHAPE, "max_len %d not in (0, 128]", max_len);
  BLM_REQUIRE((ld % 4) == 0 && (ldo % 4) == 0 && (ldd % 4) == 0 && aligned16(qkv) && aligned16(dout) && aligned16(dqkv),
              BLM_ERR_ALIGN, "attention-backward operands must be 16-byte aligned with leading dimensions %% 4 == 0");
  static bool attr = false;
  if (!attr) {
    int rc = train_init();
    if (rc != BLM_OK) return rc;
    attr = true;
  }
  mha_causal_bwd_kernel<<<static_cast<unsigned>(nseq * nhead), 256, mha_bwd_smem_bytes(max_len), as_stream(stream)>>>(
      qkv, ld, dout, ldo, seq_offsets, nhead, max_len, q_scale, dqkv, ldd);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_gpmix_dcoef(const float* z, const float* dh, int64_t ld, int64_t M, int64_t N, int32_t accumulate,
                    float* dcoef, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(z && dh && dcoef && M > 0 && N > 0 && ld >= N, BLM_ERR_ARG, "bad gpmix_dcoef arguments");
  gpmix_dcoef_kernel<<<tgrid((N + 31) / 32, 1, 8), 256, 0, as_stream(stream)>>>(z, dh, ld, M, N, accumulate, dcoef);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_vnoise_fwd(const float* f, const float* rho, const float* eps, int32_t eps_mode, uint64_t seed,
                   uint64_t stream_id, float noise_std, const float* resid, int64_t B, int32_t T, int32_t d,
                   float* fp, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(f && rho && fp && B > 0 && T > 0 && d > 0 && (d % 4) == 0, BLM_ERR_ARG, "bad vnoise arguments");
  BLM_REQUIRE(eps_mode == BLM_EPS_PHILOX || (eps_mode == BLM_EPS_PTR && eps), BLM_ERR_ARG, "bad eps_mode %d", eps_mode);
  vnoise_fwd_kernel<<<tgrid(B * T * d / 4, 256, 8), 256, 0, as_stream(stream)>>>(f, rho, eps, eps_mode, seed, stream_id,
                                                                              noise_std, resid, B * T, T, d, fp);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_vnoise_bwd(const float* dfp, const float* f, const float* rho, const float* mean_p, const float* eps,
                   int32_t eps_mode, uint64_t seed, uint64_t stream_id, float noise_std, int64_t B, int32_t T,
                   int32_t d, float kl_scale, float* df, float* drho, float* dmean_p, float* klpart,
                   blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(f && rho && mean_p && df && drho && dmean_p && klpart && B > 0 && T > 0 && d > 0, BLM_ERR_ARG,
              "bad vnoise_bwd arguments");
  BLM_REQUIRE(eps_mode == BLM_EPS_PHILOX || (eps_mode == BLM_EPS_PTR && eps), BLM_ERR_ARG, "bad eps_mode %d", eps_mode);
  const float klc = kl_scale / (static_cast<float>(B) * static_cast<float>(T) * static_cast<float>(d));
  vnoise_bwd_kernel<<<(T * d + 127) / 128, 128, 0, as_stream(stream)>>>(dfp, f, rho, mean_p, eps, eps_mode, seed, stream_id,
                                                                      noise_std, static_cast<int>(B), T, d, klc, df, drho,
                                                                      dmean_p, klpart);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_dropout(const float* x, int64_t n, const blm_dropout_desc* drop, const float* resid, float* out_f32,
                blm_bf16* out_hi, blm_bf16* out_lo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(drop && n > 0 && (n % 4) == 0 && (out_f32 || out_hi), BLM_ERR_ARG, "bad dropout arguments");
  BLM_REQUIRE(drop->p >= 0.0f && drop->p < 1.0f, BLM_ERR_ARG, "dropout probability %g not in [0, 1)", drop->p);
  BLM_REQUIRE(!out_lo || out_hi, BLM_ERR_ARG, "out_lo requires out_hi");
  BLM_REQUIRE(aligned16(x) && aligned16(drop->mask) && aligned16(resid) && aligned16(out_f32) && aligned16(out_hi) &&
                  aligned16(out_lo), BLM_ERR_ALIGN, "dropout pointers must be 16-byte aligned");
  const DropParams dp = make_drop_params(drop->mask, drop->p, drop->seed, drop->seed_dev, drop->stream_id);
  dropout_kernel<<<tgrid(n / 4, 256, 8), 256, 0, as_stream(stream)>>>(x, n / 4, dp, resid, out_f32,
                                                                       reinterpret_cast<__nv_bfloat16*>(out_hi),
                                                                       reinterpret_cast<__nv_bfloat16*>(out_lo));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_embed_bwd(const float* dx, const int32_t* tokens, float scale, int64_t M, int32_t d, float* dE,
                  blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(dx && tokens && dE && M > 0 && d > 0, BLM_ERR_ARG, "bad embed_bwd arguments");
  embed_bwd_kernel<<<tgrid(M * 32, 256, 8), 256, 0, as_stream(stream)>>>(dx, tokens, scale, M, d, dE);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_kl_gauss_bwd(const float* mu, int64_t ldmu, const float* lgstd, int64_t rows, int64_t cols, float scale,
                     float* dmu, int64_t lddmu, float* dlgstd, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(mu && lgstd && dmu && dlgstd && rows > 0 && cols > 0, BLM_ERR_ARG, "bad kl_bwd arguments");
  const float coef = scale / (static_cast<float>(rows) * static_cast<float>(cols));
  kl_bwd_kernel<<<tgrid(rows * cols, 256, 8), 256, 0, as_stream(stream)>>>(mu, ldmu, lgstd, rows, cols, coef, dmu, lddmu,
                                                                        dlgstd);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_reparam_bwd(const float* G, int64_t ldg, const float* lgstd, const float* eps, int32_t eps_mode,
                    uint64_t seed, uint64_t stream_id, int64_t rows, int64_t cols, int32_t accumulate, float* dmu,
                    int64_t lddmu, float* dlgstd, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(G && lgstd && dmu && dlgstd && rows > 0 && cols > 0, BLM_ERR_ARG, "bad reparam_bwd arguments");
  BLM_REQUIRE(eps_mode == BLM_EPS_PHILOX || (eps_mode == BLM_EPS_PTR && eps), BLM_ERR_ARG, "bad eps_mode %d", eps_mode);
  BLM_REQUIRE((cols % 4) == 0 && (ldg % 4) == 0 && (lddmu % 4) == 0, BLM_ERR_SHAPE, "reparam_bwd needs cols, ld %% 4 == 0");
  BLM_REQUIRE(aligned16(G) && aligned16(lgstd) && aligned16(eps) && aligned16(dmu) && aligned16(dlgstd), BLM_ERR_ALIGN,
              "reparam_bwd pointers must be 16-byte aligned");
  reparam_bwd_kernel<<<tgrid(rows * cols / 4, 256, 8), 256, 0, as_stream(stream)>>>(
      G, ldg, lgstd, eps, eps_mode, seed, stream_id, rows, cols, accumulate, dmu, lddmu, dlgstd);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int64_t blm_reduce_workspace_bytes(void) { return static_cast<int64_t>(sizeof(blm::ReduceWs)); }

int blm_reduce(const float* x, int64_t n, int32_t squares, float scale, int32_t accumulate, float* out,
               void* workspace, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(x && out && workspace && n > 0, BLM_ERR_ARG, "bad reduce arguments");
  int grid = tgrid(n / 4 + 1, 256, 4);
  if (grid > 1024) grid = 1024;
  if (aligned16(x) && n >= 4)
    reduce_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(x, n, squares, scale, accumulate, out,
                                                             reinterpret_cast<ReduceWs*>(workspace));
  else
    reduce_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(x, n, squares, scale, accumulate, out,
                                                              reinterpret_cast<ReduceWs*>(workspace));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_sgd_momentum(float* p, const float* g, float* v, int64_t n, float lr, float momentum, const float* norm_sq,
                     float max_norm, float grad_scale, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(p && g && v && n > 0, BLM_ERR_ARG, "bad sgd arguments");
  if (aligned16(p) && aligned16(g) && aligned16(v))
    sgd_momentum_kernel<true><<<tgrid(n / 4 + 1, 256, 8), 256, 0, as_stream(stream)>>>(p, g, v, n, lr, momentum, norm_sq,
                                                                                      max_norm, grad_scale, nullptr, nullptr);
  else
    sgd_momentum_kernel<false><<<tgrid(n, 256, 8), 256, 0, as_stream(stream)>>>(p, g, v, n, lr, momentum, norm_sq, max_norm,
                                                                                grad_scale, nullptr, nullptr);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_sgd_momentum_split(float* p, const float* g, float* v, int64_t n, float lr, float momentum, const float* norm_sq,
                           float max_norm, float grad_scale, blm_bf16* out_hi, blm_bf16* out_lo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(p && g && v && out_hi && n > 0, BLM_ERR_ARG, "bad sgd arguments");
  const bool vec = aligned16(p) && aligned16(g) && aligned16(v) && (reinterpret_cast<uintptr_t>(out_hi) & 7u) == 0 &&
                   (reinterpret_cast<uintptr_t>(out_lo) & 7u) == 0;
  if (vec)
    sgd_momentum_kernel<true><<<tgrid(n / 4 + 1, 256, 8), 256, 0, as_stream(stream)>>>(
        p, g, v, n, lr, momentum, norm_sq, max_norm, grad_scale, reinterpret_cast<__nv_bfloat16*>(out_hi),
        reinterpret_cast<__nv_bfloat16*>(out_lo));
  else
    sgd_momentum_kernel<false><<<tgrid(n, 256, 8), 256, 0, as_stream(stream)>>>(
        p, g, v, n, lr, momentum, norm_sq, max_norm, grad_scale, reinterpret_cast<__nv_bfloat16*>(out_hi),
        reinterpret_cast<__nv_bfloat16*>(out_lo));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_lstm_gates_act(float* gates, const float* c0, int64_t T, int64_t B, int64_t H, float* c_all, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(gates && c0 && c_all && T > 0 && B > 0 && H > 0, BLM_ERR_ARG, "bad lstm_gates_act arguments");
  lstm_gates_act_kernel<<<static_cast<unsigned>((B * H + 127) / 128), 128, 0, as_stream(stream)>>>(
      gates, c0, static_cast<int>(T), static_cast<int>(B), static_cast<int>(H), c_all);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_lstm_bwd_step(const float* gates_t, const float* c_prev, const float* c_t, const float* dout_t,
                      const float* dh_rec, float* dc, int32_t dc_is_zero, int64_t B, int64_t H, float* dgates_f32,
                      blm_bf16* dgates_hi, blm_bf16* dgates_lo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(gates_t && c_prev && c_t && dout_t && dc && dgates_f32 && dgates_hi && B > 0 && H > 0, BLM_ERR_ARG,
              "bad lstm_bwd_step arguments");
  lstm_bwd_step_kernel<<<static_cast<unsigned>((B * H + 255) / 256), 256, 0, as_stream(stream)>>>(
      gates_t, c_prev, c_t, dout_t, dh_rec, dc, dc_is_zero, static_cast<int>(B), static_cast<int>(H), dgates_f32,
      reinterpret_cast<__nv_bfloat16*>(dgates_hi), reinterpret_cast<__nv_bfloat16*>(dgates_lo));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

}  // extern "C"
