// HBM-bound support kernels: fp32 -> bf16 (hi, lo) split, embedding (+ positional
// encoding), LayerNorm, reparameterised-weight materialisation with Philox noise,
// and the fused KL(q || N(0,1)) reduction.  All of them are one pass over their
// operands with 128-bit accesses; grids are sized in multiples of the SM count.
#include "blm_host.h"
#include "blm_ptx.cuh"
#include "blm_philox.cuh"

namespace blm {

__device__ __forceinline__ void store_split(float4 x, __nv_bfloat16* hi, __nv_bfloat16* lo, long long i) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y),
                      h2 = __float2bfloat16_rn(x.z), h3 = __float2bfloat16_rn(x.w);
  uint2 hv;
  hv.x = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
  hv.y = static_cast<uint32_t>(__bfloat16_as_ushort(h2)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h3)) << 16);
  *reinterpret_cast<uint2*>(hi + i) = hv;
  if (lo) {
    uint2 lv;
    lv.x = pack_bf16x2(x.x - __bfloat162float(h0), x.y - __bfloat162float(h1));
    lv.y = pack_bf16x2(x.z - __bfloat162float(h2), x.w - __bfloat162float(h3));
    *reinterpret_cast<uint2*>(lo + i) = lv;
  }
}

__global__ void split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo, long long n4, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    store_split(v, hi, lo, i * 4);
  }
  // tail (n not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = n4 * 4; i < n; ++i) {
      const float v = x[i];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[i] = h;
      if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

__global__ void sigma_kernel(const float* __restrict__ lgstd, __nv_bfloat16* __restrict__ out, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __float2bfloat16_rn(expf(lgstd[i]));
}

// one warp per token row; d % 4 == 0
__global__ void embed_kernel(const int* __restrict__ tok, const int* __restrict__ pos,
                             const float* __restrict__ emb, const float* __restrict__ pe, float scale,
                             long long M, int d, float* __restrict__ out_f32,
                             __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (long long m = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
    const float* e = emb + static_cast<long long>(__ldg(tok + m)) * d;
    const float* pr = pe ? pe + static_cast<long long>(__ldg(pos + m)) * d : nullptr;
    for (int c = lane * 4; c < d; c += 128) {
      float4 v = __ldg(reinterpret_cast<const float4*>(e + c));
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      if (pr) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(pr + c));
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      const long long o = m * d + c;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = v;
      if (out_hi) store_split(v, out_hi, out_lo, o);
    }
  }
}

// one warp per row, row cached in registers (d <= 1024), two-pass mean / variance
// exactly like nn.LayerNorm (biased variance, eps inside the rsqrt).
template <int MAXV>  // float4 chunks per lane
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, long long M, int d,
                                 float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi,
                                 __nv_bfloat16* __restrict__ out_lo) {
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const float inv_d = 1.0f / static_cast<float>(d);
  for (long long m = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
    const float* xr = x + m * d;
    float4 v[MAXV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        v[i] = __ldg(reinterpret_cast<const float4*>(xr + c));
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        const float a = v[i].x - mean, b = v[i].y - mean, e = v[i].z - mean, f = v[i].w - mean;
        q += (a * a + b * b) + (e * e + f * f);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < d) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        float4 y;
        y.x = (v[i].x - mean) * rstd * g.x + b.x;
        y.y = (v[i].y - mean) * rstd * g.y + b.y;
        y.z = (v[i].z - mean) * rstd * g.z + b.z;
        y.w = (v[i].w - mean) * rstd * g.w + b.w;
        const long long o = m * d + c;
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = y;
        if (out_hi) store_split(y, out_hi, out_lo, o);
      }
    }
  }
}

__global__ void philox_normal_kernel(uint64_t seed, uint64_t stream, long long n, float scale, float* __restrict__ out) {
  const long long n4 = (n + 3) / 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 z = philox_normal4(seed, stream, static_cast<uint64_t>(i));
    const float zz[4] = {z.x, z.y, z.z, z.w};
    for (int j = 0; j < 4; ++j)
      if (i * 4 + j < n) out[i * 4 + j] = zz[j] * scale;
  }
}

// w = mu + exp(lgstd) * eps over a [rows, cols] view (cols % 4 == 0)
__global__ void reparam_kernel(const float* __restrict__ mu, long long ldmu, const float* __restrict__ lgstd,
                               const float* __restrict__ eps, int eps_mode, uint64_t seed, uint64_t stream,
                               long long rows, long long cols, float* __restrict__ out_f32,
                               __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const long long c4 = cols / 4;
  const long long n4 = rows * c4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long r = i / c4, c = (i - r * c4) * 4;
    float4 w = __ldg(reinterpret_cast<const float4*>(mu + r * ldmu + c));
    if (eps_mode != BLM_EPS_NONE) {
      const float4 ls = __ldg(reinterpret_cast<const float4*>(lgstd + r * cols + c));
      float4 e;
      if (eps_mode == BLM_EPS_PTR) {
        e = __ldg(reinterpret_cast<const float4*>(eps + r * cols + c));
      } else {
        e = philox_normal4(seed, stream, static_cast<uint64_t>(i));  // dense index (r*cols+c)/4 == i
      }
      w.x = reparam_value(w.x, ls.x, e.x);
      w.y = reparam_value(w.y, ls.y, e.y);
      w.z = reparam_value(w.z, ls.z, e.z);
      w.w = reparam_value(w.w, ls.w, e.w);
    }
    const long long o = r * cols + c;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = w;
    if (out_hi) store_split(w, out_hi, out_lo, o);
  }
}

// ---------------------------------------------------------------------- KL
// sum over a [rows, cols] view of  mu^2 - 2 rho + exp(2 rho) [- 1];
// per-thread fp32 partials over 128-bit loads, warp shuffle, one double per
// block, last block folds the block sums (deterministic order) and applies
// scale * 0.5 / (rows * cols).
struct KlWorkspace {
  unsigned int counter;
  unsigned int pad;
  double partial[1024];
};

__global__ void __launch_bounds__(256) kl_kernel(const float* __restrict__ mu, long long ldmu,
                                                 const float* __restrict__ lgstd, long long rows,
                                                 long long cols, int minus_one, float scale,
                                                 int accumulate, float* __restrict__ out,
                                                 KlWorkspace* __restrict__ ws) {
  const long long c4 = cols / 4;
  const long long n4 = rows * c4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float acc0 = 0.0f, acc1 = 0.0f;
  const bool dense = (ldmu == cols);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    long long moff;
    if (dense) {
      moff = i * 4;
    } else {
      const long long r = i / c4;
      moff = r * ldmu + (i - r * c4) * 4;
    }
    const float4 m = __ldg(reinterpret_cast<const float4*>(mu + moff));
    const float4 s = __ldg(reinterpret_cast<const float4*>(lgstd) + i);
    acc0 += (m.x * m.x - 2.0f * s.x + expf(2.0f * s.x)) + (m.y * m.y - 2.0f * s.y + expf(2.0f * s.y));
    acc1 += (m.z * m.z - 2.0f * s.z + expf(2.0f * s.z)) + (m.w * m.w - 2.0f * s.w + expf(2.0f * s.w));
  }
  // scalar tail when cols % 4 != 0 is excluded by the host check
  float v = warp_sum(acc0 + acc1);
  __shared__ float wsum[8];
  __shared__ bool is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) wsum[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int w = 0; w < 8; ++w) b += static_cast<double>(wsum[w]);
    ws->partial[blockIdx.x] = b;
    __threadfence();
    const unsigned int done = atomicAdd(&ws->counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    double t = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32) t += ws->partial[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      const double n = static_cast<double>(rows) * static_cast<double>(cols);
      double mean = t / n;
      if (minus_one) mean -= 1.0;
      const float r = static_cast<float>(0.5 * mean * static_cast<double>(scale));
      out[0] = accumulate ? out[0] + r : r;
      ws->counter = 0;  // ready for the next launch on this stream
    }
  }
}

// Monte-Carlo predictive over K posterior samples: out[m] = -log( 1/K sum_k exp(-nll[k, m]) )
__global__ void mc_combine_kernel(const float* __restrict__ nll, long long K, long long M, float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long m = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; m < M; m += stride) {
    float mn = INFINITY;
    for (long long k = 0; k < K; ++k) mn = fminf(mn, nll[k * M + m]);
    float s = 0.0f;
    for (long long k = 0; k < K; ++k) s += expf(mn - nll[k * M + m]);
    out[m] = mn - logf(s / static_cast<float>(K));
  }
}

// GP-LSTM cell update (GPLSTMCell.Gplstm, model.py:1743-1777, gpnn_type <= 3, gate_type 1..4):
// acc5 [B, 5H] holds the four pre-activations i, f, g, o (W_ih x + b_ih + W_hh h + b_ih -- bias_ih twice, the
// reference's quirk) and, in the fifth block, the GP unit's pre-activation z = W_g [x; h] + b_g.  The gate
// `gate_type` is REPLACED by gp = sum_i coef[i, u] act_i(z), acts = (sigmoid, tanh, relu)[:n_act]
// (model.py:1692-1697); the others keep sigmoid / tanh.  Rows past their length keep (h, c).
__global__ void gp_lstm_cell_kernel(const float* __restrict__ acc5, long long ld, const float* __restrict__ coef,
                                    int n_act, int gate_type, const int* __restrict__ lengths, int t, long long B,
                                    int H, float* __restrict__ c, float* __restrict__ h,
                                    __nv_bfloat16* __restrict__ h_hi, __nv_bfloat16* __restrict__ h_lo,
                                    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi,
                                    __nv_bfloat16* __restrict__ out_lo, const float* __restrict__ c_in = nullptr) {
  // gate_type 0: plain cell update from the four pre-activation blocks (no GP block is read); c_in: the cell state the
  // update starts from, when it is not c itself (GP-LSTM gate type 5: c passed through the GP unit first)
  const long long n = B * H;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long b = i / H;
    const int u = static_cast<int>(i - b * H);
    const bool live = t < __ldg(lengths + b);
    float hv = h[i];
    if (live) {
      const float* a = acc5 + b * ld + u;
      float gp = 0.0f;
      if (gate_type != 0) {
        const float z = a[4ll * H];
        gp = __ldg(coef + u) * (1.0f / (1.0f + expf(-z)));
        if (n_act > 1) gp += __ldg(coef + H + u) * tanhf(z);
        if (n_act > 2) gp += __ldg(coef + 2 * H + u) * fmaxf(z, 0.0f);
      }
      const float ig = gate_type == 1 ? gp : 1.0f / (1.0f + expf(-a[0]));
      const float fg = gate_type == 2 ? gp : 1.0f / (1.0f + expf(-a[H]));
      const float gg = gate_type == 3 ? gp : tanhf(a[2ll * H]);
      const float og = gate_type == 4 ? gp : 1.0f / (1.0f + expf(-a[3ll * H]));
      const float cv = fg * (c_in ? c_in[i] : c[i]) + ig * gg;
      hv = og * tanhf(cv);
      c[i] = cv;
      h[i] = hv;
    }
    const __nv_bfloat16 hh = __float2bfloat16_rn(hv);
    h_hi[i] = hh;
    if (h_lo) h_lo[i] = __float2bfloat16_rn(hv - __bfloat162float(hh));
    const float ov = live ? hv : 0.0f;
    if (out_f32) out_f32[i] = ov;
    if (out_hi) {
      const __nv_bfloat16 oh = __float2bfloat16_rn(ov);
      out_hi[i] = oh;
      if (out_lo) out_lo[i] = __float2bfloat16_rn(ov - __bfloat162float(oh));
    }
  }
}

static int grid_for(long long work_items, int threads, int per_sm) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace blm

extern "C" {

int blm_split_bf16(const float* x, blm_bf16* hi, blm_bf16* lo, int64_t n, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(x && hi && n > 0, BLM_ERR_ARG, "bad split arguments");
  BLM_REQUIRE(aligned16(x) && aligned16(hi) && aligned16(lo), BLM_ERR_ALIGN, "split pointers must be 16-byte aligned");
  const long long n4 = n / 4;
  split_kernel<<<grid_for(n4 > 0 ? n4 : 1, 256, 8), 256, 0, as_stream(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), n4, n);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_sigma_bf16(const float* lgstd, blm_bf16* sigma, int64_t n, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(lgstd && sigma && n > 0, BLM_ERR_ARG, "bad sigma arguments");
  sigma_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(lgstd, reinterpret_cast<__nv_bfloat16*>(sigma), n);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_embed(const int32_t* tokens, const int32_t* pos, const float* emb, const float* pe, float scale,
              int64_t M, int32_t d, float* out_f32, blm_bf16* out_hi, blm_bf16* out_lo,
              blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(tokens && emb && M > 0 && d > 0, BLM_ERR_ARG, "bad embed arguments");
  BLM_REQUIRE(!pe || pos, BLM_ERR_ARG, "positional table without positions");
  BLM_REQUIRE((d % 4) == 0, BLM_ERR_SHAPE, "embedding width %d must be a multiple of 4", d);
  BLM_REQUIRE(out_f32 || out_hi, BLM_ERR_ARG, "no output buffer");
  BLM_REQUIRE(aligned16(emb) && aligned16(pe) && aligned16(out_f32) && aligned16(out_hi) && aligned16(out_lo),
              BLM_ERR_ALIGN, "embed pointers must be 16-byte aligned");
  embed_kernel<<<grid_for(M * 32, 256, 8), 256, 0, as_stream(stream)>>>(
      tokens, pos, emb, pe, scale, M, d, out_f32, reinterpret_cast<__nv_bfloat16*>(out_hi),
      reinterpret_cast<__nv_bfloat16*>(out_lo));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_layernorm(const float* x, const float* gamma, const float* beta, float eps, int64_t M, int32_t d,
                  float* out_f32, blm_bf16* out_hi, blm_bf16* out_lo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(x && gamma && beta && M > 0 && d > 0, BLM_ERR_ARG, "bad layernorm arguments");
  BLM_REQUIRE((d % 4) == 0 && d <= 2048, BLM_ERR_SHAPE, "layernorm width %d must be a multiple of 4 and <= 2048", d);
  BLM_REQUIRE(out_f32 || out_hi, BLM_ERR_ARG, "no output buffer");
  BLM_REQUIRE(aligned16(x) && aligned16(gamma) && aligned16(beta) && aligned16(out_f32) &&
                  aligned16(out_hi) && aligned16(out_lo),
              BLM_ERR_ALIGN, "layernorm pointers must be 16-byte aligned");
  const int grid = grid_for(M * 32, 256, 8);
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(out_lo);
  cudaStream_t st = as_stream(stream);
  if (d <= 512)
    layernorm_kernel<4><<<grid, 256, 0, st>>>(x, gamma, beta, eps, M, d, out_f32, oh, ol);
  else if (d <= 1024)
    layernorm_kernel<8><<<grid, 256, 0, st>>>(x, gamma, beta, eps, M, d, out_f32, oh, ol);
  else
    layernorm_kernel<16><<<grid, 256, 0, st>>>(x, gamma, beta, eps, M, d, out_f32, oh, ol);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_reparam(const float* mu, int64_t ldmu, const float* lgstd, const float* eps, int32_t eps_mode,
                uint64_t seed, uint64_t stream_id, int64_t rows, int64_t cols, float* out_f32,
                blm_bf16* out_hi, blm_bf16* out_lo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(mu && rows > 0 && cols > 0, BLM_ERR_ARG, "bad reparam arguments");
  BLM_REQUIRE(eps_mode == BLM_EPS_NONE || lgstd, BLM_ERR_ARG, "sampling needs lgstd");
  BLM_REQUIRE(eps_mode != BLM_EPS_PTR || eps, BLM_ERR_ARG, "BLM_EPS_PTR needs eps");
  BLM_REQUIRE(eps_mode >= BLM_EPS_NONE && eps_mode <= BLM_EPS_PHILOX, BLM_ERR_ARG, "bad eps_mode %d", eps_mode);
  BLM_REQUIRE((cols % 4) == 0 && (ldmu % 4) == 0 && ldmu >= cols, BLM_ERR_SHAPE,
              "reparam needs cols %% 4 == 0 and ldmu %% 4 == 0 (cols=%lld ldmu=%lld)", (long long)cols,
              (long long)ldmu);
  BLM_REQUIRE(out_f32 || out_hi, BLM_ERR_ARG, "no output buffer");
  BLM_REQUIRE(aligned16(mu) && aligned16(lgstd) && aligned16(eps) && aligned16(out_f32) &&
                  aligned16(out_hi) && aligned16(out_lo),
              BLM_ERR_ALIGN, "reparam pointers must be 16-byte aligned");
  reparam_kernel<<<grid_for(rows * cols / 4, 256, 8), 256, 0, as_stream(stream)>>>(
      mu, ldmu, lgstd, eps, eps_mode, seed, stream_id, rows, cols, out_f32,
      reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_philox_normal_scaled(uint64_t seed, uint64_t stream_id, int64_t n, float scale, float* out, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(out && n > 0, BLM_ERR_ARG, "bad philox arguments");
  philox_normal_kernel<<<grid_for((n + 3) / 4, 256, 8), 256, 0, as_stream(stream)>>>(seed, stream_id, n, scale, out);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_philox_normal(uint64_t seed, uint64_t stream_id, int64_t n, float* out, blm_stream stream) {
  return blm_philox_normal_scaled(seed, stream_id, n, 1.0f, out, stream);
}

int blm_mc_combine(const float* nll, int64_t K, int64_t M, float* out, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(nll && out && K > 0 && M > 0, BLM_ERR_ARG, "bad mc_combine arguments");
  mc_combine_kernel<<<grid_for(M, 256, 8), 256, 0, as_stream(stream)>>>(nll, K, M, out);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int64_t blm_kl_workspace_bytes(void) { return static_cast<int64_t>(sizeof(blm::KlWorkspace)); }

int blm_kl_gauss(const float* mu, int64_t ldmu, const float* lgstd, int64_t rows, int64_t cols,
                 int32_t minus_one, float scale, int32_t accumulate, float* out, void* workspace,
                 blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(mu && lgstd && out && workspace && rows > 0 && cols > 0, BLM_ERR_ARG, "bad KL arguments");
  BLM_REQUIRE((cols % 4) == 0 && (ldmu % 4) == 0 && ldmu >= cols, BLM_ERR_SHAPE,
              "KL needs cols %% 4 == 0 and ldmu %% 4 == 0 (cols=%lld ldmu=%lld)", (long long)cols,
              (long long)ldmu);
  BLM_REQUIRE(aligned16(mu) && aligned16(lgstd) && aligned16(workspace), BLM_ERR_ALIGN,
              "KL pointers must be 16-byte aligned");
  // 4 resident CTAs of 256 threads per SM, each thread 4+ independent 128-bit loads in flight
  int grid = grid_for(rows * cols / 4 / 4 + 1, 256, 4);
  if (grid > 1024) grid = 1024;
  kl_kernel<<<grid, 256, 0, as_stream(stream)>>>(mu, ldmu, lgstd, rows, cols, minus_one, scale, accumulate,
                                                out, reinterpret_cast<KlWorkspace*>(workspace));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_gp_lstm_cell(const float* acc5, int64_t ld, const float* coef, int32_t n_act, int32_t gate_type,
                     const int32_t* lengths, int32_t t, int64_t B, int32_t H, float* c, float* h, blm_bf16* h_hi,
                     blm_bf16* h_lo, float* out_f32, blm_bf16* out_hi, blm_bf16* out_lo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(acc5 && coef && lengths && c && h && h_hi && B > 0 && H > 0 && ld >= 5ll * H, BLM_ERR_ARG,
              "bad gp_lstm_cell arguments");
  BLM_REQUIRE(n_act >= 1 && n_act <= 3 && gate_type >= 1 && gate_type <= 4, BLM_ERR_ARG,
              "gp_lstm_cell: n_act=%d gate_type=%d out of range", n_act, gate_type);
  gp_lstm_cell_kernel<<<grid_for(B * H, 256, 8), 256, 0, as_stream(stream)>>>(
      acc5, ld, coef, n_act, gate_type, lengths, t, B, H, c, h, reinterpret_cast<__nv_bfloat16*>(h_hi),
      reinterpret_cast<__nv_bfloat16*>(h_lo), out_f32, reinterpret_cast<__nv_bfloat16*>(out_hi),
      reinterpret_cast<__nv_bfloat16*>(out_lo));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_lstm_cell_step(const float* acc4, int64_t ld, const float* c_in, const int32_t* lengths, int32_t t, int64_t B,
                       int32_t H, float* c, float* h, blm_bf16* h_hi, blm_bf16* h_lo, float* out_f32, blm_bf16* out_hi,
                       blm_bf16* out_lo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(acc4 && lengths && c && h && h_hi && B > 0 && H > 0 && ld >= 4ll * H, BLM_ERR_ARG,
              "bad lstm_cell_step arguments");
  gp_lstm_cell_kernel<<<grid_for(B * H, 256, 8), 256, 0, as_stream(stream)>>>(
      acc4, ld, nullptr, 0, 0, lengths, t, B, H, c, h, reinterpret_cast<__nv_bfloat16*>(h_hi),
      reinterpret_cast<__nv_bfloat16*>(h_lo), out_f32, reinterpret_cast<__nv_bfloat16*>(out_hi),
      reinterpret_cast<__nv_bfloat16*>(out_lo), c_in);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

}  // extern "C"
