// Training-mode pieces of the Variational / GP LSTM cells (model.py:2426-2579, 1674-1777).
//
// Variational cell (VLSTMCell + VNN, vnn_type 1, training): after every step h <- h + n_t with n_t = e_t exp(rho),
// e_t ~ N(0, 0.1^2) of shape (1, H) shared by the batch rows (model.py:2506-2507, 2557-2577).  The noise is additive
// and does not depend on h, so W_hh (h_{t-1} + n_{t-1}) = W_hh h_{t-1} + W_hh n_{t-1}: the second term is a per-step
// bias row folded into the hoisted input gates, and the persistent recurrence kernel runs unchanged on the pure h.
// These kernels build n_t, add / sum per-timestep rows over the batch, and evaluate the VNN KL
//   KL = mean_{B,H}(h^2 - 2 rho + exp(2 h) - 1) / 2,  h = the pure hidden of the last step (model.py:2545-2551)
// with its gradients.  GP cell: one step of the backward recurrence (gate replaced by the GP mixture).
#include "blm_host.h"
#include "blm_ptx.cuh"

namespace blm {

static int cgrid(long long items, int threads, int per_sm) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms() > 0 ? num_sms() : 148) * per_sm;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

// n[t, j] = e[t, j] * exp(rho[j])
__global__ void vnn_noise_kernel(const float* __restrict__ e, const float* __restrict__ rho, long long T, int H,
                                 float* __restrict__ n) {
  const long long tot = T * H;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < tot; i += stride)
    n[i] = e[i] * __expf(__ldg(rho + (i % H)));
}

// out[g*B + b, :] = x[g*B + b, :] + r[g, :]   (rows are group-major: group = timestep)
__global__ void rowgroup_add_kernel(const float* x, const float* __restrict__ r, long long G, long long B, int W,
                                    float* out_f32, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const long long n4 = G * B * W / 4;
  const int w4 = W / 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long row = i / w4;
    const int c4 = static_cast<int>(i - row * w4);
    const long long g = row / B;
    float4 v = *(reinterpret_cast<const float4*>(x) + i);
    const float4 a = __ldg(reinterpret_cast<const float4*>(r) + g * w4 + c4);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
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

// out[g, :] = sum_b x[g*B + b, :]   (fixed order: deterministic)
__global__ void rowgroup_sum_kernel(const float* __restrict__ x, long long G, long long B, int W, float* __restrict__ out) {
  const long long tot = G * W;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < tot; i += stride) {
    const long long g = i / W;
    const int c = static_cast<int>(i - g * W);
    const float* p = x + g * B * W + c;
    float s = 0.0f;
    for (long long b = 0; b < B; ++b) s += p[b * W];
    out[i] = s;
  }
}

// VNN KL and gradients, one block (B*H is a few 10^4): kl_out[0] += 0.5 (mean(h^2 + exp(2h)) - mean(2 rho) - 1);
// dh[b, j] += kl_scale (h + exp(2h)) / (B H);  drho[j] += -kl_scale / H
__global__ void __launch_bounds__(1024) vnn_kl_kernel(const float* __restrict__ h, const float* __restrict__ rho, long long B,
                                                      int H, float kl_scale, float* __restrict__ kl_out,
                                                      float* __restrict__ dh, float* __restrict__ drho) {
  __shared__ double red[32];
  const long long n = B * H;
  double acc = 0.0;
  const float ch = kl_scale / static_cast<float>(n);
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = h[i];
    const float e2 = __expf(2.0f * v);
    acc += static_cast<double>(v) * v + e2;
    if (dh) dh[i] += ch * (v + e2);
  }
  double racc = 0.0;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    racc += rho[j];
    if (drho) drho[j] -= kl_scale / static_cast<float>(H);
  }
  acc = acc / static_cast<double>(n) - 2.0 * racc / H;
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0 && kl_out) kl_out[0] += static_cast<float>(0.5 * (v - 1.0));
  }
}

// drho[j] += exp(rho[j]) sum_t dn[t, j] e[t, j]
__global__ void vnn_drho_kernel(const float* __restrict__ dn, const float* __restrict__ e, const float* __restrict__ rho,
                                long long T, int H, float* __restrict__ drho) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  float s = 0.0f;
  for (long long t = 0; t < T; ++t) s = fmaf(dn[t * H + j], e[t * H + j], s);
  drho[j] += __expf(rho[j]) * s;
}

// One step of the GP-LSTM backward recurrence (GPLSTMCell.Gplstm, model.py:1743-1777), thread per unit u, fixed-order
// loop over the batch rows (the coefficient gradient is a sum over rows).  acc5 [B, 5H]: pre-activations i, f, g, o
// and the GP unit's z; the gate `gate_type` was REPLACED by gp = sum_k coef[k, u] act_k(z), acts = (sigmoid, tanh,
// relu)[:n_act], so its own linear pre-activation gets no gradient.  dh = dout (+ dh_rec); dc carries dL/dc.
__global__ void gp_lstm_bwd_step_kernel(const float* __restrict__ acc5, long long ld, const float* __restrict__ coef,
                                        int n_act, int gate_type, const float* __restrict__ c_prev,
                                        const float* __restrict__ c_t, const float* __restrict__ dout,
                                        const float* __restrict__ dh_rec, float* __restrict__ dc, int dc_is_zero,
                                        long long B, int H, float* __restrict__ dacc, __nv_bfloat16* __restrict__ dacc_hi,
                                        __nv_bfloat16* __restrict__ dacc_lo, long long ldd, float* __restrict__ dcoef) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= H) return;
  const float k0 = coef[u], k1 = n_act > 1 ? coef[H + u] : 0.0f, k2 = n_act > 2 ? coef[2 * H + u] : 0.0f;
  float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
  for (long long b = 0; b < B; ++b) {
    const float* a = acc5 + b * ld + u;
    const long long i = b * H + u;
    const float z = a[4ll * H];
    const float sg = 1.0f / (1.0f + expf(-z)), th = tanhf(z), rl = fmaxf(z, 0.0f);
    const float gp = k0 * sg + k1 * th + k2 * rl;
    const float dgp = k0 * sg * (1.0f - sg) + k1 * (1.0f - th * th) + k2 * (z > 0.0f ? 1.0f : 0.0f);
    const float ig = gate_type == 1 ? gp : 1.0f / (1.0f + expf(-a[0]));
    const float fg = gate_type == 2 ? gp : 1.0f / (1.0f + expf(-a[H]));
    const float gg = gate_type == 3 ? gp : tanhf(a[2ll * H]);
    const float og = gate_type == 4 ? gp : 1.0f / (1.0f + expf(-a[3ll * H]));
    const float tc = tanhf(c_t[i]);
    const float dh = dout[i] + (dh_rec ? dh_rec[i] : 0.0f);
    const float dct = dh * og * (1.0f - tc * tc) + (dc_is_zero ? 0.0f : dc[i]);
    const float dog = dh * tc, dfg = dct * c_prev[i], dig = dct * gg, dgg = dct * ig;
    dc[i] = dct * fg;
    float d[5];
    d[0] = gate_type == 1 ? 0.0f : dig * ig * (1.0f - ig);
    d[1] = gate_type == 2 ? 0.0f : dfg * fg * (1.0f - fg);
    d[2] = gate_type == 3 ? 0.0f : dgg * (1.0f - gg * gg);
    d[3] = gate_type == 4 ? 0.0f : dog * og * (1.0f - og);
    const float dgate = gate_type == 1 ? dig : gate_type == 2 ? dfg : gate_type == 3 ? dgg : dog;
    d[4] = dgate * dgp;
    g0 = fmaf(dgate, sg, g0);
    g1 = fmaf(dgate, th, g1);
    g2 = fmaf(dgate, rl, g2);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const long long o = b * ldd + static_cast<long long>(k) * H + u;
      dacc[o] = d[k];
      const __nv_bfloat16 hi = __float2bfloat16_rn(d[k]);
      dacc_hi[o] = hi;
      if (dacc_lo) dacc_lo[o] = __float2bfloat16_rn(d[k] - __bfloat162float(hi));
    }
  }
  dcoef[u] += g0;
  if (n_act > 1) dcoef[H + u] += g1;
  if (n_act > 2) dcoef[2 * H + u] += g2;
}

// Backward of the GP unit of the LSTM cells when it stands alone in front of the gates (gate types 5-7):
//   h = sum_i coef[i, n] act_i(z), acts (sigmoid, tanh, relu)   (model.py:1787, 1893-1899)
//   dz = dh . sum_i coef[i, n] act_i'(z),   dcoef[i, n] += sum_m dh[m, n] act_i(z[m, n])
// One block per 32 columns walks all M rows (M = T B of a fine-tune step: thousands at most): fixed order, no atomics.
__global__ void __launch_bounds__(256) gp3_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ z,
                                                      const float* __restrict__ coef, long long ld, long long M, int N,
                                                      float* __restrict__ dz, __nv_bfloat16* __restrict__ dz_hi,
                                                      __nv_bfloat16* __restrict__ dz_lo, float* __restrict__ dcoef) {
  __shared__ float red[8][3][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
  if (n < N) {
    const float c0 = __ldg(coef + n), c1 = __ldg(coef + N + n), c2 = __ldg(coef + 2 * N + n);
    for (long long m = ty; m < M; m += 8) {
      const long long o = m * ld + n;
      const float zz = z[o], d = dh[o];
      const float sg = 1.0f / (1.0f + expf(-zz)), th = tanhf(zz);
      a0 += d * sg;
      a1 += d * th;
      a2 += d * fmaxf(zz, 0.0f);
      const float v = d * (c0 * sg * (1.0f - sg) + c1 * (1.0f - th * th) + (zz > 0.0f ? c2 : 0.0f));
      dz[o] = v;
      if (dz_hi) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        dz_hi[o] = h;
        if (dz_lo) dz_lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
  }
  red[ty][0][tx] = a0;
  red[ty][1][tx] = a1;
  red[ty][2][tx] = a2;
  __syncthreads();
  if (ty < 3 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][ty][tx];
    dcoef[static_cast<long long>(ty) * N + n] += t;
  }
}

}  // namespace blm

extern "C" {

int blm_vnn_noise(const float* e, const float* rho, int64_t T, int32_t H, float* n, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(e && rho && n && T > 0 && H > 0, BLM_ERR_ARG, "bad vnn_noise arguments");
  vnn_noise_kernel<<<cgrid(T * H, 256, 8), 256, 0, as_stream(stream)>>>(e, rho, T, H, n);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_rowgroup_add(const float* x, const float* r, int64_t G, int64_t B, int32_t W, float* out_f32, blm_bf16* out_hi,
                     blm_bf16* out_lo, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(x && r && G > 0 && B > 0 && W > 0 && (W % 4) == 0 && (out_f32 || out_hi), BLM_ERR_ARG, "bad rowgroup_add arguments");
  BLM_REQUIRE(!out_lo || out_hi, BLM_ERR_ARG, "out_lo requires out_hi");
  BLM_REQUIRE(aligned16(x) && aligned16(r) && aligned16(out_f32) && aligned16(out_hi) && aligned16(out_lo), BLM_ERR_ALIGN,
              "rowgroup_add pointers must be 16-byte aligned");
  rowgroup_add_kernel<<<cgrid(G * B * W / 4, 256, 8), 256, 0, as_stream(stream)>>>(
      x, r, G, B, W, out_f32, reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo));
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_rowgroup_sum(const float* x, int64_t G, int64_t B, int32_t W, float* out, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(x && out && G > 0 && B > 0 && W > 0, BLM_ERR_ARG, "bad rowgroup_sum arguments");
  rowgroup_sum_kernel<<<cgrid(G * W, 256, 8), 256, 0, as_stream(stream)>>>(x, G, B, W, out);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_vnn_kl(const float* h, const float* rho, int64_t B, int32_t H, float kl_scale, float* kl_out, float* dh, float* drho,
               blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(h && rho && B > 0 && H > 0, BLM_ERR_ARG, "bad vnn_kl arguments");
  vnn_kl_kernel<<<1, 1024, 0, as_stream(stream)>>>(h, rho, B, H, kl_scale, kl_out, dh, drho);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_vnn_drho(const float* dn, const float* e, const float* rho, int64_t T, int32_t H, float* drho, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(dn && e && rho && drho && T > 0 && H > 0, BLM_ERR_ARG, "bad vnn_drho arguments");
  vnn_drho_kernel<<<(H + 127) / 128, 128, 0, as_stream(stream)>>>(dn, e, rho, T, H, drho);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_gp_lstm_bwd_step(const float* acc5, int64_t ld, const float* coef, int32_t n_act, int32_t gate_type,
                         const float* c_prev, const float* c_t, const float* dout, const float* dh_rec, float* dc,
                         int32_t dc_is_zero, int64_t B, int32_t H, float* dacc, blm_bf16* dacc_hi, blm_bf16* dacc_lo,
                         int64_t ldd, float* dcoef, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(acc5 && coef && c_prev && c_t && dout && dc && dacc && dacc_hi && dcoef && B > 0 && H > 0 && ld >= 5ll * H &&
                  ldd >= 5ll * H, BLM_ERR_ARG, "bad gp_lstm_bwd_step arguments");
  BLM_REQUIRE(n_act >= 1 && n_act <= 3 && gate_type >= 1 && gate_type <= 4, BLM_ERR_ARG,
              "gp_lstm_bwd_step: n_act=%d gate_type=%d out of range", n_act, gate_type);
  gp_lstm_bwd_step_kernel<<<(H + 127) / 128, 128, 0, as_stream(stream)>>>(
      acc5, ld, coef, n_act, gate_type, c_prev, c_t, dout, dh_rec, dc, dc_is_zero, B, H, dacc,
      reinterpret_cast<__nv_bfloat16*>(dacc_hi), reinterpret_cast<__nv_bfloat16*>(dacc_lo), ldd, dcoef);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_gp3_bwd(const float* dh, const float* z, const float* coef, int64_t ld, int64_t M, int32_t N, float* dz,
                blm_bf16* dz_hi, blm_bf16* dz_lo, float* dcoef, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(dh && z && coef && dz && dcoef && M > 0 && N > 0 && ld >= N, BLM_ERR_ARG, "bad gp3_bwd arguments");
  BLM_REQUIRE(!dz_lo || dz_hi, BLM_ERR_ARG, "dz_lo requires dz_hi");
  gp3_bwd_kernel<<<static_cast<unsigned>((N + 31) / 32), 256, 0, as_stream(stream)>>>(
      dh, z, coef, ld, M, N, dz, reinterpret_cast<__nv_bfloat16*>(dz_hi), reinterpret_cast<__nv_bfloat16*>(dz_lo), dcoef);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

}  // extern "C"
