// Persistent LSTM recurrence for sm_100a  (SURVEY.md 8 row a6 / north_star (b)).
//
// One cooperative launch runs all T steps of one layer for a batch of B sequences advanced in
// lock step.  CTA (j, bi) owns hidden units [U j, U j + U) and batch rows block bi: its 4U rows of
// W_hh (4 gates x U units, bf16 hi [+ lo]) are TMA-loaded ONCE into 128B-swizzled shared memory
// and stay resident for every timestep -- HBM never sees W_hh again.  U = 16 (128 KB of W) with the
// batch split in two in bf16 mode, U = 8 (hi + lo = 128 KB) and no split in the precise mode: what a
// CTA must ingest per step is h_{t-1} of its batch block, so the 2-D split halves the L2 -> SM
// traffic that bounds the step (every CTA reading all of h: 101 us per step at B = 2048, r01k).
// Per step each CTA
//   * streams h_{t-1} [rows, H] (bf16 hi [+ lo], double-buffered in global, L2 resident) through a
//     TMA ring and issues tcgen05.mma  D[128 x 4U] += h_tile[128 x 64] . W_slice[4U x 64]^T into
//     one 4U-column TMEM accumulator per 128-row batch tile;
//   * epilogue warps read the accumulator (one batch row per thread), add the hoisted input
//     projection gates_x[t], apply the four gate nonlinearities (i, f, o = sigmoid, g = tanh),
//     update c and h for their 8 units, and publish h_t (fp32 state, layer output, bf16 operand
//     for the next step);
//   * a grid-wide barrier (one atomic per CTA, bounded spin) separates the steps.
// Rows past their own length (right padding) keep (h, c) untouched, so the final state is the
// state after each row's last valid token (the hidden carry of score.py:271-274).
//
// Kernels in this file:
//   lstm_layer_kernel<U, CL, SUB>  the form described above; used for batches of up to two 128-row tiles and for the
//       hypothesis-#0 chains (one row per session), where it loads only the real rows, all K blocks at once, and
//       both epilogue warp sets share the one tile (7.9 us per step at 12 rows);
//   lstm_pair_kernel<U, SUB>       from three tiles up: cta_group::2 pairs (M = 256), per-tile step flags instead of the
//       grid barrier, gate operands requested one 8-unit group ahead (20.9 us per step at B = 2048 in bf16 mode with
//       U = 16; 97.6 us in precise mode with U = 8 and hi + lo slices resident).
// The input projection gates_x and the running cell state are addressed in 32-row blocks (rows32_f4): one batch row per
// thread is what the tensor-memory accumulator dictates, and row-major fp32 would make every warp access 32 lines.
#include <stdlib.h>
#include <string.h>

#include "blm_host.h"
#include "blm_ptx.cuh"

namespace blm {

constexpr int kLStages = 5;      // h-tile ring depth with one K block per stage (16 KB stages)
constexpr int kLStages2 = 3;     // ... with two K blocks per stage (32 KB stages): fewer, larger round trips
constexpr int kLThreads = 384;   // 4 control warps + 8 epilogue warps (two per scheduler, alternate tiles)
constexpr int kLMaxTiles = 16;   // per CTA: 16 x 32 (U = 8) or 8 x 64 (U = 16) TMEM columns = 512
constexpr int kLABytes = 128 * 64 * 2;

struct LstmParams {
  CUtensorMap tmH[2][2];  // [buffer][hi, lo] : h state [B, H] bf16, box 128 x 64 (box 32 x 64 in the cluster kernel)
  CUtensorMap tmW[2];     // [hi, lo]         : W_hh [4H, H] bf16, box U x 64
  int a_box_bytes;        // bytes one h K block brings: box rows x 128 (a one-tile batch loads only its real rows)
  int whole_k;            // 1: small one-tile batch -- ALL K blocks of h_{t-1} are loaded at once, packed at a_box_bytes
                          // stride, and the MMAs run back to back: no ring round trips inside a step
  int kb_stagger;         // 1: stagger the K-block order per CTA (experiment, BLM_LSTM_STAGGER=1)
  int unit_blocks;        // H / U
  int tiles_per_cta;      // 128-row batch tiles per CTA (batch split)
  const float* gates_x;   // [T, B, 4H]; gx_rows32: in 32-row blocks, [ceil(T B / 32)][H][32 rows][4 floats]
  int gx_rows32;
  float* c_ws;            // running cell state, private to the launch, always in 32-row blocks [ceil(B / 32)][H / 4][32][4]
  const float* h0;
  const float* c0;
  const int* lengths;
  int T, B, H;
  int nsplit;             // 1: bf16, 3: hi*hi + hi*lo + lo*hi
  int kblocks, m_tiles;
  float* out_f32;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* hT;
  float* cT;
  float* c_seq;           // optional [T, B, H]: the cell state after every live step (state snapshots of long chains)
  __nv_bfloat16* hbuf[2][2];
  unsigned int* barrier;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (cooperative, co-resident) grid arrive; bounded spin like mbar_wait
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    unsigned int polls = 0;
    while (ld_acquire_u32(ctr) < target) {
      if (((++polls) & 0xfffu) == 0u && (clock64() - t0) > 8000000000LL) {
        printf("blm: LSTM grid barrier timed out (block %d target %u)\n", (int)blockIdx.x, target);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// producer-side wait on a monotonically increasing counter (bounded spin)
__device__ __forceinline__ void flag_wait(const unsigned int* ctr, unsigned int target) {
  const long long t0 = clock64();
  unsigned int polls = 0;
  while (ld_acquire_u32(ctr) < target) {
    if (((++polls) & 0xfffu) == 0u && (clock64() - t0) > 8000000000LL) {
      printf("blm: LSTM tile flag timed out (block %d target %u)\n", (int)blockIdx.x, target);
      __trap();
    }
  }
}

// sigmoid / tanh on the MUFU pipe: ex2.approx (2 ulp) + rcp.approx (1 ulp); |x| is clamped where both have
// saturated in fp32.  The libdevice expf / tanhf / IEEE division of the first version cost ~150 issue slots per
// hidden unit and made the single-warp-per-scheduler epilogue, not the recurrent product, the step's bound.
__device__ __forceinline__ float fast_sigmoid(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(fmaxf(x, -30.0f), 30.0f) * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float fast_tanh(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(fmaxf(x, -15.0f), 15.0f) * -2.8853900817779268f));  // exp(-2x)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return (1.0f - e) * r;
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void store_h8(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&h)[8]) {
  uint32_t a[4], b[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat16 x = __float2bfloat16_rn(h[2 * q]), y = __float2bfloat16_rn(h[2 * q + 1]);
    a[q] = static_cast<uint32_t>(__bfloat16_as_ushort(x)) | (static_cast<uint32_t>(__bfloat16_as_ushort(y)) << 16);
    b[q] = pack_bf16x2(h[2 * q] - __bfloat162float(x), h[2 * q + 1] - __bfloat162float(y));
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(a[0], a[1], a[2], a[3]);
  if (lo) *reinterpret_cast<uint4*>(lo) = make_uint4(b[0], b[1], b[2], b[3]);
}

// One batch row per thread is what the tensor-memory accumulator dictates, and with row-major [rows, cols] fp32 tensors
// every warp-level 16-byte access then touches 32 different lines (32 LSU wavefronts per instruction): the gates_x and
// cell-state accesses alone were 24 of 38 us per step at B = 2048.  In the 32-row-block layout the same instruction
// reads 512 contiguous bytes.
__device__ __forceinline__ long long rows32_f4(long long row, int col, int ncols) {
  return ((row >> 5) * (ncols >> 2) + (col >> 2)) * 32 + (row & 31);
}
__device__ __forceinline__ const float4* gx_f4(const LstmParams& p, long long m, int col) {
  const float4* base = reinterpret_cast<const float4*>(p.gates_x);
  return p.gx_rows32 ? base + rows32_f4(m, col, 4 * p.H) : base + ((m * 4 * p.H + col) >> 2);
}
__device__ __forceinline__ float4* c_f4(const LstmParams& p, int b, int col) {
  return reinterpret_cast<float4*>(p.c_ws) + rows32_f4(b, col, p.H);
}

// One batch row, kU hidden units starting at `unit0`: v holds the recurrent product [4 gates][kU units] of the row.
// Adds the hoisted input projection, applies the gates, updates (c, h) and publishes h_t (fp32 state, layer output,
// bf16 operand of step t + 1).  Rows past their length keep their state and carry the operand to the other buffer.
template <int kU>
__device__ __forceinline__ void cell_update(const LstmParams& p, const float (&v)[4 * kU], int t, int b, int unit0,
                                            int cur, int nxt) {
  const int H = p.H, B = p.B;
  const int len = min(__ldg(p.lengths + b), p.T);
  const bool live = t < len;
  const bool last = t + 1 == len;   // the carried h leaves the kernel once, after the row's last token
  const long long m = static_cast<long long>(t) * B + b;
#pragma unroll
  for (int u8 = 0; u8 < kU; u8 += 8) {   // groups of 8 units: 16-byte bf16 stores
    const long long o = static_cast<long long>(b) * H + unit0 + u8;
    const long long ot = (static_cast<long long>(t) * B + b) * H + unit0 + u8;
    __nv_bfloat16* nh = p.hbuf[nxt][0] + o;
    __nv_bfloat16* nl = p.nsplit == 3 ? p.hbuf[nxt][1] + o : nullptr;
    if (live) {
      float a[4][8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 x0 = __ldg(gx_f4(p, m, g * H + unit0 + u8));
        const float4 x1 = __ldg(gx_f4(p, m, g * H + unit0 + u8 + 4));
        const float* vv = v + g * kU + u8;
        a[g][0] = vv[0] + x0.x; a[g][1] = vv[1] + x0.y; a[g][2] = vv[2] + x0.z; a[g][3] = vv[3] + x0.w;
        a[g][4] = vv[4] + x1.x; a[g][5] = vv[5] + x1.y; a[g][6] = vv[6] + x1.z; a[g][7] = vv[7] + x1.w;
      }
      float c[8], h[8];
      float4* cw0 = c_f4(p, b, unit0 + u8);
      float4* cw1 = c_f4(p, b, unit0 + u8 + 4);
      *reinterpret_cast<float4*>(c) = *cw0;
      *reinterpret_cast<float4*>(c + 4) = *cw1;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float ig = fast_sigmoid(a[0][u]);
        const float fg = fast_sigmoid(a[1][u]);
        const float gg = fast_tanh(a[2][u]);
        const float og = fast_sigmoid(a[3][u]);
        c[u] = fmaf(fg, c[u], ig * gg);
        h[u] = og * fast_tanh(c[u]);
      }
      *cw0 = *reinterpret_cast<float4*>(c);
      *cw1 = *reinterpret_cast<float4*>(c + 4);
      if (last) {
        *reinterpret_cast<float4*>(p.cT + o) = *reinterpret_cast<float4*>(c);
        *reinterpret_cast<float4*>(p.cT + o + 4) = *reinterpret_cast<float4*>(c + 4);
        *reinterpret_cast<float4*>(p.hT + o) = *reinterpret_cast<float4*>(h);
        *reinterpret_cast<float4*>(p.hT + o + 4) = *reinterpret_cast<float4*>(h + 4);
      }
      store_h8(nh, nl, h);
      if (p.c_seq) {
        *reinterpret_cast<float4*>(p.c_seq + ot) = *reinterpret_cast<float4*>(c);
        *reinterpret_cast<float4*>(p.c_seq + ot + 4) = *reinterpret_cast<float4*>(c + 4);
      }
      if (p.out_f32) {
        *reinterpret_cast<float4*>(p.out_f32 + ot) = *reinterpret_cast<float4*>(h);
        *reinterpret_cast<float4*>(p.out_f32 + ot + 4) = *reinterpret_cast<float4*>(h + 4);
      }
      if (p.out_hi) store_h8(p.out_hi + ot, p.out_lo ? p.out_lo + ot : nullptr, h);
    } else {
      // padded step: state unchanged; carry the bf16 operand into the other buffer
      *reinterpret_cast<uint4*>(nh) = *reinterpret_cast<const uint4*>(p.hbuf[cur][0] + o);
      if (nl) *reinterpret_cast<uint4*>(nl) = *reinterpret_cast<const uint4*>(p.hbuf[cur][1] + o);
      if (p.out_f32) {
        *reinterpret_cast<float4*>(p.out_f32 + ot) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(p.out_f32 + ot + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (p.out_hi) {
        *reinterpret_cast<uint4*>(p.out_hi + ot) = make_uint4(0, 0, 0, 0);
        if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + ot) = make_uint4(0, 0, 0, 0);
      }
    }
  }
}

// h0 -> the bf16 operand of step 0 and the running (h, c) state, rows [row_lo, row_hi) x units [unit0, unit0 + kU)
template <int kU>
__device__ __forceinline__ void publish_initial_state(const LstmParams& p, int row_lo, int row_hi, int unit0) {
  const int H = p.H;
  for (int idx = threadIdx.x; idx < (row_hi - row_lo) * (kU / 8); idx += kLThreads) {
    const int b = row_lo + idx / (kU / 8);
    const long long o = static_cast<long long>(b) * H + unit0 + (idx % (kU / 8)) * 8;
    float h[8], c[8];
    *reinterpret_cast<float4*>(h) = __ldg(reinterpret_cast<const float4*>(p.h0 + o));
    *reinterpret_cast<float4*>(h + 4) = __ldg(reinterpret_cast<const float4*>(p.h0 + o + 4));
    *reinterpret_cast<float4*>(c) = __ldg(reinterpret_cast<const float4*>(p.c0 + o));
    *reinterpret_cast<float4*>(c + 4) = __ldg(reinterpret_cast<const float4*>(p.c0 + o + 4));
    *reinterpret_cast<float4*>(p.hT + o) = *reinterpret_cast<float4*>(h);
    *reinterpret_cast<float4*>(p.hT + o + 4) = *reinterpret_cast<float4*>(h + 4);
    *reinterpret_cast<float4*>(p.cT + o) = *reinterpret_cast<float4*>(c);
    *reinterpret_cast<float4*>(p.cT + o + 4) = *reinterpret_cast<float4*>(c + 4);
    const int col = unit0 + (idx % (kU / 8)) * 8;
    *c_f4(p, b, col) = *reinterpret_cast<float4*>(c);
    *c_f4(p, b, col + 4) = *reinterpret_cast<float4*>(c + 4);
    store_h8(p.hbuf[0][0] + o, p.nsplit == 3 ? p.hbuf[0][1] + o : nullptr, h);
  }
}

// The pair kernel's gate math, software-pipelined: the operands of a group of 8 hidden units that do NOT come from the
// tensor core -- gates_x (4 gates x 8 units) and the cell state -- are requested one group ahead (and, for the first
// group of a step, before the accumulator is even waited for), so their HBM / L2 latency runs under the previous
// group's arithmetic instead of being exposed four times per tile (8 of 31 us per step at B = 2048).
struct GateIn {
  float4 x[8];   // gates_x: [gate][2 x float4]
  float4 c[2];   // cell state of the 8 units
};
__device__ __forceinline__ void load_gate_in(const LstmParams& p, GateIn& in, int t, int b, int unit, bool live) {
  if (!live) return;
  const long long m = static_cast<long long>(t) * p.B + b;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    in.x[2 * g] = __ldg(gx_f4(p, m, g * p.H + unit));
    in.x[2 * g + 1] = __ldg(gx_f4(p, m, g * p.H + unit + 4));
  }
  in.c[0] = *c_f4(p, b, unit);
  in.c[1] = *c_f4(p, b, unit + 4);
}
// acc: the recurrent product [gate][8 units] of the row.  Same arithmetic and stores as cell_update.
__device__ __forceinline__ void cell_group(const LstmParams& p, const GateIn& in, const float (&acc)[32], int t, int b,
                                           int unit, bool live, bool last, int cur, int nxt) {
  const int H = p.H;
  const long long o = static_cast<long long>(b) * H + unit;
  const long long ot = (static_cast<long long>(t) * p.B + b) * H + unit;
  __nv_bfloat16* nh = p.hbuf[nxt][0] + o;
  __nv_bfloat16* nl = p.nsplit == 3 ? p.hbuf[nxt][1] + o : nullptr;
  if (live) {
    float c[8], h[8];
    *reinterpret_cast<float4*>(c) = in.c[0];
    *reinterpret_cast<float4*>(c + 4) = in.c[1];
    const float* x = reinterpret_cast<const float*>(in.x);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float ig = fast_sigmoid(acc[u] + x[u]);
      const float fg = fast_sigmoid(acc[8 + u] + x[8 + u]);
      const float gg = fast_tanh(acc[16 + u] + x[16 + u]);
      const float og = fast_sigmoid(acc[24 + u] + x[24 + u]);
      c[u] = fmaf(fg, c[u], ig * gg);
      h[u] = og * fast_tanh(c[u]);
    }
    *c_f4(p, b, unit) = *reinterpret_cast<float4*>(c);
    *c_f4(p, b, unit + 4) = *reinterpret_cast<float4*>(c + 4);
    if (last) {
      *reinterpret_cast<float4*>(p.cT + o) = *reinterpret_cast<float4*>(c);
      *reinterpret_cast<float4*>(p.cT + o + 4) = *reinterpret_cast<float4*>(c + 4);
      *reinterpret_cast<float4*>(p.hT + o) = *reinterpret_cast<float4*>(h);
      *reinterpret_cast<float4*>(p.hT + o + 4) = *reinterpret_cast<float4*>(h + 4);
    }
    store_h8(nh, nl, h);
    if (p.c_seq) {
      *reinterpret_cast<float4*>(p.c_seq + ot) = *reinterpret_cast<float4*>(c);
      *reinterpret_cast<float4*>(p.c_seq + ot + 4) = *reinterpret_cast<float4*>(c + 4);
    }
    if (p.out_f32) {
      *reinterpret_cast<float4*>(p.out_f32 + ot) = *reinterpret_cast<float4*>(h);
      *reinterpret_cast<float4*>(p.out_f32 + ot + 4) = *reinterpret_cast<float4*>(h + 4);
    }
    if (p.out_hi) store_h8(p.out_hi + ot, p.out_lo ? p.out_lo + ot : nullptr, h);
  } else {
    *reinterpret_cast<uint4*>(nh) = *reinterpret_cast<const uint4*>(p.hbuf[cur][0] + o);
    if (nl) *reinterpret_cast<uint4*>(nl) = *reinterpret_cast<const uint4*>(p.hbuf[cur][1] + o);
    if (p.out_f32) {
      *reinterpret_cast<float4*>(p.out_f32 + ot) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(p.out_f32 + ot + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (p.out_hi) {
      *reinterpret_cast<uint4*>(p.out_hi + ot) = make_uint4(0, 0, 0, 0);
      if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + ot) = make_uint4(0, 0, 0, 0);
    }
  }
}

// kCL > 1: kCL CTAs with the same batch block (consecutive unit blocks) form a cluster; every h tile is
// fetched ONCE per cluster -- CTA r loads rows [32 r, 32 r + 32) of the 128-row tile with TMA multicast
// into all kCL CTAs -- and a ring slot is released by multicast tcgen05.commit from all kCL MMA warps.
// The L2 -> SM traffic of the step (every CTA needs all of h_{t-1} of its batch block) drops kCL-fold.
// kSub: K blocks per ring stage.  The ring's throughput is bytes in flight / round-trip time, and the round
// trip (commit -> producer wake-up -> TMA issue -> L2 -> mbarrier -> MMA) is mostly fixed cost: 32 KB stages
// move twice the bytes per trip in the same shared memory (3 x 32 KB instead of 5 x 16 KB).
template <int kU, int kCL, int kSub>
__global__ void __launch_bounds__(kLThreads, 1) lstm_layer_kernel(const __grid_constant__ LstmParams p) {
  constexpr int kNStages = kSub == 2 ? kLStages2 : kLStages;
  constexpr int kStageBytes = kSub * kLABytes;
  constexpr int kWholeStride = kNStages * kStageBytes / 2 / 1024 * 1024;   // whole_k mode: two half-ring areas
  static_assert(kSub == 1 || kCL == 1, "the cluster experiment keeps one K block per stage");
  constexpr int kLN = 4 * kU;            // MMA N: 4 gates x U units
  constexpr int kLWTile = kLN * 64 * 2;  // one K block of the W slice
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023u) & ~static_cast<uintptr_t>(1023u));
  const int w_bytes = p.kblocks * kLWTile;
  uint8_t* sW[2] = {smem, smem + w_bytes};
  uint8_t* sA = smem + (p.nsplit == 3 ? 2 : 1) * w_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + kNStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kNStages;
  uint64_t* tfull_bar = empty_bar + kNStages;  // [kLMaxTiles]
  uint64_t* w_bar = tfull_bar + kLMaxTiles;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x % p.unit_blocks;   // unit block
  const int bi = blockIdx.x / p.unit_blocks;  // batch block
  const int H = p.H, B = p.B;
  const int mt0 = bi * p.tiles_per_cta;                           // first 128-row tile of this CTA
  const int n_mt = min(p.tiles_per_cta, p.m_tiles - mt0);         // its tile count
  const int row_lo = mt0 * 128, row_hi = min(B, (mt0 + n_mt) * 128);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kNStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kCL);
    }
    for (int s = 0; s < kLMaxTiles; ++s) mbar_init(&tfull_bar[s], 1);
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (kCL > 1) cluster_sync_all();  // peers' barriers exist before any multicast lands
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = kCL > 1 ? cluster_ctarank() : 0u;

  // resident W slice: rows {g*H + 8j + u}, one 8-row TMA box per (gate, K block) = one swizzle atom
  if (warp == 0 && lane == 0) {
    const int parts = p.nsplit == 3 ? 2 : 1;
    mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(parts * w_bytes));
    for (int part = 0; part < parts; ++part)
      for (int kb = 0; kb < p.kblocks; ++kb)
        for (int g = 0; g < 4; ++g)
          tma_load_2d(sW[part] + kb * kLWTile + g * (kU * 128), &p.tmW[part], w_bar, kb * 64, g * H + j * kU,
                      kEvictFirst);
  }

  // prologue: publish h_{-1} = h0 as the bf16 operand and seed the running (h, c) state, own columns
  publish_initial_state<kU>(p, row_lo, row_hi, j * kU);
  fence_proxy_async_all();
  unsigned int bar_n = 0;
  grid_barrier(p.barrier, (++bar_n) * gridDim.x);
  mbar_wait(w_bar, 0);

  // every CTA (cluster) walks the K blocks of a tile from a different starting block: in lock step all
  // CTAs of a batch block would otherwise request the same 16 KB of h from the same L2 slices at once
  const int kb_rot = p.kb_stagger ? ((j / kCL) % p.kblocks) : 0;
  int stage = 0;
  uint32_t phase = 0;  // ring position, advanced identically by producer and MMA threads
  const int a_parts = p.nsplit == 3 ? 2 : 1;

  for (int t = 0; t < p.T; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    if (warp == 0) {
      if (lane == 0) {
        fence_proxy_async_all();  // h_{t-1} was written with generic stores by other SMs
        if (kCL == 1 && p.whole_k) {
          // the hypothesis-#0 chains: a dozen rows, latency is everything.  Two half-ring areas alternate; the 16 KB an
          // M = 128 descriptor spans past a K block's few real rows overlaps the following K blocks (finite values that
          // only reach accumulator rows nobody reads)
          for (int part = 0; part < a_parts; ++part) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.kblocks * p.a_box_bytes));
            for (int kb = 0; kb < p.kblocks; ++kb)
              tma_load_2d(sA + stage * kWholeStride + kb * p.a_box_bytes, &p.tmH[cur][part], &full_bar[stage], kb * 64,
                          mt0 * 128, kEvictNormal);
            if (++stage == 2) {
              stage = 0;
              phase ^= 1u;
            }
          }
        } else
        for (int mt = 0; mt < n_mt; ++mt)
          for (int part = 0; part < a_parts; ++part)
            for (int kbi = 0; kbi < p.kblocks; kbi += kSub) {
              const int kb = (kbi + kb_rot) % p.kblocks;  // staggered K order (experiment, kSub == 1 only)
              mbar_wait(&empty_bar[stage], phase ^ 1u);
              mbar_arrive_expect_tx(&full_bar[stage], kCL > 1 ? kStageBytes : kSub * p.a_box_bytes);
              if constexpr (kCL > 1) {
                constexpr int kQ = 128 / kCL;  // rows of the tile this CTA fetches for the whole cluster
                tma_load_2d_multicast(sA + stage * kStageBytes + crank * (kQ * 128), &p.tmH[cur][part], &full_bar[stage],
                                      kb * 64, (mt0 + mt) * 128 + static_cast<int>(crank) * kQ,
                                      static_cast<uint16_t>((1u << kCL) - 1u), kEvictNormal);
              } else {
#pragma unroll
                for (int sub = 0; sub < kSub; ++sub)
                  tma_load_2d(sA + stage * kStageBytes + sub * kLABytes, &p.tmH[cur][part], &full_bar[stage],
                              (kb + sub) * 64, (mt0 + mt) * 128, kEvictNormal);
              }
              if (++stage == kNStages) {
                stage = 0;
                phase ^= 1u;
              }
            }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, kLN);  // N = 32 or 64
        tcgen05_fence_after();
        if (kCL == 1 && p.whole_k) {
          uint32_t accum = 0;
          for (int part = 0; part < a_parts; ++part) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            for (int kb = 0; kb < p.kblocks; ++kb) {
              const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * kWholeStride + kb * p.a_box_bytes));
              const uint64_t dw_hi = umma_desc_sw128(smem_u32(sW[0] + kb * kLWTile));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_ss(tmem_base, da + static_cast<uint64_t>(2 * k), dw_hi + static_cast<uint64_t>(2 * k), idesc, accum);
                accum = 1;
              }
              if (part == 0 && p.nsplit == 3) {  // h_hi . W_lo
                const uint64_t dw_lo = umma_desc_sw128(smem_u32(sW[1] + kb * kLWTile));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_ss(tmem_base, da + static_cast<uint64_t>(2 * k), dw_lo + static_cast<uint64_t>(2 * k), idesc, 1u);
              }
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == 2) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(&tfull_bar[0]);
        } else
        for (int mt = 0; mt < n_mt; ++mt) {
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(mt * kLN);
          uint32_t accum = 0;
          for (int part = 0; part < a_parts; ++part)
            for (int kbi = 0; kbi < p.kblocks; kbi += kSub) {
              const int kb0 = (kbi + kb_rot) % p.kblocks;
              mbar_wait(&full_bar[stage], phase);
              tcgen05_fence_after();
#pragma unroll
              for (int sub = 0; sub < kSub; ++sub) {
                const int kb = kb0 + sub;
                const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * kStageBytes + sub * kLABytes));
                const uint64_t dw_hi = umma_desc_sw128(smem_u32(sW[0] + kb * kLWTile));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16_ss(tmem_d, da + static_cast<uint64_t>(2 * k), dw_hi + static_cast<uint64_t>(2 * k), idesc, accum);
                  accum = 1;
                }
                if (part == 0 && p.nsplit == 3) {  // h_hi . W_lo
                  const uint64_t dw_lo = umma_desc_sw128(smem_u32(sW[1] + kb * kLWTile));
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem_d, da + static_cast<uint64_t>(2 * k), dw_lo + static_cast<uint64_t>(2 * k), idesc, 1u);
                }
              }
              if constexpr (kCL > 1)
                umma_commit_multicast(&empty_bar[stage], static_cast<uint16_t>((1u << kCL) - 1u));
              else
                umma_commit(&empty_bar[stage]);
              if (++stage == kNStages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          umma_commit(&tfull_bar[mt]);
        }
      }
      __syncwarp();
    } else if (warp >= 4) {
      const int lane_grp = warp & 3;
      if (kCL == 1 && n_mt == 1) {
        // one tile (the hypothesis-#0 chains, B <= 128): the step is pure latency, so both epilogue warp sets work on
        // the tile, one 8-unit group each, and request gates_x / the cell state BEFORE they wait for the accumulator
        const int b = mt0 * 128 + lane_grp * 32 + lane;
        const int len = b < B ? min(__ldg(p.lengths + b), p.T) : 0;
        for (int u8 = ((warp - 4) >> 2) * 8; u8 < kU; u8 += 16) {
          GateIn gin;
          if (b < B) load_gate_in(p, gin, t, b, j * kU + u8, t < len);
          mbar_wait(&tfull_bar[0], static_cast<uint32_t>(t & 1));
          tcgen05_fence_after();
          float acc[32];
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 4; ++g)
            tmem_ld_32x8(tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(g * kU + u8),
                         acc + 8 * g);
          tmem_ld_wait();
          if (b < B) cell_group(p, gin, acc, t, b, j * kU + u8, t < len, t + 1 == len, cur, nxt);
        }
      } else
      for (int mt = (warp - 4) >> 2; mt < n_mt; mt += 2) {
        mbar_wait(&tfull_bar[mt], static_cast<uint32_t>(t & 1));
        tcgen05_fence_after();
        float v[kLN];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kLN / 32; ++q)
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(mt * kLN + q * 32),
                        *reinterpret_cast<float(*)[32]>(v + q * 32));
        tmem_ld_wait();
        const int b = (mt0 + mt) * 128 + lane_grp * 32 + lane;
        if (b < B) cell_update<kU>(p, v, t, b, j * kU, cur, nxt);
      }
      fence_proxy_async_all();  // h_t must be visible to the TMA (async proxy) reads of step t+1
    }
    tcgen05_fence_before();
    if (t + 1 < p.T) {
      grid_barrier(p.barrier, (++bar_n) * gridDim.x);
    } else {
      __syncthreads();
    }
    tcgen05_fence_after();
  }

  if constexpr (kCL > 1) cluster_sync_all();  // no multicast / remote arrive may target a CTA that has exited
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair form (bf16 mode, batch >= 2 tiles).  The step of the kernel above is bound by the bytes of h_{t-1} a CTA
// can keep in flight next to its resident W slice (96 KB ring over a ~2 us TMA -> MMA -> commit round trip), so the
// lever is bytes per CTA, not the load path.  Two CTAs of a 2-CTA cluster (one TPC) own 32 hidden units together:
// `tcgen05.mma.cta_group::2` with M = 256, N = 128 reads each CTA's OWN 128 rows of h and each CTA's OWN 64 rows of
// the W slice (4 gates x 16 units, the same 128 KB as before) and leaves, in each CTA's tensor memory, its 128 batch
// rows x all 128 gate columns of the pair.  Per step a CTA therefore streams h for HALF the batch rows (one 128-row
// tile per 256-row pair tile) and still applies the gates for the same number of (row, unit) cells.
//   warp 0  producer in both CTAs: own h rows, `cp.async.bulk.tensor...cta_group::2` completing on the LEADER's
//           full barrier (the leader arms expect_tx for both CTAs' bytes)
//   warp 1  leader only: MMA issue; multicast commits free the ring slot in both CTAs and publish the accumulator
//           of a pair tile to both CTAs' epilogue warps
//   warps 4-11  gates for the CTA's own rows: accumulator columns [64 c, 64 c + 64) are the units of CTA c
template <int kU, int kSub>
__global__ void __launch_bounds__(kLThreads, 1) lstm_pair_kernel(const __grid_constant__ LstmParams p) {
  constexpr int kNStages = kSub == 2 ? kLStages2 : kLStages;
  constexpr int kStageBytes = kSub * kLABytes;
  constexpr int kLN = 8 * kU;            // MMA N of the pair: 2 CTAs x 4 gates x kU units
  constexpr int kLWTile = 4 * kU * 64 * 2;  // one K block of this CTA's W slice (4 kU rows), per part
  constexpr int GPT = kU / 4;            // groups of 8 units per pair tile (2 CTAs x kU / 8)
  constexpr int TPW = 512 / kLN / 2;     // pair tiles per epilogue warp set (tiles w, w + 2, ...)
  constexpr int NQ = GPT * TPW;          // groups per step and warp: 8 either way
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023u) & ~static_cast<uintptr_t>(1023u));
  const int w_bytes = p.kblocks * kLWTile;
  uint8_t* sW[2] = {smem, smem + w_bytes};                 // hi, lo (precise mode: kU = 8, both resident)
  const int a_parts = p.nsplit == 3 ? 2 : 1;
  uint8_t* sA = smem + a_parts * w_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + kNStages * kStageBytes);  // used in the leader only
  uint64_t* empty_bar = full_bar + kNStages;
  uint64_t* tfull_bar = empty_bar + kNStages;  // [kLMaxTiles]
  uint64_t* w_bar = tfull_bar + kLMaxTiles;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int j = blockIdx.x % p.unit_blocks;   // this CTA's 16-unit block (even: leader, odd: peer)
  const int bi = blockIdx.x / p.unit_blocks;  // batch block of the pair
  const int B = p.B;
  const int mt0 = bi * p.tiles_per_cta;                    // first 256-row pair tile
  const int n_mt = min(p.tiles_per_cta, p.m_tiles - mt0);  // pair tiles (p.m_tiles counts 256-row tiles here)
  const int row_lo = mt0 * 256, row_hi = min(B, (mt0 + n_mt) * 256);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kNStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kLMaxTiles; ++s) mbar_init(&tfull_bar[s], 1);
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(a_parts * w_bytes));
    for (int part = 0; part < a_parts; ++part)
      for (int kb = 0; kb < p.kblocks; ++kb)
        for (int g = 0; g < 4; ++g)
          tma_load_2d(sW[part] + kb * kLWTile + g * (kU * 128), &p.tmW[part], w_bar, kb * 64, g * p.H + j * kU,
                      kEvictFirst);
  }
  publish_initial_state<kU>(p, row_lo, row_hi, j * kU);
  fence_proxy_async_all();
  unsigned int bar_n = 0;
  grid_barrier(p.barrier, (++bar_n) * gridDim.x);
  mbar_wait(w_bar, 0);
  cluster_sync_all();  // the peer's W slice is resident before the leader's first MMA reads it

  int stage = 0;
  uint32_t phase = 0;
  const uint32_t mask2 = 0x3;
  // No grid barrier between the steps.  Step t + 1 of pair tile mt needs h_t of that tile's 256 rows only, from every
  // CTA of the batch block; one counter per (batch block, pair tile), on its own 128-byte line, carries exactly that
  // dependency, so the loads and MMAs of tile 0 of the next step run under the gate math of the last tiles of this
  // one.  Hazards: the operand is double-buffered -- h_{t+1} of a tile overwrites the buffer h_{t-1} was read from, and
  // every CTA's reads of it (its step-t MMAs of that tile) completed before its own arrival for (t, tile), which all
  // step-(t+1) MMAs of the tile wait for; the tile's accumulator is re-used only after this CTA's own epilogue warps
  // arrived (they arrive after their tcgen05.ld completed).
  unsigned int* tile_flag = p.barrier + 64 + (bi * 8) * 32;
  const unsigned int flag_per_step = static_cast<unsigned int>(p.unit_blocks) * 4u;

  // epilogue warps: warp w of a lane group takes pair tiles w and w + 2; one batch row per thread
  const int lane_grp = warp & 3;
  const int epi_w = (warp - 4) >> 2;
  const int epi_b0 = (mt0 + epi_w) * 256 + static_cast<int>(rank) * 128 + lane_grp * 32 + lane;   // row in tile w
  int len_t[TPW];   // length of this thread's row in each of its tiles
  GateIn gin[2];
  if (warp >= 4) {
#pragma unroll
    for (int ti = 0; ti < TPW; ++ti) {
      const int b = epi_b0 + 2 * ti * 256;
      len_t[ti] = (epi_w + 2 * ti < n_mt && b < B) ? min(__ldg(p.lengths + b), p.T) : 0;
    }
    if (epi_w < n_mt) load_gate_in(p, gin[0], 0, epi_b0, (j & ~1) * kU, 0 < len_t[0]);
  }

  for (int t = 0; t < p.T; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    if (warp == 0) {
      if (lane == 0) {
        for (int mt = 0; mt < n_mt; ++mt) {
          // rows of pair tile mt of h_{t-1} are complete once every epilogue warp that owns a slice of them (4 warps in
          // each of the unit_blocks CTAs of this batch block) has arrived t times
          if (t > 0) flag_wait(tile_flag + mt * 32, static_cast<unsigned int>(t) * flag_per_step);
          fence_proxy_async_all();  // h_{t-1} was written with generic stores by other SMs
          for (int part = 0; part < a_parts; ++part)
            for (int kb = 0; kb < p.kblocks; kb += kSub) {
              mbar_wait(&empty_bar[stage], phase ^ 1u);
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
              const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
#pragma unroll
              for (int sub = 0; sub < kSub; ++sub)
                tma_load_2d_2sm(sA + stage * kStageBytes + sub * kLABytes, &p.tmH[cur][part], bar, (kb + sub) * 64,
                                (mt0 + mt) * 256 + static_cast<int>(rank) * 128, kEvictNormal);
              if (++stage == kNStages) {
                stage = 0;
                phase ^= 1u;
              }
            }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      if (leader && lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(256, kLN);
        tcgen05_fence_after();
        for (int mt = 0; mt < n_mt; ++mt) {
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(mt * kLN);
          uint32_t accum = 0;
          for (int part = 0; part < a_parts; ++part)
            for (int kb0 = 0; kb0 < p.kblocks; kb0 += kSub) {
              mbar_wait(&full_bar[stage], phase);
              tcgen05_fence_after();
#pragma unroll
              for (int sub = 0; sub < kSub; ++sub) {
                const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * kStageBytes + sub * kLABytes));
                const uint64_t dw = umma_desc_sw128(smem_u32(sW[0] + (kb0 + sub) * kLWTile));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16_ss_2sm(tmem_d, da + static_cast<uint64_t>(2 * k), dw + static_cast<uint64_t>(2 * k), idesc, accum);
                  accum = 1;
                }
                if (part == 0 && p.nsplit == 3) {  // h_hi . W_lo (part 1 is h_lo . W_hi)
                  const uint64_t dl = umma_desc_sw128(smem_u32(sW[1] + (kb0 + sub) * kLWTile));
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16_ss_2sm(tmem_d, da + static_cast<uint64_t>(2 * k), dl + static_cast<uint64_t>(2 * k), idesc, 1u);
                }
              }
              umma_commit_2sm(&empty_bar[stage], mask2);
              if (++stage == kNStages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          umma_commit_2sm(&tfull_bar[mt], mask2);
        }
      }
      __syncwarp();
    } else if (warp >= 4) {
      // groups of the step, in order: tiles w, w + 2, .. x (leader's units, peer's units) x 8-unit groups
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        constexpr int kG8 = kU / 8;
        const int ti = q / GPT, half = (q % GPT) / kG8, u8 = ((q % GPT) % kG8) * 8;
        const int mt = epi_w + 2 * ti;
        if (mt >= n_mt) break;
        const int b = epi_b0 + 2 * ti * 256;
        const int len = len_t[ti];
        const int unit = ((j & ~1) + half) * kU + u8;
        // request the operands of the next group: next q of this step, or the first group of step t + 1
        {
          const int qn = q + 1;
          const int tin = qn < NQ ? qn / GPT : 0;
          const bool wrap = qn == NQ || epi_w + 2 * tin >= n_mt;
          const int halfn = (qn % GPT) / kG8, u8n = ((qn % GPT) % kG8) * 8;
          const int tn = wrap ? t + 1 : t;
          const int bn = wrap ? epi_b0 : epi_b0 + 2 * tin * 256;
          const int lenn = wrap ? len_t[0] : len_t[tin];
          const int unitn = wrap ? (j & ~1) * kU : ((j & ~1) + halfn) * kU + u8n;
          load_gate_in(p, gin[(q + 1) & 1], tn, bn, unitn, tn < lenn);
        }
        if ((q % GPT) == 0) {
          mbar_wait(&tfull_bar[mt], static_cast<uint32_t>(t & 1));
          tcgen05_fence_after();
        }
        float acc[32];
        __syncwarp();
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                              static_cast<uint32_t>(mt * kLN + half * (4 * kU) + u8);
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld_32x8(tcol + static_cast<uint32_t>(g * kU), acc + 8 * g);
        tmem_ld_wait();
        if (b < B) cell_group(p, gin[q & 1], acc, t, b, unit, t < len, t + 1 == len, cur, nxt);
        if ((q % GPT) == GPT - 1) {
          // publish this warp's 32 rows x 2 kU units of h_t: visible to the async proxy (TMA reads of step t + 1) and
          // at gpu scope before the flag moves
          tcgen05_fence_before();
          fence_proxy_async_all();
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(tile_flag + mt * 32, 1u);
        }
      }
    }
  }
  __syncthreads();

  cluster_sync_all();  // no multicast commit / remote completion may target a CTA that has exited
  if (warp == 2) tmem_dealloc_2sm<512>(tmem_base);
}

static size_t lstm_smem_bytes(int kblocks, int nsplit, int U, int sub = 1) {
  const int stages = sub == 2 ? kLStages2 : kLStages;
  return static_cast<size_t>((nsplit == 3 ? 2 : 1) * kblocks * (4 * U * 128) + stages * sub * kLABytes +
                             (2 * stages + kLMaxTiles + 1) * 8 + 16 + 1024);
}

int lstm_init() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(lstm_layer_kernel<8, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lstm_smem_bytes(16, 3, 8))));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(lstm_layer_kernel<16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lstm_smem_bytes(16, 1, 16))));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(lstm_layer_kernel<8, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lstm_smem_bytes(16, 3, 8, 2))));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(lstm_layer_kernel<16, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lstm_smem_bytes(16, 1, 16, 2))));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(lstm_layer_kernel<16, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lstm_smem_bytes(16, 1, 16))));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(lstm_pair_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lstm_smem_bytes(16, 1, 16))));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(lstm_pair_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lstm_smem_bytes(16, 1, 16, 2))));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(lstm_pair_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lstm_smem_bytes(16, 3, 8))));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(lstm_pair_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(lstm_smem_bytes(16, 3, 8, 2))));
  return BLM_OK;
}

}  // namespace blm

extern "C" {

int64_t blm_lstm_workspace_bytes(int64_t B, int64_t H) {
  // barrier counter + per-tile step flags (4 KB header) + 2 buffers x (hi, lo) x [B, H] bf16 + the running cell state
  // in 32-row blocks
  return 4096 + 4 * B * H * 2 + (B + 31) / 32 * 32 * H * 4;
}

int blm_lstm_layer(const float* gates_x, const blm_bf16* w_hh_hi, const blm_bf16* w_hh_lo, const float* h0,
                   const float* c0, const int32_t* lengths, int64_t T, int64_t B, int64_t H, float* out_f32,
                   blm_bf16* out_hi, blm_bf16* out_lo, float* hT, float* cT, void* workspace, blm_stream stream) {
  return blm_lstm_layer_seq(gates_x, 0, w_hh_hi, w_hh_lo, h0, c0, lengths, T, B, H, out_f32, out_hi, out_lo, hT, cT, nullptr,
                            workspace, stream);
}

int blm_lstm_layer_seq(const float* gates_x, int32_t gx_rows32, const blm_bf16* w_hh_hi, const blm_bf16* w_hh_lo, const float* h0,
                       const float* c0, const int32_t* lengths, int64_t T, int64_t B, int64_t H, float* out_f32,
                       blm_bf16* out_hi, blm_bf16* out_lo, float* hT, float* cT, float* c_seq, void* workspace,
                       blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(num_sms() > 0, BLM_ERR_ARCH, "blm_init() has not been called");
  BLM_REQUIRE(gates_x && w_hh_hi && h0 && c0 && lengths && hT && cT && workspace, BLM_ERR_ARG,
              "null LSTM argument");
  BLM_REQUIRE(T >= 1 && B >= 1 && T < (1 << 20), BLM_ERR_SHAPE, "bad LSTM shape T=%lld B=%lld", (long long)T,
              (long long)B);
  BLM_REQUIRE(H >= 8 && (H % 8) == 0 && H <= 1024, BLM_ERR_SHAPE,
              "hidden size %lld must be a multiple of 8 and <= 1024 (W_hh slice must fit in shared memory)",
              (long long)H);
  BLM_REQUIRE(H / 8 <= num_sms(), BLM_ERR_SHAPE, "hidden size %lld needs more CTAs than SMs", (long long)H);
  BLM_REQUIRE(B <= 128 * kLMaxTiles, BLM_ERR_SHAPE, "batch %lld exceeds %d rows per launch", (long long)B,
              128 * kLMaxTiles);
  BLM_REQUIRE(!out_lo || out_hi, BLM_ERR_ARG, "out_lo requires out_hi");
  BLM_REQUIRE(aligned16(gates_x) && aligned16(h0) && aligned16(c0) && aligned16(hT) && aligned16(cT) &&
                  aligned16(out_f32) && aligned16(out_hi) && aligned16(out_lo) && aligned16(workspace),
              BLM_ERR_ALIGN, "LSTM pointers must be 16-byte aligned");

  LstmParams p;
  memset(&p, 0, sizeof(p));
  p.nsplit = w_hh_lo ? 3 : 1;
  p.T = static_cast<int>(T);
  p.B = static_cast<int>(B);
  p.H = static_cast<int>(H);
  p.kblocks = static_cast<int>((H + 63) / 64);
  p.m_tiles = static_cast<int>((B + 127) / 128);
  p.gates_x = gates_x;
  p.h0 = h0;
  p.c0 = c0;
  p.lengths = lengths;
  p.out_f32 = out_f32;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(out_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(out_lo);
  p.hT = hT;
  p.cT = cT;
  p.c_seq = c_seq;
  BLM_REQUIRE(aligned16(c_seq), BLM_ERR_ALIGN, "c_seq must be 16-byte aligned");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  p.barrier = reinterpret_cast<unsigned int*>(ws);
  __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(ws + 4096);
  p.c_ws = reinterpret_cast<float*>(ws + 4096 + 4 * B * H * 2);
  p.gx_rows32 = gx_rows32 ? 1 : 0;
  // bf16 mode: 16 units per CTA and the batch split over two CTA rows (halves the per-step h ingest);
  // precise mode keeps 8 units per CTA (hi + lo slices fill the same 128 KB) and no batch split
  static const bool force_u8 = getenv("BLM_LSTM_U8") != nullptr;          // A/B switches for profiling
  // Two experiments kept behind switches because they measured as exact no-ops (67.7 us per step at B = 2048
  // in all four combinations, profiles/r01v): TMA multicast of the h tiles inside 4-CTA clusters and a
  // staggered K-block order.  The step is therefore NOT bound by L2 bandwidth or hot lines but by latency:
  // 80 KB of h in flight per SM (5-stage ring next to the 128 KB resident W slice) over a ~2 us
  // TMA + MMA + commit round trip = ~40 GB/s per SM.  The fix is structural (cta_group::2 pairs sharing the
  // resident slice so that each CTA needs half the rows), not a different load path.
  static const bool no_cluster = getenv("BLM_LSTM_CLUSTER") == nullptr;
  static const bool no_stagger = getenv("BLM_LSTM_STAGGER") == nullptr;
  p.kb_stagger = no_stagger ? 0 : 1;
  const int U = (!w_hh_lo && (H % 16) == 0 && !force_u8) ? 16 : 8;
  const int nb = (U == 16 && p.m_tiles >= 2 && 2 * (H / U) <= num_sms()) ? 2 : 1;
  const int CL = (U == 16 && !no_cluster && ((H / U) % 4) == 0 && p.m_tiles >= 2) ? 4 : 1;
  // One-tile batches (the hypothesis-#0 chains: one row per session) load only their real rows: a 128-row box would
  // move 256 KB per CTA and step for a dozen rows, and that stream, not the barrier, was the chain's step time.  The
  // rows of the shared-memory tile the box does not cover keep whatever they held; they only feed accumulator rows
  // >= B, which nobody reads.
  const int box_rows = (p.m_tiles == 1 && getenv("BLM_LSTM_FULL_BOX") == nullptr) ? static_cast<int>((B + 7) / 8 * 8) : 128 / CL;
  p.a_box_bytes = box_rows * 128;
  for (int buf = 0; buf < 2; ++buf)
    for (int part = 0; part < 2; ++part) {
      p.hbuf[buf][part] = hb + (static_cast<int64_t>(buf) * 2 + part) * B * H;
      int rc = encode_tmap_bf16(&p.tmH[buf][part], p.hbuf[buf][part], B, H, H, box_rows);
      if (rc != BLM_OK) return rc;
    }
  p.unit_blocks = static_cast<int>(H / U);
  p.tiles_per_cta = (p.m_tiles + nb - 1) / nb;
  BLM_REQUIRE(p.tiles_per_cta * 4 * U <= 512, BLM_ERR_SHAPE, "batch %lld exceeds the TMEM accumulators of one launch",
              (long long)B);
  int rc = encode_tmap_bf16(&p.tmW[0], w_hh_hi, 4 * H, H, H, U);
  if (rc != BLM_OK) return rc;
  rc = encode_tmap_bf16(&p.tmW[1], w_hh_lo ? w_hh_lo : w_hh_hi, 4 * H, H, H, U);
  if (rc != BLM_OK) return rc;

  cudaStream_t st = as_stream(stream);
  BLM_CHECK_CUDA(cudaMemsetAsync(p.barrier, 0, 4096, st));
  void* args[] = {&p};
  static const bool no_sub2 = getenv("BLM_LSTM_NO_SUB2") != nullptr;  // A/B switch for profiling
  const int sub = (!no_sub2 && (p.kblocks % 2) == 0 && !p.kb_stagger) ? 2 : 1;
  // CTA pairs (lstm_pair_kernel): three or more 128-row tiles (below that a CTA already streams one tile); bf16 mode
  // with 16 units per CTA and the batch split in two, precise mode with 8 units per CTA (hi + lo slices resident).
  // Read per call so that a test can compare both kernels in one process.
  const int pair_tiles = static_cast<int>((B + 255) / 256);
  const int pair_nb = (pair_tiles >= 2 && 2 * (H / U) <= num_sms()) ? 2 : 1;
  const int pair_tpc = (pair_tiles + pair_nb - 1) / pair_nb;
  const bool pair = CL == 1 && !p.kb_stagger && (H % (2 * U)) == 0 && (U == 16 || w_hh_lo) && p.m_tiles >= 3 &&
                    pair_tpc * 8 * U <= 512 && H / U <= num_sms() && getenv("BLM_LSTM_NO_PAIR") == nullptr;
  if (pair) {
    p.m_tiles = pair_tiles;
    p.tiles_per_cta = pair_tpc;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(p.unit_blocks * pair_nb));
    cfg.blockDim = dim3(kLThreads);
    cfg.dynamicSmemBytes = lstm_smem_bytes(p.kblocks, p.nsplit, U, sub);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;  // the tile flags need every CTA resident
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    // Nsight Compute cannot replay a launch that is both clustered and cooperative (LaunchFailed).  For a capture only,
    // BLM_LSTM_NO_COOP=1 drops the cooperative attribute; residency then rests on the grid (<= 148 CTAs, one per SM)
    // being alone on the device, which a profiling run is.
    cfg.numAttrs = getenv("BLM_LSTM_NO_COOP") ? 1 : 2;
    if (U == 16) {
      if (sub == 2)
        BLM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_pair_kernel<16, 2>, p));
      else
        BLM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_pair_kernel<16, 1>, p));
    } else {
      if (sub == 2)
        BLM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_pair_kernel<8, 2>, p));
      else
        BLM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_pair_kernel<8, 1>, p));
    }
    return BLM_OK;
  }
  {
    const int ring_half = (sub == 2 ? kLStages2 * 2 : kLStages) * kLABytes / 2 / 1024 * 1024;
    p.whole_k = p.m_tiles == 1 && CL == 1 && !p.kb_stagger && (p.kblocks - 1) * p.a_box_bytes + kLABytes <= ring_half &&
                getenv("BLM_LSTM_NO_WHOLE_K") == nullptr;
  }
  const dim3 grid(static_cast<unsigned>(p.unit_blocks * nb)), block(kLThreads);
  if (CL > 1) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = lstm_smem_bytes(p.kblocks, p.nsplit, U);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;  // the per-step grid barrier needs every CTA resident
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    BLM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lstm_layer_kernel<16, 4, 1>, p));
    return BLM_OK;
  }
  void* fn;
  if (U == 16)
    fn = sub == 2 ? reinterpret_cast<void*>(lstm_layer_kernel<16, 1, 2>) : reinterpret_cast<void*>(lstm_layer_kernel<16, 1, 1>);
  else
    fn = sub == 2 ? reinterpret_cast<void*>(lstm_layer_kernel<8, 1, 2>) : reinterpret_cast<void*>(lstm_layer_kernel<8, 1, 1>);
  BLM_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, lstm_smem_bytes(p.kblocks, p.nsplit, U, sub), st));
  return BLM_OK;
}

}  // extern "C"
