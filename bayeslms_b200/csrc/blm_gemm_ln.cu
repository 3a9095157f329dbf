// blm_gemm_ln: out = LayerNorm(resid + A B^T + bias) * gamma + beta in one tcgen05 kernel (sm_100a).
//
// The post-LN sublayer tails of the Transformer block (o_net -> +x -> norm1, linear2 -> +x -> norm2,
// model.py:1040-1046) are a GEMM with N = d_model followed by a row reduction over exactly those N
// columns.  d_model = 512 fp32 columns is the whole TMEM of an SM, so one CTA owns a full 128 x N row
// block: every epilogue thread owns one accumulator ROW, the LayerNorm statistics are thread-local, and
// the pre-LayerNorm sum (an [M, N] fp32 round trip through HBM plus a second launch in the unfused path)
// never leaves the SM.
//
// One persistent CTA per SM, 12 warps:
//   warp 0   TMA producer: A K-blocks (128 x 64) into a 2-slot ring, B half K-blocks (256 x 64) into a
//            4-slot ring; sub-stage = (K block, N half), full barrier per B slot (A rides on half 0)
//   warp 1   MMA issuer: tcgen05.mma 128 x 256 x 16 into TMEM columns [256 half, +256)
//   warp 2   TMEM allocator (512 columns: ONE accumulator stage -- the epilogue of a tile does not overlap
//            the next tile's MMAs, but the TMA ring keeps filling underneath it)
//   warp 4.. 8 epilogue warps; warp w owns TMEM lanes [32 (w % 4), +32) and the column half (w - 4) / 4:
//     pass 1   x = acc + bias + resid (resid chunks [32 rows x 32 cols] stream in by TMA, double
//              buffered), x written back to TMEM (tcgen05.st), shifted sum / sum of squares per half row
//     merge    the two warps sharing a row exchange (mean, M2) through shared memory (Chan)
//     pass 2   y = (x - mean) rstd gamma + beta -> fp32 chunks -> swizzled staging -> TMA store
//     pass 3   the same y packed to bf16, [32 rows x 64 cols] tiles -> TMA store
// All global traffic of the epilogue is TMA: a row-per-thread access pattern costs 32 LSU wavefronts per
// instruction and was the measured bound of the fp32-output GEMM epilogues (r01 profiles).
#include <stdlib.h>
#include <string.h>

#include "blm_gemm_common.cuh"

namespace blm {

struct GemmLnParams {
  CUtensorMap tmA, tmB;   // bf16 [M, K] box 128 x 64; bf16 [N, K] box 256 x 64
  CUtensorMap tmR;        // fp32 resid [M, N], box 32 rows x 32 cols
  CUtensorMap tmO;        // fp32 out   [M, N], box 32 rows x 32 cols
  CUtensorMap tmH;        // bf16 out   [M, N], box 32 rows x 64 cols
  int M, N, kblocks, m_tiles;
  int halves;             // N halves of 256 columns (1 or 2)
  int chunks_per_warp;    // 32-column chunks per epilogue warp = N / 64
  const float* bias;
  const float* gamma;
  const float* beta;
  float eps;
  int has_f32, has_hi;
};

namespace ln {
constexpr int kASlots = 2, kBSlots = 4;
constexpr int kABytes = kBM * kBK * 2;        // 16 KB
constexpr int kBBytes = 256 * kBK * 2;        // 32 KB
constexpr int kAOff = 0;
constexpr int kBOff = kASlots * kABytes;                  // 32 KB
constexpr int kStgOff = kBOff + kBSlots * kBBytes;        // 160 KB; 8 warps x 2 buffers x 4 KB
constexpr int kXchOff = kStgOff + 8 * 8192;               // 224 KB; float2 [2][128]
constexpr int kBarOff = kXchOff + 2 * 128 * 8;
// bfull[4] bempty[4] aempty[2] tfull tempty rfull[8][2]
constexpr int kNumBars = kBSlots * 2 + kASlots + 2 + 16;
constexpr int kSmemBytes = kBarOff + kNumBars * 8 + 16;
static_assert(kSmemBytes <= 232448, "shared memory budget of one sm_100 CTA");
}  // namespace ln

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// at most one committed bulk store of this thread may still be reading its shared-memory source
__device__ __forceinline__ void bulk_wait_group_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// at most N of this thread's most recent bulk stores may still be reading their shared-memory sources
template <int N>
__device__ __forceinline__ void bulk_wait_group_read_n() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(384, 1) gemm_ln_kernel(const __grid_constant__ GemmLnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("blm: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bfull = reinterpret_cast<uint64_t*>(smem + ln::kBarOff);
  uint64_t* bempty = bfull + ln::kBSlots;
  uint64_t* aempty = bempty + ln::kBSlots;
  uint64_t* tfull = aempty + ln::kASlots;
  uint64_t* tempty = tfull + 1;
  uint64_t* rfull = tempty + 1;  // [8 warps][2 buffers]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmR);
    tma_prefetch_desc(&p.tmO);
    tma_prefetch_desc(&p.tmH);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < ln::kBSlots; ++s) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
    }
    for (int s = 0; s < ln::kASlots; ++s) mbar_init(&aempty[s], 1);
    mbar_init(tfull, 1);
    mbar_init(tempty, 8);  // one arrive per epilogue warp
    for (int s = 0; s < 16; ++s) mbar_init(&rfull[s], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------ TMA producer
    if (lane == 0) {
      int bs = 0, as = 0;
      uint32_t bph = 0, aph = 0;
      for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x) {
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&aempty[as], aph ^ 1u);
          for (int h = 0; h < p.halves; ++h) {
            mbar_wait(&bempty[bs], bph ^ 1u);
            mbar_arrive_expect_tx(&bfull[bs], static_cast<uint32_t>(ln::kBBytes + (h == 0 ? ln::kABytes : 0)));
            if (h == 0)
              tma_load_2d(smem + ln::kAOff + as * ln::kABytes, &p.tmA, &bfull[bs], kb * kBK, t * kBM, kEvictNormal);
            tma_load_2d(smem + ln::kBOff + bs * ln::kBBytes, &p.tmB, &bfull[bs], kb * kBK, h * 256, kEvictLast);
            if (++bs == ln::kBSlots) {
              bs = 0;
              bph ^= 1u;
            }
          }
          if (++as == ln::kASlots) {
            as = 0;
            aph ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // -------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, 256);
      int bs = 0, as = 0;
      uint32_t bph = 0, tph = 0;
      for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x) {
        mbar_wait(tempty, tph ^ 1u);  // the epilogue has drained the previous tile
        tph ^= 1u;
        tcgen05_fence_after();
        for (int kb = 0; kb < p.kblocks; ++kb) {
          const uint64_t da = umma_desc_sw128(smem_u32(smem + ln::kAOff + as * ln::kABytes));
          for (int h = 0; h < p.halves; ++h) {
            mbar_wait(&bfull[bs], bph);
            tcgen05_fence_after();
            const uint64_t db = umma_desc_sw128(smem_u32(smem + ln::kBOff + bs * ln::kBBytes));
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(h * 256);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_ss(tmem_d, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                           (kb | k) != 0 ? 1u : 0u);
            umma_commit(&bempty[bs]);
            if (++bs == ln::kBSlots) {
              bs = 0;
              bph ^= 1u;
            }
          }
          umma_commit(&aempty[as]);
          if (++as == ln::kASlots) as = 0;
        }
        umma_commit(tfull);
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ---------------------------------------------------------- epilogue
    const int ew = warp - kEpiWarp0;
    const int lane_grp = warp & 3;
    const int col_grp = ew >> 2;
    const int nch = p.chunks_per_warp;              // 32-column chunks of this warp
    const int col_base = col_grp * nch * 32;        // first column of this warp
    uint8_t* buf0 = smem + ln::kStgOff + ew * 8192;
    uint8_t* buf1 = buf0 + 4096;
    uint64_t* rb = rfull + ew * 2;
    float2* xch = reinterpret_cast<float2*>(smem + ln::kXchOff);
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(col_base);
    const int sw = lane & 7;
    uint32_t rph0 = 0, rph1 = 0, tph = 0;
    const float inv_half = 1.0f / static_cast<float>(nch * 32);

    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x) {
      const int m0 = t * kBM + lane_grp * 32;
      const bool rows_ok = m0 < p.M;  // warp-uniform: this warp has at least one valid row
      // residual chunks 0 and 1 are fetched while the MMAs of this tile still run
      if (rows_ok && lane == 0) {
        bulk_wait_group_read0();  // the staging buffers were the sources of the previous tile's stores
        mbar_arrive_expect_tx(&rb[0], 4096u);
        tma_load_2d(buf0, &p.tmR, &rb[0], col_base, m0, kEvictFirst);
        if (nch > 1) {
          mbar_arrive_expect_tx(&rb[1], 4096u);
          tma_load_2d(buf1, &p.tmR, &rb[1], col_base + 32, m0, kEvictFirst);
        }
      }
      mbar_wait(tfull, tph);
      tph ^= 1u;
      tcgen05_fence_after();

      // ---- pass 1: x = acc + bias + resid -> TMEM; shifted sums over this warp's half row
      float shift = 0.0f, s1 = 0.0f, s2 = 0.0f;
      for (int c = 0; c < nch; ++c) {
        float v[32];
        __syncwarp();
        tmem_ld_32x32(tlane + static_cast<uint32_t>(c * 32), v);
        float4 r[8];
        if (rows_ok) {
          uint8_t* b = (c & 1) ? buf1 : buf0;
          if (c & 1) {
            mbar_wait(&rb[1], rph1);
            rph1 ^= 1u;
          } else {
            mbar_wait(&rb[0], rph0);
            rph0 ^= 1u;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) r[q] = *reinterpret_cast<const float4*>(b + lane * 128 + ((q ^ sw) << 4));
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) r[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
        const int col0 = col_base + c * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + q);
          v[4 * q] += bb.x + r[q].x;
          v[4 * q + 1] += bb.y + r[q].y;
          v[4 * q + 2] += bb.z + r[q].z;
          v[4 * q + 3] += bb.w + r[q].w;
        }
        // refill only after every lane has consumed its loads (see the pair kernel below)
        if (rows_ok) {
          __syncwarp();
          if (lane == 0 && c + 2 < nch) {
            mbar_arrive_expect_tx(&rb[c & 1], 4096u);
            tma_load_2d((c & 1) ? buf1 : buf0, &p.tmR, &rb[c & 1], col_base + (c + 2) * 32, m0, kEvictFirst);
          }
        }
        if (c == 0) shift = v[0];
        float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float d0 = v[j] - shift, d1 = v[j + 1] - shift;
          a0 += d0;
          a1 += d1;
          q0 = fmaf(d0, d0, q0);
          q1 = fmaf(d1, d1, q1);
        }
        s1 += a0 + a1;
        s2 += q0 + q1;
        tmem_st_32x32(tlane + static_cast<uint32_t>(c * 32), v);
      }
      tmem_st_wait();

      // ---- merge the two half rows (Chan): mean, M2 over n = nch * 32 elements each
      const float mean_w = shift + s1 * inv_half;
      const float m2_w = fmaxf(s2 - s1 * s1 * inv_half, 0.0f);
      const int row = lane_grp * 32 + lane;
      xch[col_grp * 128 + row] = make_float2(mean_w, m2_w);
      epi_bar_sync(256);
      const float2 o = xch[(col_grp ^ 1) * 128 + row];
      const float nh = static_cast<float>(nch * 32);
      const float dm = o.x - mean_w;
      const float mean = 0.5f * (mean_w + o.x);
      const float var = (m2_w + o.y + dm * dm * (0.5f * nh)) / (2.0f * nh);
      const float rstd = rsqrtf(var + p.eps);
      const float nmr = -mean * rstd;

      // ---- pass 2: fp32 output chunks through the two staging buffers
      if (p.has_f32) {
        for (int c = 0; c < nch; ++c) {
          float v[32];
          __syncwarp();
          tmem_ld_32x32(tlane + static_cast<uint32_t>(c * 32), v);
          const int col0 = col_base + c * 32;
          uint8_t* b = (c & 1) ? buf1 : buf0;
          if (lane == 0) bulk_wait_group_read1();  // the store issued two chunks ago has left this buffer
          tmem_ld_wait();
          __syncwarp();
          if (rows_ok) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col0) + q);
              const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + col0) + q);
              float4 y;
              y.x = fmaf(fmaf(v[4 * q], rstd, nmr), g.x, be.x);
              y.y = fmaf(fmaf(v[4 * q + 1], rstd, nmr), g.y, be.y);
              y.z = fmaf(fmaf(v[4 * q + 2], rstd, nmr), g.z, be.z);
              y.w = fmaf(fmaf(v[4 * q + 3], rstd, nmr), g.w, be.w);
              *reinterpret_cast<float4*>(b + lane * 128 + ((q ^ sw) << 4)) = y;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&p.tmO, b, col0, m0);
              bulk_commit_group();
            }
          }
        }
      }
      // ---- pass 3: bf16 output, [32 rows x 64 columns] per TMA store
      if (p.has_hi) {
        for (int c = 0; c < nch; c += 2) {
          float va[32], vb[32];
          __syncwarp();
          tmem_ld_32x32(tlane + static_cast<uint32_t>(c * 32), va);
          tmem_ld_32x32(tlane + static_cast<uint32_t>(c * 32 + 32), vb);
          const int col0 = col_base + c * 32;
          uint8_t* b = (c & 2) ? buf1 : buf0;
          if (lane == 0) bulk_wait_group_read1();
          tmem_ld_wait();
          __syncwarp();
          if (rows_ok) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col0) + q);
              const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + col0) + q);
              va[4 * q] = fmaf(fmaf(va[4 * q], rstd, nmr), g.x, be.x);
              va[4 * q + 1] = fmaf(fmaf(va[4 * q + 1], rstd, nmr), g.y, be.y);
              va[4 * q + 2] = fmaf(fmaf(va[4 * q + 2], rstd, nmr), g.z, be.z);
              va[4 * q + 3] = fmaf(fmaf(va[4 * q + 3], rstd, nmr), g.w, be.w);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + 32) + q);
              const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + col0 + 32) + q);
              vb[4 * q] = fmaf(fmaf(vb[4 * q], rstd, nmr), g.x, be.x);
              vb[4 * q + 1] = fmaf(fmaf(vb[4 * q + 1], rstd, nmr), g.y, be.y);
              vb[4 * q + 2] = fmaf(fmaf(vb[4 * q + 2], rstd, nmr), g.z, be.z);
              vb[4 * q + 3] = fmaf(fmaf(vb[4 * q + 3], rstd, nmr), g.w, be.w);
            }
            stage_chunk_bf16(va, b, lane, 0);
            stage_chunk_bf16(vb, b, lane, 1);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&p.tmH, b, col0, m0);
              bulk_commit_group();
            }
          }
        }
      }
      // every TMEM read of this tile has completed: hand the accumulator back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
    }
    if (lane == 0) bulk_wait_group0();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// ===================================================================================================
// v2 (N = 512): a CTA PAIR splits the row block by columns -- CTA r of the 2-CTA cluster owns columns
// [256 r, 256 r + 256) of the same 128 rows.  Each CTA then runs the plain GEMM pipeline on a 128 x 256 tile
// (4-stage 48 KB ring, TWO 256-column TMEM accumulator stages), so the LayerNorm epilogue of tile i overlaps
// the MMAs of tile i + 1 -- what the one-CTA kernel above cannot do with all 512 TMEM columns in one
// accumulator (measured: 315 us fused vs 279 us unfused at K = 4096).  The price is one exchange per tile:
// the two warps of a CTA that share a row merge their quarter-row (mean, M2) through shared memory, then
// the pair swaps half-row statistics through DISTRIBUTED shared memory (st.shared::cluster + a remote
// mbarrier arrive, 1 KB per tile), double buffered by tile parity.  The pair is otherwise uncoupled: each CTA
// has its own TMA producer, MMA issuer (cta_group::1) and TMEM.
namespace ln2 {
constexpr int kNH = 256;                       // columns per CTA
constexpr int kABytes = kBM * kBK * 2;         // 16 KB
constexpr int kBBytes = kNH * kBK * 2;         // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
template <int STAGES, int NBUF>
struct Smem {
  static constexpr int kStgOff = STAGES * kStageBytes;        // 8 warps x NBUF x 4 KB, 1024-B aligned
  static constexpr int kXchOff = kStgOff + 8 * NBUF * 4096;   // remote half-row statistics: float2 [2 parities][128]
  static constexpr int kBarOff = kXchOff + 2 * 128 * 8;
  // full[STAGES] empty[STAGES] tfull[2] tempty[2] rfull[8][NBUF] xbar[2]
  static constexpr int kNumBars = 2 * STAGES + 4 + 8 * NBUF + 2;
  static constexpr int kBytes = kBarOff + kNumBars * 8 + 16;
  static_assert(kBytes <= 232448, "shared memory budget of one sm_100 CTA");
};
}  // namespace ln2

__device__ __forceinline__ void st_cluster_f32x2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t ok = 0, polls = 0;
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (((++polls) & 0x3fffu) == 0u && (clock64() - t0) > 8000000000LL) {
      printf("blm: cluster mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

template <int STAGES, int NBUF>
__global__ void __launch_bounds__(384, 1) gemm_ln2_kernel(const __grid_constant__ GemmLnParams p) {
  using L = ln2::Smem<STAGES, NBUF>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("blm: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* rfull = tempty + 2;          // [8 warps][NBUF]
  uint64_t* xbar = rfull + 8 * NBUF;     // [2 parities]: 4 remote arrivals (the peer's column-group-0 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int ncol0 = static_cast<int>(rank) * ln2::kNH;   // first output column of this CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmR);
    tma_prefetch_desc(&p.tmO);
    tma_prefetch_desc(&p.tmH);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 8);
      mbar_init(&xbar[s], 4);
    }
    for (int s = 0; s < 8 * NBUF; ++s) mbar_init(&rfull[s], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers exist before anyone arrives on them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair_id; t < p.m_tiles; t += n_pairs) {
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], ln2::kStageBytes);
          uint8_t* st = smem + stage * ln2::kStageBytes;
          tma_load_2d(st, &p.tmA, &full[stage], kb * kBK, t * kBM, kEvictNormal);
          tma_load_2d(st + ln2::kABytes, &p.tmB, &full[stage], kb * kBK, ncol0, kEvictLast);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // -------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, ln2::kNH);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = pair_id; t < p.m_tiles; t += n_pairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * ln2::kNH);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t st = smem_u32(smem + stage * ln2::kStageBytes);
          const uint64_t da = umma_desc_sw128(st), db = umma_desc_sw128(st + ln2::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16_ss(tmem_d, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ---------------------------------------------------------- epilogue
    constexpr int kCh = ln2::kNH / 64;              // 4 chunks of 32 columns per warp
    const int ew = warp - kEpiWarp0;
    const int lane_grp = warp & 3;
    const int col_grp = ew >> 2;
    const int col_base = ncol0 + col_grp * (ln2::kNH / 2);   // first global column of this warp
    uint8_t* bufs = smem + L::kStgOff + ew * NBUF * 4096;
    uint64_t* rb = rfull + ew * NBUF;
    float2* xch = reinterpret_cast<float2*>(smem + L::kXchOff);
    const uint32_t xch_peer = mapa_shared(smem_u32(xch), rank ^ 1u);
    const uint32_t xbar_peer[2] = {mapa_shared(smem_u32(&xbar[0]), rank ^ 1u), mapa_shared(smem_u32(&xbar[1]), rank ^ 1u)};
    // local exchange slot of this warp's rows: the first 256 bytes of the PARTNER-visible staging buffer
    float2* lx_mine = reinterpret_cast<float2*>(bufs);
    float2* lx_partner = reinterpret_cast<float2*>(smem + L::kStgOff + (ew ^ 4) * NBUF * 4096);
    const int sw = lane & 7;
    const int row = lane_grp * 32 + lane;
    uint32_t rph[NBUF];
#pragma unroll
    for (int c = 0; c < NBUF; ++c) rph[c] = 0u;
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    constexpr float kInvQ = 1.0f / 128.0f;   // quarter row = 128 columns

    for (int t = pair_id; t < p.m_tiles; t += n_pairs, ++it) {
      const int m0 = t * kBM + lane_grp * 32;
      const bool rows_ok = m0 < p.M;
      const int par = it & 1;
      const uint32_t xph = static_cast<uint32_t>(it >> 1) & 1u;
      // prefetch the first residual chunk(s) while this tile's MMAs still run
      if (rows_ok && lane == 0) {
        bulk_wait_group_read0();
#pragma unroll
        for (int c = 0; c < NBUF; ++c) {
          mbar_arrive_expect_tx(&rb[c], 4096u);
          tma_load_2d(bufs + c * 4096, &p.tmR, &rb[c], col_base + c * 32, m0, kEvictFirst);
        }
      }
      mbar_wait(&tfull[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t tlane = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                             static_cast<uint32_t>(acc * ln2::kNH + col_grp * (ln2::kNH / 2));

      // ---- pass 1
      float shift = 0.0f, s1 = 0.0f, s2 = 0.0f;
#pragma unroll 1
      for (int c = 0; c < kCh; ++c) {
        float v[32];
        __syncwarp();
        tmem_ld_32x32(tlane + static_cast<uint32_t>(c * 32), v);
        float4 r[8];
        if (rows_ok) {
          const int bi = c % NBUF;
          uint8_t* b = bufs + bi * 4096;
          mbar_wait(&rb[bi], rph[bi]);
          rph[bi] ^= 1u;
#pragma unroll
          for (int q = 0; q < 8; ++q) r[q] = *reinterpret_cast<const float4*>(b + lane * 128 + ((q ^ sw) << 4));
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) r[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
        const int col0 = col_base + c * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + q);
          v[4 * q] += bb.x + r[q].x;
          v[4 * q + 1] += bb.y + r[q].y;
          v[4 * q + 2] += bb.z + r[q].z;
          v[4 * q + 3] += bb.w + r[q].w;
        }
        // The buffer is refilled only after every lane has CONSUMED its shared-memory loads (the adds above wait
        // on their scoreboards).  Issuing the TMA right after the loads were merely issued raced: with the tensor
        // core saturating shared-memory bandwidth under this epilogue (K = 4096 mainloop), queued LDS were
        // overtaken by the next chunk's TMA write (sporadic wrong rows, found by the float64 parity test).
        if (rows_ok) {
          __syncwarp();
          if (lane == 0 && c + NBUF < kCh) {
            const int bi = c % NBUF;
            mbar_arrive_expect_tx(&rb[bi], 4096u);
            tma_load_2d(bufs + bi * 4096, &p.tmR, &rb[bi], col_base + (c + NBUF) * 32, m0, kEvictFirst);
          }
        }
        if (c == 0) shift = v[0];
        float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float d0 = v[j] - shift, d1 = v[j + 1] - shift;
          a0 += d0;
          a1 += d1;
          q0 = fmaf(d0, d0, q0);
          q1 = fmaf(d1, d1, q1);
        }
        s1 += a0 + a1;
        s2 += q0 + q1;
        tmem_st_32x32(tlane + static_cast<uint32_t>(c * 32), v);
      }
      tmem_st_wait();

      // ---- quarter rows -> half row (the partner warp of this CTA), then half rows across the pair
      const float mean_q = shift + s1 * kInvQ;
      const float m2_q = fmaxf(s2 - s1 * s1 * kInvQ, 0.0f);
      __syncwarp();   // every lane has consumed the residual chunk that shared this buffer
      lx_mine[lane] = make_float2(mean_q, m2_q);
      epi_bar_sync(256);
      const float2 o = lx_partner[lane];
      const float dq = o.x - mean_q;
      const float mean_h = 0.5f * (mean_q + o.x);
      const float m2_h = m2_q + o.y + dq * dq * 64.0f;            // n_a n_b / (n_a + n_b) = 128 * 128 / 256
      if (col_grp == 0) {
        st_cluster_f32x2(xch_peer + static_cast<uint32_t>((par * 128 + row) * 8), mean_h, m2_h);
        fence_acq_rel_cluster();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(xbar_peer[par]);
      }
      mbar_wait_cluster(&xbar[par], xph);
      const float2 oh = xch[par * 128 + row];
      const float dh = oh.x - mean_h;
      const float mean = 0.5f * (mean_h + oh.x);
      const float var = (m2_h + oh.y + dh * dh * 128.0f) * (1.0f / 512.0f);   // 256 * 256 / 512
      const float rstd = rsqrtf(var + p.eps);
      const float nmr = -mean * rstd;
      epi_bar_sync(256);   // both partners have read the local slots before they become staging again

      // ---- pass 2: fp32 output chunks
      if (p.has_f32) {
#pragma unroll 1
        for (int c = 0; c < kCh; ++c) {
          float v[32];
          __syncwarp();
          tmem_ld_32x32(tlane + static_cast<uint32_t>(c * 32), v);
          const int col0 = col_base + c * 32;
          uint8_t* b = bufs + (c % NBUF) * 4096;
          if (lane == 0) bulk_wait_group_read_n<NBUF - 1>();   // the store that last used this buffer has left it
          tmem_ld_wait();
          __syncwarp();
          if (rows_ok) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col0) + q);
              const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + col0) + q);
              float4 y;
              y.x = fmaf(fmaf(v[4 * q], rstd, nmr), g.x, be.x);
              y.y = fmaf(fmaf(v[4 * q + 1], rstd, nmr), g.y, be.y);
              y.z = fmaf(fmaf(v[4 * q + 2], rstd, nmr), g.z, be.z);
              y.w = fmaf(fmaf(v[4 * q + 3], rstd, nmr), g.w, be.w);
              *reinterpret_cast<float4*>(b + lane * 128 + ((q ^ sw) << 4)) = y;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&p.tmO, b, col0, m0);
              bulk_commit_group();
            }
          }
        }
      }
      // ---- pass 3: bf16 output
      if (p.has_hi) {
#pragma unroll 1
        for (int c = 0; c < kCh; c += 2) {
          float va[32], vb[32];
          __syncwarp();
          tmem_ld_32x32(tlane + static_cast<uint32_t>(c * 32), va);
          tmem_ld_32x32(tlane + static_cast<uint32_t>(c * 32 + 32), vb);
          const int col0 = col_base + c * 32;
          uint8_t* b = bufs + ((c >> 1) % NBUF) * 4096;
          if (lane == 0) bulk_wait_group_read_n<NBUF - 1>();   // the store that last used this buffer has left it
          tmem_ld_wait();
          __syncwarp();
          if (rows_ok) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col0) + q);
              const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + col0) + q);
              va[4 * q] = fmaf(fmaf(va[4 * q], rstd, nmr), g.x, be.x);
              va[4 * q + 1] = fmaf(fmaf(va[4 * q + 1], rstd, nmr), g.y, be.y);
              va[4 * q + 2] = fmaf(fmaf(va[4 * q + 2], rstd, nmr), g.z, be.z);
              va[4 * q + 3] = fmaf(fmaf(va[4 * q + 3], rstd, nmr), g.w, be.w);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + 32) + q);
              const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + col0 + 32) + q);
              vb[4 * q] = fmaf(fmaf(vb[4 * q], rstd, nmr), g.x, be.x);
              vb[4 * q + 1] = fmaf(fmaf(vb[4 * q + 1], rstd, nmr), g.y, be.y);
              vb[4 * q + 2] = fmaf(fmaf(vb[4 * q + 2], rstd, nmr), g.z, be.z);
              vb[4 * q + 3] = fmaf(fmaf(vb[4 * q + 3], rstd, nmr), g.w, be.w);
            }
            stage_chunk_bf16(va, b, lane, 0);
            stage_chunk_bf16(vb, b, lane, 1);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&p.tmH, b, col0, m0);
              bulk_commit_group();
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if (lane == 0) bulk_wait_group0();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be writing statistics into this CTA's shared memory
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int STAGES, int NBUF>
static int launch_ln2(const GemmLnParams& p, cudaStream_t st) {
  int pairs = num_sms() / 2;
  if (pairs > p.m_tiles) pairs = p.m_tiles;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(2 * pairs));
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = ln2::Smem<STAGES, NBUF>::kBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BLM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_ln2_kernel<STAGES, NBUF>, p));
  return BLM_OK;
}

int gemm_ln_init() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ln::kSmemBytes));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_ln2_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ln2::Smem<4, 1>::kBytes));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_ln2_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ln2::Smem<3, 2>::kBytes));
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_ln2_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ln2::Smem<2, 4>::kBytes));
  return BLM_OK;
}

}  // namespace blm

extern "C" int blm_gemm_ln(const blm_gemm_ln_desc* d, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(d != nullptr, BLM_ERR_ARG, "null descriptor");
  BLM_REQUIRE(num_sms() > 0, BLM_ERR_ARCH, "blm_init() has not been called");
  BLM_REQUIRE(d->M > 0 && d->M < (1ll << 31) && d->K > 0, BLM_ERR_SHAPE, "bad shape M=%lld K=%lld", (long long)d->M,
              (long long)d->K);
  BLM_REQUIRE(d->N == 128 || d->N == 256 || d->N == 384 || d->N == 512, BLM_ERR_SHAPE,
              "blm_gemm_ln needs N in {128, 256, 384, 512}, got %lld", (long long)d->N);
  BLM_REQUIRE(d->A && d->B && d->resid && d->gamma && d->beta, BLM_ERR_ARG, "null operand");
  BLM_REQUIRE(d->out_f32 || d->out_hi, BLM_ERR_ARG, "no output buffer");
  BLM_REQUIRE(aligned16(d->bias) && aligned16(d->gamma) && aligned16(d->beta), BLM_ERR_ALIGN,
              "bias / gamma / beta must be 16-byte aligned");
  GemmLnParams p;
  memset(&p, 0, sizeof(p));
  int rc = encode_tmap_bf16(&p.tmA, d->A, d->M, d->K, d->lda, kBM);
  if (rc != BLM_OK) return rc;
  rc = encode_tmap_bf16(&p.tmB, d->B, d->N, d->K, d->ldb, 256);
  if (rc != BLM_OK) return rc;
  rc = encode_tmap_f32(&p.tmR, d->resid, d->M, d->N, d->ldr, 32);
  if (rc != BLM_OK) return rc;
  // unused output maps still have to be valid descriptors (they are prefetched): alias the residual
  rc = encode_tmap_f32(&p.tmO, d->out_f32 ? d->out_f32 : d->resid, d->M, d->N, d->out_f32 ? d->ldc : d->ldr, 32);
  if (rc != BLM_OK) return rc;
  if (d->out_hi) {
    rc = encode_tmap_bf16(&p.tmH, d->out_hi, d->M, d->N, d->ldc, 32);
    if (rc != BLM_OK) return rc;
  } else {
    p.tmH = p.tmA;
  }
  p.M = static_cast<int>(d->M);
  p.N = static_cast<int>(d->N);
  p.kblocks = static_cast<int>((d->K + kBK - 1) / kBK);
  p.m_tiles = static_cast<int>((d->M + kBM - 1) / kBM);
  p.halves = d->N > 256 ? 2 : 1;
  p.chunks_per_warp = static_cast<int>(d->N / 64);
  p.bias = d->bias;
  p.gamma = d->gamma;
  p.beta = d->beta;
  p.eps = d->eps;
  p.has_f32 = d->out_f32 != nullptr;
  p.has_hi = d->out_hi != nullptr;
  // N = 512: CTA pairs split the columns (two TMEM stages, epilogue overlapped); BLM_GEMM_LN_V1=1 keeps the
  // one-CTA full-row kernel for A/B
  static const bool force_v1 = getenv("BLM_GEMM_LN_V1") != nullptr;
  if (d->N == 512 && !force_v1 && num_sms() >= 2) {
    rc = encode_tmap_bf16(&p.tmB, d->B, d->N, d->K, d->ldb, ln2::kNH);
    if (rc != BLM_OK) return rc;
    // deep ring + one staging buffer when the mainloop hides the epilogue (K = 4096), shallower ring + double
    // buffered staging when the epilogue is the longer leg (K = 512)
    if (p.kblocks > 16) return launch_ln2<4, 1>(p, as_stream(stream));
    // <2, 4> (every residual chunk of the tile prefetched, two ring stages) measured slower: 90.6 vs 85.0 us at
    // M = 52833 -- the K = 512 mainloop needs the third ring stage more than the epilogue needs the buffers
    static const bool deep_stg = getenv("BLM_GEMM_LN_24") != nullptr;   // A/B switch
    if (deep_stg) return launch_ln2<2, 4>(p, as_stream(stream));
    return launch_ln2<3, 2>(p, as_stream(stream));
  }
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  gemm_ln_kernel<<<grid, 384, ln::kSmemBytes, as_stream(stream)>>>(p);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}
