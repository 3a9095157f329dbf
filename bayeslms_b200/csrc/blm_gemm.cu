// tcgen05 / TMEM / TMA GEMM family for sm_100a.
//
//   blm_gemm       C = epilogue(sum_s A_s B_s^T)        (a),(c): linear layers
//   blm_vocab_nll  nll = LSE(h E^T + b) - (h E^T + b)[t]  (d): logits stay in TMEM
//
// One persistent CTA per SM, (4 + EW) warps:
//   warp 0   TMA producer   (lane 0): A/B tiles -> 128B-swizzled smem ring
//   warp 1   MMA issuer     (lane 0): tcgen05.mma 128 x BN x 16, fp32 accum in TMEM
//   warp 2   TMEM allocator
//   warp 4.. EW epilogue warps (8 or 16): warp w reads TMEM lanes [32 (w % 4), +32) -- one
//            accumulator row per thread -- and owns column group (w - 4) / 4 of the tile: fused
//            bias / scale / GELU / GP-mix / residual / (hi,lo) split, or the online
//            log-sum-exp + target gather of the vocabulary sweep.  Two or four epilogue warps per
//            scheduler hide each other's TMEM-load, MUFU and dependency latencies (with one, the
//            epilogue issued 1 instruction in 5 cycles and ran 2.5x longer than the MMAs, r01d
//            profile); the tile's bias is staged once in shared memory instead of being fetched
//            from L2 by every thread for every chunk.
// Two TMEM accumulator stages let the MMA of tile i+1 overlap the epilogue of
// tile i.  Ragged M/N/K edges rely on TMA zero fill; the epilogue masks rows
// >= M and columns >= N.
#include <stdlib.h>
#include <string.h>

#include "blm_gemm_common.cuh"
#include "blm_philox.cuh"

namespace blm {

// ---- generate-once sampled weights (blm_gemm_sampled): W~ is built INSIDE the GEMM launch, once ----------
// A tile-stationary scheme regenerates every W~ tile for each group of M tiles that consumes it (128x for
// the FFN weight at M = 65536: ~270 M normals, which makes the generator warps, not the tensor pipe, the
// bound -- 0.37 of the tensor roofline).  The whole sampled weight is 4 MB: it fits the
// 126 MB L2 thirty times over.  So each CTA of the persistent grid draws 1/grid of W~ exactly once (Philox
// in registers, mu / sigma read once with 16-byte loads), writes it as bf16 to an L2-resident scratch
// tensor, and a grid-wide arrival counter gates the first B-tile TMA load of every CTA; from there on the
// kernel is the plain pipelined GEMM.  Same noise indexing and rounding as the tile-fused kernel:
// element (n, k) is lane (nK + k) % 8 of Philox counter (nK + k) / 8; W~ = bf16(fma(sigma, eps, mu)).
__device__ __forceinline__ void generate_weights(const GemmParams& p) {
  const long long groups = static_cast<long long>(p.N) * p.gen_K / 8;  // gen_K % 8 == 0
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const long long e0 = g * 8;
    const int n = static_cast<int>(e0 / p.gen_K), k = static_cast<int>(e0 - static_cast<long long>(n) * p.gen_K);
    float e[8];
    if (p.gen_eps) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.gen_eps + e0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.gen_eps + e0 + 4));
      e[0] = a.x, e[1] = a.y, e[2] = a.z, e[3] = a.w, e[4] = b.x, e[5] = b.y, e[6] = b.z, e[7] = b.w;
    } else {
      const Normal8 z = philox_normal8(p.gen_seed, p.gen_stream, static_cast<uint64_t>(g));
#pragma unroll
      for (int j = 0; j < 8; ++j) e[j] = z.v[j];
    }
    uint32_t o[4];
    if (p.gen_mu32) {
      const float* mp = p.gen_mu32 + static_cast<long long>(n) * p.gen_ldmu32 + k;
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(mp)), m1 = __ldg(reinterpret_cast<const float4*>(mp + 4));
      const float4 l0 = __ldg(reinterpret_cast<const float4*>(p.gen_lgstd32 + e0));
      const float4 l1 = __ldg(reinterpret_cast<const float4*>(p.gen_lgstd32 + e0 + 4));
      o[0] = pack_bf16x2(reparam_value(m0.x, l0.x, e[0]), reparam_value(m0.y, l0.y, e[1]));
      o[1] = pack_bf16x2(reparam_value(m0.z, l0.z, e[2]), reparam_value(m0.w, l0.w, e[3]));
      o[2] = pack_bf16x2(reparam_value(m1.x, l1.x, e[4]), reparam_value(m1.y, l1.y, e[5]));
      o[3] = pack_bf16x2(reparam_value(m1.z, l1.z, e[6]), reparam_value(m1.w, l1.w, e[7]));
    } else {
      const uint4 m = __ldg(reinterpret_cast<const uint4*>(p.gen_mu + static_cast<long long>(n) * p.gen_ldmu + k));
      const uint4 sg = __ldg(reinterpret_cast<const uint4*>(p.gen_sigma + e0));
      const uint32_t mw[4] = {m.x, m.y, m.z, m.w}, sw[4] = {sg.x, sg.y, sg.z, sg.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w0 = fmaf(__uint_as_float(sw[j] << 16), e[2 * j], __uint_as_float(mw[j] << 16));
        const float w1 = fmaf(__uint_as_float(sw[j] & 0xffff0000u), e[2 * j + 1], __uint_as_float(mw[j] & 0xffff0000u));
        o[j] = pack_bf16x2(w0, w1);
      }
    }
    *reinterpret_cast<uint4*>(p.gen_wt + e0) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  // generic-proxy global writes -> visible to the TMA loads (async proxy) of every CTA
  asm volatile("fence.proxy.async;" ::: "memory");
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(p.gen_sync, 1u);
}

// TMA producer thread: wait until every CTA of the (co-resident, persistent) grid has published its share
__device__ __forceinline__ void wait_generated(const GemmParams& p) {
  unsigned int seen;
  const long long t0 = clock64();
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.gen_sync) : "memory");
    if (seen < gridDim.x && (clock64() - t0) > 8000000000LL) {
      printf("blm: generate-once barrier timed out (block %d, %u of %u)\n", (int)blockIdx.x, seen, gridDim.x);
      __trap();
    }
  } while (seen < gridDim.x);
  asm volatile("fence.proxy.async;" ::: "memory");
  // the last CTA to leave re-arms the counters for the next launch (nobody is still polling by then)
  if (atomicAdd(p.gen_sync + 1, 1u) == gridDim.x - 1) {
    p.gen_sync[0] = 0u;
    p.gen_sync[1] = 0u;
    __threadfence();
  }
}

// CHUNK: the tensor core adds into its fp32 accumulator with truncation, one truncation per
// 16-wide K step, which shrinks every output by ~2e-8 x (K steps) relative (measured: -2e-5 at
// K = 3 x 4096).  With CHUNK the MMA warp closes the TMEM accumulator every p.chunk_kb K blocks and
// the epilogue warps sum the chunks in fp32 registers (round to nearest), so the bias is bounded by
// the chunk length, not by K (the precise bf16x3 mode uses chunks of 128 K elements).
template <int BN, int STAGES, int EPI, int ACT, int ARES, int EW, int CHUNK = 0, int STG = 0>
__global__ void __launch_bounds__((4 + EW) * 32, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
  using L = SmemLayout<BN, STAGES, ARES>;
  static_assert(EW == 8 || EW == 16, "epilogue warps");
  static_assert(!CHUNK || (BN == 128 && EW == 8 && EPI == EPI_STORE && ARES == 0),
                "chunked accumulation keeps a 128 x 128 tile in the registers of 8 epilogue warps");
  constexpr int kColGroups = EW / 4;                 // column groups of the tile, one per 4 warps
  constexpr int kChunks = BN / 32 / kColGroups;      // 32-column chunks per epilogue warp
  static_assert(kChunks >= 2 && (kChunks % 2) == 0, "the chunk loop is unrolled by two");  // (EW = 16: 2 chunks)
  constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                            : (2 * BN <= 256) ? 256 : 512;
  static_assert(2 * BN <= 512, "two accumulator stages must fit in TMEM");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {  // SWIZZLE_128B tiles need 1024-B aligned bases
    if (threadIdx.x == 0) printf("blm: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* ring = smem + L::kResBytes;
  float* sbias = reinterpret_cast<float*>(smem + L::kBiasOffset);  // [2][BN]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* afull_bar = tempty_bar + 2;
  uint64_t* aempty_bar = afull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (p.gen_wt) generate_weights(p);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.tmA[s]);
      tma_prefetch_desc(&p.tmB[s]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EW);  // one arrive per epilogue warp
    }
    mbar_init(afull_bar, 1);
    mbar_init(aempty_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int total_kb = 0;
  for (int s = 0; s < p.nseg; ++s) total_kb += p.kblocks[s];
  const bool x3 = CHUNK && p.x3;        // precise-mode triples: one stage = A.hi, A.lo, B.hi, B.lo of a K block
  if (x3) total_kb /= 3;                // K blocks of the distinct operands (each feeds three products)

  if (warp == 0) {
    // ------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      if (p.gen_wt) wait_generated(p);
      if constexpr (CHUNK) {
        if (x3) {
          // The split product hi*lo + lo*hi + hi*hi re-uses every operand tile twice.  Loading the three segments one
          // after the other (six tile loads per K block) made the precise mode bound by the L2 -> shared-memory
          // traffic of its 128 x 128 tiles (4.05x the bf16 step instead of 3x); here a stage of 64 KB = two ring slots
          // carries the four distinct tiles once.
          constexpr int kTile = L::kABytes;                       // 16 KB: [128 x 64] bf16 (BN == 128)
          constexpr int kStages3 = STAGES / 2;
          for (int w = blockIdx.x; w < p.num_works; w += gridDim.x) {
            const int m_tile = w / p.n_groups;
            const int grp = w - m_tile * p.n_groups;
            const int n0 = grp * p.tiles_per_group;
            const int n1 = min(p.n_tiles, n0 + p.tiles_per_group);
            for (int n = n0; n < n1; ++n)
              for (int ls = 0; ls < p.nseg; ls += 3)
                for (int kb = 0; kb < p.kblocks[ls]; ++kb) {
                  mbar_wait(&empty_bar[stage], phase ^ 1u);
                  mbar_arrive_expect_tx(&full_bar[stage], 4 * kTile);
                  uint8_t* st = ring + stage * 4 * kTile;
                  // segment ls = (A.hi, B.lo), ls + 1 = (A.lo, B.hi): their tensor maps name the four tiles
                  const CUtensorMap* ta[2] = {&p.tmA[ls], &p.tmA[ls + 1]};
                  const CUtensorMap* tb[2] = {&p.tmB[ls + 1], &p.tmB[ls]};
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    if (!p.a_mn) {
                      tma_load_2d(st + h * kTile, ta[h], &full_bar[stage], kb * kBK, m_tile * kBM, kEvictNormal);
                    } else {
#pragma unroll
                      for (int j = 0; j < kBM / 64; ++j)
                        tma_load_2d(st + h * kTile + j * 8192, ta[h], &full_bar[stage], m_tile * kBM + 64 * j, kb * kBK, kEvictNormal);
                    }
                    if (!p.b_mn) {
                      tma_load_2d(st + (2 + h) * kTile, tb[h], &full_bar[stage], kb * kBK, n * BN, kEvictLast);
                    } else {
#pragma unroll
                      for (int j = 0; j < BN / 64; ++j)
                        tma_load_2d(st + (2 + h) * kTile + j * 8192, tb[h], &full_bar[stage], n * BN + 64 * j, kb * kBK, kEvictLast);
                    }
                  }
                  if (++stage == kStages3) {
                    stage = 0;
                    phase ^= 1u;
                  }
                }
          }
        }
      }
      if (!x3)
      for (int w = blockIdx.x; w < p.num_works; w += gridDim.x) {
        const int m_tile = w / p.n_groups;
        const int grp = w - m_tile * p.n_groups;
        const int n0 = grp * p.tiles_per_group;
        const int n1 = min(p.n_tiles, n0 + p.tiles_per_group);
        if constexpr (ARES > 0) {
          // the previous work's MMAs must have drained the resident A before it is overwritten
          mbar_wait(aempty_bar, a_phase ^ 1u);
          mbar_arrive_expect_tx(afull_bar, static_cast<uint32_t>(total_kb * L::kABytes));
          int idx = 0;
          for (int s = 0; s < p.nseg; ++s)
            for (int kb = 0; kb < p.kblocks[s]; ++kb, ++idx)
              tma_load_2d(smem + idx * L::kABytes, &p.tmA[s], afull_bar, kb * kBK, m_tile * kBM, kEvictFirst);
          a_phase ^= 1u;
        }
        for (int n = n0; n < n1; ++n) {
          for (int s = 0; s < p.nseg; ++s) {
            const int kbs = p.kblocks[s];
            for (int kb = 0; kb < kbs; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1u);
              mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
              uint8_t* st = ring + stage * L::kStageBytes;
              // activations stream once; weights are re-read by every M tile and stay L2 resident
              if constexpr (ARES == 0) {
                if (!p.a_mn) {
                  tma_load_2d(st, &p.tmA[s], &full_bar[stage], kb * kBK, m_tile * kBM, kEvictNormal);
                } else {   // MN-major: [64 reduction rows x 64 elements] boxes, one per 64 rows of the tile
#pragma unroll
                  for (int j = 0; j < kBM / 64; ++j)
                    tma_load_2d(st + j * 8192, &p.tmA[s], &full_bar[stage], m_tile * kBM + 64 * j, kb * kBK, kEvictNormal);
                }
                st += L::kABytes;
              }
              if (!p.b_mn) {
                tma_load_2d(st, &p.tmB[s], &full_bar[stage], kb * kBK, n * BN, kEvictLast);
              } else {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_2d(st + j * 8192, &p.tmB[s], &full_bar[stage], n * BN + 64 * j, kb * kBK, kEvictLast);
              }
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // -------------------------------------------------------- MMA issuer
    if (lane == 0) {
      // A / B formats (bits [7,10) / [10,13)): 1 = bf16, 0 = f16.  Mixed A = f16 with B = bf16 is an illegal
      // instruction on sm_100a (tried), so the fp16 mode switches both operands
      const uint32_t idesc = (umma_idesc_bf16(kBM, BN) & ~(p.a_f16 ? ((1u << 7) | (1u << 10)) : 0u)) |
                             (p.a_mn ? (1u << 15) : 0u) | (p.b_mn ? (1u << 16) : 0u);   // bits 15 / 16: A / B MN-major
      const uint64_t a_step = p.a_mn ? 128u : 2u, b_step = p.b_mn ? 128u : 2u;          // per K = 16 step, in 16-byte units
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if constexpr (CHUNK) {
        if (x3) {
          constexpr int kTile = L::kABytes;
          constexpr int kStages3 = STAGES / 2;
          for (int w = blockIdx.x; w < p.num_works; w += gridDim.x) {
            const int m_tile = w / p.n_groups;
            const int grp = w - m_tile * p.n_groups;
            const int n0 = grp * p.tiles_per_group;
            const int n1 = min(p.n_tiles, n0 + p.tiles_per_group);
            for (int n = n0; n < n1; ++n)
              for (int kb0 = 0; kb0 < total_kb; kb0 += p.chunk_kb) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
                const int kb1 = min(total_kb, kb0 + p.chunk_kb);
                for (int kb = kb0; kb < kb1; ++kb) {
                  mbar_wait(&full_bar[stage], phase);
                  tcgen05_fence_after();
                  const uint32_t st = smem_u32(ring + stage * 4 * kTile);
                  uint64_t da[2], db[2];
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    da[h] = p.a_mn ? umma_desc_sw128_mn(st + h * kTile) : umma_desc_sw128(st + h * kTile);
                    db[h] = p.b_mn ? umma_desc_sw128_mn(st + (2 + h) * kTile) : umma_desc_sw128(st + (2 + h) * kTile);
                  }
                  // small terms first (the accumulator truncates relative to the running sum): hi*lo, lo*hi, hi*hi
                  const int ia[3] = {0, 1, 0}, ib[3] = {1, 0, 0};
#pragma unroll
                  for (int t3 = 0; t3 < 3; ++t3)
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                      umma_bf16_ss(tmem_d, da[ia[t3]] + static_cast<uint64_t>(k) * a_step, db[ib[t3]] + static_cast<uint64_t>(k) * b_step,
                                   idesc, ((kb - kb0) | k | t3) != 0 ? 1u : 0u);
                  umma_commit(&empty_bar[stage]);
                  if (++stage == kStages3) {
                    stage = 0;
                    phase ^= 1u;
                  }
                }
                umma_commit(&tfull_bar[acc]);
                if (++acc == 2) {
                  acc = 0;
                  acc_phase ^= 1u;
                }
              }
          }
        }
      }
      if (!x3)
      for (int w = blockIdx.x; w < p.num_works; w += gridDim.x) {
        const int m_tile = w / p.n_groups;
        const int grp = w - m_tile * p.n_groups;
        const int n0 = grp * p.tiles_per_group;
        const int n1 = min(p.n_tiles, n0 + p.tiles_per_group);
        if constexpr (ARES > 0) {
          mbar_wait(afull_bar, a_phase);
          a_phase ^= 1u;
        }
        for (int n = n0; n < n1; ++n) {
          const int chunk_kb = CHUNK ? p.chunk_kb : total_kb;
          for (int kb0 = 0; kb0 < total_kb; kb0 += chunk_kb) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
            const int kb1 = min(total_kb, kb0 + chunk_kb);
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(&full_bar[stage], phase);
              tcgen05_fence_after();
              const uint32_t st = smem_u32(ring + stage * L::kStageBytes);
              const uint32_t sa = ARES > 0 ? smem_u32(smem + kb * L::kABytes) : st;
              const uint32_t sb = ARES > 0 ? st : st + L::kABytes;
              const uint64_t da = p.a_mn ? umma_desc_sw128_mn(sa) : umma_desc_sw128(sa);
              const uint64_t db = p.b_mn ? umma_desc_sw128_mn(sb) : umma_desc_sw128(sb);
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                // K-major: +32 bytes per 16-element K step inside the 128-byte swizzle row; MN-major: +2 atoms
                umma_bf16_ss(tmem_d, da + static_cast<uint64_t>(k) * a_step, db + static_cast<uint64_t>(k) * b_step,
                             idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
              }
              umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
              }
            }
            umma_commit(&tfull_bar[acc]);  // accumulator (chunk) complete
            if (++acc == 2) {
              acc = 0;
              acc_phase ^= 1u;
            }
          }
        }
        if constexpr (ARES > 0) umma_commit(aempty_bar);  // resident A free once this work's MMAs retire
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ---------------------------------------------------------- epilogue
    const int lane_grp = warp & 3;  // TMEM lanes [32*lane_grp, +32) belong to this warp
    const int col_grp = (warp - kEpiWarp0) >> 2;
    const int etid = threadIdx.x - kEpiWarp0 * 32;
    const int row_in_tile = lane_grp * 32 + lane;
    const int c0 = col_grp * kChunks;  // first chunk of this warp inside the tile
    static_assert(!STG || (EPI == EPI_STORE && ((ARES == 0 && (EW == 8 || STG == 2)) || (ARES > 0 && STG == 2 && EW == 16))),
                  "store staging: 8 warps x 4 KB, or 16 warps x 2 KB (TMA store only; also the A-resident variant)");
    float4* stg = STG ? reinterpret_cast<float4*>(smem + L::kStgOffset + (warp - kEpiWarp0) * (EW == 16 ? 2048 : 4096))
                      : nullptr;  // store-transpose / TMA-store staging of this warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < p.num_works; w += gridDim.x) {
      const int m_tile = w / p.n_groups;
      const int grp = w - m_tile * p.n_groups;
      const int n0 = grp * p.tiles_per_group;
      const int n1 = min(p.n_tiles, n0 + p.tiles_per_group);
      const int m = m_tile * kBM + row_in_tile;
      const bool row_ok = m < p.M;
      const bool warp_rows_ok = m_tile * kBM + lane_grp * 32 < p.M;  // warp-uniform: any valid row in this warp
      (void)warp_rows_ok;

      NllState st{-INFINITY, 0.0f, -INFINITY, -1};
      if constexpr (EPI == EPI_NLL) {
        if (row_ok) st.tgt = __ldg(p.targets + m);
      }

      for (int n = n0; n < n1; ++n) {
        // this tile's bias: fetched while the MMAs still run, staged in shared memory for all rows
        float breg[(BN + EW * 32 - 1) / (EW * 32)];
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < (BN + EW * 32 - 1) / (EW * 32); ++i) {
            const int col = n * BN + etid + i * EW * 32;
            breg[i] = (etid + i * EW * 32 < BN && col < p.N) ? __ldg(p.bias + col) : 0.0f;
          }
        }
        if constexpr (CHUNK) {
          // sum the K chunks of this tile in registers: this warp owns 64 columns of its 32 rows
          float accr[2][32];
#pragma unroll
          for (int j = 0; j < 32; ++j) accr[0][j] = accr[1][j] = 0.0f;
          float* sbc = sbias;  // one bias buffer: a tile's bias is staged once, at its first chunk
          for (int kb0 = 0; kb0 < total_kb; kb0 += p.chunk_kb) {
            mbar_wait(&tfull_bar[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                                   static_cast<uint32_t>(acc * BN + c0 * 32);
            float va[32], vb[32];
            __syncwarp();
            tmem_ld_32x32(taddr, va);
            tmem_ld_32x32(taddr + 32u, vb);
            tmem_ld_wait();
            tcgen05_fence_before();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              accr[0][j] += va[j];
              accr[1][j] += vb[j];
            }
            if (++acc == 2) {
              acc = 0;
              acc_phase ^= 1u;
            }
          }
          if (p.bias) {
            epi_bar_sync(EW * 32);  // every warp is done with the previous tile's bias
#pragma unroll
            for (int i = 0; i < (BN + EW * 32 - 1) / (EW * 32); ++i)
              if (etid + i * EW * 32 < BN) sbc[etid + i * EW * 32] = breg[i];
            epi_bar_sync(EW * 32);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int col0 = n * BN + (c0 + c) * 32;
            if (col0 < p.N && m_tile * kBM + lane_grp * 32 < p.M)
              store_chunk<ACT, STG>(p, accr[c], m, row_ok, lane, col0, sbc + (c0 + c) * 32, stg);
          }
          continue;
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tcgen05_fence_after();
        float* sb = sbias + acc * BN;
        if (p.bias) {
          // sbias[acc] was last read for tile n-2, which every epilogue warp finished before it
          // passed the barrier of tile n-1
#pragma unroll
          for (int i = 0; i < (BN + EW * 32 - 1) / (EW * 32); ++i)
            if (etid + i * EW * 32 < BN) sb[etid + i * EW * 32] = breg[i];
          epi_bar_sync(EW * 32);
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                               static_cast<uint32_t>(acc * BN + c0 * 32);
        if constexpr (EW == 16 && STG == 2) {
          // Sixteen epilogue warps (four per scheduler): the packed-fp16 activations are issue / latency bound on
          // eight (273 us with GELU vs 246 us without an epilogue), so thread-level parallelism replaces the
          // two-chunk register pipeline: one 32-column chunk in registers (96-register budget at 640 threads),
          // 2 KB of staging per warp, [32 x 32] bf16 TMA stores (64B swizzle).
          float v[32];
#pragma unroll 1
          for (int c = 0; c < kChunks; ++c) {
            __syncwarp();
            tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
            tmem_ld_wait();
            if (c + 1 == kChunks) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            const int col0 = n * BN + (c0 + c) * 32;
            if (col0 < p.N && warp_rows_ok) {
              store_chunk<ACT, STG>(p, v, m, row_ok, lane, col0, sb + (c0 + c) * 32, stg);
              if (lane == 0) bulk_wait_group_read0();
              __syncwarp();
              stage_chunk_bf16_sw64(v, reinterpret_cast<uint8_t*>(stg), lane);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&p.tmC, stg, col0, m - lane);
                bulk_commit_group();
              }
            }
          }
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1u;
          }
          continue;
        }
        // two register buffers: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed
        float va[32], vb[32];
        __syncwarp();
        tmem_ld_32x32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < kChunks; c += 2) {
          tmem_ld_wait();
          __syncwarp();
          tmem_ld_32x32(taddr + static_cast<uint32_t>((c + 1) * 32), vb);
          {
            const int col0 = n * BN + (c0 + c) * 32;
            if (col0 < p.N) {
              if constexpr (EPI == EPI_STORE) {
                if (warp_rows_ok) {
                  store_chunk<ACT, STG>(p, va, m, row_ok, lane, col0, sb + (c0 + c) * 32, stg);
                  if constexpr (STG == 2) {
                    // the previous TMA store of this warp has finished reading the staging tile
                    if (lane == 0) bulk_wait_group_read0();
                    __syncwarp();
                    stage_chunk_bf16(va, reinterpret_cast<uint8_t*>(stg), lane, 0);
                  }
                }
              } else {
                nll_chunk(p, va, col0, st, sb + (c0 + c) * 32);
              }
            }
          }
          tmem_ld_wait();
          __syncwarp();
          if (c + 2 < kChunks) {
            tmem_ld_32x32(taddr + static_cast<uint32_t>((c + 2) * 32), va);
          } else {
            // every column this warp owns is in registers: hand the stage back to the MMA warp
            tcgen05_fence_before();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          {
            const int col0 = n * BN + (c0 + c + 1) * 32;
            if (col0 < p.N) {
              if constexpr (EPI == EPI_STORE) {
                if (warp_rows_ok) {
                  store_chunk<ACT, STG>(p, vb, m, row_ok, lane, col0, sb + (c0 + c + 1) * 32, stg);
                  if constexpr (STG == 2) stage_chunk_bf16(vb, reinterpret_cast<uint8_t*>(stg), lane, 1);
                }
              } else {
                nll_chunk(p, vb, col0, st, sb + (c0 + c + 1) * 32);
              }
            }
            if constexpr (EPI == EPI_STORE && STG == 2) {
              // one 32-row x 64-column bf16 tile per chunk pair leaves through the TMA store path
              if (warp_rows_ok && n * BN + (c0 + c) * 32 < p.N) {
                fence_proxy_async_smem();  // this lane's generic-proxy writes -> visible to the TMA engine
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&p.tmC, stg, n * BN + (c0 + c) * 32, m - lane);
                  bulk_commit_group();
                }
              }
            }
          }
        }
        if constexpr (EPI == EPI_NLL) {
          // Canonical vocabulary segments: the online log-sum-exp of a row restarts every p.seg_tiles tiles and leaves
          // one (max, sum) partial per (segment, column group), whatever group of tiles this work item covers.  The
          // partition of the vocabulary into work groups depends on M (it is chosen to fill whole waves), the
          // segments do not: a row's NLL is therefore bit-identical for every batch composition -- what makes the
          // scores of a list independent of how it is sharded over GPUs (tools/multi_gpu_parity.py).
          if ((n + 1) % p.seg_tiles == 0 || n + 1 == n1) {
            if (row_ok) {
              const long long o = static_cast<long long>((n / p.seg_tiles) * kColGroups + col_grp) * p.M + m;
              p.part_max[o] = st.run_max;
              p.part_sum[o] = st.run_sum;
            }
            st.run_max = -INFINITY;
            st.run_sum = 0.0f;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      if constexpr (EPI == EPI_NLL) {
        if (row_ok)   // the target logit is found by exactly one (group, column group): merged with max, order-free
          p.part_tgt[static_cast<long long>(grp * kColGroups + col_grp) * p.M + m] = st.tgt_logit;
      }
    }
  }

  if constexpr (STG == 2) {
    if (warp >= kEpiWarp0 && lane == 0) bulk_wait_group0();  // every committed TMA store has completed
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// nll[m] = ln2 * (gmax2 + log2(sum_g sum_g * 2^(max2_g - gmax2))) - target logit
__global__ void nll_merge_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                 const float* __restrict__ part_tgt, int groups, int tgt_groups, int M,
                                 float* __restrict__ nll, float* __restrict__ lse) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float gmax = -INFINITY, tgt = -INFINITY;
  for (int g = 0; g < groups; ++g) gmax = fmaxf(gmax, part_max[static_cast<long long>(g) * M + m]);
  for (int g = 0; g < tgt_groups; ++g) tgt = fmaxf(tgt, part_tgt[static_cast<long long>(g) * M + m]);
  float s = 0.0f;
  for (int g = 0; g < groups; ++g)
    s += part_sum[static_cast<long long>(g) * M + m] * exp2f(part_max[static_cast<long long>(g) * M + m] - gmax);
  constexpr float kLn2 = 0.6931471805599453f;
  const float l = kLn2 * (gmax + log2f(s));  // partial maxima are in the log2 domain
  nll[m] = l - tgt;
  if (lse) lse[m] = l;
}

__global__ void segment_sum_kernel(const float* __restrict__ x, const int* __restrict__ offs,
                                   long long nseg, float* __restrict__ out) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (i >= nseg) return;
  const int lane = threadIdx.x & 31;
  const int b = offs[i], e = offs[i + 1];
  float s = 0.0f;
  for (int t = b + lane; t < e; t += 32) s += x[t];
  s = warp_sum(s);
  if (lane == 0) out[i] = s;
}

// ------------------------------------------------------------------ host
constexpr int kStages256 = 4;
constexpr int kStages128 = 6;
constexpr int kNllAres = 8;        // resident A: 8 K blocks = K <= 512 (128 KB)
constexpr int kNllAresStages = 3;  // + 3 x 32 KB of streamed vocabulary tiles

// epilogue warps for the 128 x 256 tiles: 8 (default) or 16 (BLM_EPI_WARPS=16, A/B switch for profiling)
static int epi_warps256() {
  static const int ew = [] {
    const char* e = getenv("BLM_EPI_WARPS");
    return (e && atoi(e) == 16) ? 16 : 8;
  }();
  return ew;
}

template <int BN, int STAGES, int EPI, int ACT, int ARES, int EW>
static int set_smem_attr1() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, STAGES, EPI, ACT, ARES, EW>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SmemLayout<BN, STAGES, ARES>::kDynBytes));
  return BLM_OK;
}

template <int BN, int STAGES, int EPI, int ACT, int ARES = 0>
static int set_smem_attr() {
  int rc = set_smem_attr1<BN, STAGES, EPI, ACT, ARES, 8>();
  if constexpr (BN == 256) {
    if (rc == BLM_OK) rc = set_smem_attr1<BN, STAGES, EPI, ACT, ARES, 16>();
  }
  return rc;
}

template <int ACT>
static int set_smem_attr_chunk() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<128, kStages128, EPI_STORE, ACT, 0, 8, 1>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SmemLayout<128, kStages128, 0>::kDynBytes));
  return BLM_OK;
}

// Every GEMM launch goes through here.  A launch that draws its sampled weights first (generate_weights /
// wait_generated: a grid-wide spin barrier) must have ALL its CTAs resident at once: it is launched cooperatively, so
// the driver either co-schedules the whole grid -- also when kernels of other streams occupy SMs -- or refuses the
// launch with an error; it can no longer start partially and spin until the watchdog trap.
template <typename Kernel>
static int launch_gemm(Kernel kernel, int grid, int threads, int smem, cudaStream_t st, const GemmParams& p) {
  if (p.gen_wt) {
    int per_sm = 0;
    BLM_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    BLM_REQUIRE(static_cast<long long>(per_sm) * num_sms() >= grid, BLM_ERR_SHAPE,
                "generate-once sampled GEMM: a grid of %d CTAs cannot be co-resident (%d per SM x %d SMs)", grid, per_sm,
                num_sms());
    void* args[] = {const_cast<GemmParams*>(&p)};
    BLM_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(grid), dim3(threads), args, smem, st));
    return BLM_OK;
  }
  kernel<<<grid, threads, smem, st>>>(p);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

template <int ACT>
static int launch_chunk(const GemmParams& p, cudaStream_t st) {
  const int grid = p.num_works < num_sms() ? p.num_works : num_sms();
  return launch_gemm(gemm_kernel<128, kStages128, EPI_STORE, ACT, 0, 8, 1>, grid, (4 + 8) * 32, SmemLayout<128, kStages128, 0>::kDynBytes, st, p);
}

// transposed-store (STG) variants: the fp32-output GEMMs (QKV in training, o_net, FFN2) and the GELU-gradient
// dgrad (dz1 = (dF W2) * gelu'(z1): fp32 + bf16 outputs and the saved pre-activation read, all row-coalesced)
template <int BN, int STAGES, int CHUNK, int ACT = BLM_ACT_NONE>
static int set_smem_attr_stg() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, STAGES, EPI_STORE, ACT, 0, 8, CHUNK, 1>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SmemLayout<BN, STAGES, 0>::kDynBytes));
  return BLM_OK;
}

template <int BN, int STAGES, int CHUNK, int ACT = BLM_ACT_NONE>
static int launch_stg(const GemmParams& p, cudaStream_t st) {
  const int grid = p.num_works < num_sms() ? p.num_works : num_sms();
  return launch_gemm(gemm_kernel<BN, STAGES, EPI_STORE, ACT, 0, 8, CHUNK, 1>, grid, (4 + 8) * 32, SmemLayout<BN, STAGES, 0>::kDynBytes, st, p);
}

// TMA-store (STG == 2) variants: bf16-hi-only outputs of the forward GEMMs (QKV, FFN1)
template <int BN, int STAGES, int ACT>
static int set_smem_attr_tma() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, STAGES, EPI_STORE, ACT, 0, 8, 0, 2>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SmemLayout<BN, STAGES, 0>::kDynBytes));
  return BLM_OK;
}

// 16 epilogue warps + [32 x 32] TMA stores: the packed-fp16 activation epilogues (p.tmC: 32 x 32 box, 64B swizzle)
template <int ACT>
static int set_smem_attr_tma16() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<256, kStages256, EPI_STORE, ACT, 0, 16, 0, 2>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<256, kStages256, 0>::kDynBytes));
  return BLM_OK;
}

template <int ACT>
static int launch_tma16(const GemmParams& p, cudaStream_t st) {
  const int grid = p.num_works < num_sms() ? p.num_works : num_sms();
  return launch_gemm(gemm_kernel<256, kStages256, EPI_STORE, ACT, 0, 16, 0, 2>, grid, (4 + 16) * 32, SmemLayout<256, kStages256, 0>::kDynBytes, st, p);
}

// A-resident form of the 16-epilogue-warp TMA-store kernel for K <= 512 (QKV, FFN1): the 128 x 512 activation
// tile stays in shared memory while the CTA sweeps a group of N tiles, only weight tiles stream (2 x 32 KB ring)
constexpr int kAresStoreStages = 2;
template <int ACT>
static int set_smem_attr_tma16_ares() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<256, kAresStoreStages, EPI_STORE, ACT, kNllAres, 16, 0, 2>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SmemLayout<256, kAresStoreStages, kNllAres>::kDynBytes));
  return BLM_OK;
}

template <int ACT>
static int launch_tma16_ares(const GemmParams& p, cudaStream_t st) {
  const int grid = p.num_works < num_sms() ? p.num_works : num_sms();
  return launch_gemm(gemm_kernel<256, kAresStoreStages, EPI_STORE, ACT, kNllAres, 16, 0, 2>, grid, (4 + 16) * 32, SmemLayout<256, kAresStoreStages, kNllAres>::kDynBytes, st, p);
}

template <int BN, int STAGES, int ACT>
static int launch_tma(const GemmParams& p, cudaStream_t st) {
  const int grid = p.num_works < num_sms() ? p.num_works : num_sms();
  return launch_gemm(gemm_kernel<BN, STAGES, EPI_STORE, ACT, 0, 8, 0, 2>, grid, (4 + 8) * 32, SmemLayout<BN, STAGES, 0>::kDynBytes, st, p);
}

int gemm_init() {
  int rc;
  if ((rc = set_smem_attr_tma16_ares<BLM_ACT_GELU_FAST>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma16_ares<BLM_ACT_NONE>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma16<BLM_ACT_GELU_FAST>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma16<BLM_ACT_GPMIX_FAST>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma<256, kStages256, BLM_ACT_NONE>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma<256, kStages256, BLM_ACT_GELU>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma<256, kStages256, BLM_ACT_GELU_FAST>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma<256, kStages256, BLM_ACT_GPMIX_FAST>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma<128, kStages128, BLM_ACT_GPMIX_FAST>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma<256, kStages256, BLM_ACT_GPMIX>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma<128, kStages128, BLM_ACT_GPMIX>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma<128, kStages128, BLM_ACT_NONE>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma<128, kStages128, BLM_ACT_GELU>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_tma<128, kStages128, BLM_ACT_GELU_FAST>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_stg<256, kStages256, 0>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_stg<128, kStages128, 0>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_stg<128, kStages128, 1>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_stg<256, kStages256, 0, BLM_ACT_GELU_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_stg<128, kStages128, 0, BLM_ACT_GELU_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_stg<128, kStages128, 1, BLM_ACT_GELU_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_stg<128, kStages128, 1, BLM_ACT_GELU>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_stg<128, kStages128, 1, BLM_ACT_GPMIX>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_chunk<BLM_ACT_NONE>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_chunk<BLM_ACT_GELU>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_chunk<BLM_ACT_GPMIX>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_chunk<BLM_ACT_SOFTMAX_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_chunk<BLM_ACT_GELU_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr_chunk<BLM_ACT_GPMIX_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kStages256, EPI_STORE, BLM_ACT_NONE>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kStages256, EPI_STORE, BLM_ACT_GELU>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kStages256, EPI_STORE, BLM_ACT_GPMIX>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<128, kStages128, EPI_STORE, BLM_ACT_NONE>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<128, kStages128, EPI_STORE, BLM_ACT_GELU>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<128, kStages128, EPI_STORE, BLM_ACT_GPMIX>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kStages256, EPI_STORE, BLM_ACT_SOFTMAX_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<128, kStages128, EPI_STORE, BLM_ACT_SOFTMAX_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kStages256, EPI_STORE, BLM_ACT_GELU_FAST>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<128, kStages128, EPI_STORE, BLM_ACT_GELU_FAST>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kStages256, EPI_STORE, BLM_ACT_GELU_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<128, kStages128, EPI_STORE, BLM_ACT_GELU_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kStages256, EPI_STORE, BLM_ACT_GPMIX_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<128, kStages128, EPI_STORE, BLM_ACT_GPMIX_GRAD>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kStages256, EPI_NLL, BLM_ACT_NONE>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kNllAresStages, EPI_NLL, BLM_ACT_NONE, kNllAres>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kNllAresStages, EPI_STORE, BLM_ACT_NONE, kNllAres>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kNllAresStages, EPI_STORE, BLM_ACT_GELU, kNllAres>()) != BLM_OK) return rc;
  if ((rc = set_smem_attr<256, kNllAresStages, EPI_STORE, BLM_ACT_GPMIX, kNllAres>()) != BLM_OK) return rc;
  return BLM_OK;
}

template <int BN, int STAGES, int EPI, int ACT, int ARES = 0>
static int launch(const GemmParams& p, cudaStream_t st) {
  const int grid = p.num_works < num_sms() ? p.num_works : num_sms();
  constexpr int smem = SmemLayout<BN, STAGES, ARES>::kDynBytes;
  if constexpr (BN == 256) {
    if (epi_warps256() == 16) {
      return launch_gemm(gemm_kernel<BN, STAGES, EPI, ACT, ARES, 16>, grid, (4 + 16) * 32, smem, st, p);
    }
  }
  return launch_gemm(gemm_kernel<BN, STAGES, EPI, ACT, ARES, 8>, grid, (4 + 8) * 32, smem, st, p);
}

int gemm2_store(GemmParams p, int act, cudaStream_t st, int tma_store);   // blm_gemm2.cu: CTA-pair (cta_group::2) kernels
int gemm2_nll(GemmParams p, int groups, cudaStream_t st);

// CTA-pair path switch: BLM_GEMM2=0 disables, =1 enables (default set below after measurement)
static bool use_gemm2() {
  static const bool on = [] {
    const char* e = getenv("BLM_GEMM2");
    return e ? atoi(e) != 0 : false;
  }();
  return on;
}

static int fill_segments(GemmParams& p, int nseg, const blm_bf16* const* A, const blm_bf16* const* B,
                         const int64_t* K, const int64_t* lda, const int64_t* ldb, int64_t M,
                         int64_t N, int BN, bool a_mn = false, bool b_mn = false) {
  BLM_REQUIRE(nseg >= 1 && nseg <= BLM_MAX_SEG, BLM_ERR_ARG, "nseg=%d out of range", nseg);
  p.nseg = nseg;
  for (int s = 0; s < nseg; ++s) {
    BLM_REQUIRE(A[s] && B[s], BLM_ERR_ARG, "segment %d has a null operand", s);
    BLM_REQUIRE(K[s] > 0, BLM_ERR_SHAPE, "K[%d]=%lld must be positive", s, (long long)K[s]);
    // MN-major operands are [K_s, M] / [K_s, N] row-major: boxes of 64 reduction rows x 64 contiguous elements
    int rc = a_mn ? encode_tmap_bf16(&p.tmA[s], A[s], K[s], M, lda[s], 64) : encode_tmap_bf16(&p.tmA[s], A[s], M, K[s], lda[s], kBM);
    if (rc != BLM_OK) return rc;
    rc = b_mn ? encode_tmap_bf16(&p.tmB[s], B[s], K[s], N, ldb[s], 64) : encode_tmap_bf16(&p.tmB[s], B[s], N, K[s], ldb[s], BN);
    if (rc != BLM_OK) return rc;
    p.kblocks[s] = static_cast<int>((K[s] + kBK - 1) / kBK);
  }
  return BLM_OK;
}

}  // namespace blm

namespace blm {

// gen != null: generate-once sampled weights (see generate_weights); d->B[0] is then the scratch tensor
int gemm_impl(const blm_gemm_desc* d, const GemmGen* gen, blm_stream stream) {
  BLM_REQUIRE(d != nullptr, BLM_ERR_ARG, "null descriptor");
  BLM_REQUIRE(num_sms() > 0, BLM_ERR_ARCH, "blm_init() has not been called");
  BLM_REQUIRE(d->M > 0 && d->N > 0 && d->M < (1ll << 31) && d->N < (1ll << 31), BLM_ERR_SHAPE,
              "bad GEMM shape M=%lld N=%lld", (long long)d->M, (long long)d->N);
  BLM_REQUIRE(d->out_f32 || d->out_hi, BLM_ERR_ARG, "no output buffer");
  BLM_REQUIRE(!d->out_lo || d->out_hi, BLM_ERR_ARG, "out_lo requires out_hi");
  BLM_REQUIRE((d->ldc % 8) == 0 && d->ldc >= d->N, BLM_ERR_ALIGN, "ldc=%lld", (long long)d->ldc);
  BLM_REQUIRE(aligned16(d->out_f32) && aligned16(d->out_hi) && aligned16(d->out_lo) && aligned16(d->resid),
              BLM_ERR_ALIGN, "output / residual pointers must be 16-byte aligned");
  BLM_REQUIRE(!d->resid || ((d->ldr % 4) == 0 && d->ldr >= d->N), BLM_ERR_ALIGN, "ldr=%lld",
              (long long)d->ldr);
  BLM_REQUIRE(d->act >= BLM_ACT_NONE && d->act <= BLM_ACT_GPMIX_FAST, BLM_ERR_ARG, "unknown activation %d",
              d->act);
  BLM_REQUIRE((d->act != BLM_ACT_GELU_FAST && d->act != BLM_ACT_GPMIX_FAST) ||
                  (d->out_hi && !d->out_lo && !d->out_f32 && !d->out_pre && d->k_chunk == 0),
              BLM_ERR_ARG, "the packed-fp16 activations are for bf16-hi-only outputs of the fast mode");
  BLM_REQUIRE(d->act != BLM_ACT_GPMIX_FAST || ((d->N % 2) == 0 && !d->resid && aligned16(d->coef)), BLM_ERR_ARG,
              "BLM_ACT_GPMIX_FAST needs an even N, no residual and a 16-byte aligned coef");
  BLM_REQUIRE(!d->a_f16 || (d->nseg == 1 && d->k_chunk == 0), BLM_ERR_ARG, "fp16 operands: one segment, no k_chunk");
  const bool grad_act = d->act == BLM_ACT_GELU_GRAD || d->act == BLM_ACT_GPMIX_GRAD;
  BLM_REQUIRE(!grad_act || (d->aux && (d->ldaux % 4) == 0 && d->ldaux >= d->N && aligned16(d->aux)), BLM_ERR_ARG,
              "the activation-gradient epilogues need aux (16-byte aligned, ldaux %% 4 == 0)");
  BLM_REQUIRE(aligned16(d->out_pre), BLM_ERR_ALIGN, "out_pre must be 16-byte aligned");
  BLM_REQUIRE(d->act != BLM_ACT_SOFTMAX_GRAD || (d->lse && d->targets), BLM_ERR_ARG,
              "softmax-grad epilogue needs lse and targets");
  BLM_REQUIRE((d->act != BLM_ACT_GPMIX && d->act != BLM_ACT_GPMIX_GRAD && d->act != BLM_ACT_GPMIX_FAST) || d->coef, BLM_ERR_ARG,
              "GP-mix epilogue needs coef");

  // Tile choice: 128x256 tiles unless that leaves most SMs idle, then 128x128.
  const int m_tiles = static_cast<int>((d->M + kBM - 1) / kBM);
  const int n_tiles256 = static_cast<int>((d->N + 255) / 256);
  BLM_REQUIRE(d->k_chunk >= 0 && (d->k_chunk % kBK) == 0, BLM_ERR_ARG, "k_chunk=%d must be a multiple of %d",
              d->k_chunk, kBK);
  long long total_k = 0;
  for (int s = 0; s < d->nseg && s < BLM_MAX_SEG; ++s) total_k += (d->K[s] + kBK - 1) / kBK * kBK;
  const bool chunked = d->k_chunk > 0 && total_k > d->k_chunk;
  bool use256 = !chunked && (d->N >= 256) && (static_cast<long long>(m_tiles) * n_tiles256 >= num_sms());
  if (use256) {
    // few waves: compare the wave counts of both tile widths (a 128-wide tile costs a little more than half a 256-wide
    // one).  QKV of the fine-tune step: 25 x 6 = 150 wide tiles on 148 SMs are TWO waves, 300 narrow ones 2.03 narrow
    // waves = 1.65 wide-tile times.  From 4 wide waves up the wide tile always wins.
    const long long sms = num_sms();
    const long long t256 = static_cast<long long>(m_tiles) * n_tiles256;
    const long long t128 = static_cast<long long>(m_tiles) * ((d->N + 127) / 128);
    const long long w256 = (t256 + sms - 1) / sms, w128 = (t128 + sms - 1) / sms;
    static const bool no_wave = getenv("BLM_GEMM_NO_WAVE_FIT") != nullptr;   // A/B switch
    if (!no_wave && w256 <= 3 && 55 * w128 < 100 * w256) use256 = false;
  }
  const int BN = use256 ? 256 : 128;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_segments(p, d->nseg, d->A, d->B, d->K, d->lda, d->ldb, d->M, d->N, BN, d->a_mn != 0, d->b_mn != 0);
  if (rc != BLM_OK) return rc;
  p.a_mn = d->a_mn != 0;
  p.b_mn = d->b_mn != 0;
  p.M = static_cast<int>(d->M);
  p.N = static_cast<int>(d->N);
  p.m_tiles = m_tiles;
  p.n_tiles = static_cast<int>((d->N + BN - 1) / BN);
  p.n_groups = p.n_tiles;
  p.tiles_per_group = 1;
  p.num_works = p.m_tiles * p.n_tiles;
  p.bias = d->bias;
  p.coef = d->coef;
  p.col_scale = d->col_scale;
  p.col_scale_cols = d->col_scale_cols;
  p.resid = d->resid;
  p.ldr = d->ldr;
  p.out_f32 = d->out_f32;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(d->out_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(d->out_lo);
  p.ldc = d->ldc;
  p.lse = d->lse;
  p.targets = d->targets;
  p.grad_scale = d->grad_scale;
  // store transpose: pays for fp32 outputs (QKV fp32 228 -> 149 us at M = 65536); for bf16-only outputs the
  // extra shared-memory round trip costs more issue slots than the wider stores save (FFN1 479 -> 536 us)
  static const char* stg_env = getenv("BLM_STG");  // A/B switch for profiling: 0 = never, 1 = always
  p.use_stg = stg_env ? atoi(stg_env) : (d->out_f32 != nullptr || (d->k_chunk > 0 && d->out_lo != nullptr));
  p.f32_rows32 = d->f32_rows32;
  if (d->drop && (d->drop->mask || d->drop->p > 0.0f)) {
    BLM_REQUIRE((d->N % 32) == 0 && d->act != BLM_ACT_SOFTMAX_GRAD && !d->f32_rows32, BLM_ERR_ARG,
                "fused dropout needs N %% 32 == 0 (N = %lld) and a storing epilogue other than the softmax gradient",
                (long long)d->N);
    BLM_REQUIRE(d->drop->p >= 0.0f && d->drop->p < 1.0f, BLM_ERR_ARG, "dropout p = %f", (double)d->drop->p);
    p.drop_on = 1;
    p.drop = make_drop_params(d->drop->mask, d->drop->p, d->drop->seed, d->drop->seed_dev, d->drop->stream_id);
  }
  if (d->f32_rows32) {
    BLM_REQUIRE(d->out_f32 && !d->out_pre && d->ldc == d->N && (d->N % 4) == 0, BLM_ERR_ARG,
                "f32_rows32 needs out_f32 with ldc == N, N %% 4 == 0 and no out_pre");
    p.use_stg = 0;   // one row per thread IS the coalesced store of this layout
  }
  p.out_pre = d->out_pre;
  p.aux = d->aux;
  p.ldaux = d->ldaux;
  p.a_f16 = d->a_f16;
  p.fast_act = d->fast_act;
  if (gen) {
    p.gen_mu = reinterpret_cast<const __nv_bfloat16*>(gen->mu);
    p.gen_ldmu = gen->ldmu;
    p.gen_sigma = reinterpret_cast<const __nv_bfloat16*>(gen->sigma);
    p.gen_eps = gen->eps;
    p.gen_mu32 = gen->mu32;
    p.gen_ldmu32 = gen->ldmu32;
    p.gen_lgstd32 = gen->lgstd32;
    p.gen_seed = gen->seed;
    p.gen_stream = gen->stream_id;
    p.gen_K = static_cast<int>(d->K[0]);
    p.gen_wt = reinterpret_cast<__nv_bfloat16*>(gen->wt);
    p.gen_sync = gen->sync;
  }
  cudaStream_t st = as_stream(stream);
  // CTA-pair kernel: one bf16 segment, bf16-only output, forward activations, enough 256 x 256 tiles
  if (!gen && !d->a_f16 && !d->a_mn && !d->b_mn && use_gemm2() && !chunked && d->nseg == 1 && !d->out_f32 && !d->out_pre && d->N >= 256 &&
      (d->act == BLM_ACT_NONE || d->act == BLM_ACT_GELU || d->act == BLM_ACT_GELU_FAST) &&
      static_cast<long long>((d->M + 255) / 256) * ((d->N + 255) / 256) >= num_sms() / 2) {
    GemmParams p2 = p;
    rc = fill_segments(p2, d->nseg, d->A, d->B, d->K, d->lda, d->ldb, d->M, d->N, 128);  // B boxes of 128 rows
    if (rc != BLM_OK) return rc;
    static const bool tma2 = [] {
      const char* e = getenv("BLM_GEMM2_TMA_STORE");
      return e ? atoi(e) != 0 : true;
    }();
    const bool ts = tma2 && d->out_hi && !d->out_lo && !d->resid;
    static const bool pair16 = [] {
      const char* e = getenv("BLM_EPI16");
      return e ? atoi(e) != 0 : true;
    }();
    if (ts && pair16 && d->act == BLM_ACT_GELU_FAST) {
      rc = encode_tmap_bf16_box32(&p2.tmC, d->out_hi, d->M, d->N, d->ldc);
      if (rc != BLM_OK) return rc;
      return gemm2_store(p2, d->act, st, 2);
    }
    if (ts) {
      rc = encode_tmap_bf16(&p2.tmC, d->out_hi, d->M, d->N, d->ldc, 32);
      if (rc != BLM_OK) return rc;
    }
    return gemm2_store(p2, d->act, st, ts ? 1 : 0);
  }
  // bf16-hi-only output of a forward GEMM: the tile leaves through TMA stores (row-per-thread 16-byte stores to
  // rows 8 KB apart back up the LSU / L2 request queues; BLM_TMA_STORE=0 is the A/B switch)
  static const bool tma_store_on = [] {
    const char* e = getenv("BLM_TMA_STORE");
    return e ? atoi(e) != 0 : true;
  }();
  const bool tma_ok = tma_store_on && !chunked && d->out_hi && !d->out_lo && !d->out_f32 && !d->out_pre && !d->resid && !p.drop_on;
  blm_gemm_desc downgraded;
  if (d->act == BLM_ACT_GPMIX_FAST && !tma_ok) {   // the packed variant only exists on the TMA-store path
    downgraded = *d;
    downgraded.act = BLM_ACT_GPMIX;
    d = &downgraded;
  }
  if (tma_ok &&
      (d->act == BLM_ACT_NONE || d->act == BLM_ACT_GELU || d->act == BLM_ACT_GELU_FAST || d->act == BLM_ACT_GPMIX ||
       d->act == BLM_ACT_GPMIX_FAST)) {
    static const bool ew16 = [] {
      const char* e = getenv("BLM_EPI16");   // A/B switch: 0 keeps the packed-fp16 epilogues on 8 warps
      return e ? atoi(e) != 0 : true;
    }();
    static const bool ares16 = getenv("BLM_GEMM_ARES16") != nullptr;   // A/B switch (opt-in experiment)
    if (ares16 && ew16 && BN == 256 && d->nseg == 1 && p.kblocks[0] <= kNllAres && p.n_tiles >= 4 &&
        (d->act == BLM_ACT_GELU_FAST || d->act == BLM_ACT_NONE)) {
      // groups of N tiles per work item: as few A reloads as possible at >= 95 % wave efficiency
      int best_g = p.n_tiles;
      for (int g = 1; g <= p.n_tiles; g *= 2) {
        const long long works = static_cast<long long>(p.m_tiles) * g;
        const long long rounds = (works + num_sms() - 1) / num_sms();
        if (static_cast<double>(works) / (rounds * num_sms()) >= 0.95) {
          best_g = g;
          break;
        }
      }
      p.n_groups = best_g;
      p.tiles_per_group = (p.n_tiles + best_g - 1) / best_g;
      p.n_groups = (p.n_tiles + p.tiles_per_group - 1) / p.tiles_per_group;
      p.num_works = p.m_tiles * p.n_groups;
      rc = encode_tmap_bf16_box32(&p.tmC, d->out_hi, d->M, d->N, d->ldc);
      if (rc != BLM_OK) return rc;
      return d->act == BLM_ACT_GELU_FAST ? launch_tma16_ares<BLM_ACT_GELU_FAST>(p, st) : launch_tma16_ares<BLM_ACT_NONE>(p, st);
    }
    if (ew16 && BN == 256 && (d->act == BLM_ACT_GELU_FAST || d->act == BLM_ACT_GPMIX_FAST)) {
      rc = encode_tmap_bf16_box32(&p.tmC, d->out_hi, d->M, d->N, d->ldc);
      if (rc != BLM_OK) return rc;
      return d->act == BLM_ACT_GELU_FAST ? launch_tma16<BLM_ACT_GELU_FAST>(p, st) : launch_tma16<BLM_ACT_GPMIX_FAST>(p, st);
    }
    rc = encode_tmap_bf16(&p.tmC, d->out_hi, d->M, d->N, d->ldc, 32);
    if (rc != BLM_OK) return rc;
    if (BN == 256) {
      switch (d->act) {
        case BLM_ACT_NONE: return launch_tma<256, kStages256, BLM_ACT_NONE>(p, st);
        case BLM_ACT_GELU: return launch_tma<256, kStages256, BLM_ACT_GELU>(p, st);
        case BLM_ACT_GPMIX: return launch_tma<256, kStages256, BLM_ACT_GPMIX>(p, st);
        case BLM_ACT_GPMIX_FAST: return launch_tma<256, kStages256, BLM_ACT_GPMIX_FAST>(p, st);
        default: return launch_tma<256, kStages256, BLM_ACT_GELU_FAST>(p, st);
      }
    }
    switch (d->act) {
      case BLM_ACT_NONE: return launch_tma<128, kStages128, BLM_ACT_NONE>(p, st);
      case BLM_ACT_GELU: return launch_tma<128, kStages128, BLM_ACT_GELU>(p, st);
      case BLM_ACT_GPMIX: return launch_tma<128, kStages128, BLM_ACT_GPMIX>(p, st);
      case BLM_ACT_GPMIX_FAST: return launch_tma<128, kStages128, BLM_ACT_GPMIX_FAST>(p, st);
      default: return launch_tma<128, kStages128, BLM_ACT_GELU_FAST>(p, st);
    }
  }
  const bool gstg = p.use_stg && d->act == BLM_ACT_GELU_GRAD && !d->out_pre;
  const bool stg = p.use_stg && d->act == BLM_ACT_NONE;
  if (chunked) {
    p.chunk_kb = d->k_chunk / kBK;
    // precise-mode triples (A.hi, B.lo), (A.lo, B.hi), (A.hi, B.hi) over one K range: load the four distinct tiles once
    // per K block and issue the three products from them (BLM_GEMM_NO_X3=1 is the A/B switch)
    static const bool no_x3 = getenv("BLM_GEMM_NO_X3") != nullptr;
    bool triples = !no_x3 && d->nseg % 3 == 0;
    for (int s = 0; triples && s < d->nseg; s += 3)
      triples = d->A[s] == d->A[s + 2] && d->B[s + 1] == d->B[s + 2] && d->A[s] != d->A[s + 1] && d->K[s] == d->K[s + 1] &&
                d->K[s] == d->K[s + 2] && d->lda[s] == d->lda[s + 1] && d->lda[s] == d->lda[s + 2] &&
                d->ldb[s] == d->ldb[s + 1] && d->ldb[s] == d->ldb[s + 2];
    p.x3 = triples ? 1 : 0;
    if (stg) return launch_stg<128, kStages128, 1>(p, st);
    if (gstg) return launch_stg<128, kStages128, 1, BLM_ACT_GELU_GRAD>(p, st);
    // forward activations with a hi + lo (or fp32) output: the transposed stores pay here as well -- two bf16 tensors
    // through row-per-thread 16-byte stores were 19 % of the precise FFN1 (813 vs 639 us hi only; 687 us transposed)
    if (p.use_stg && !d->out_pre && d->act == BLM_ACT_GELU) return launch_stg<128, kStages128, 1, BLM_ACT_GELU>(p, st);
    if (p.use_stg && !d->out_pre && d->act == BLM_ACT_GPMIX) return launch_stg<128, kStages128, 1, BLM_ACT_GPMIX>(p, st);
    switch (d->act) {
      case BLM_ACT_NONE: return launch_chunk<BLM_ACT_NONE>(p, st);
      case BLM_ACT_GELU: return launch_chunk<BLM_ACT_GELU>(p, st);
      case BLM_ACT_GPMIX: return launch_chunk<BLM_ACT_GPMIX>(p, st);
      case BLM_ACT_GELU_GRAD: return launch_chunk<BLM_ACT_GELU_GRAD>(p, st);
      case BLM_ACT_GPMIX_GRAD: return launch_chunk<BLM_ACT_GPMIX_GRAD>(p, st);
      default: return launch_chunk<BLM_ACT_SOFTMAX_GRAD>(p, st);
    }
  }
  // Opt-in experiment (BLM_GEMM_ARES=1): keep the A tile resident in shared memory while the CTA sweeps
  // all N tiles of its rows, as the NLL kernel does.  Measured SLOWER for the storing GEMMs (QKV 267 vs
  // 218 us, FFN1 525 vs 432 us at M = 65536, profiles/r01g): the tile loads were not the bound, the
  // row-per-thread stores were (fixed by the store transpose), and a single resident A buffer stalls
  // every work boundary.
  static const bool no_ares = getenv("BLM_GEMM_ARES") == nullptr;
  if (BN == 256 && !no_ares && !d->a_mn && !d->b_mn && d->nseg == 1 && p.kblocks[0] <= kNllAres && p.n_tiles >= 4 && p.m_tiles >= num_sms() &&
      (d->act == BLM_ACT_NONE || d->act == BLM_ACT_GELU || d->act == BLM_ACT_GPMIX)) {
    p.n_groups = 1;
    p.tiles_per_group = p.n_tiles;
    p.num_works = p.m_tiles;
    switch (d->act) {
      case BLM_ACT_NONE: return launch<256, kNllAresStages, EPI_STORE, BLM_ACT_NONE, kNllAres>(p, st);
      case BLM_ACT_GELU: return launch<256, kNllAresStages, EPI_STORE, BLM_ACT_GELU, kNllAres>(p, st);
      default: return launch<256, kNllAresStages, EPI_STORE, BLM_ACT_GPMIX, kNllAres>(p, st);
    }
  }
  if (stg) return BN == 256 ? launch_stg<256, kStages256, 0>(p, st) : launch_stg<128, kStages128, 0>(p, st);
  if (gstg)
    return BN == 256 ? launch_stg<256, kStages256, 0, BLM_ACT_GELU_GRAD>(p, st)
                     : launch_stg<128, kStages128, 0, BLM_ACT_GELU_GRAD>(p, st);
  if (BN == 256) {
    switch (d->act) {
      case BLM_ACT_NONE: return launch<256, kStages256, EPI_STORE, BLM_ACT_NONE>(p, st);
      case BLM_ACT_GELU: return launch<256, kStages256, EPI_STORE, BLM_ACT_GELU>(p, st);
      case BLM_ACT_GELU_FAST: return launch<256, kStages256, EPI_STORE, BLM_ACT_GELU_FAST>(p, st);
      case BLM_ACT_GPMIX: return launch<256, kStages256, EPI_STORE, BLM_ACT_GPMIX>(p, st);
      case BLM_ACT_GELU_GRAD: return launch<256, kStages256, EPI_STORE, BLM_ACT_GELU_GRAD>(p, st);
      case BLM_ACT_GPMIX_GRAD: return launch<256, kStages256, EPI_STORE, BLM_ACT_GPMIX_GRAD>(p, st);
      default: return launch<256, kStages256, EPI_STORE, BLM_ACT_SOFTMAX_GRAD>(p, st);
    }
  }
  switch (d->act) {
    case BLM_ACT_NONE: return launch<128, kStages128, EPI_STORE, BLM_ACT_NONE>(p, st);
    case BLM_ACT_GELU: return launch<128, kStages128, EPI_STORE, BLM_ACT_GELU>(p, st);
    case BLM_ACT_GELU_FAST: return launch<128, kStages128, EPI_STORE, BLM_ACT_GELU_FAST>(p, st);
    case BLM_ACT_GPMIX: return launch<128, kStages128, EPI_STORE, BLM_ACT_GPMIX>(p, st);
    case BLM_ACT_GELU_GRAD: return launch<128, kStages128, EPI_STORE, BLM_ACT_GELU_GRAD>(p, st);
    case BLM_ACT_GPMIX_GRAD: return launch<128, kStages128, EPI_STORE, BLM_ACT_GPMIX_GRAD>(p, st);
    default: return launch<128, kStages128, EPI_STORE, BLM_ACT_SOFTMAX_GRAD>(p, st);
  }
}

}  // namespace blm

extern "C" {

int blm_gemm(const blm_gemm_desc* d, blm_stream stream) { return blm::gemm_impl(d, nullptr, stream); }

// vocabulary groups: (m_tile, group) work items should fill WHOLE waves of the persistent grid.  One group per m
// tile left 7 % of the chip idle at 413 m tiles (2.79 waves) and 13.5 % at 512 (3.46): the group count is the
// smallest one (<= 8, or what it takes to have one work item per SM) whose last wave is at least 97 % full, else
// the best found.  A group costs one (max, sum, target) partial per row and one reload of the resident 128 KB
// hidden tile per work item -- noise next to the 256 KB vocabulary tiles it streams.
// canonical segment of the vocabulary sweep = 4 tiles = 1024 columns (see the EPI_NLL epilogue).  Measured at
// M = 65536: 1447 us with 4-tile segments, 1504 us with 2, 1423 us with BLM_NLL_SEG_TILES=0 (the A/B switch: one segment
// per work group, i.e. the batch-composition-DEPENDENT rounding of the first version).
static int nll_seg_tiles() {
  static const int v = [] {
    const char* e = getenv("BLM_NLL_SEG_TILES");
    return e ? atoi(e) : 4;
  }();
  return v;
}

static int nll_tiles_per_group(int n_tiles, int g) {
  const int tpg = (n_tiles + g - 1) / g;
  const int s = nll_seg_tiles();
  return s > 0 ? (tpg + s - 1) / s * s : tpg;   // groups are whole segments
}

static int nll_groups(int64_t M, int64_t V) {
  const int m_tiles = static_cast<int>((M + blm::kBM - 1) / blm::kBM);
  const int n_tiles = static_cast<int>((V + 255) / 256);
  int sms = blm::num_sms() > 0 ? blm::num_sms() : 148;
  int g_min = (sms + m_tiles - 1) / m_tiles;
  if (g_min > n_tiles) g_min = n_tiles;
  if (g_min < 1) g_min = 1;
  int g_max = g_min > 8 ? g_min : 8;
  if (g_max > n_tiles) g_max = n_tiles;
  static const bool no_balance = getenv("BLM_NLL_NO_BALANCE") != nullptr;   // A/B switch
  if (no_balance) return g_min;
  int best = g_min;
  double best_eff = 0.0;
  for (int g = g_min; g <= g_max; ++g) {
    const int tpg = nll_tiles_per_group(n_tiles, g);
    const int used = (n_tiles + tpg - 1) / tpg;
    const long long works = static_cast<long long>(m_tiles) * used;
    const long long waves = (works + sms - 1) / sms;
    const double eff = static_cast<double>(works) / static_cast<double>(waves * sms);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = g;
    }
    if (eff >= 0.97) break;
  }
  return best;
}

int64_t blm_vocab_nll_workspace_bytes(int64_t M, int64_t V) {
  // per row and epilogue column group (<= 4): one (max, sum) partial per canonical segment of the vocabulary and one
  // target-logit partial per work group
  const int64_t n_tiles = (V + 255) / 256;
  const int64_t segs = nll_seg_tiles() > 0 ? (n_tiles + nll_seg_tiles() - 1) / nll_seg_tiles() : 8;
  const int64_t per_row = 4 * (2 * segs + static_cast<int64_t>(nll_groups(M, V)));
  const int64_t legacy = 3 * 4 * static_cast<int64_t>(nll_groups(M, V));          // (the CTA-pair variant's layout)
  return (per_row > legacy ? per_row : legacy) * M * static_cast<int64_t>(sizeof(float)) + 64;
}

int blm_vocab_nll(const blm_vocab_nll_desc* d, blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(d != nullptr, BLM_ERR_ARG, "null descriptor");
  BLM_REQUIRE(num_sms() > 0, BLM_ERR_ARCH, "blm_init() has not been called");
  BLM_REQUIRE(d->M > 0 && d->V > 0 && d->M < (1ll << 31) && d->V < (1ll << 31), BLM_ERR_SHAPE,
              "bad shape M=%lld V=%lld", (long long)d->M, (long long)d->V);
  BLM_REQUIRE(d->targets && d->nll && d->workspace, BLM_ERR_ARG, "null targets / nll / workspace");
  BLM_REQUIRE(aligned16(d->workspace), BLM_ERR_ALIGN, "workspace must be 16-byte aligned");
  const int groups = nll_groups(d->M, d->V);
  const int col_groups = epi_warps256() / 4;
  BLM_REQUIRE(d->workspace_bytes >= blm_vocab_nll_workspace_bytes(d->M, d->V) - 64, BLM_ERR_ARG,
              "workspace too small: %lld bytes", (long long)d->workspace_bytes);

  GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_segments(p, d->nseg, d->H, d->E, d->K, d->ldh, d->lde, d->M, d->V, 256);
  if (rc != BLM_OK) return rc;
  p.M = static_cast<int>(d->M);
  p.N = static_cast<int>(d->V);
  p.m_tiles = static_cast<int>((d->M + kBM - 1) / kBM);
  p.n_tiles = static_cast<int>((d->V + 255) / 256);
  p.tiles_per_group = nll_tiles_per_group(p.n_tiles, groups);
  const int used_groups = (p.n_tiles + p.tiles_per_group - 1) / p.tiles_per_group;  // no empty groups
  p.n_groups = used_groups;
  p.num_works = p.m_tiles * used_groups;
  p.seg_tiles = nll_seg_tiles() > 0 ? nll_seg_tiles() : p.tiles_per_group;
  p.bias = d->bias;
  p.targets = d->targets;
  float* ws = reinterpret_cast<float*>(d->workspace);
  p.part_max = ws;
  const int n_segs = (p.n_tiles + p.seg_tiles - 1) / p.seg_tiles;
  const int parts = n_segs * col_groups;            // (max, sum) partials: per canonical segment
  const int tgt_parts = used_groups * col_groups;   // target-logit partials: per work group
  p.part_sum = ws + static_cast<int64_t>(parts) * d->M;
  p.part_tgt = ws + 2 * static_cast<int64_t>(parts) * d->M;
  cudaStream_t st = as_stream(stream);
  int total_kb = 0;
  for (int i = 0; i < p.nseg; ++i) total_kb += p.kblocks[i];
  // K <= 512 in one segment (the bf16 Transformer case): keep the hidden-state tile resident in
  // shared memory for the whole vocabulary sweep, so only the embedding tiles stream from L2
  static const bool no_ares = getenv("BLM_NLL_NO_ARES") != nullptr;  // A/B switch for profiling
  if (use_gemm2() && p.nseg == 1 && total_kb <= kNllAres && d->M >= 256 * 8) {
    // CTA-pair kernel: each CTA streams only its half of every vocabulary tile
    GemmParams p2 = p;
    rc = fill_segments(p2, d->nseg, d->H, d->E, d->K, d->ldh, d->lde, d->M, d->V, 128);
    if (rc != BLM_OK) return rc;
    const int m_pairs = static_cast<int>((d->M + 255) / 256);
    int g2 = (num_sms() / 2 + m_pairs - 1) / m_pairs;
    if (g2 > p.n_tiles) g2 = p.n_tiles;
    if (g2 > groups) g2 = groups;   // the workspace was sized for `groups`
    if (g2 < 1) g2 = 1;
    const int tpg = (p.n_tiles + g2 - 1) / g2;
    const int used = (p.n_tiles + tpg - 1) / tpg;
    const int parts2 = used * col_groups;
    p2.part_max = ws;
    p2.part_sum = ws + static_cast<int64_t>(parts2) * d->M;
    p2.part_tgt = ws + 2 * static_cast<int64_t>(parts2) * d->M;
    rc = gemm2_nll(p2, g2, st);
    if (rc != BLM_OK) return rc;
    nll_merge_kernel<<<static_cast<int>((d->M + 255) / 256), 256, 0, st>>>(p2.part_max, p2.part_sum, p2.part_tgt, parts2, parts2,
                                                                         p.M, d->nll, d->lse);
    BLM_CHECK_CUDA(cudaGetLastError());
    return BLM_OK;
  }
  if (total_kb <= kNllAres && !no_ares)
    rc = launch<256, kNllAresStages, EPI_NLL, BLM_ACT_NONE, kNllAres>(p, st);
  else
    rc = launch<256, kStages256, EPI_NLL, BLM_ACT_NONE>(p, st);
  if (rc != BLM_OK) return rc;
  const int threads = 256;
  const int blocks = static_cast<int>((d->M + threads - 1) / threads);
  nll_merge_kernel<<<blocks, threads, 0, st>>>(p.part_max, p.part_sum, p.part_tgt, parts, tgt_parts, p.M,
                                               d->nll, d->lse);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

int blm_segment_sum(const float* x, const int32_t* seg_offsets, int64_t nseg, float* out,
                    blm_stream stream) {
  using namespace blm;
  BLM_REQUIRE(x && seg_offsets && out && nseg > 0, BLM_ERR_ARG, "bad segment_sum arguments");
  const int threads = 256;
  const long long blocks = (nseg * 32 + threads - 1) / threads;
  segment_sum_kernel<<<static_cast<unsigned>(blocks), threads, 0, as_stream(stream)>>>(
      x, seg_offsets, nseg, out);
  BLM_CHECK_CUDA(cudaGetLastError());
  return BLM_OK;
}

}  // extern "C"
