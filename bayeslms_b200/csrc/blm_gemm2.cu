// CTA-pair (cta_group::2) form of the tcgen05 GEMM family for sm_100a.
//
// Two CTAs of a 2-CTA cluster (one TPC) work on ONE 256 x 256 output tile: `tcgen05.mma.cta_group::2`
// with M = 256 is issued by the leader CTA only and reads, at the same shared-memory offsets of both
// CTAs, each CTA's own 128 rows of A and each CTA's own HALF (128 of 256 rows) of the B tile; each CTA's
// tensor memory receives its 128 accumulator rows x 256 columns.  Per CTA and K block that is 16 KB of
// A + 16 KB of B instead of 16 + 32 KB: the K = 512 shapes of this model (QKV, FFN1, vocabulary
// projection) are bound by the bytes a CTA can keep in flight from L2, so halving the
// B traffic per SM is what moves them.
//
// Protocol (per CTA unless noted):
//   warp 0  TMA producer: own A rows + own B half, `cp.async.bulk.tensor...cta_group::2`, completing
//           bytes on the LEADER's full barrier (the leader arms expect_tx for both CTAs' bytes)
//   warp 1  (leader only) MMA issuer; `tcgen05.commit...multicast::cluster` releases the ring slot in
//           both CTAs and publishes the accumulator to both CTAs' epilogues
//   warp 2  TMEM allocator (cta_group::2 allocation, both CTAs)
//   warp 4-11  epilogue on the CTA's own 128 rows (same code as the 1-CTA kernel); the peer's epilogue
//           warps release the accumulator stage with a remote mbarrier arrive on the leader's barrier.
// EPI_NLL keeps each CTA's hidden-state tile resident (ARES), as in the 1-CTA kernel.
#include <stdlib.h>
#include <string.h>

#include "blm_gemm_common.cuh"

namespace blm {

constexpr int k2BN = 256;      // N of the pair's tile; each CTA stages k2BN / 2 rows of B
constexpr int k2EW = 8;        // epilogue warps per CTA
constexpr int k2Threads = (4 + k2EW) * 32;
constexpr int k2AB = kBM * kBK * 2;          // 16 KB: one A tile / one B half tile
constexpr int k2Ares = 8;                    // resident A: K <= 512

template <int STAGES, int ARES>
struct Smem2 {
  static constexpr int kResBytes = ARES * k2AB;
  static constexpr int kStageBytes = (ARES ? 0 : k2AB) + k2AB;
  static constexpr int kStgOffset = kResBytes + STAGES * kStageBytes;   // 1024-B aligned: TMA-store staging, 4 KB per warp
  static constexpr int kStgBytes = ARES ? 0 : k2EW * 4096;
  static constexpr int kBiasOffset = kStgOffset + kStgBytes;
  static constexpr int kBarOffset = kBiasOffset + 2 * k2BN * 4;
  // full[STAGES] empty[STAGES] tfull[2] tempty[2] afull aempty + tmem slot
  static constexpr int kBytes = kBarOffset + (2 * STAGES + 6) * 8 + 16;
  static constexpr int kDynBytes = kBytes;
};

// STG == 2: bf16-hi-only output leaves through TMA stores (as in the 1-CTA kernel, blm_gemm.cu): row-per-thread
// 16-byte stores cost 32 LSU wavefronts per instruction and were the bound of the K = 512 shapes.
template <int STAGES, int EPI, int ACT, int ARES, int STG = 0, int EW = k2EW>
__global__ void __launch_bounds__((4 + EW) * 32, 1) gemm2_kernel(const __grid_constant__ GemmParams p) {
  using L = Smem2<STAGES, ARES>;
  static_assert(!STG || (EPI == EPI_STORE && ARES == 0), "TMA-store staging belongs to the storing kernels");
  static_assert(EW == 8 || (EW == 16 && STG == 2), "16 epilogue warps: TMA-store kernels only (2 KB staging per warp)");
  constexpr int kChunks = k2BN / 32 / (EW / 4);  // 4 (EW = 8) or 2 (EW = 16) chunks of 32 columns per epilogue warp
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("blm: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* ring = smem + L::kResBytes;
  float* sbias = reinterpret_cast<float*>(smem + L::kBiasOffset);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);  // used in the LEADER only
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;                                    // used in the LEADER only
  uint64_t* afull_bar = tempty_bar + 2;                                    // LEADER only
  uint64_t* aempty_bar = afull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's arrive.expect_tx; bytes of both CTAs
      mbar_init(&empty_bar[s], 1);  // one multicast commit from the leader's MMA warp
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * EW);  // epilogue warps of both CTAs
    }
    mbar_init(afull_bar, 1);
    mbar_init(aempty_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_kb = p.kblocks[0];
  const uint32_t mask2 = 0x3;

  if (warp == 0) {
    // ------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      for (int w = pair_id; w < p.num_works; w += n_pairs) {
        const int m_tile = w / p.n_groups;  // index of the 256-row pair tile
        const int grp = w - m_tile * p.n_groups;
        const int n0 = grp * p.tiles_per_group;
        const int n1 = min(p.n_tiles, n0 + p.tiles_per_group);
        const int my_m = m_tile * 2 * kBM + static_cast<int>(rank) * kBM;
        if constexpr (ARES > 0) {
          mbar_wait(aempty_bar, a_phase ^ 1u);
          if (leader) mbar_arrive_expect_tx(afull_bar, static_cast<uint32_t>(2 * total_kb * k2AB));
          const uint32_t bar = mapa_shared(smem_u32(afull_bar), 0);
          for (int kb = 0; kb < total_kb; ++kb)
            tma_load_2d_2sm(smem + kb * k2AB, &p.tmA[0], bar, kb * kBK, my_m, kEvictFirst);
          a_phase ^= 1u;
        }
        for (int n = n0; n < n1; ++n) {
          for (int kb = 0; kb < total_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            uint8_t* st = ring + stage * L::kStageBytes;
            if constexpr (ARES == 0) {
              tma_load_2d_2sm(st, &p.tmA[0], bar, kb * kBK, my_m, kEvictNormal);
              st += k2AB;
            }
            tma_load_2d_2sm(st, &p.tmB[0], bar, kb * kBK, n * k2BN + static_cast<int>(rank) * (k2BN / 2), kEvictLast);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // -------------------------------------------------------- MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBM, k2BN);
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = pair_id; w < p.num_works; w += n_pairs) {
        const int m_tile = w / p.n_groups;
        const int grp = w - m_tile * p.n_groups;
        const int n0 = grp * p.tiles_per_group;
        const int n1 = min(p.n_tiles, n0 + p.tiles_per_group);
        if constexpr (ARES > 0) {
          mbar_wait(afull_bar, a_phase);
          a_phase ^= 1u;
        }
        for (int n = n0; n < n1; ++n) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * k2BN);
          for (int kb = 0; kb < total_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            const uint32_t st = smem_u32(ring + stage * L::kStageBytes);
            const uint32_t sa = ARES > 0 ? smem_u32(smem + kb * k2AB) : st;
            const uint32_t sb = ARES > 0 ? st : st + k2AB;
            const uint64_t da = umma_desc_sw128(sa);
            const uint64_t db = umma_desc_sw128(sb);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_ss_2sm(tmem_d, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                               (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage], mask2);  // both CTAs' producers may refill the slot
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit_2sm(&tfull_bar[acc], mask2);  // both CTAs' epilogues
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1u;
          }
        }
        if constexpr (ARES > 0) umma_commit_2sm(aempty_bar, mask2);
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0) {
    // ---------------------------------------------------------- epilogue (own 128 rows)
    const int lane_grp = warp & 3;
    const int col_grp = (warp - kEpiWarp0) >> 2;
    const int etid = threadIdx.x - kEpiWarp0 * 32;
    const int row_in_tile = lane_grp * 32 + lane;
    const int c0 = col_grp * kChunks;
    const uint32_t tempty_leader[2] = {mapa_shared(smem_u32(&tempty_bar[0]), 0), mapa_shared(smem_u32(&tempty_bar[1]), 0)};
    uint8_t* stg = STG ? smem + L::kStgOffset + (warp - kEpiWarp0) * (EW == 16 ? 2048 : 4096) : nullptr;
    (void)stg;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = pair_id; w < p.num_works; w += n_pairs) {
      const int m_tile = w / p.n_groups;
      const int grp = w - m_tile * p.n_groups;
      const int n0 = grp * p.tiles_per_group;
      const int n1 = min(p.n_tiles, n0 + p.tiles_per_group);
      const int m_base = m_tile * 2 * kBM + static_cast<int>(rank) * kBM;
      const int m = m_base + row_in_tile;
      const bool row_ok = m < p.M;
      const bool warp_rows_ok = m_base + lane_grp * 32 < p.M;
      (void)warp_rows_ok;
      NllState st{-INFINITY, 0.0f, -INFINITY, -1};
      if constexpr (EPI == EPI_NLL) {
        if (row_ok) st.tgt = __ldg(p.targets + m);
      }
      for (int n = n0; n < n1; ++n) {
        float breg[1];
        if (p.bias && etid < k2BN) {
          const int col = n * k2BN + etid;
          breg[0] = col < p.N ? __ldg(p.bias + col) : 0.0f;
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tcgen05_fence_after();
        float* sb = sbias + acc * k2BN;
        if (p.bias) {
          if (etid < k2BN) sb[etid] = breg[0];
          epi_bar_sync(EW * 32);
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                               static_cast<uint32_t>(acc * k2BN + c0 * 32);
        if constexpr (EW == 16) {
          // sixteen epilogue warps, one chunk in registers, [32 x 32] TMA stores (see blm_gemm.cu)
          float v[32];
#pragma unroll 1
          for (int c = 0; c < kChunks; ++c) {
            __syncwarp();
            tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
            tmem_ld_wait();
            if (c + 1 == kChunks) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_remote(tempty_leader[acc]);
            }
            const int col0 = n * k2BN + (c0 + c) * 32;
            if (col0 < p.N && warp_rows_ok) {
              store_chunk<ACT, STG>(p, v, m, row_ok, lane, col0, sb + (c0 + c) * 32);
              if (lane == 0) bulk_wait_group_read0();
              __syncwarp();
              stage_chunk_bf16_sw64(v, stg, lane);
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
        float va[32], vb[32];
        __syncwarp();
        tmem_ld_32x32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < kChunks; c += 2) {
          tmem_ld_wait();
          __syncwarp();
          tmem_ld_32x32(taddr + static_cast<uint32_t>((c + 1) * 32), vb);
          {
            const int col0 = n * k2BN + (c0 + c) * 32;
            if (col0 < p.N) {
              if constexpr (EPI == EPI_STORE) {
                if (warp_rows_ok) {
                  store_chunk<ACT, STG>(p, va, m, row_ok, lane, col0, sb + (c0 + c) * 32);
                  if constexpr (STG == 2) {
                    if (lane == 0) bulk_wait_group_read0();  // the previous TMA store has left the staging tile
                    __syncwarp();
                    stage_chunk_bf16(va, stg, lane, 0);
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
            tcgen05_fence_before();
            if (lane == 0) mbar_arrive_remote(tempty_leader[acc]);  // the leader's MMA warp owns the stage
          }
          {
            const int col0 = n * k2BN + (c0 + c + 1) * 32;
            if (col0 < p.N) {
              if constexpr (EPI == EPI_STORE) {
                if (warp_rows_ok) {
                  store_chunk<ACT, STG>(p, vb, m, row_ok, lane, col0, sb + (c0 + c + 1) * 32);
                  if constexpr (STG == 2) stage_chunk_bf16(vb, stg, lane, 1);
                }
              } else {
                nll_chunk(p, vb, col0, st, sb + (c0 + c + 1) * 32);
              }
            }
            if constexpr (EPI == EPI_STORE && STG == 2) {
              if (warp_rows_ok && n * k2BN + (c0 + c) * 32 < p.N) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&p.tmC, stg, n * k2BN + (c0 + c) * 32, m - lane);
                  bulk_commit_group();
                }
              }
            }
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      if constexpr (EPI == EPI_NLL) {
        if (row_ok) {
          const long long o = static_cast<long long>(grp * (EW / 4) + col_grp) * p.M + m;
          p.part_max[o] = st.run_max;
          p.part_sum[o] = st.run_sum;
          p.part_tgt[o] = st.tgt_logit;
        }
      }
    }
  }

  if constexpr (STG == 2) {
    if (warp >= kEpiWarp0 && lane == 0) bulk_wait_group0();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

constexpr int k2Stages = 6;       // 6 x 32 KB (A tile + B half)
constexpr int k2NllStages = 6;    // 128 KB resident A + 6 x 16 KB of B halves

template <int STAGES, int EPI, int ACT, int ARES, int STG = 0, int EW = k2EW>
static int set_attr2() {
  BLM_CHECK_CUDA(cudaFuncSetAttribute(gemm2_kernel<STAGES, EPI, ACT, ARES, STG, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Smem2<STAGES, ARES>::kDynBytes));
  return BLM_OK;
}

int gemm2_init() {
  int rc;
  if ((rc = set_attr2<k2Stages, EPI_STORE, BLM_ACT_NONE, 0>()) != BLM_OK) return rc;
  if ((rc = set_attr2<k2Stages, EPI_STORE, BLM_ACT_GELU, 0>()) != BLM_OK) return rc;
  if ((rc = set_attr2<k2Stages, EPI_STORE, BLM_ACT_GELU_FAST, 0>()) != BLM_OK) return rc;
  if ((rc = set_attr2<k2Stages, EPI_STORE, BLM_ACT_NONE, 0, 2>()) != BLM_OK) return rc;
  if ((rc = set_attr2<k2Stages, EPI_STORE, BLM_ACT_GELU, 0, 2>()) != BLM_OK) return rc;
  if ((rc = set_attr2<k2Stages, EPI_STORE, BLM_ACT_GELU_FAST, 0, 2>()) != BLM_OK) return rc;
  if ((rc = set_attr2<k2Stages, EPI_STORE, BLM_ACT_GELU_FAST, 0, 2, 16>()) != BLM_OK) return rc;
  if ((rc = set_attr2<k2NllStages, EPI_NLL, BLM_ACT_NONE, k2Ares>()) != BLM_OK) return rc;
  return BLM_OK;
}

template <int STAGES, int EPI, int ACT, int ARES, int STG = 0, int EW = k2EW>
static int launch2(const GemmParams& p, cudaStream_t st) {
  int pairs = num_sms() / 2;
  if (pairs > p.num_works) pairs = p.num_works;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(2 * pairs));
  cfg.blockDim = dim3((4 + EW) * 32);
  cfg.dynamicSmemBytes = Smem2<STAGES, ARES>::kDynBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BLM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm2_kernel<STAGES, EPI, ACT, ARES, STG, EW>, p));
  return BLM_OK;
}

// Pair-tile launch of a storing GEMM (called by blm_gemm when the shape qualifies): p is filled for 128-row
// tiles by the caller; only the tile bookkeeping changes here.
// tma_store: 0 = direct stores, 1 = [32 x 64] TMA stores on 8 epilogue warps (p.tmC box 64 x 32),
//            2 = GELU_FAST on 16 epilogue warps with [32 x 32] TMA stores (p.tmC box 32 x 32, 64B swizzle)
int gemm2_store(GemmParams p, int act, cudaStream_t st, int tma_store) {
  p.m_tiles = (p.M + 2 * kBM - 1) / (2 * kBM);
  p.n_tiles = (p.N + k2BN - 1) / k2BN;
  p.n_groups = p.n_tiles;
  p.tiles_per_group = 1;
  p.num_works = p.m_tiles * p.n_tiles;
  if (tma_store == 2) return launch2<k2Stages, EPI_STORE, BLM_ACT_GELU_FAST, 0, 2, 16>(p, st);
  if (tma_store) {   // p.tmC: bf16 [M, N], box 32 rows x 64 columns (encoded by the caller)
    switch (act) {
      case BLM_ACT_NONE: return launch2<k2Stages, EPI_STORE, BLM_ACT_NONE, 0, 2>(p, st);
      case BLM_ACT_GELU: return launch2<k2Stages, EPI_STORE, BLM_ACT_GELU, 0, 2>(p, st);
      default: return launch2<k2Stages, EPI_STORE, BLM_ACT_GELU_FAST, 0, 2>(p, st);
    }
  }
  switch (act) {
    case BLM_ACT_NONE: return launch2<k2Stages, EPI_STORE, BLM_ACT_NONE, 0>(p, st);
    case BLM_ACT_GELU: return launch2<k2Stages, EPI_STORE, BLM_ACT_GELU, 0>(p, st);
    default: return launch2<k2Stages, EPI_STORE, BLM_ACT_GELU_FAST, 0>(p, st);
  }
}

int gemm2_nll(GemmParams p, int groups, cudaStream_t st) {
  p.m_tiles = (p.M + 2 * kBM - 1) / (2 * kBM);
  p.n_tiles = (p.N + k2BN - 1) / k2BN;
  p.tiles_per_group = (p.n_tiles + groups - 1) / groups;
  p.n_groups = (p.n_tiles + p.tiles_per_group - 1) / p.tiles_per_group;
  p.num_works = p.m_tiles * p.n_groups;
  return launch2<k2NllStages, EPI_NLL, BLM_ACT_NONE, k2Ares>(p, st);
}

}  // namespace blm
