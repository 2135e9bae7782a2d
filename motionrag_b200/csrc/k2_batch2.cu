// K2 (CTA-pair form) — the batched tcgen05 scan with cta_group::2: two CTAs of a cluster (one
// TPC) run ONE M=256 x N=256 x K=16 MMA stream. Each CTA stages its own 128 query rows (A half)
// and 128 of the 256 database rows of the tile (B half), so per k-block an SM pulls 32 KB from
// L2 instead of 48 KB and the ring is 6 stages deep instead of 4 — the single-CTA kernel is
// limited by L2->SM traffic and TMA latency, not by the tensor pipe. Roles per CTA as in
// k2_batch.cu; differences:
//   * TMA loads of both CTAs credit the LEADER's full barrier (tma_load_2d_pair);
//   * only the leader's warp 1 issues tcgen05.mma.cta_group::2; its commits are multicast to
//     the empty / tmem-full barriers of BOTH CTAs;
//   * the peer's epilogue warps release accumulators by arriving on the leader's tmem-empty
//     barrier through the cluster (mapa + mbarrier.arrive.shared::cluster);
//   * each CTA's TMEM holds the 128 x 256 fp32 block of its own query rows, so the fused
//     top-32 epilogue is unchanged.
// Work item = (pair of query tiles, database chunk).
#include "k2_common.cuh"

namespace mrag {

constexpr int kStages2 = 6;
constexpr uint32_t kABytes2 = kBM * kBK * 2;        // 16 KB: this CTA's 128 query rows
constexpr uint32_t kBBytes2 = (kBN / 2) * kBK * 2;  // 16 KB: this CTA's half of the DB tile
constexpr uint32_t kStageBytes2 = kABytes2 + kBBytes2;

struct K2Smem2 {
  static constexpr uint32_t kTiles = 0;
  static constexpr uint32_t kEpiStage = kStages2 * kStageBytes2;
  static constexpr uint32_t kBars = kEpiStage + 2 * kEpiWarps * 32 * 32 * 4;
  static constexpr uint32_t kTotal = kBars + 256;
};

template <int SETS, int KC, bool EXTRA>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2_threads(SETS), 1)
    k2_batch2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                     const K2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K2Smem2::kBars);
  uint64_t* full_bar = bars;                       // [kStages2]  (used in the leader only)
  uint64_t* empty_bar = bars + kStages2;           // [kStages2]  (each CTA its own)
  uint64_t* tfull_bar = bars + 2 * kStages2;       // [2]         (each CTA its own)
  uint64_t* tempty_bar = bars + 2 * kStages2 + 2;  // [2]         (used in the leader only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages2 + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = clk();
  long long w_a = 0, w_b = 0, busy = 0, served = 0;  // role-specific wait / work counters
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int pair_id = blockIdx.x >> 1;
  const int kblocks = a.dim / kBK;
  const int unit = pair_id;
  K2Seg sg;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_db);
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full_bar[s], 1);   // leader producer's arrive.expect_tx (+ tx bytes of both CTAs)
      mbar_init(&empty_bar[s], 1);  // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * kEpiWarps);  // epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      const uint64_t pol_q = policy_evict_last();
      const uint64_t pol_db = policy_evict_normal();
      int stage = 0;
      uint32_t phase = 0;
      for (int si = 0; k2_segment(a, unit, si, sg); ++si) {
        const int q_row0 = (sg.m * 2 + int(cta_rank)) * kBM;
        for (int t = sg.t0; t < sg.t1; ++t) {
          ++served;
          const int db_row0 = t * kBN + int(cta_rank) * (kBN / 2);
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait_timed(&empty_bar[stage], phase ^ 1, w_a);
            uint8_t* sa = smem + K2Smem2::kTiles + stage * kStageBytes2;
            uint8_t* sb = sa + kABytes2;
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes2);
            tma_load_2d_pair(sa, &tm_q, &full_bar[stage], kb * kBK, q_row0, pol_q);
            tma_load_2d_pair(sb, &tm_db, &full_bar[stage], kb * kBK, db_row0, pol_db);
            if (++stage == kStages2) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA, one thread) =================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int si = 0; k2_segment(a, unit, si, sg); ++si) {
        for (int t = sg.t0; t < sg.t1; ++t) {
          mbar_wait_timed(&tempty_bar[acc], acc_phase ^ 1, w_b);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + uint32_t(acc) * kBN;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait_timed(&full_bar[stage], phase, w_a);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + K2Smem2::kTiles + stage * kStageBytes2);
            const uint64_t da = umma_desc_k_sw128(sa);
            const uint64_t db = umma_desc_k_sw128(sa + kABytes2);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k)
              umma_bf16_pair(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc,
                             (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
            if (kb == kblocks - 1) umma_commit_pair(&tfull_bar[acc]);
            if (++stage == kStages2) {
              stage = 0;
              phase ^= 1;
            }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // ================= epilogue (both CTAs, own 128 query rows) =================
    const int quarter = warp & 3;      // TMEM lane quarter this warp may read
    const int set = (warp - 2) >> 2;   // which accumulator buffer (tile parity) this warp serves
    float* stg = reinterpret_cast<float*>(smem + K2Smem2::kEpiStage) + (warp - 2) * 32 * 32;
    uint32_t n = 0;                    // running tile count of this CTA, same in every role
    uint32_t ph0 = 0u, ph1 = 0u;
    for (int si = 0; k2_segment(a, unit, si, sg); ++si) {
      const int t0 = sg.t0, t1 = sg.t1;
      const int q_row = (sg.m * 2 + int(cta_rank)) * kBM + quarter * 32 + lane;

      TopList<KC> top;
      top.reset();
      const bool live = q_row < a.nq;  // padding rows of the last query tile keep no state
      uint32_t* gthr_q = a.gthr + (live ? q_row : 0);
      EpiExtra xe{nullptr, nullptr, -1};
      if constexpr (EXTRA) {
        xe.bias = a.ex.row_bias;
        xe.groups = a.ex.row_group;
        if (live && a.ex.row_group != nullptr && a.ex.exclude_group != nullptr) xe.excl = a.ex.exclude_group[q_row];
      }
      for (int t = t0; t < t1; ++t, ++n) {
        const int acc = int(n & 1u);
        if (SETS == 2 && acc != set) continue;
        if (live) top.refresh(gthr_q);
        mbar_wait_timed(&tfull_bar[acc], acc ? ph1 : ph0, w_a);
        tc_fence_after();
        const long long t_busy = clk();
        const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc) * kBN;
        epilogue_tile<SETS == 1, KC, EXTRA>(top, t_addr, int64_t(t) * kBN, a.n_rows, stg, lane, a.debug, xe);
        if (live) top.publish(gthr_q);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);
        if (acc) ph1 ^= 1; else ph0 ^= 1;
        busy += clk() - t_busy;
        ++served;
      }
      if (live) top.store(a.cand + ((int64_t(q_row) * a.runs + sg.run) * SETS + set) * KC);
    }
  }

  if (a.stats != nullptr && lane == 0) {
    unsigned long long* st = a.stats + size_t(blockIdx.x) * 8;
    if (warp == 0) {
      st[6] = (unsigned long long)served;
      st[0] = (unsigned long long)(clk() - t_start);
      st[1] = (unsigned long long)w_a;
    } else if (warp == 1) {
      st[2] = (unsigned long long)w_a;
      st[3] = (unsigned long long)w_b;
    } else if (warp == 2) {
      st[4] = (unsigned long long)w_a;
      st[5] = (unsigned long long)busy;
      st[7] = (unsigned long long)served;
    }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the pair's MMAs / remote arrives may still land
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
}

// ---- host side ------------------------------------------------------------------------------
K2Plan k2_plan_pair(int64_t n_rows, int nq, int sm_count) {
  K2Plan p{};
  p.m_tiles = (nq + kBM - 1) / kBM;
  p.n_tiles = int((n_rows + kBN - 1) / kBN);
  const int pairs = sm_count / 2;
  k2_assign(p, pairs, (p.m_tiles + 1) / 2);
  p.grid = 2 * pairs;
  p.epi_sets = k2_epi_sets();
  return p;
}

cudaError_t launch_k2_batch_pair(const void* q_bf16, int q_rows_padded, const void* db_bf16,
                                 int64_t db_rows_padded, int64_t n_rows, int dim, int nq,
                                 const K2Plan& plan, uint64_t* cand, uint32_t* gthr, const K2Extra& ex,
                            cudaStream_t st) {
  CUtensorMap tm_q, tm_db;
  if (!make_tmap(&tm_q, q_bf16, q_rows_padded, dim, kBM) ||
      !make_tmap(&tm_db, db_bf16, db_rows_padded, dim, kBN / 2))
    return cudaErrorInvalidValue;
  K2Args a;
  a.nq = nq;
  a.dim = dim;
  a.n_rows = n_rows;
  k2_fill_args(a, plan);
  a.cand = cand;
  a.gthr = gthr;
  a.ex = ex;
  const bool extra = ex.row_bias != nullptr || (ex.row_group != nullptr && ex.exclude_group != nullptr);
  a.debug = k2_debug_mode();
  a.stats = k2_stats_alloc(plan.grid);
  const size_t smem = K2Smem2::kTotal + 1024;
  cudaError_t e = cudaErrorInvalidValue;
  auto go = [&](auto kern, int threads) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e == cudaSuccess) kern<<<plan.grid, threads, smem, st>>>(tm_q, tm_db, a);
  };
  if (plan.kc == 16) extra ? go(k2_batch2_kernel<1, 16, true>, k2_threads(1)) : go(k2_batch2_kernel<1, 16, false>, k2_threads(1));
  else extra ? go(k2_batch2_kernel<1, 32, true>, k2_threads(1)) : go(k2_batch2_kernel<1, 32, false>, k2_threads(1));
  if (e != cudaSuccess) return e;
  note_launch();
  cudaError_t le = cudaGetLastError();
  k2_stats_report(a.stats, plan.grid, st, "pair");
  return le;
}

}  // namespace mrag
