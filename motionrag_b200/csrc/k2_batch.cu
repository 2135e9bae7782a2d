// K2 — batched similarity scan on the 5th-gen tensor cores with top-k fused in the epilogue.
//
// For large query batches the LanceDB per-query flat scan (reference src/data/rag.py:54,
// driven one call per annotation by src/data/datamodule.py:257-262) is a dense
// [nq x dim] x [dim x n_rows] contraction. Here queries sit on the MMA M axis (128 per CTA
// tile), database rows on N (256 per tile), dim is walked in 64-element (128-byte, SWIZZLE_128B)
// k-blocks:
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the Q tile slice and the DB tile
//               slice into a 4-stage shared-memory ring (mbarrier full/empty pairs)
//   warp 1      tcgen05.mma issuer (one elected thread), fp32 accumulators in TMEM,
//               two 128x256 accumulators (all 512 TMEM columns) so MMA of tile t+1 overlaps
//               the epilogue of tile t; tcgen05.commit releases smem stages / publishes tiles
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns at a time; every thread owns one
//               query row, compares its scores against a private running threshold and keeps
//               a sorted top-32 in registers — the score matrix never reaches HBM.
// Work items are (query tile, database chunk) pairs walked persistently, query tile fastest,
// so CTAs that run side by side read the same database tiles out of L2 and HBM sees each
// tile about once. Output: 32 candidate keys per (query, chunk); K3 merges and re-scores in
// fp32 from the master rows.
//
// Algorithmic flops: 2 * nq * n_rows * dim per launch.
#include "k2_common.cuh"

namespace mrag {

constexpr int kStages = 4;
constexpr uint32_t kABytes = kBM * kBK * 2;  // 16 KB
constexpr uint32_t kBBytes = kBN * kBK * 2;  // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;

struct K2Smem {
  // offsets into the 1024-byte aligned dynamic shared memory window
  static constexpr uint32_t kTiles = 0;
  static constexpr uint32_t kEpiStage = kStages * kStageBytes;               // 4 x 32 x 32 floats
  static constexpr uint32_t kBars = kEpiStage + 2 * kEpiWarps * 32 * 32 * 4;
  static constexpr uint32_t kTotal = kBars + 256;
};

template <int SETS, int KC, bool EXTRA>
__global__ void __launch_bounds__(k2_threads(SETS), 1)
    k2_batch_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                    const K2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K2Smem::kBars);
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = clk();
  long long w_a = 0, w_b = 0, busy = 0, served = 0;  // role-specific wait / work counters
  const int kblocks = a.dim / kBK;
  const int unit = blockIdx.x;
  K2Seg sg;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_db);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const uint64_t pol_q = policy_evict_last();    // 6 MB of queries: keep in L2
      const uint64_t pol_db = policy_evict_normal(); // shared by the CTAs of the same chunk
      int stage = 0;
      uint32_t phase = 0;
      for (int si = 0; k2_segment(a, unit, si, sg); ++si) {
        const int m = sg.m;
        for (int t = sg.t0; t < sg.t1; ++t) {
          ++served;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait_timed(&empty_bar[stage], phase ^ 1, w_a);
            uint8_t* sa = smem + K2Smem::kTiles + stage * kStageBytes;
            uint8_t* sb = sa + kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            tma_load_2d(sa, &tm_q, &full_bar[stage], kb * kBK, m * kBM, pol_q);
            tma_load_2d(sb, &tm_db, &full_bar[stage], kb * kBK, t * kBN, pol_db);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int si = 0; k2_segment(a, unit, si, sg); ++si) {
        for (int t = sg.t0; t < sg.t1; ++t) {
          mbar_wait_timed(&tempty_bar[acc], acc_phase ^ 1, w_b);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + uint32_t(acc) * kBN;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait_timed(&full_bar[stage], phase, w_a);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + K2Smem::kTiles + stage * kStageBytes);
            const uint64_t da = umma_desc_k_sw128(sa);
            const uint64_t db = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // +32 bytes per K step inside the 128-byte swizzle row (field is addr >> 4)
              umma_bf16(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc,
                        (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
            if (kb == kblocks - 1) umma_commit(&tfull_bar[acc]);
            if (++stage == kStages) {
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
    // ================= epilogue: fused top-32 per query row =================
    const int quarter = warp & 3;      // TMEM lane quarter this warp may read
    const int set = (warp - 2) >> 2;   // which accumulator buffer (tile parity) this warp serves
    float* stg = reinterpret_cast<float*>(smem + K2Smem::kEpiStage) + (warp - 2) * 32 * 32;
    uint32_t n = 0;                    // running tile count of this CTA, same in every role
    uint32_t ph0 = 0u, ph1 = 0u;
    for (int si = 0; k2_segment(a, unit, si, sg); ++si) {
      const int t0 = sg.t0, t1 = sg.t1;
      const int q_row = sg.m * kBM + quarter * 32 + lane;

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
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
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
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---- host side ------------------------------------------------------------------------------
bool k2_supported(int dim) { return dim % kBK == 0 && dim >= kBK && dim <= 4096; }

K2Plan k2_plan(int64_t n_rows, int nq, int sm_count) {
  K2Plan p{};
  p.m_tiles = (nq + kBM - 1) / kBM;
  p.n_tiles = int((n_rows + kBN - 1) / kBN);
  k2_assign(p, sm_count, p.m_tiles);
  p.grid = sm_count;  // every unit runs: each (query, run) cell is written by exactly one unit
  p.epi_sets = k2_epi_sets();
  return p;
}

cudaError_t launch_k2_batch(const void* q_bf16, int q_rows_padded, const void* db_bf16,
                            int64_t db_rows_padded, int64_t n_rows, int dim, int nq,
                            const K2Plan& plan, uint64_t* cand, uint32_t* gthr, const K2Extra& ex,
                            cudaStream_t st) {
  CUtensorMap tm_q, tm_db;
  if (!make_tmap(&tm_q, q_bf16, q_rows_padded, dim, kBM) ||
      !make_tmap(&tm_db, db_bf16, db_rows_padded, dim, kBN))
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
  const size_t smem = K2Smem::kTotal + 1024;
  cudaError_t e = cudaErrorInvalidValue;
  auto go = [&](auto kern, int threads) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e == cudaSuccess) kern<<<plan.grid, threads, smem, st>>>(tm_q, tm_db, a);
  };
  if (plan.kc == 16) extra ? go(k2_batch_kernel<1, 16, true>, k2_threads(1)) : go(k2_batch_kernel<1, 16, false>, k2_threads(1));
  else extra ? go(k2_batch_kernel<1, 32, true>, k2_threads(1)) : go(k2_batch_kernel<1, 32, false>, k2_threads(1));
  if (e != cudaSuccess) return e;
  note_launch();
  cudaError_t le = cudaGetLastError();
  k2_stats_report(a.stats, plan.grid, st, "single");
  return le;
}

}  // namespace mrag
