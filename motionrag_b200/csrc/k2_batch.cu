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
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace mrag {

constexpr int kBM = 128;             // queries per tile (UMMA M)
constexpr int kBN = 256;             // database rows per tile (UMMA N)
constexpr int kBK = 64;              // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;           // K per tcgen05.mma for 16-bit inputs
constexpr int kStages = 4;
constexpr int kK2Threads = 192;      // 6 warps
constexpr int kEpiWarps = 4;
constexpr uint32_t kABytes = kBM * kBK * 2;  // 16 KB
constexpr uint32_t kBBytes = kBN * kBK * 2;  // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr uint32_t kTmemCols = 512;

struct K2Smem {
  // offsets into the 1024-byte aligned dynamic shared memory window
  static constexpr uint32_t kTiles = 0;
  static constexpr uint32_t kEpiStage = kStages * kStageBytes;               // 4 x 32 x 32 floats
  static constexpr uint32_t kBars = kEpiStage + kEpiWarps * 32 * 32 * 4;
  static constexpr uint32_t kTotal = kBars + 256;
};

struct K2Args {
  int nq;
  int dim;
  int64_t n_rows;
  int m_tiles, n_tiles, chunks, tiles_per_chunk;
  uint64_t* cand;  // [nq][chunks][32]
};

__global__ void __launch_bounds__(kK2Threads, 1)
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
  const int kblocks = a.dim / kBK;
  const int total_items = a.m_tiles * a.chunks;

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
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int m = item % a.m_tiles, chunk = item / a.m_tiles;
        const int t0 = chunk * a.tiles_per_chunk;
        const int t1 = min(a.n_tiles, t0 + a.tiles_per_chunk);
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
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
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int chunk = item / a.m_tiles;
        const int t0 = chunk * a.tiles_per_chunk;
        const int t1 = min(a.n_tiles, t0 + a.tiles_per_chunk);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + uint32_t(acc) * kBN;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&full_bar[stage], phase);
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
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int ew = warp - 2;
    float* stg = reinterpret_cast<float*>(smem + K2Smem::kEpiStage) + ew * 32 * 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int m = item % a.m_tiles, chunk = item / a.m_tiles;
      const int t0 = chunk * a.tiles_per_chunk;
      const int t1 = min(a.n_tiles, t0 + a.tiles_per_chunk);
      const int q_row = m * kBM + quarter * 32 + lane;

      float ls[kK2Cand];
      int li[kK2Cand];
#pragma unroll
      for (int i = 0; i < kK2Cand; ++i) {
        ls[i] = -INFINITY;
        li[i] = kInvalidIdx;
      }
      float thr = -INFINITY;

      for (int t = t0; t < t1; ++t) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const int64_t row_base = int64_t(t) * kBN;
        const bool ragged = row_base + kBN > a.n_rows;  // last tile: TMA zero-filled rows
        const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc) * kBN;
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_addr + c * 32, v);
          tmem_ld_wait();
          const int col0 = int(row_base) + c * 32;
          if (ragged) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (int64_t(col0) + j >= a.n_rows) v[j] = 0xff800000u;  // -inf
          }
          uint32_t hits = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) hits |= (__uint_as_float(v[j]) > thr) ? (1u << j) : 0u;
          if (__any_sync(0xffffffffu, hits != 0)) {
            // rare path: park the 32 scores (column-major per lane, conflict-free) and let
            // each thread walk its own hit mask
#pragma unroll
            for (int j = 0; j < 32; ++j) stg[j * 32 + lane] = __uint_as_float(v[j]);
            while (hits) {
              const int j = __ffs(hits) - 1;
              hits &= hits - 1;
              float cv = stg[j * 32 + lane];
              if (cv > thr) {
                int ci = col0 + j;
                // bubble the new entry down a descending list; strict '>' keeps the earlier
                // (lower) row index ahead on equal scores
#pragma unroll
                for (int i = 0; i < kK2Cand; ++i) {
                  const bool sw = cv > ls[i];
                  const float ts = ls[i];
                  const int ti = li[i];
                  ls[i] = sw ? cv : ts;
                  li[i] = sw ? ci : ti;
                  cv = sw ? ts : cv;
                  ci = sw ? ti : ci;
                }
                thr = ls[kK2Cand - 1];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }

      if (q_row < a.nq) {
        uint64_t* dst = a.cand + (int64_t(q_row) * a.chunks + chunk) * kK2Cand;
#pragma unroll
        for (int i = 0; i < kK2Cand; ++i) dst[i] = make_sim_key(ls[i], li[i]);
      }
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// [rows][dim] bf16 row-major; box = 64 elements x box_rows rows, 128-byte swizzle
static bool make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int dim, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {cuuint64_t(dim), cuuint64_t(rows)};
  cuuint64_t gstride[1] = {cuuint64_t(dim) * 2};
  cuuint32_t box[2] = {cuuint32_t(kBK), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool k2_supported(int dim) { return dim % kBK == 0 && dim >= kBK && dim <= 4096; }

K2Plan k2_plan(int64_t n_rows, int nq, int sm_count) {
  K2Plan p;
  p.m_tiles = (nq + kBM - 1) / kBM;
  p.n_tiles = int((n_rows + kBN - 1) / kBN);
  // choose the chunk count that minimises the makespan (items per CTA x tiles per item);
  // prefer fewer chunks on ties (fewer candidates, fewer list restarts)
  const int max_chunks = p.n_tiles < 160 ? p.n_tiles : 160;
  int64_t best_cost = -1;
  int best = 1;
  for (int c = 1; c <= max_chunks; ++c) {
    const int tpc = (p.n_tiles + c - 1) / c;
    const int eff_chunks = (p.n_tiles + tpc - 1) / tpc;
    if (eff_chunks != c) continue;
    const int64_t items = int64_t(p.m_tiles) * c;
    const int64_t waves = (items + sm_count - 1) / sm_count;
    const int64_t cost = waves * tpc;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = c;
    }
  }
  p.chunks = best;
  p.tiles_per_chunk = (p.n_tiles + best - 1) / best;
  const int64_t items = int64_t(p.m_tiles) * p.chunks;
  p.grid = int(items < sm_count ? items : sm_count);
  return p;
}

cudaError_t launch_k2_batch(const void* q_bf16, int q_rows_padded, const void* db_bf16,
                            int64_t db_rows_padded, int64_t n_rows, int dim, int nq,
                            const K2Plan& plan, uint64_t* cand, cudaStream_t st) {
  CUtensorMap tm_q, tm_db;
  if (!make_tmap(&tm_q, q_bf16, q_rows_padded, dim, kBM) ||
      !make_tmap(&tm_db, db_bf16, db_rows_padded, dim, kBN))
    return cudaErrorInvalidValue;
  K2Args a;
  a.nq = nq;
  a.dim = dim;
  a.n_rows = n_rows;
  a.m_tiles = plan.m_tiles;
  a.n_tiles = plan.n_tiles;
  a.chunks = plan.chunks;
  a.tiles_per_chunk = plan.tiles_per_chunk;
  a.cand = cand;
  const size_t smem = K2Smem::kTotal + 1024;
  cudaError_t e = cudaFuncSetAttribute(k2_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       int(smem));
  if (e != cudaSuccess) return e;
  k2_batch_kernel<<<plan.grid, kK2Threads, smem, st>>>(tm_q, tm_db, a);
  note_launch();
  return cudaGetLastError();
}

}  // namespace mrag
