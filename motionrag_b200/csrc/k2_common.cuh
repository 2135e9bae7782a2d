// Pieces shared by the single-CTA (k2_batch.cu) and CTA-pair (k2_batch2.cu) tcgen05 scan kernels.
#pragma once
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace mrag {

constexpr int kBM = 128;     // queries per CTA tile (UMMA M per CTA)
constexpr int kBN = 256;     // database rows per tile (UMMA N)
constexpr int kBK = 64;      // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;   // K per tcgen05.mma for 16-bit inputs
constexpr int kK2Threads = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kEpiWarps = 4;
constexpr uint32_t kTmemCols = 512;  // two 128 x 256 fp32 accumulators

struct K2Args {
  int nq;
  int dim;
  int64_t n_rows;
  int m_tiles, n_tiles, chunks, tiles_per_chunk;
  uint64_t* cand;  // [nq][chunks][32]
  // [nq] running lower bound of each query's global 32nd-best score (ordered-uint encoding,
  // zero-initialised), shared by every CTA / chunk working on that query: a CTA's local 32nd
  // best is such a bound, so scores below it can never reach the global top-32 and are
  // dropped before the insertion path. Makes the epilogue's insert count ~ln(N) per query
  // instead of ~chunks * ln(N / chunks).
  uint32_t* gthr;
  int debug;  // profiling only (MRAG_K2_DEBUG): 1 = epilogue skips TMEM reads, 2 = reads but never inserts
};

// running top-32 of one query row, sorted descending, held in registers
struct TopList {
  float ls[kK2Cand];
  int li[kK2Cand];
  float thr;        // max(local 32nd best, global bound): scores must beat it to be inserted
  float gbound;     // last global bound read
  float published;  // last local 32nd best pushed to the global bound
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < kK2Cand; ++i) {
      ls[i] = -INFINITY;
      li[i] = kInvalidIdx;
    }
    thr = -INFINITY;
    gbound = -INFINITY;
    published = -INFINITY;
  }
  // fold in the global bound. A score EQUAL to the bound may still belong to the global
  // top-32 (ties go to the lower row index), so the bound is applied as "strictly below is
  // dropped": nextafter(bound, -inf) turns that into the strict '>' test used everywhere.
  __device__ __forceinline__ void refresh(const uint32_t* gthr_q) {
    const uint32_t o = *reinterpret_cast<const volatile uint32_t*>(gthr_q);
    if (o != 0u) {
      const float g = ordered_to_f32(o - 1u);  // one ulp below the bound in ordered space
      gbound = fmaxf(gbound, g);
      thr = fmaxf(thr, gbound);
    }
  }
  __device__ __forceinline__ void publish(uint32_t* gthr_q) {
    const float mine = ls[kK2Cand - 1];
    if (mine > published) {
      published = mine;
      atomicMax(gthr_q, f32_to_ordered(mine));
    }
  }
  // bubble a new entry down the list; strict '>' keeps the earlier (lower) row ahead on ties
  __device__ __forceinline__ void insert(float cv, int ci) {
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
    thr = fmaxf(ls[kK2Cand - 1], gbound);
  }
  __device__ __forceinline__ void store(uint64_t* dst) const {
#pragma unroll
    for (int i = 0; i < kK2Cand; ++i) dst[i] = make_sim_key(ls[i], li[i]);
  }
};

// One 128 x 256 accumulator tile: this warp's 32 query rows (TMEM lanes) x 256 columns, read
// 32 columns at a time; thread = query row. `stg` is this warp's private [32][32] float scratch.
__device__ __forceinline__ void epilogue_tile(TopList& top, uint32_t t_addr, int64_t row_base,
                                              int64_t n_rows, float* stg, int lane, int debug) {
  if (debug == 1) return;
  if (debug == 2) top.thr = INFINITY;
  const bool ragged = row_base + kBN > n_rows;  // last tile: rows past the table are zero-filled
#pragma unroll 1
  for (int c = 0; c < kBN / 32; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(t_addr + c * 32, v);
    tmem_ld_wait();
    const int col0 = int(row_base) + c * 32;
    if (ragged) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (int64_t(col0) + j >= n_rows) v[j] = 0xff800000u;  // -inf
    }
    uint32_t hits = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) hits |= (__uint_as_float(v[j]) > top.thr) ? (1u << j) : 0u;
    if (__any_sync(0xffffffffu, hits != 0)) {
      // rare path: park the 32 scores (column-major per lane, conflict-free) and let each
      // thread walk its own hit mask
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[j * 32 + lane] = __uint_as_float(v[j]);
      while (hits) {
        const int j = __ffs(hits) - 1;
        hits &= hits - 1;
        const float cv = stg[j * 32 + lane];
        if (cv > top.thr) top.insert(cv, col0 + j);
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int k2_debug_mode() {
  const char* e = getenv("MRAG_K2_DEBUG");
  return e ? atoi(e) : 0;
}

inline EncodeTiledFn get_encode_fn() {
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
inline bool make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int dim, int box_rows) {
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

}  // namespace mrag
