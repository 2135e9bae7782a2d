// Pieces shared by the single-CTA (k2_batch.cu) and CTA-pair (k2_batch2.cu) tcgen05 scan kernels.
#pragma once
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace mrag {

constexpr int kBM = 128;     // queries per CTA tile (UMMA M per CTA)
constexpr int kBN = 256;     // database rows per tile (UMMA N)
constexpr int kBK = 64;      // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;   // K per tcgen05.mma for 16-bit inputs
// warp 0 TMA, warp 1 MMA, then SETS x 4 epilogue warps (one warp per TMEM lane quarter). With
// SETS = 2 the sets alternate tiles (set s serves accumulator s) and keep their own top lists.
constexpr int kEpiWarps = 4;
constexpr int k2_threads(int sets) { return 64 + 128 * sets; }
constexpr uint32_t kTmemCols = 512;  // two 128 x 256 fp32 accumulators

struct K2Args {
  int nq;
  int dim;
  int64_t n_rows;
  int m_tiles, n_tiles;
  // work assignment (see k2_segment): m_groups query groups (tiles, or tile pairs for the CTA-
  // pair kernel) x n_tiles database tiles over n_units units (CTAs or CTA pairs)
  int m_groups, n_units, fixed_per_group, quota, lr_t0, bt, nb;
  int runs;        // candidate runs per query = fixed_per_group + nb
  uint64_t* cand;  // [nq][runs][SETS][KC]
  // [nq] running lower bound of each query's global 32nd-best score (ordered-uint encoding,
  // zero-initialised), shared by every CTA / chunk working on that query: a CTA's local KC-th
  // best is such a bound, so scores below it can never reach the global top-KC and are
  // dropped before the insertion path. Makes the epilogue's insert count ~ln(N) per query
  // instead of ~chunks * ln(N / chunks).
  uint32_t* gthr;
  // profiling only (MRAG_K2_STATS=1): per CTA {total, producer wait-empty, mma wait-full,
  // mma wait-tmem-empty, epilogue wait-tmem-full, epilogue busy, tiles} in SM cycles (epilogue
  // numbers are from the first epilogue warp)
  unsigned long long* stats;
  int debug;  // profiling only (MRAG_K2_DEBUG): 1 = epilogue skips TMEM reads, 2 = reads but never inserts
  K2Extra ex;  // row bias / pre-filter (kernels instantiated with EXTRA only)
};

// Work assignment. "Fixed" units keep ONE query group for their whole life and walk a contiguous
// tile range [j*quota, (j+1)*quota): their top lists are never reset, so the insert count per
// query is ~ln(rows of the range) and the shared bound tightens quickly; the m_groups units with
// the same j walk the same database tiles side by side, so HBM sees each tile once. The units
// left over when n_units is not a multiple of m_groups ("floaters") take the remaining tile range
// [lr_t0, n_tiles) in blocks of bt tiles, query group fastest (they restart their lists per block
// but start from the tight shared bound). With fewer units than groups everything is a floater.
struct K2Seg {
  int m, t0, t1, run;
};
__device__ __forceinline__ bool k2_segment(const K2Args& a, int unit, int si, K2Seg& s) {
  const int n_fixed = a.fixed_per_group * a.m_groups;
  if (unit < n_fixed) {
    if (si != 0) return false;
    const int j = unit / a.m_groups;
    s.m = unit - j * a.m_groups;
    s.t0 = min(a.lr_t0, j * a.quota);
    s.t1 = min(a.lr_t0, s.t0 + a.quota);
    s.run = j;
    return true;
  }
  const int e = a.n_units - n_fixed;
  const int id = (unit - n_fixed) + si * e;
  if (id >= a.nb * a.m_groups) return false;
  const int b = id / a.m_groups;
  s.m = id - b * a.m_groups;
  s.t0 = a.lr_t0 + b * a.bt;
  s.t1 = min(a.n_tiles, s.t0 + a.bt);
  s.run = a.fixed_per_group + b;
  return true;
}

// host: fill the assignment fields of a plan for `units` units and `groups` query groups
inline void k2_assign(K2Plan& p, int units, int groups) {
  const int T = p.n_tiles;
  int f = units / groups;
  if (f > T) f = T;
  p.m_groups = groups;
  p.n_units = units;
  p.fixed_per_group = f;
  p.quota = f > 0 ? int((int64_t(groups) * T + units - 1) / units) : 0;
  if (f > 0 && int64_t(f) * p.quota > T) p.quota = (T + f - 1) / f;  // never hand out more than exists
  p.lr_t0 = f > 0 ? (int64_t(f) * p.quota < T ? f * p.quota : T) : 0;
  const int lr = T - p.lr_t0;
  const int e = units - f * groups;
  p.bt = 1;
  p.nb = 0;
  if (lr > 0 && e > 0) {
    // floater makespan for block size bt: ceil(blocks * groups / e) * bt. Take the LARGEST block
    // (fewest runs / list restarts) whose makespan stays within 1.5 % of the better of the fixed
    // units' quota and the best achievable.
    int64_t best = -1;
    for (int bt = 4; bt <= 128; ++bt) {
      const int nb = (lr + bt - 1) / bt;
      if (nb > 200) continue;
      const int64_t cost = ((int64_t(nb) * groups + e - 1) / e) * bt;
      if (best < 0 || cost < best) best = cost;
    }
    if (best >= 0) {
      int64_t limit = best > p.quota ? best : p.quota;
      limit += limit * 15 / 1000;
      for (int bt = 4; bt <= 128; ++bt) {
        const int nb = (lr + bt - 1) / bt;
        if (nb > 200) continue;
        const int64_t cost = ((int64_t(nb) * groups + e - 1) / e) * bt;
        if (cost <= limit) {
          p.bt = bt;
          p.nb = nb;
        }
      }
    }
    if (best < 0) {  // very long leftover: cap the block count instead
      p.nb = 200;
      p.bt = (lr + 199) / 200;
    }
  }
  p.chunks = p.fixed_per_group + p.nb;  // runs per query (per epilogue set)
  p.tiles_per_chunk = p.quota > p.bt ? p.quota : p.bt;
}
inline void k2_fill_args(K2Args& a, const K2Plan& p) {
  a.m_tiles = p.m_tiles;
  a.n_tiles = p.n_tiles;
  a.m_groups = p.m_groups;
  a.n_units = p.n_units;
  a.fixed_per_group = p.fixed_per_group;
  a.quota = p.quota;
  a.lr_t0 = p.lr_t0;
  a.bt = p.bt;
  a.nb = p.nb;
  a.runs = p.chunks;
}

// running top-KC of one query row, sorted descending, held in registers
template <int KC>
struct TopList {
  float ls[KC];
  int li[KC];
  float thr;        // max(local 32nd best, global bound): scores must beat it to be inserted
  float gbound;     // last global bound read
  float published;  // last local 32nd best pushed to the global bound
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < KC; ++i) {
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
    const float mine = ls[KC - 1];
    if (mine > published) {
      published = mine;
      atomicMax(gthr_q, f32_to_ordered(mine));
    }
  }
  // bubble a new entry down the list; strict '>' keeps the earlier (lower) row ahead on ties
  __device__ __forceinline__ void insert(float cv, int ci) {
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      const bool sw = cv > ls[i];
      const float ts = ls[i];
      const int ti = li[i];
      ls[i] = sw ? cv : ts;
      li[i] = sw ? ci : ti;
      cv = sw ? ts : cv;
      ci = sw ? ti : ci;
    }
    thr = fmaxf(ls[KC - 1], gbound);
  }
  __device__ __forceinline__ void store(uint64_t* dst) const {
#pragma unroll
    for (int i = 0; i < KC; ++i) dst[i] = make_sim_key(ls[i], li[i]);
  }
};

// 32 scores of one query row (32 consecutive database rows starting at col0).
// Fast path: one max-tree + one compare rejects the whole group (the common case once the
// threshold is warm). Slow path, only for lanes that have a candidate: park the scores in this
// warp's scratch (column-major per lane, conflict-free) and walk the hit mask.
// EXTRA: a per-row bias (uniform 128-bit loads: every lane reads the same 32 floats) is added to
// the scores before selection, and rows of the query's excluded group (`excl` >= 0) are dropped on
// the insertion path (pre-filter) — both off the fast reject path's critical dependencies.
struct EpiExtra {
  const float* bias;
  const int32_t* groups;
  int excl;
};
template <int KC, bool EXTRA>
__device__ __forceinline__ void process_group(TopList<KC>& top, const uint32_t (&v)[32], int col0,
                                              bool ragged, int64_t n_rows, float* stg, int lane,
                                              const EpiExtra& xe) {
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  if constexpr (EXTRA) {
    if (xe.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(xe.bias + col0);  // col0 % 32 == 0
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);
        f[4 * j + 0] += b.x;
        f[4 * j + 1] += b.y;
        f[4 * j + 2] += b.z;
        f[4 * j + 3] += b.w;
      }
    }
  }
  if (ragged) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (int64_t(col0) + j >= n_rows) f[j] = -INFINITY;  // rows past the table
  }
  float m[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) m[j] = fmaxf(f[j], f[j + 16]);
#pragma unroll
  for (int w = 8; w > 0; w >>= 1)
#pragma unroll
    for (int j = 0; j < w; ++j) m[j] = fmaxf(m[j], m[j + w]);
  if (m[0] > top.thr) {
    uint32_t hits = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      stg[j * 32 + lane] = f[j];
      hits |= (f[j] > top.thr) ? (1u << j) : 0u;
    }
    while (hits) {
      const int j = __ffs(hits) - 1;
      hits &= hits - 1;
      const float cv = stg[j * 32 + lane];
      if constexpr (EXTRA) {
        if (xe.excl >= 0 && __ldg(xe.groups + col0 + j) == xe.excl) continue;  // pre-filtered row
      }
      if (cv > top.thr) top.insert(cv, col0 + j);
    }
  }
}

// One 128 x 256 accumulator tile: this warp's 32 query rows (TMEM lanes) x 256 columns, read
// 32 columns at a time with the next tcgen05.ld in flight while the current group is scanned;
// thread = query row. `stg` is this warp's private [32][32] float scratch.
template <bool DOUBLE_BUFFER, int KC, bool EXTRA>
__device__ __forceinline__ void epilogue_tile(TopList<KC>& top, uint32_t t_addr, int64_t row_base,
                                              int64_t n_rows, float* stg, int lane, int debug,
                                              const EpiExtra& xe) {
  if (debug == 1) return;
  if (debug == 2) top.thr = INFINITY;
  const bool ragged = row_base + kBN > n_rows;  // last tile: rows past the table are zero-filled
  if constexpr (!DOUBLE_BUFFER) {  // two warps per scheduler already hide the TMEM latency
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(t_addr + c * 32, v);
      tmem_ld_wait();
      process_group<KC, EXTRA>(top, v, int(row_base) + c * 32, ragged, n_rows, stg, lane, xe);
    }
    return;
  }
  uint32_t va[32], vb[32];
  tmem_ld_32x32(t_addr, va);
#pragma unroll 1
  for (int c = 0; c < kBN / 32; c += 2) {
    tmem_ld_wait();
    tmem_ld_32x32(t_addr + (c + 1) * 32, vb);
    process_group<KC, EXTRA>(top, va, int(row_base) + c * 32, ragged, n_rows, stg, lane, xe);
    tmem_ld_wait();
    if (c + 2 < kBN / 32) tmem_ld_32x32(t_addr + (c + 2) * 32, va);
    process_group<KC, EXTRA>(top, vb, int(row_base) + (c + 1) * 32, ragged, n_rows, stg, lane, xe);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int k2_epi_sets() { return 1; }  // (a second epilogue warp set measured no faster; removed)

__device__ __forceinline__ long long clk() { return clock64(); }
// mbar_wait that charges the waited cycles to `acc`
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, long long& acc) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clk();
  while (!mbar_try_wait(bar, parity)) {
  }
  acc += clk() - t0;
}

// host: allocate / print the stats buffer when MRAG_K2_STATS=1
inline unsigned long long* k2_stats_alloc(int n_ctas) {
  static const bool on = [] { const char* e = getenv("MRAG_K2_STATS"); return e && e[0] == '1'; }();
  if (!on) return nullptr;
  unsigned long long* p = nullptr;
  if (cudaMalloc(&p, size_t(n_ctas) * 8 * 8) != cudaSuccess) return nullptr;
  cudaMemset(p, 0, size_t(n_ctas) * 8 * 8);
  return p;
}
inline void k2_stats_report(unsigned long long* dev, int n_ctas, cudaStream_t st, const char* what) {
  if (!dev) return;
  cudaStreamSynchronize(st);
  unsigned long long* h = new unsigned long long[size_t(n_ctas) * 8];
  cudaMemcpy(h, dev, size_t(n_ctas) * 64, cudaMemcpyDeviceToHost);
  double s[8] = {0};
  for (int c = 0; c < n_ctas; ++c)
    for (int j = 0; j < 8; ++j) s[j] += double(h[c * 8 + j]);
  for (int j = 0; j < 8; ++j) s[j] /= n_ctas;
  const double tiles = s[6] > 0 ? s[6] : 1;
  fprintf(stderr,
          "[k2 stats %s] per CTA: total %.0f cyc, tiles %.0f (%.0f cyc/tile) | producer wait-empty %.1f%% | "
          "mma wait-full %.1f%% wait-tmem-empty %.1f%% | epilogue(w0) wait-tmem-full %.1f%% busy %.1f%% "
          "(%.0f cyc/tile served)\n",
          what, s[0], tiles, s[0] / tiles, 100 * s[1] / s[0], 100 * s[2] / s[0], 100 * s[3] / s[0],
          100 * s[4] / s[0], 100 * s[5] / s[0], s[7] > 0 ? s[5] / s[7] : 0.0);
  delete[] h;
  cudaFree(dev);
}

inline int k2_debug_mode() {
  static const int mode = [] { const char* e = getenv("MRAG_K2_DEBUG"); return e ? atoi(e) : 0; }();
  return mode;
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
