// K1 — HBM-streaming similarity scan with fused top-k (few queries, nq <= 4).
//
// Replaces the LanceDB flat scan behind table.search(vec, 'text_embedding').limit(k)
// (reference src/data/rag.py:54): every database row is read exactly once with 128-bit
// coalesced, L1-bypassing loads; one warp owns a row at a time (lane l holds the 16-byte
// slices l, l+32, ...), dots are finished with a 5-step xor-shuffle, and each warp keeps a
// running top-KC in registers (lane i = slot i) guarded by a warp-uniform threshold so the
// score vector never leaves the SM. Per-CTA lists are merged (warp shuffle sort + binary-search
// ranks) and written as u64 keys (score descending, index ascending). For a single query the last
// CTA to finish then runs the K3 body itself (select, fp32 re-score, filter, cross-GPU exchange),
// so the whole search is one launch; otherwise K3 follows as its own kernel.
//
// Algorithmic bytes: n_rows * dim * sizeof(T) per launch (T = float or bf16).
#include "common.cuh"
#include "k3_body.cuh"
#include "kernels.h"

namespace mrag {

// CTA shapes: 256 threads, 2-3 CTAs per SM for the multi-query variants; the single-query variants run
// ONE 512-thread CTA per SM (same 16 warps per SM, half as many candidate runs, and a fused tail with 16
// warps to spread the select / re-rank work over)

template <typename T>
struct Elt;
template <>
struct Elt<float> {
  static constexpr int kPerVec = 4;
};
template <>
struct Elt<__nv_bfloat16> {
  static constexpr int kPerVec = 8;
};

// partial dot of one 16-byte database slice against the matching query slice(s)
__device__ __forceinline__ float dot_slice_f32(const uint4& v, const float4& q, float acc) {
  acc = fmaf(__uint_as_float(v.x), q.x, acc);
  acc = fmaf(__uint_as_float(v.y), q.y, acc);
  acc = fmaf(__uint_as_float(v.z), q.z, acc);
  acc = fmaf(__uint_as_float(v.w), q.w, acc);
  return acc;
}
__device__ __forceinline__ float dot_slice_bf16(const uint4& v, const float4& qa, const float4& qb,
                                                float acc) {
  acc = fmaf(bf16lo_to_f32(v.x), qa.x, acc);
  acc = fmaf(bf16hi_to_f32(v.x), qa.y, acc);
  acc = fmaf(bf16lo_to_f32(v.y), qa.z, acc);
  acc = fmaf(bf16hi_to_f32(v.y), qa.w, acc);
  acc = fmaf(bf16lo_to_f32(v.z), qb.x, acc);
  acc = fmaf(bf16hi_to_f32(v.z), qb.y, acc);
  acc = fmaf(bf16lo_to_f32(v.w), qb.z, acc);
  acc = fmaf(bf16hi_to_f32(v.w), qb.w, acc);
  return acc;
}

// "a is a worse list entry than b": lower score, then higher index, then higher lane
__device__ __forceinline__ bool worse(float sa, int ia, int la, float sb, int ib, int lb) {
  if (sa != sb) return sa < sb;
  if (ia != ib) return ia > ib;
  return la > lb;
}

// shared memory of the fused tail exists only in the single-query instantiations
template <int Q>
struct K1TailSmem {};
template <>
struct K1TailSmem<1> {
  K3Smem k3;
};

template <typename T, int D, int Q, int R, int MINB, int NT>
__global__ void __launch_bounds__(NT, MINB)
    k1_stream_kernel(const T* __restrict__ db, int64_t n_rows, const float* __restrict__ queries,
                     uint64_t* __restrict__ cand, int kc, int64_t rows_per_cta, const K1Extra ex) {
  constexpr int kK1Threads = NT, kK1Warps = NT / 32;
  constexpr int EPV = Elt<T>::kPerVec;        // elements per 16-byte vector
  constexpr int STEPS = D / (32 * EPV);       // vectors per lane per row
  constexpr int QV = (EPV == 4) ? 1 : 2;      // float4 query slices per database vector
  static_assert(D % (32 * EPV) == 0, "dim must be a multiple of 32 vectors");

  // queries staged so that lane l's float4 slices are contiguous across lanes (conflict-free)
  __shared__ __align__(16) float q_s[Q * D];
  __shared__ uint64_t merge_keys[kK1Threads];  // 8 sorted lists of 32 keys (one per warp)
  __shared__ K1TailSmem<Q> tail;
  __shared__ int is_last_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // programmatic dependent launch: let the (1-4 block) K3 grid become resident now; it parks
  // in griddepcontrol.wait until this grid has completed and its candidate keys are visible
  asm volatile("griddepcontrol.launch_dependents;");
  if constexpr (Q == 1) {
    if (ex.k3.stamps != nullptr && tid == 0) atomicMin(ex.k3.stamps + 0, global_timer_ns());
  }

  const int64_t row0 = int64_t(blockIdx.x) * rows_per_cta;
  const int64_t row1 = min(n_rows, row0 + rows_per_cta);
  const uint4* __restrict__ dbv = reinterpret_cast<const uint4*>(db);
  constexpr int VPR = D / EPV;  // vectors per row
  const float* __restrict__ bias = ex.row_bias;

  uint4 v[R][STEPS];
  float bz[R];
  // all loads of one step (R rows) are issued before any math: R * STEPS 128-bit loads in flight
  auto load_rows = [&](int64_t base) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = min(base + r, n_rows - 1);  // clamp: tail rows are masked below
      const uint4* p = dbv + row * VPR + lane;
#pragma unroll
      for (int j = 0; j < STEPS; ++j) v[r][j] = ld_stream_v4(p + j * 32);
      bz[r] = (bias != nullptr) ? __ldg(bias + row) : 0.f;
    }
  };
  int64_t base = row0 + int64_t(warp) * R;
  // the first rows are requested before the queries are staged: the HBM latency of the first
  // step overlaps the staging (matters for small shards, where a CTA runs only a few steps)
  if (base < row1) load_rows(base);

  for (int e = tid; e < Q * D; e += kK1Threads) {
    int q = e / D, d = e % D;
    int vec = d / EPV, c = d % EPV;           // vec = step*32 + lane
    int step = vec >> 5, ln = vec & 31;
    int half = c >> 2, cc = c & 3;
    int pos = ((step * QV + half) * 32 + ln) * 4 + cc;
    q_s[q * D + pos] = queries[e];
  }
  // pre-filter (`video != own` applied BEFORE the top-k): rows of the excluded group never enter
  // the lists, so they hold the best eligible rows whatever the size of the group
  int excl[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q)
    excl[q] = (ex.row_group != nullptr && ex.exclude_group != nullptr) ? ex.exclude_group[q] : -1;
  __syncthreads();

  // warp-distributed running top-kc per query: lane i holds slot i
  float ls[Q];
  int li[Q];
  float thr[Q];
  int minlane[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    ls[q] = (lane < kc) ? -INFINITY : INFINITY;
    li[q] = kInvalidIdx;
    thr[q] = -INFINITY;
    minlane[q] = kc - 1;
  }

  while (base < row1) {
    float acc[R][Q];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = 0; q < Q; ++q) acc[r][q] = 0.f;

#pragma unroll
    for (int j = 0; j < STEPS; ++j) {
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float4* qp = reinterpret_cast<const float4*>(q_s + q * D) + (j * QV) * 32 + lane;
        float4 qa = qp[0];
        if constexpr (EPV == 4) {
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r][q] = dot_slice_f32(v[r][j], qa, acc[r][q]);
        } else {
          float4 qb = qp[32];
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r][q] = dot_slice_bf16(v[r][j], qa, qb, acc[r][q]);
        }
      }
    }
    float bcur[R];
#pragma unroll
    for (int r = 0; r < R; ++r) bcur[r] = bz[r];
    const int64_t cur = base;
    base += int64_t(kK1Warps) * R;
    if (base < row1) load_rows(base);  // next step's loads fly while this step is reduced / inserted

#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        float a = acc[r][q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        acc[r][q] = a + bcur[r];
      }

#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = cur + r;
      if (row < row1) {  // warp-uniform
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float s = acc[r][q];
          // rows arrive in ascending order inside a warp, so an equal score never displaces
          if (s > thr[q]) {  // warp-uniform
            if (excl[q] >= 0 && __ldg(ex.row_group + row) == excl[q]) continue;  // pre-filtered row
            if (lane == minlane[q]) {
              ls[q] = s;
              li[q] = int(row);
            }
            float ws = ls[q];
            int wi = li[q], wl = lane;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              float os = __shfl_xor_sync(0xffffffffu, ws, off);
              int oi = __shfl_xor_sync(0xffffffffu, wi, off);
              int ol = __shfl_xor_sync(0xffffffffu, wl, off);
              if (worse(os, oi, ol, ws, wi, wl)) {
                ws = os;
                wi = oi;
                wl = ol;
              }
            }
            thr[q] = ws;
            minlane[q] = wl;
          }
        }
      }
    }
  }

  if constexpr (Q == 1) {
    if (ex.k3.stamps != nullptr && tid == 0) atomicMax(ex.k3.stamps + 1, global_timer_ns());
  }
  // Everything above only READ global memory (rows, bias, groups, the query: inputs that were complete before
  // this kernel was enqueued). When the single-query form is launched with the programmatic-stream-serialization
  // attribute this grid may have started while its predecessor in the stream — normally the previous search,
  // whose last CTA is still selecting / re-ranking / exchanging — was running: wait for it to complete (and
  // its writes to be visible) before the first global write. A no-op for ordinary launches.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // CTA merge: every warp sorts its 32 slots with shuffles; the sorted lists are folded four at a time
  // (bitonic "keep the 32 smallest" merges), then warp 0 folds the partial results into the best kc —
  // three barriers per query, no sorting network in shared memory
#pragma unroll 1
  for (int q = 0; q < Q; ++q) {
    const uint64_t own = warp_sort_u64((lane < kc) ? make_sim_key(ls[q], li[q]) : kEmptyKey, lane);
    if (q > 0) __syncthreads();  // previous query's lists have been folded
    merge_keys[tid] = own;
    __syncthreads();
    if (warp < kK1Warps / 4) {
      uint64_t acc = merge_keys[(4 * warp) * 32 + lane];
#pragma unroll
      for (int j = 1; j < 4; ++j) acc = warp_merge_keep32(acc, merge_keys[(4 * warp + j) * 32 + 31 - lane], lane);
      __syncwarp();
      merge_keys[(4 * warp) * 32 + lane] = acc;
    }
    __syncthreads();
    if (warp == 0) {
      uint64_t acc = merge_keys[lane];
#pragma unroll
      for (int j = 1; j < kK1Warps / 4; ++j) acc = warp_merge_keep32(acc, merge_keys[(4 * j) * 32 + 31 - lane], lane);
      if (lane < kc) cand[(int64_t(q) * gridDim.x + blockIdx.x) * kc + lane] = acc;
    }
  }

  // fused tail (single query): the last CTA to finish selects, re-scores in fp32, filters and emits
  // (and runs the cross-GPU exchange of a row-sharded store) — the whole search is ONE launch
  if constexpr (Q == 1) {
    if (ex.ticket != nullptr) {
      if (warp == 0) {
        // warp 0 wrote this CTA's candidate keys; the release half of the ticket (cumulative over what
        // the warp barrier ordered before it) publishes them device-wide, the acquire half plus the CTA
        // barrier below makes every other CTA's keys visible to the whole last CTA
        __syncwarp();
        if (lane == 0) {
          int old;
          asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(ex.ticket) : "memory");
          is_last_s = (old == int(gridDim.x) - 1) ? 1 : 0;
        }
      }
      __syncthreads();
      if (is_last_s) {
        if (ex.k3.stamps != nullptr && tid == 0) ex.k3.stamps[2] = global_timer_ns();
        k3_body<kK1Threads>(ex.k3, 0, tail.k3);
        if (tid == 0) *ex.ticket = 0;  // re-armed for the next call / graph replay
        if (ex.k3.stamps != nullptr && tid == 0) ex.k3.stamps[8] = global_timer_ns();
      }
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
// (rows in flight per warp, resident CTAs per SM), measured on B200 (profiles/r1_k1_*): fp32 (2, 2);
// bf16 768-d single query (6, 2); other bf16 shapes 4 rows, 3 CTAs unless 3-4 queries or 1024-d rows
// (4 rows in flight per warp do not fit 80 registers without spilling there)
template <typename T, int D, int Q, int R, int MINB, int NT>
static cudaError_t launch_cfg(const void* db, int64_t n_rows, const float* queries, uint64_t* cand,
                              int kc, int grid, const K1Extra& ex, cudaStream_t st) {
  const int64_t quantum = int64_t(NT / 32) * R;
  int64_t rows_per_cta = (n_rows + grid - 1) / grid;
  rows_per_cta = (rows_per_cta + quantum - 1) / quantum * quantum;
  K1Extra e = ex;
  if (Q != 1) e.ticket = nullptr;
  const T* dbt = static_cast<const T*>(db);
  if (Q == 1 && e.ticket != nullptr && e.overlap) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, k1_stream_kernel<T, D, Q, R, MINB, NT>, dbt, n_rows, queries, cand, kc,
                                        rows_per_cta, e);
    note_launch();
    return le != cudaSuccess ? le : cudaGetLastError();
  }
  k1_stream_kernel<T, D, Q, R, MINB, NT><<<grid, NT, 0, st>>>(dbt, n_rows, queries, cand, kc, rows_per_cta, e);
  note_launch();
  return cudaGetLastError();
}

template <typename T, int D, int Q>
static cudaError_t launch_one(const void* db, int64_t n_rows, const float* queries, uint64_t* cand,
                              int kc, int grid, const K1Extra& ex, cudaStream_t st) {
  constexpr bool F = sizeof(T) == 4;
  if constexpr (Q == 1)
    return launch_cfg<T, D, Q, F ? 2 : (D == 768 ? 6 : 4), 1, 512>(db, n_rows, queries, cand, kc, grid, ex, st);
  else  // `else` matters: without it the 256-thread form is instantiated for Q == 1 as well (8 dead kernels)
    return launch_cfg<T, D, Q, F ? 2 : 4, (F || Q >= 3 || D >= 1024) ? 2 : 3, 256>(db, n_rows, queries, cand, kc, grid, ex, st);
}

template <typename T, int D>
static cudaError_t launch_q(const void* db, int64_t n_rows, const float* queries, int nq,
                            uint64_t* cand, int kc, int grid, const K1Extra& ex, cudaStream_t st) {
  switch (nq) {
    case 1: return launch_one<T, D, 1>(db, n_rows, queries, cand, kc, grid, ex, st);
    case 2: return launch_one<T, D, 2>(db, n_rows, queries, cand, kc, grid, ex, st);
    case 3: return launch_one<T, D, 3>(db, n_rows, queries, cand, kc, grid, ex, st);
    case 4: return launch_one<T, D, 4>(db, n_rows, queries, cand, kc, grid, ex, st);
    default: return cudaErrorInvalidValue;
  }
}

template <typename T>
static cudaError_t launch_d(const void* db, int64_t n_rows, int dim, const float* queries, int nq,
                            uint64_t* cand, int kc, int grid, const K1Extra& ex, cudaStream_t st) {
  switch (dim) {
    case 256: return launch_q<T, 256>(db, n_rows, queries, nq, cand, kc, grid, ex, st);
    case 512: return launch_q<T, 512>(db, n_rows, queries, nq, cand, kc, grid, ex, st);
    case 768: return launch_q<T, 768>(db, n_rows, queries, nq, cand, kc, grid, ex, st);
    case 1024: return launch_q<T, 1024>(db, n_rows, queries, nq, cand, kc, grid, ex, st);
    default: return cudaErrorInvalidValue;
  }
}

bool k1_supported(int dim, int nq) {
  return (dim == 256 || dim == 512 || dim == 768 || dim == 1024) && nq >= 1 && nq <= 4;
}

int k1_grid(int64_t n_rows, int elt_bytes, int dim, int nq, int sm_count, bool spare_sm) {
  const int r = elt_bytes == 4 ? 2 : ((dim == 768 && nq == 1) ? 6 : 4);
  const int minb = nq == 1 ? 1 : ((elt_bytes == 4 || nq >= 3 || dim >= 1024) ? 2 : 3);
  const int64_t quantum = int64_t(nq == 1 ? 16 : 8) * r;
  // one resident wave; small tables get fewer CTAs
  int64_t want = (n_rows + quantum - 1) / quantum;
  int64_t cap = int64_t(sm_count) * minb;
  if (spare_sm && nq == 1 && cap > 1) cap -= 1;
  return int(want < cap ? (want < 1 ? 1 : want) : cap);
}

cudaError_t launch_k1_stream(const void* db, int elt_bytes, int64_t n_rows, int dim,
                             const float* queries, int nq, uint64_t* cand, int kc, int grid,
                             const K1Extra& ex, cudaStream_t st) {
  if (elt_bytes == 4) return launch_d<float>(db, n_rows, dim, queries, nq, cand, kc, grid, ex, st);
  return launch_d<__nv_bfloat16>(db, n_rows, dim, queries, nq, cand, kc, grid, ex, st);
}

}  // namespace mrag
