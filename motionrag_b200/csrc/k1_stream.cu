// K1 — HBM-streaming similarity scan with fused top-k (few queries, nq <= 4).
//
// Replaces the LanceDB flat scan behind table.search(vec, 'text_embedding').limit(k)
// (reference src/data/rag.py:54): every database row is read exactly once with 128-bit
// coalesced, L1-bypassing loads; one warp owns a row at a time (lane l holds the 16-byte
// slices l, l+32, ...), dots are finished with a 5-step xor-shuffle, and each warp keeps a
// running top-KC in registers (lane i = slot i) guarded by a warp-uniform threshold so the
// score vector never leaves the SM. Per-CTA lists are merged with a shared-memory bitonic
// sort and written as u64 keys (score descending, index ascending); K3 finishes the job.
//
// Algorithmic bytes: n_rows * dim * sizeof(T) per launch (T = float or bf16).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace mrag {

constexpr int kK1Threads = 256;
constexpr int kK1Warps = kK1Threads / 32;

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

template <typename T, int D, int Q, int R, int MINB>
__global__ void __launch_bounds__(kK1Threads, MINB)
    k1_stream_kernel(const T* __restrict__ db, int64_t n_rows, const float* __restrict__ queries,
                     uint64_t* __restrict__ cand, int kc, int64_t rows_per_cta) {
  constexpr int EPV = Elt<T>::kPerVec;        // elements per 16-byte vector
  constexpr int STEPS = D / (32 * EPV);       // vectors per lane per row
  constexpr int QV = (EPV == 4) ? 1 : 2;      // float4 query slices per database vector
  static_assert(D % (32 * EPV) == 0, "dim must be a multiple of 32 vectors");

  // queries staged so that lane l's float4 slices are contiguous across lanes (conflict-free)
  __shared__ __align__(16) float q_s[Q * D];
  __shared__ uint64_t merge_keys[kK1Threads];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // programmatic dependent launch: let the (1-4 block) K3 grid become resident now; it parks
  // in griddepcontrol.wait until this grid has completed and its candidate keys are visible
  asm volatile("griddepcontrol.launch_dependents;");
  for (int e = tid; e < Q * D; e += kK1Threads) {
    int q = e / D, d = e % D;
    int vec = d / EPV, c = d % EPV;           // vec = step*32 + lane
    int step = vec >> 5, ln = vec & 31;
    int half = c >> 2, cc = c & 3;
    int pos = ((step * QV + half) * 32 + ln) * 4 + cc;
    q_s[q * D + pos] = queries[e];
  }
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

  const int64_t row0 = int64_t(blockIdx.x) * rows_per_cta;
  const int64_t row1 = min(n_rows, row0 + rows_per_cta);
  const uint4* __restrict__ dbv = reinterpret_cast<const uint4*>(db);
  constexpr int VPR = D / EPV;  // vectors per row

  for (int64_t base = row0 + int64_t(warp) * R; base < row1; base += int64_t(kK1Warps) * R) {
    uint4 v[R][STEPS];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int64_t row = min(base + r, n_rows - 1);  // clamp: tail rows are masked below
      const uint4* p = dbv + row * VPR + lane;
#pragma unroll
      for (int j = 0; j < STEPS; ++j) v[r][j] = ld_stream_v4(p + j * 32);
    }
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
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        float a = acc[r][q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        acc[r][q] = a;
      }

#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = base + r;
      if (row < row1) {  // warp-uniform
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float s = acc[r][q];
          // rows arrive in ascending order inside a warp, so an equal score never displaces
          if (s > thr[q]) {  // warp-uniform
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

  // CTA merge: 8 warps x 32 slots -> sorted, keep the best kc
#pragma unroll 1
  for (int q = 0; q < Q; ++q) {
    __syncthreads();
    merge_keys[tid] = (lane < kc) ? make_sim_key(ls[q], li[q]) : kEmptyKey;
    bitonic_sort_smem(merge_keys, kK1Threads, tid, kK1Threads);
    if (tid < kc) cand[(int64_t(q) * gridDim.x + blockIdx.x) * kc + tid] = merge_keys[tid];
  }
}

// ---- host side ------------------------------------------------------------------------------
// (rows in flight per warp, resident CTAs per SM). Variant 0 is the shipped configuration;
// MRAG_K1_VARIANT selects the others for tuning runs.
struct K1Variant {
  int r, minb;
};
static const K1Variant kVariantsF32[] = {{2, 2}, {4, 2}, {2, 3}, {2, 4}, {1, 4}, {3, 2}};
static const K1Variant kVariantsBF16[] = {{6, 2}, {8, 2}, {4, 2}, {4, 4}, {2, 4}, {4, 3}};

static int k1_variant_index() {
  const char* e = getenv("MRAG_K1_VARIANT");
  int v = e ? atoi(e) : 0;
  return (v < 0 || v > 5) ? 0 : v;
}
static K1Variant k1_variant(int elt_bytes) {
  return elt_bytes == 4 ? kVariantsF32[k1_variant_index()] : kVariantsBF16[k1_variant_index()];
}

template <typename T, int D, int Q, int R, int MINB>
static cudaError_t launch_cfg(const void* db, int64_t n_rows, const float* queries, uint64_t* cand,
                              int kc, int grid, cudaStream_t st) {
  const int64_t quantum = int64_t(kK1Warps) * R;
  int64_t rows_per_cta = (n_rows + grid - 1) / grid;
  rows_per_cta = (rows_per_cta + quantum - 1) / quantum * quantum;
  k1_stream_kernel<T, D, Q, R, MINB><<<grid, kK1Threads, 0, st>>>(
      static_cast<const T*>(db), n_rows, queries, cand, kc, rows_per_cta);
  note_launch();
  return cudaGetLastError();
}

template <typename T, int D, int Q>
static cudaError_t launch_one(const void* db, int64_t n_rows, const float* queries, uint64_t* cand,
                              int kc, int grid, cudaStream_t st) {
  constexpr bool F = sizeof(T) == 4;
  if constexpr (D == 768 && Q == 1) {  // tuning variants exist for the headline shape only
    switch (k1_variant_index()) {
      case 1: return launch_cfg<T, D, Q, F ? 4 : 8, 2>(db, n_rows, queries, cand, kc, grid, st);
      case 2: return launch_cfg<T, D, Q, F ? 2 : 4, F ? 3 : 2>(db, n_rows, queries, cand, kc, grid, st);
      case 3: return launch_cfg<T, D, Q, F ? 2 : 4, 4>(db, n_rows, queries, cand, kc, grid, st);
      case 4: return launch_cfg<T, D, Q, F ? 1 : 2, 4>(db, n_rows, queries, cand, kc, grid, st);
      case 5: return launch_cfg<T, D, Q, F ? 3 : 4, F ? 2 : 3>(db, n_rows, queries, cand, kc, grid, st);
      default: break;
    }
  }
  if constexpr (D == 768 && Q == 1) return launch_cfg<T, D, Q, F ? 2 : 6, 2>(db, n_rows, queries, cand, kc, grid, st);
  // bf16 with 3-4 queries or 1024-d rows: 4 rows in flight per warp do not fit 80 registers without spilling
  return launch_cfg<T, D, Q, F ? 2 : 4, (F || Q >= 3 || D >= 1024) ? 2 : 3>(db, n_rows, queries, cand, kc, grid, st);
}

template <typename T, int D>
static cudaError_t launch_q(const void* db, int64_t n_rows, const float* queries, int nq,
                            uint64_t* cand, int kc, int grid, cudaStream_t st) {
  switch (nq) {
    case 1: return launch_one<T, D, 1>(db, n_rows, queries, cand, kc, grid, st);
    case 2: return launch_one<T, D, 2>(db, n_rows, queries, cand, kc, grid, st);
    case 3: return launch_one<T, D, 3>(db, n_rows, queries, cand, kc, grid, st);
    case 4: return launch_one<T, D, 4>(db, n_rows, queries, cand, kc, grid, st);
    default: return cudaErrorInvalidValue;
  }
}

template <typename T>
static cudaError_t launch_d(const void* db, int64_t n_rows, int dim, const float* queries, int nq,
                            uint64_t* cand, int kc, int grid, cudaStream_t st) {
  switch (dim) {
    case 256: return launch_q<T, 256>(db, n_rows, queries, nq, cand, kc, grid, st);
    case 512: return launch_q<T, 512>(db, n_rows, queries, nq, cand, kc, grid, st);
    case 768: return launch_q<T, 768>(db, n_rows, queries, nq, cand, kc, grid, st);
    case 1024: return launch_q<T, 1024>(db, n_rows, queries, nq, cand, kc, grid, st);
    default: return cudaErrorInvalidValue;
  }
}

bool k1_supported(int dim, int nq) {
  return (dim == 256 || dim == 512 || dim == 768 || dim == 1024) && nq >= 1 && nq <= 4;
}

int k1_grid(int64_t n_rows, int elt_bytes, int dim, int nq, int sm_count) {
  // the headline shape (768-d, one query) has tuned variants; other shapes use (2,2) / (4,3 or 2)
  K1Variant v = (dim == 768 && nq == 1) ? k1_variant(elt_bytes)
                                        : (elt_bytes == 4 ? K1Variant{2, 2} : K1Variant{4, (nq >= 3 || dim >= 1024) ? 2 : 3});
  const int64_t quantum = int64_t(kK1Warps) * v.r;
  // one resident wave; small tables get fewer CTAs
  int64_t want = (n_rows + quantum - 1) / quantum;
  int64_t cap = int64_t(sm_count) * v.minb;
  return int(want < cap ? (want < 1 ? 1 : want) : cap);
}

cudaError_t launch_k1_stream(const void* db, int elt_bytes, int64_t n_rows, int dim,
                             const float* queries, int nq, uint64_t* cand, int kc, int grid,
                             cudaStream_t st) {
  if (elt_bytes == 4) return launch_d<float>(db, n_rows, dim, queries, nq, cand, kc, grid, st);
  return launch_d<__nv_bfloat16>(db, n_rows, dim, queries, nq, cand, kc, grid, st);
}

}  // namespace mrag
