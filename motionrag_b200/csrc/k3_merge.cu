// K0 (store preparation), K3 (candidate merge + exact fp32 re-score + filter) and the
// cross-shard merge.
//
// K3 turns the per-CTA / per-chunk candidate keys written by K1 / K2 into the record list the
// reference returns from RAGDatabase.vector_search (src/data/rag.py:54-61): at most k rows,
// ascending `_distance`, computed in fp32 from the master rows with LanceDB's own formulas
// (l2 = sum (q-d)^2, cosine = 1 - cos, dot = 1 - q.d), ties broken by lowest row index, and
// the `video != "<own>"` filter of src/data/datamodule.py:235 applied as a post-filter
// (LanceDB 0.14 default) or pre-filter.
#include "common.cuh"
#include "kernels.h"

namespace mrag {

// ---- K0 -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k0_prepare_rows_kernel(float* __restrict__ rows, __nv_bfloat16* __restrict__ shadow, int64_t n,
                           int dim, int normalise, unsigned int* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  float4* p = reinterpret_cast<float4*>(rows + row * dim);
  const int nv = dim >> 2;
  float scale = 1.f;
  float ss = 0.f;
  for (int i = lane; i < nv; i += 32) {
    float4 v = p[i];
    ss = fmaf(v.x, v.x, ss);
    ss = fmaf(v.y, v.y, ss);
    ss = fmaf(v.z, v.z, ss);
    ss = fmaf(v.w, v.w, ss);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  // x / max(|x|, eps), the torch.nn.functional.normalize convention
  if (normalise) scale = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0 && stats != nullptr) {
    // stats[0] = max | |row|^2 - 1 | over non-zero rows as stored (float bits, >= 0 so uint order
    // works); stats[1] = number of all-zero rows (LanceDB's on_bad_vectors='fill' produces them)
    if (ss == 0.f) {
      atomicAdd(&stats[1], 1u);
    } else {
      const float stored = normalise ? ss * scale * scale : ss;
      atomicMax(&stats[0], __float_as_uint(fabsf(stored - 1.f)));
    }
  }
  uint2* o = reinterpret_cast<uint2*>(shadow + row * dim);
  for (int i = lane; i < nv; i += 32) {
    float4 v = p[i];
    if (normalise) {
      v.x *= scale;
      v.y *= scale;
      v.z *= scale;
      v.w *= scale;
      p[i] = v;
    }
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    o[i] = pk;
  }
}

__global__ void __launch_bounds__(256)
    k0_cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n4) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(in)[i];
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&a);
  pk.y = *reinterpret_cast<uint32_t*>(&b);
  reinterpret_cast<uint2*>(out)[i] = pk;
}

cudaError_t launch_prepare_rows(float* rows_f32, void* rows_bf16, int64_t n, int dim,
                                bool normalise, unsigned int* stats, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const int64_t threads = n * 32;
  const int64_t blocks = (threads + 255) / 256;
  k0_prepare_rows_kernel<<<unsigned(blocks), 256, 0, st>>>(
      rows_f32, static_cast<__nv_bfloat16*>(rows_bf16), n, dim, normalise ? 1 : 0, stats);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_cast_queries_bf16(const float* q, void* q_bf16, int nq, int dim,
                                     cudaStream_t st) {
  const int64_t n4 = int64_t(nq) * dim / 4;
  k0_cast_bf16_kernel<<<unsigned((n4 + 255) / 256), 256, 0, st>>>(
      q, static_cast<__nv_bfloat16*>(q_bf16), n4);
  note_launch();
  return cudaGetLastError();
}

// ---- shared tail: filter + emit the first entries of a sorted list (executed by warp 0) ----
// entry(j) -> valid?, distance, global index, group; list length <= 64, sorted ascending.
struct Emit {
  float dist;
  int64_t idx;
  int32_t group;
  bool valid;
};

template <typename EntryFn>
__device__ __forceinline__ void emit_filtered(EntryFn entry, int n_sorted, int k, int filter_mode,
                                              int exclude, float* out_dist, int64_t* out_idx,
                                              int32_t* out_group, int lane) {
  // two entries per lane: j0 = lane, j1 = lane + 32
  Emit e0 = entry(lane, lane < n_sorted);
  Emit e1 = entry(lane + 32, lane + 32 < n_sorted);
  const bool excl0 = (filter_mode != 0) && (exclude >= 0) && e0.valid && (e0.group == exclude);
  const bool excl1 = (filter_mode != 0) && (exclude >= 0) && e1.valid && (e1.group == exclude);
  bool keep0, keep1;
  if (filter_mode == 1) {  // post-filter: only the k nearest are eligible at all
    keep0 = e0.valid && (lane < k) && !excl0;
    keep1 = e1.valid && (lane + 32 < k) && !excl1;
  } else {
    keep0 = e0.valid && !excl0;
    keep1 = e1.valid && !excl1;
  }
  const uint32_t b0 = __ballot_sync(0xffffffffu, keep0);
  const uint32_t b1 = __ballot_sync(0xffffffffu, keep1);
  const uint32_t lt = (1u << lane) - 1u;
  const int pos0 = __popc(b0 & lt);
  const int pos1 = __popc(b0) + __popc(b1 & lt);
  if (keep0 && pos0 < k) {
    out_dist[pos0] = e0.dist;
    out_idx[pos0] = e0.idx;
    if (out_group) out_group[pos0] = e0.group;
  }
  if (keep1 && pos1 < k) {
    out_dist[pos1] = e1.dist;
    out_idx[pos1] = e1.idx;
    if (out_group) out_group[pos1] = e1.group;
  }
  const int total = min(k, __popc(b0) + __popc(b1));
  for (int j = total + lane; j < k; j += 32) {
    out_dist[j] = INFINITY;
    out_idx[j] = -1;
    if (out_group) out_group[j] = -1;
  }
}

// ---- cross-GPU exchange fused into K3 (row-sharded stores) -------------------------------------
// Every rank owns an exchange buffer that all peers have mapped (CUDA IPC over NVLink):
//   records: [slot 2][src rank][query][ idx i64 x k_cap | dist f32 x k_cap | group i32 x k_cap ]
//   flags  : [slot 2][src rank][query] u32 epoch
// A K3 block (one query) stores its shard's top-k record into the same (slot, own rank, query)
// cell of EVERY rank's buffer with plain NVLink stores, fences at system scope, raises the
// matching flags, then waits for the flags of all source ranks in its OWN buffer and merges the
// world * k candidates locally. No collective library call, no extra launch: the exchange
// overlaps with the other queries' blocks. Slots alternate with the epoch, so a rank can run at
// most one call ahead of the slowest peer, which cannot still be reading the slot being
// rewritten.
struct XchgArgs {
  int world, rank;      // world <= 1 disables the exchange
  int nq_cap, k_cap;
  uint32_t epoch;
  char* const* bufs;    // device array [world]: exchange-buffer base of every rank
};
__host__ __device__ inline size_t xchg_rec_bytes(int k_cap) { return size_t(k_cap) * 16; }
__host__ __device__ inline size_t xchg_flags_offset(int world, int nq_cap, int k_cap) {
  return size_t(2) * world * nq_cap * xchg_rec_bytes(k_cap);
}
__host__ __device__ inline size_t xchg_total_bytes(int world, int nq_cap, int k_cap) {
  return xchg_flags_offset(world, nq_cap, k_cap) + size_t(2) * world * nq_cap * 4;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- K3 -------------------------------------------------------------------------------------
// Candidates arrive as `n_runs` runs of `run_len` keys, each run sorted best-first (one run per
// K1 CTA / K2 chunk). A full sort of up to 16 K keys is shared-memory-bandwidth bound (~80 us on
// one SM), so selection is done on the run heads instead: with R = rerank, the R-th best key
// overall can be no worse than T = the R-th best run head, hence only keys <= T (in key order)
// can matter, and they all live in the (exactly R, keys are unique) runs whose head is <= T.
// Sort <= 1024 heads -> T -> compact the qualifying keys (<= R * run_len <= 2048) -> sort those.
constexpr int kMaxRerank = 64;
constexpr int kMaxRuns = 1024;
constexpr int kMaxSel = 2048;
constexpr int kK3Threads = 1024;  // 32 warps: one re-rank candidate / one qualifying run per warp

__global__ void __launch_bounds__(kK3Threads)
    k3_merge_rerank_kernel(const uint64_t* __restrict__ cand, int n_runs, int run_len,
                           const float* __restrict__ db, int dim, const float* __restrict__ queries,
                           const int32_t* __restrict__ row_group,
                           const int32_t* __restrict__ exclude_group, int filter_mode, int metric,
                           int rerank, int k, int64_t index_base, float* __restrict__ out_dist,
                           int64_t* __restrict__ out_idx, int32_t* __restrict__ out_group,
                           float* __restrict__ out_margin, const XchgArgs x) {
  __shared__ uint64_t heads[kMaxRuns];      // run heads (index = run)
  __shared__ uint64_t small_sorted[256];    // output of the rank sorts
  __shared__ uint64_t sel[kMaxSel];
  __shared__ uint64_t rr_keys[kMaxRerank];  // (ordered distance << 32) | local row
  __shared__ float rr_dot[kMaxRerank];      // true q.d of candidate slot c
  __shared__ uint32_t rr_row[kMaxRerank];   // local row of candidate slot c
  __shared__ float q_norm_s;
  __shared__ int n_sel_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = blockIdx.x;
  const uint64_t* src = cand + int64_t(q) * n_runs * run_len;
  // launched with programmatic stream serialization: wait here for the scan kernel's results
  asm volatile("griddepcontrol.wait;" ::: "memory");

  int heads_pad = 64;
  while (heads_pad < n_runs) heads_pad <<= 1;
  for (int r = tid; r < heads_pad; r += kK3Threads) {
    const uint64_t h = (r < n_runs) ? src[int64_t(r) * run_len] : kEmptyKey;
    heads[r] = h;
  }
  if (tid == 0) n_sel_s = 0;
  if (tid < kMaxRerank) rr_keys[tid] = kEmptyKey;
  uint64_t T = kEmptyKey;  // select everything unless there are more runs than needed
  __shared__ uint64_t T_s;
  if (n_runs > rerank) {
    // T = the rerank-th best run head: rank by counting (no sorting network, one barrier)
    __syncthreads();
    for (int i = tid; i < n_runs; i += kK3Threads) {
      const uint64_t mine = heads[i];
      int r = 0;
      for (int j = 0; j < n_runs; ++j) {
        const uint64_t o = heads[j];
        r += (o < mine) || (o == mine && j < i);
      }
      if (r == rerank - 1) T_s = mine;
    }
    __syncthreads();
    T = T_s;
  } else {
    __syncthreads();
  }
  // compact keys <= T from qualifying runs: one warp per run, lane = position in the run
  for (int r = warp; r < n_runs; r += kK3Threads / 32) {
    if (heads[r] > T) continue;  // warp-uniform
    const uint64_t key = (lane < run_len) ? src[int64_t(r) * run_len + lane] : kEmptyKey;
    const bool take = (key <= T) && (uint32_t(key) < uint32_t(kInvalidIdx));
    const uint32_t m = __ballot_sync(0xffffffffu, take);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(&n_sel_s, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      if (pos < kMaxSel) sel[pos] = key;
    }
  }
  __syncthreads();
  const int n_sel = min(n_sel_s, kMaxSel);
  int sel_pad = 64;
  while (sel_pad < n_sel) sel_pad <<= 1;
  if (n_sel <= 256) {
    rank_sort_smem(sel, small_sorted, n_sel, tid, kK3Threads);
    for (int i = tid; i < n_sel; i += kK3Threads) sel[i] = small_sorted[i];
    __syncthreads();
  } else {
    for (int i = n_sel + tid; i < sel_pad; i += kK3Threads) sel[i] = kEmptyKey;
    bitonic_sort_smem(sel, sel_pad, tid, kK3Threads);
  }

  // exact fp32 distances for the best `rerank` candidates: one warp per candidate
  const float4* qv = reinterpret_cast<const float4*>(queries + int64_t(q) * dim);
  const int nv = dim >> 2;
  const int n_rr = min(rerank, min(n_sel, kMaxRerank));
  for (int c = warp; c < n_rr; c += kK3Threads / 32) {
    const uint64_t key = sel[c];
    const uint32_t idx = uint32_t(key);
    const float4* dv = reinterpret_cast<const float4*>(db + int64_t(idx) * dim);
    float l2 = 0.f, dot = 0.f, qq = 0.f, dd = 0.f;
    for (int i = lane; i < nv; i += 32) {
      const float4 a = qv[i];
      const float4 b = dv[i];
      float t;
      t = a.x - b.x; l2 = fmaf(t, t, l2);
      t = a.y - b.y; l2 = fmaf(t, t, l2);
      t = a.z - b.z; l2 = fmaf(t, t, l2);
      t = a.w - b.w; l2 = fmaf(t, t, l2);
      dot = fmaf(a.x, b.x, dot); dot = fmaf(a.y, b.y, dot);
      dot = fmaf(a.z, b.z, dot); dot = fmaf(a.w, b.w, dot);
      qq = fmaf(a.x, a.x, qq); qq = fmaf(a.y, a.y, qq);
      qq = fmaf(a.z, a.z, qq); qq = fmaf(a.w, a.w, qq);
      dd = fmaf(b.x, b.x, dd); dd = fmaf(b.y, b.y, dd);
      dd = fmaf(b.z, b.z, dd); dd = fmaf(b.w, b.w, dd);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      l2 += __shfl_xor_sync(0xffffffffu, l2, off);
      dot += __shfl_xor_sync(0xffffffffu, dot, off);
      qq += __shfl_xor_sync(0xffffffffu, qq, off);
      dd += __shfl_xor_sync(0xffffffffu, dd, off);
    }
    float dist;
    if (metric == 0) dist = l2;
    else if (metric == 1) dist = 1.f - dot / fmaxf(sqrtf(qq) * sqrtf(dd), 1e-30f);
    else dist = 1.f - dot;
    if (lane == 0) {
      rr_keys[c] = (uint64_t(f32_to_ordered(dist)) << 32) | idx;
      rr_dot[c] = dot;
      rr_row[c] = idx;
      if (c == 0) q_norm_s = sqrtf(qq);
    }
  }
  rank_sort_smem(rr_keys, small_sorted, kMaxRerank, tid, kK3Threads);
  if (tid < kMaxRerank) rr_keys[tid] = small_sorted[tid];
  __syncthreads();

  // exactness certificate of the bf16 scan (see mrag_search_params.out_margin)
  if (out_margin != nullptr && tid == 0) {
    float margin = INFINITY;
    if (x.world > 1 || filter_mode == 2) {
      margin = __int_as_float(0x7fc00000);  // NaN: not defined for sharded / pre-filtered searches
    } else if (n_rr == rerank && n_rr > 0) {
      // rows outside the re-ranked set may exist (with fewer than `rerank` candidates no run was
      // full, so every row was a candidate and was re-ranked): the re-ranked set is the global
      // top-`rerank` by scan score (rerank <= run length), so their score is <= the weakest one
      const float weakest = sim_key_score(sel[n_rr - 1]);
      const int kth = min(k, n_rr) - 1;
      const uint32_t row = uint32_t(rr_keys[kth]);
      float dk = -INFINITY;
      for (int c = 0; c < n_rr; ++c)
        if (rr_row[c] == row) dk = rr_dot[c];
      margin = (dk - weakest) / fmaxf(q_norm_s, 1e-30f);
    }
    out_margin[q] = margin;
  }

  const int exclude = (exclude_group != nullptr) ? exclude_group[q] : -1;
  auto entry = [&](int j, bool in_range) {
    Emit e;
    e.valid = false;
    e.dist = INFINITY;
    e.idx = -1;
    e.group = -1;
    if (in_range && j < kMaxRerank) {
      const uint64_t key = rr_keys[j];
      const uint32_t idx = uint32_t(key);
      if (key != kEmptyKey && idx < uint32_t(kInvalidIdx)) {
        e.valid = true;
        e.dist = ordered_to_f32(uint32_t(key >> 32));
        e.idx = index_base + int64_t(idx);
        e.group = (row_group != nullptr) ? row_group[idx] : -1;
      }
    }
    return e;
  };
  if (x.world <= 1) {
    if (warp == 0)
      emit_filtered(entry, n_rr, k, filter_mode, exclude, out_dist + int64_t(q) * k,
                    out_idx + int64_t(q) * k, out_group ? out_group + int64_t(q) * k : nullptr, lane);
    return;
  }

  // ---- row-sharded: publish this shard's top-k to every rank, wait for theirs, merge ----
  __shared__ float rec_dist[32];
  __shared__ int64_t rec_idx[32];
  __shared__ int32_t rec_grp[32];
  // the post-filter belongs after the GLOBAL top-k; a pre-filter can be applied per shard
  if (warp == 0)
    emit_filtered(entry, n_rr, k, filter_mode == 2 ? 2 : 0, exclude, rec_dist, rec_idx, rec_grp, lane);
  __syncthreads();
  const int slot = int(x.epoch & 1u);
  const size_t rec_bytes = xchg_rec_bytes(x.k_cap);
  const size_t cell = (size_t(slot) * x.world + x.rank) * x.nq_cap + q;  // (slot, src = me, query)
  const size_t flags_off = xchg_flags_offset(x.world, x.nq_cap, x.k_cap);
  for (int t = tid; t < x.world * k; t += kK3Threads) {
    const int r = t / k, j = t % k;
    char* rec = x.bufs[r] + cell * rec_bytes;
    reinterpret_cast<int64_t*>(rec)[j] = rec_idx[j];
    reinterpret_cast<float*>(rec + size_t(x.k_cap) * 8)[j] = rec_dist[j];
    reinterpret_cast<int32_t*>(rec + size_t(x.k_cap) * 12)[j] = rec_grp[j];
  }
  // The record stores above are ordered before the flag by the CTA barrier followed by a
  // system-scope RELEASE store (release is cumulative over what happened-before it in this
  // CTA); a separate __threadfence_system() per thread would only add a second fence round trip.
  __syncthreads();
  if (tid < x.world)
    st_release_sys(reinterpret_cast<uint32_t*>(x.bufs[tid] + flags_off) + cell, x.epoch);
  if (tid < x.world) {
    const size_t src_cell = (size_t(slot) * x.world + tid) * x.nq_cap + q;
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(x.bufs[x.rank] + flags_off) + src_cell;
    while (ld_acquire_sys(flag) != x.epoch) {
    }
  }
  __syncthreads();
  // merge world * k candidates; slot order == global row order among equal distances
  const char* mine = x.bufs[x.rank];
  const int total = x.world * k;  // <= 256
  uint64_t* mk = sel;             // reuse: 256 keys
  if (tid < 256) {
    uint64_t key = kEmptyKey;
    if (tid < total) {
      const int r = tid / k, j = tid % k;
      const char* rec = mine + ((size_t(slot) * x.world + r) * x.nq_cap + q) * rec_bytes;
      const int64_t gi = __ldcv(reinterpret_cast<const long long*>(rec) + j);
      const float gd = __ldcv(reinterpret_cast<const float*>(rec + size_t(x.k_cap) * 8) + j);
      if (gi >= 0) key = (uint64_t(f32_to_ordered(gd)) << 32) | uint32_t(tid);
    }
    mk[tid] = key;
  }
  rank_sort_smem(mk, small_sorted, 256, tid, kK3Threads);
  if (tid < 256) mk[tid] = small_sorted[tid];
  __syncthreads();
  if (warp == 0) {
    auto gentry = [&](int j, bool in_range) {
      Emit e;
      e.valid = false;
      e.dist = INFINITY;
      e.idx = -1;
      e.group = -1;
      if (in_range) {
        const uint64_t kk = mk[j];
        if (kk != kEmptyKey) {
          const int t = int(uint32_t(kk));
          const int r = t / k, jj = t % k;
          const char* rec = mine + ((size_t(slot) * x.world + r) * x.nq_cap + q) * rec_bytes;
          e.valid = true;
          e.idx = __ldcv(reinterpret_cast<const long long*>(rec) + jj);
          e.dist = __ldcv(reinterpret_cast<const float*>(rec + size_t(x.k_cap) * 8) + jj);
          e.group = __ldcv(reinterpret_cast<const int*>(rec + size_t(x.k_cap) * 12) + jj);
        }
      }
      return e;
    };
    emit_filtered(gentry, min(total, 64), k, filter_mode, exclude, out_dist + int64_t(q) * k,
                  out_idx + int64_t(q) * k, out_group ? out_group + int64_t(q) * k : nullptr, lane);
  }
}

size_t exchange_bytes(int world, int nq_cap, int k_cap) { return xchg_total_bytes(world, nq_cap, k_cap); }

cudaError_t launch_k3_merge_rerank(const uint64_t* cand, int n_runs, int run_len,
                                   const float* db_f32, int dim, const float* queries, int nq,
                                   const int32_t* row_group, const int32_t* exclude_group,
                                   int filter_mode, int metric, int rerank, int k,
                                   int64_t index_base, float* out_dist, int64_t* out_idx,
                                   int32_t* out_group, float* out_margin, const ExchangeDesc* xd,
                                   cudaStream_t st) {
  XchgArgs x{};
  if (xd != nullptr && xd->world > 1) {
    if (nq > xd->nq_cap || k > xd->k_cap || xd->k_cap > 32 || xd->world * k > 256)
      return cudaErrorInvalidValue;
    x.world = xd->world;
    x.rank = xd->rank;
    x.nq_cap = xd->nq_cap;
    x.k_cap = xd->k_cap;
    x.epoch = xd->epoch;
    x.bufs = reinterpret_cast<char* const*>(xd->bufs_dev);
  }
  if (n_runs < 1 || n_runs > kMaxRuns || run_len < 1 || run_len > 32 || rerank < 1 ||
      rerank > kMaxRerank || int64_t(rerank) * run_len > kMaxSel)
    return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(nq));
  cfg.blockDim = dim3(kK3Threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, k3_merge_rerank_kernel, cand, n_runs, run_len, db_f32, dim,
                                      queries, row_group, exclude_group, filter_mode, metric, rerank, k,
                                      index_base, out_dist, out_idx, out_group, out_margin, x);
  note_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

// ---- second stage of text_image_search: exact distances of given rows --------------------------
// The reference materialises the text hits as a temporary table and runs the image search inside it
// (src/data/rag.py:118-128). Here the candidate row ids of every query stay on the device and one
// block per query scores them against the image-embedding store: warp per candidate, the fp32
// formulas of the re-rank above, ties -> earlier candidate (= lower row of the temporary table).
constexpr int kMaxRescore = 64;

__global__ void __launch_bounds__(256)
    k3_rescore_rows_kernel(const float* __restrict__ db, int64_t n_rows, int dim, const float* __restrict__ queries,
                           const int64_t* __restrict__ cand_idx, int kc, int metric, int k_out,
                           float* __restrict__ out_dist, int64_t* __restrict__ out_idx) {
  __shared__ uint64_t keys[kMaxRescore], sorted[kMaxRescore];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = blockIdx.x;
  if (tid < kMaxRescore) keys[tid] = kEmptyKey;
  __syncthreads();
  const float4* qv = reinterpret_cast<const float4*>(queries + int64_t(q) * dim);
  const int nv = dim >> 2;
  for (int c = warp; c < kc; c += 8) {
    const int64_t row = cand_idx[int64_t(q) * kc + c];
    if (row < 0 || row >= n_rows) continue;  // warp-uniform
    const float4* dv = reinterpret_cast<const float4*>(db + row * dim);
    float l2 = 0.f, dot = 0.f, qq = 0.f, dd = 0.f;
    for (int i = lane; i < nv; i += 32) {
      const float4 a = qv[i];
      const float4 b = dv[i];
      float t;
      t = a.x - b.x; l2 = fmaf(t, t, l2);
      t = a.y - b.y; l2 = fmaf(t, t, l2);
      t = a.z - b.z; l2 = fmaf(t, t, l2);
      t = a.w - b.w; l2 = fmaf(t, t, l2);
      dot = fmaf(a.x, b.x, dot); dot = fmaf(a.y, b.y, dot);
      dot = fmaf(a.z, b.z, dot); dot = fmaf(a.w, b.w, dot);
      qq = fmaf(a.x, a.x, qq); qq = fmaf(a.y, a.y, qq);
      qq = fmaf(a.z, a.z, qq); qq = fmaf(a.w, a.w, qq);
      dd = fmaf(b.x, b.x, dd); dd = fmaf(b.y, b.y, dd);
      dd = fmaf(b.z, b.z, dd); dd = fmaf(b.w, b.w, dd);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      l2 += __shfl_xor_sync(0xffffffffu, l2, off);
      dot += __shfl_xor_sync(0xffffffffu, dot, off);
      qq += __shfl_xor_sync(0xffffffffu, qq, off);
      dd += __shfl_xor_sync(0xffffffffu, dd, off);
    }
    float dist;
    if (metric == 0) dist = l2;
    else if (metric == 1) dist = 1.f - dot / fmaxf(sqrtf(qq) * sqrtf(dd), 1e-30f);
    else dist = 1.f - dot;
    if (lane == 0) keys[c] = (uint64_t(f32_to_ordered(dist)) << 32) | uint32_t(c);
  }
  rank_sort_smem(keys, sorted, kMaxRescore, tid, 256);
  for (int j = tid; j < k_out; j += 256) {
    const uint64_t key = j < kMaxRescore ? sorted[j] : kEmptyKey;
    const bool ok = key != kEmptyKey;
    out_dist[int64_t(q) * k_out + j] = ok ? ordered_to_f32(uint32_t(key >> 32)) : INFINITY;
    out_idx[int64_t(q) * k_out + j] = ok ? cand_idx[int64_t(q) * kc + int(uint32_t(key))] : -1;
  }
}

cudaError_t launch_k3_rescore_rows(const float* db_f32, int64_t n_rows, int dim, const float* queries, int nq,
                                   const int64_t* cand_idx, int kc, int metric, int k_out, float* out_dist,
                                   int64_t* out_idx, cudaStream_t st) {
  if (kc < 1 || kc > kMaxRescore || k_out < 1 || k_out > kMaxRescore || (dim & 3) != 0) return cudaErrorInvalidValue;
  k3_rescore_rows_kernel<<<nq, 256, 0, st>>>(db_f32, n_rows, dim, queries, cand_idx, kc, metric, k_out, out_dist,
                                             out_idx);
  note_launch();
  return cudaGetLastError();
}

// ---- cross-shard merge ----------------------------------------------------------------------
// Shard g's [nq][k_in] block of each field sits `stride` elements (of that field's type) after
// the base pointer; each shard's list is already sorted by (distance, index) and shards are
// ordered by ascending row range, so "lower slot" == "lower global index" among equal distances.
__global__ void __launch_bounds__(256)
    k3_merge_shards_kernel(const float* __restrict__ cand_dist, const int64_t* __restrict__ cand_idx,
                           const int32_t* __restrict__ cand_group, int64_t stride4, int64_t stride8,
                           int nshards, int nq, int k_in,
                           int k_out, const int32_t* __restrict__ exclude_group, int filter_mode,
                           float* __restrict__ out_dist, int64_t* __restrict__ out_idx,
                           int32_t* __restrict__ out_group) {
  __shared__ uint64_t keys[256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = blockIdx.x;
  const int total = nshards * k_in;  // <= 256
  uint64_t key = kEmptyKey;
  if (tid < total) {
    const int g = tid / k_in, j = tid % k_in;
    const int64_t off = int64_t(q) * k_in + j;
    if (cand_idx[g * stride8 + off] >= 0)
      key = (uint64_t(f32_to_ordered(cand_dist[g * stride4 + off])) << 32) | uint32_t(tid);
  }
  keys[tid] = key;
  bitonic_sort_smem(keys, 256, tid, 256);
  if (warp == 0) {
    const int exclude = (exclude_group != nullptr) ? exclude_group[q] : -1;
    auto entry = [&](int j, bool in_range) {
      Emit e;
      e.valid = false;
      e.dist = INFINITY;
      e.idx = -1;
      e.group = -1;
      if (in_range) {
        const uint64_t kk = keys[j];
        if (kk != kEmptyKey) {
          const int slot = int(uint32_t(kk));
          const int g = slot / k_in, jj = slot % k_in;
          const int64_t off = int64_t(q) * k_in + jj;
          e.valid = true;
          e.dist = cand_dist[g * stride4 + off];
          e.idx = cand_idx[g * stride8 + off];
          e.group = cand_group ? cand_group[g * stride4 + off] : -1;
        }
      }
      return e;
    };
    emit_filtered(entry, min(total, 64), k_out, filter_mode, exclude, out_dist + int64_t(q) * k_out,
                  out_idx + int64_t(q) * k_out,
                  out_group ? out_group + int64_t(q) * k_out : nullptr, lane);
  }
}

cudaError_t launch_k3_merge_shards(const float* cand_dist, const int64_t* cand_idx,
                                   const int32_t* cand_group, int64_t shard_stride_bytes,
                                   int nshards, int nq, int k_in,
                                   int k_out, const int32_t* exclude_group, int filter_mode,
                                   float* out_dist, int64_t* out_idx, int32_t* out_group,
                                   cudaStream_t st) {
  if (nshards * k_in > 256) return cudaErrorInvalidValue;
  const int64_t dense = int64_t(nq) * k_in;
  const int64_t stride4 = shard_stride_bytes ? shard_stride_bytes / 4 : dense;
  const int64_t stride8 = shard_stride_bytes ? shard_stride_bytes / 8 : dense;
  k3_merge_shards_kernel<<<nq, 256, 0, st>>>(cand_dist, cand_idx, cand_group, stride4, stride8,
                                             nshards, nq, k_in,
                                             k_out, exclude_group, filter_mode, out_dist, out_idx,
                                             out_group);
  note_launch();
  return cudaGetLastError();
}

}  // namespace mrag
