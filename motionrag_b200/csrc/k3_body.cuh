// K3 body — candidate select + exact fp32 re-score + filter (+ cross-GPU exchange), shared by
// the stand-alone K3 kernel (k3_merge.cu, one 1024-thread block per query) and the fused tail of
// the single-query streaming scan (k1_stream.cu: the LAST K1 CTA to finish runs this body, so a
// one-query search is one launch).
//
// It turns the per-CTA / per-chunk candidate keys written by K1 / K2 into the record list the
// reference returns from RAGDatabase.vector_search (src/data/rag.py:54-61): at most k rows,
// ascending `_distance`, computed in fp32 from the master rows with LanceDB's own formulas
// (l2 = sum (q-d)^2, cosine = 1 - cos, dot = 1 - q.d), ties broken by lowest row index, and
// the `video != "<own>"` filter of src/data/datamodule.py:235 applied as a post-filter
// (LanceDB 0.14 default; a pre-filter is applied inside the scan kernels, so the candidate lists
// already hold eligible rows only).
#pragma once
#include "common.cuh"

namespace mrag {

// ---- cross-GPU exchange over peer-mapped memory (row-sharded stores) ---------------------------
// Every rank owns an exchange buffer that all peers have mapped (CUDA IPC over NVLink):
//   records: [slot 2][src rank][query] { idx i64 x k_cap | dist f32 x k_cap | group i32 x k_cap |
//                                        score f32 x k_cap | weakest f32, |q| f32, pad x 2 }
//   flags  : [slot 2][src rank][query] u32 epoch
// The publishing side stores its shard's top-k record into the same (slot, own rank, query) cell
// of EVERY rank's buffer with plain NVLink stores, then raises the matching flags with a
// system-scope release store. The merging side waits (bounded) for the flags of all source ranks
// in its OWN buffer and merges the world * k candidates locally. `score` is the exact ranking
// score (q.d + row bias) of each entry and `weakest` the scan score of the shard's weakest
// re-ranked candidate: together they give the exactness margin of the GLOBAL result.
// Slots alternate with the epoch, so a rank can run at most one call ahead of the slowest peer,
// which cannot still be reading the slot being rewritten.
struct XchgArgs {
  int world, rank;      // world <= 1 disables the exchange
  int nq_cap, k_cap;
  uint32_t epoch;       // used when epoch_dev == nullptr
  const uint32_t* epoch_dev;  // optional: the epoch is read from device memory (graph replays)
  char* const* bufs;    // device array [world]: exchange-buffer base of every rank
  unsigned long long timeout_ns;  // bound of the flag wait
  int* err_word;        // set to 1 when a wait timed out (host-visible, see mrag_store_poll_error)
  int phase;            // 0 = publish + wait + merge, 1 = publish only, 2 = wait + merge only
};
__host__ __device__ inline size_t xchg_rec_bytes(int k_cap) { return size_t(k_cap) * 20 + 16; }
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
__device__ __forceinline__ uint64_t ldcg_u64(const uint64_t* p) {
  return __ldcg(reinterpret_cast<const unsigned long long*>(p));
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int kMaxRerank = 32;  // candidates re-scored per query (= the longest run)
constexpr int kMaxRuns = 1024;
constexpr int kMaxSel = 1024;   // rerank * run length

struct K3Params {
  const uint64_t* cand;  // [nq][n_runs][run_len] keys, each run sorted best-first
  int n_runs, run_len;
  const float* db;       // fp32 master rows
  const float* row_bias; // per-row additive term of the ranking score (l2 on non-unit rows) or null
  int dim;
  const float* queries;
  const int32_t* row_group;
  const int32_t* exclude_group;
  int filter_mode, metric, rerank, k;
  int64_t index_base;
  float* out_dist;
  int64_t* out_idx;
  int32_t* out_group;
  float* out_margin;
  XchgArgs x;
  // profiling only (MRAG_K3_STAMPS=1): globaltimer stamps of the phases of query 0 (see api.cu)
  unsigned long long* stamps;
  // single-query host calls: once the results are written, the value at *done_seq_dev is stored to
  // *done_flag (mapped pinned host memory) so the host can return without waiting for the kernel to retire
  uint32_t* done_flag;
  const uint32_t* done_seq_dev;
};

struct K3Smem {
  uint64_t heads[kMaxRuns];      // run heads (index = run)
  uint64_t sel[kMaxSel];         // keys <= T of the qualifying runs; also the 256-key merge scratch
  uint64_t lists[1024];          // per-warp sorted lists of block_top32 (NT <= 1024 keys)
  uint64_t rr_keys[kMaxRerank];  // (ordered distance << 32) | local row
  float rr_score[kMaxRerank];    // exact ranking score (q.d + bias) of candidate slot c
  uint32_t rr_row[kMaxRerank];   // local row of candidate slot c
  int qual[kMaxRerank];          // runs whose head is <= T
  int64_t rec_idx[32];
  float rec_dist[32];
  int32_t rec_grp[32];
  float rec_score[32];
  float q_norm;
  int n_sel;
  int n_qual;
  int timed_out;
};

// ---- shared tail: filter + emit the first entries of a sorted list (executed by one warp) -----
// entry(j) -> valid?, distance, global index, group, score; list length <= 64, sorted ascending.
struct Emit {
  float dist;
  int64_t idx;
  int32_t group;
  float score;
  bool valid;
};

template <typename EntryFn>
__device__ __forceinline__ void emit_filtered(EntryFn entry, int n_sorted, int k, int filter_mode,
                                              int exclude, float* out_dist, int64_t* out_idx,
                                              int32_t* out_group, float* out_score, int lane) {
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
    if (out_score) out_score[pos0] = e0.score;
  }
  if (keep1 && pos1 < k) {
    out_dist[pos1] = e1.dist;
    out_idx[pos1] = e1.idx;
    if (out_group) out_group[pos1] = e1.group;
    if (out_score) out_score[pos1] = e1.score;
  }
  const int total = min(k, __popc(b0) + __popc(b1));
  for (int j = total + lane; j < k; j += 32) {
    out_dist[j] = INFINITY;
    out_idx[j] = -1;
    if (out_group) out_group[j] = -1;
    if (out_score) out_score[j] = -INFINITY;
  }
}

// ---- exchange: wait for every rank's record of query q, merge world * k candidates, emit ------
// Called by all NT threads of the block. `sm.lists` is reused as scratch.
template <int NT>
__device__ __forceinline__ void k3_exchange_merge(const XchgArgs& x, uint32_t epoch, int q, int k,
                                                  int filter_mode, int exclude, float* out_dist,
                                                  int64_t* out_idx, int32_t* out_group,
                                                  float* out_margin, K3Smem& sm, uint32_t* done_flag = nullptr,
                                                  const uint32_t* done_seq_dev = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = int(epoch & 1u);
  const size_t rec_bytes = xchg_rec_bytes(x.k_cap);
  const size_t flags_off = xchg_flags_offset(x.world, x.nq_cap, x.k_cap);
  if (tid == 0) sm.timed_out = 0;
  __syncthreads();
  if (tid < x.world) {
    const size_t src_cell = (size_t(slot) * x.world + tid) * x.nq_cap + q;
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(x.bufs[x.rank] + flags_off) + src_cell;
    // bounded spin: a dead or desynchronised peer must surface as an error, not as a hung GPU
    const unsigned long long t0 = global_timer_ns();
    uint32_t spins = 0;
    while (ld_acquire_sys(flag) != epoch) {
      if ((++spins & 0xffu) == 0u && global_timer_ns() - t0 > x.timeout_ns) {
        sm.timed_out = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (sm.timed_out) {
    if (tid == 0 && x.err_word != nullptr) atomicExch(x.err_word, 1);
    for (int j = tid; j < k; j += NT) {
      out_dist[j] = INFINITY;
      out_idx[j] = -1;
      if (out_group) out_group[j] = -1;
    }
    if (tid == 0 && out_margin != nullptr) *out_margin = __int_as_float(0x7fc00000);
    __syncthreads();
    if (tid == 0 && done_flag != nullptr) {
      __threadfence_system();
      st_release_sys(done_flag, __ldcg(done_seq_dev));
    }
    return;
  }
  // merge world * k candidates; slot order == global row order among equal distances. Only the k
  // nearest can be returned (k <= 32): the 32 best, sorted, are enough.
  const char* mine = x.bufs[x.rank];
  const int total = x.world * k;  // <= 256
  uint64_t* mk = sm.lists;
  block_top32<NT>(
      [&](int t) {
        const int r = t / k, j = t % k;
        const char* rec = mine + ((size_t(slot) * x.world + r) * x.nq_cap + q) * rec_bytes;
        const int64_t gi = __ldcv(reinterpret_cast<const long long*>(rec) + j);
        const float gd = __ldcv(reinterpret_cast<const float*>(rec + size_t(x.k_cap) * 8) + j);
        return gi >= 0 ? ((uint64_t(f32_to_ordered(gd)) << 32) | uint32_t(t)) : kEmptyKey;
      },
      total, mk);
  if (warp == 0) {
    auto rec_of = [&](int t) {
      return mine + ((size_t(slot) * x.world + t / k) * x.nq_cap + q) * rec_bytes;
    };
    auto gentry = [&](int j, bool in_range) {
      Emit e;
      e.valid = false;
      e.dist = INFINITY;
      e.idx = -1;
      e.group = -1;
      e.score = -INFINITY;
      if (in_range) {
        const uint64_t kk = mk[j];
        if (kk != kEmptyKey) {
          const int t = int(uint32_t(kk));
          const int jj = t % k;
          const char* rec = rec_of(t);
          e.valid = true;
          e.idx = __ldcv(reinterpret_cast<const long long*>(rec) + jj);
          e.dist = __ldcv(reinterpret_cast<const float*>(rec + size_t(x.k_cap) * 8) + jj);
          e.group = __ldcv(reinterpret_cast<const int*>(rec + size_t(x.k_cap) * 12) + jj);
        }
      }
      return e;
    };
    emit_filtered(gentry, min(total, 32), k, filter_mode, exclude, out_dist, out_idx, out_group,
                  nullptr, lane);
    if (out_margin != nullptr) {
      // margin of the GLOBAL result: exact score of the k-th nearest (before the post-filter)
      // against the best row any shard may have left un-re-ranked
      float weakest = -INFINITY;
      float qn = 0.f;
      if (lane < x.world) {
        const char* rec = mine + ((size_t(slot) * x.world + lane) * x.nq_cap + q) * rec_bytes;
        const float* hdr = reinterpret_cast<const float*>(rec + size_t(x.k_cap) * 20);
        weakest = __ldcv(hdr);
        qn = __ldcv(hdr + 1);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        weakest = fmaxf(weakest, __shfl_xor_sync(0xffffffffu, weakest, off));
        qn = fmaxf(qn, __shfl_xor_sync(0xffffffffu, qn, off));
      }
      if (lane == 0) {
        int n_valid = 0;
        while (n_valid < min(min(total, k), 32) && mk[n_valid] != kEmptyKey) ++n_valid;
        float margin = INFINITY;
        if (weakest > -INFINITY && n_valid > 0) {
          const int t = int(uint32_t(mk[n_valid - 1]));
          const float sk = __ldcv(reinterpret_cast<const float*>(rec_of(t) + size_t(x.k_cap) * 16) + t % k);
          margin = (sk - weakest) / fmaxf(qn, 1e-30f);
        }
        *out_margin = margin;
      }
    }
    if (done_flag != nullptr) {
      __threadfence_system();
      __syncwarp();
      if (lane == 0) st_release_sys(done_flag, __ldcg(done_seq_dev));
    }
  }
}

// ---- the K3 body for query q ------------------------------------------------------------------
// Candidates arrive as `n_runs` runs of `run_len` keys, each run sorted best-first (one run per
// K1 CTA / K2 chunk). A full sort of up to 16 K keys is shared-memory-bandwidth bound (~80 us on
// one SM), so selection is done on the run heads instead: with R = rerank, the R-th best key
// overall can be no worse than T = the R-th best run head, hence only keys <= T (in key order)
// can matter, and they all live in the (exactly R, keys are unique) runs whose head is <= T.
// Sort <= 1024 heads -> T -> compact the qualifying keys (<= R * run_len <= 2048) -> sort those.
// Candidate keys are read with ld.global.cg: in the fused form they were written by other CTAs of
// the SAME kernel, so the non-coherent read-only path must not be used for them.
#define K3_STAMP(i)                                                         \
  do {                                                                      \
    if (p.stamps != nullptr && q == 0 && threadIdx.x == 0) p.stamps[i] = global_timer_ns(); \
  } while (0)

template <int NT>
__device__ __forceinline__ void k3_body(const K3Params& p, int q, K3Smem& sm) {
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  K3_STAMP(3);
  const int n_runs = p.n_runs, run_len = p.run_len, rerank = p.rerank, k = p.k;  // rerank <= 32
  const uint64_t* src = p.cand + int64_t(q) * n_runs * run_len;
  const XchgArgs& x = p.x;
  const uint32_t epoch = (x.world > 1 && x.epoch_dev != nullptr) ? __ldcg(x.epoch_dev) : x.epoch;
  const int exclude = (p.exclude_group != nullptr) ? p.exclude_group[q] : -1;
  float* out_dist = p.out_dist + int64_t(q) * k;
  int64_t* out_idx = p.out_idx + int64_t(q) * k;
  int32_t* out_group = p.out_group ? p.out_group + int64_t(q) * k : nullptr;
  float* out_margin = p.out_margin ? p.out_margin + q : nullptr;

  // ---- T = the rerank-th best run head: the 32 best heads, sorted (warp sorts + list merges) ----
  block_top32<NT>(
      [&](int r) {
        const uint64_t h = ldcg_u64(src + int64_t(r) * run_len);
        sm.heads[r] = h;
        return h;
      },
      n_runs, sm.lists);
  // select everything unless there are more runs than needed
  const uint64_t T = (n_runs > rerank) ? sm.lists[rerank - 1] : kEmptyKey;
  if (tid == 0) {
    sm.n_sel = 0;
    sm.n_qual = 0;
  }
  if (tid < kMaxRerank) sm.rr_keys[tid] = kEmptyKey;
  __syncthreads();
  K3_STAMP(4);
  // ---- the (<= rerank) runs whose head is <= T, then their keys <= T: all loads issued at once ----
  for (int r = tid; r < n_runs; r += NT) {
    const uint64_t h = sm.heads[r];
    if (h <= T && uint32_t(h) < uint32_t(kInvalidIdx)) {
      const int slot = atomicAdd(&sm.n_qual, 1);
      if (slot < kMaxRerank) sm.qual[slot] = r;
    }
  }
  __syncthreads();
  const int n_items = min(sm.n_qual, kMaxRerank) * 32;  // (run, position) pairs, <= 1024 by construction
  constexpr int MAXI = (kMaxRerank * 32 + NT - 1) / NT;
  uint64_t ck[MAXI];
#pragma unroll
  for (int u = 0; u < MAXI; ++u) {
    const int t = tid + u * NT;
    ck[u] = (t < n_items && (t & 31) < run_len) ? ldcg_u64(src + int64_t(sm.qual[t >> 5]) * run_len + (t & 31))
                                               : kEmptyKey;
  }
#pragma unroll
  for (int u = 0; u < MAXI; ++u) {
    if (u * NT < n_items) {  // block-uniform: whole warps take part in the ballots
      const bool take = (ck[u] <= T) && (uint32_t(ck[u]) < uint32_t(kInvalidIdx));
      const uint32_t m = __ballot_sync(0xffffffffu, take);
      int base = 0;
      if (lane == 0 && m) base = atomicAdd(&sm.n_sel, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (take) {
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (pos < kMaxSel) sm.sel[pos] = ck[u];
      }
    }
  }
  __syncthreads();
  const int n_sel = min(sm.n_sel, 1024);
  // ---- the best `rerank` of the selected keys, sorted by scan score ----
  block_top32<NT>([&](int i) { return sm.sel[i]; }, n_sel, sm.lists);
  const int n_rr = min(rerank, n_sel);
  K3_STAMP(5);

  // ---- exact fp32 distances of those candidates: a warp scores two rows at a time ----
  const float4* qv = reinterpret_cast<const float4*>(p.queries + int64_t(q) * p.dim);
  const int nv = p.dim >> 2;
  for (int c0 = warp; c0 < n_rr; c0 += 2 * NW) {
    const int c1 = c0 + NW;
    const bool two = c1 < n_rr;
    const uint32_t idx0 = uint32_t(sm.lists[c0]);
    const uint32_t idx1 = two ? uint32_t(sm.lists[c1]) : idx0;
    const float4* d0 = reinterpret_cast<const float4*>(p.db + int64_t(idx0) * p.dim);
    const float4* d1 = reinterpret_cast<const float4*>(p.db + int64_t(idx1) * p.dim);
    float l2a = 0.f, dota = 0.f, dda = 0.f, l2b = 0.f, dotb = 0.f, ddb = 0.f, qq = 0.f;
#pragma unroll 4
    for (int i = lane; i < nv; i += 32) {
      const float4 a = qv[i];
      const float4 b = d0[i];
      const float4 c = d1[i];
      float t;
      t = a.x - b.x; l2a = fmaf(t, t, l2a);
      t = a.y - b.y; l2a = fmaf(t, t, l2a);
      t = a.z - b.z; l2a = fmaf(t, t, l2a);
      t = a.w - b.w; l2a = fmaf(t, t, l2a);
      dota = fmaf(a.x, b.x, dota); dota = fmaf(a.y, b.y, dota);
      dota = fmaf(a.z, b.z, dota); dota = fmaf(a.w, b.w, dota);
      dda = fmaf(b.x, b.x, dda); dda = fmaf(b.y, b.y, dda);
      dda = fmaf(b.z, b.z, dda); dda = fmaf(b.w, b.w, dda);
      t = a.x - c.x; l2b = fmaf(t, t, l2b);
      t = a.y - c.y; l2b = fmaf(t, t, l2b);
      t = a.z - c.z; l2b = fmaf(t, t, l2b);
      t = a.w - c.w; l2b = fmaf(t, t, l2b);
      dotb = fmaf(a.x, c.x, dotb); dotb = fmaf(a.y, c.y, dotb);
      dotb = fmaf(a.z, c.z, dotb); dotb = fmaf(a.w, c.w, dotb);
      ddb = fmaf(c.x, c.x, ddb); ddb = fmaf(c.y, c.y, ddb);
      ddb = fmaf(c.z, c.z, ddb); ddb = fmaf(c.w, c.w, ddb);
      qq = fmaf(a.x, a.x, qq); qq = fmaf(a.y, a.y, qq);
      qq = fmaf(a.z, a.z, qq); qq = fmaf(a.w, a.w, qq);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      l2a += __shfl_xor_sync(0xffffffffu, l2a, off);
      dota += __shfl_xor_sync(0xffffffffu, dota, off);
      dda += __shfl_xor_sync(0xffffffffu, dda, off);
      l2b += __shfl_xor_sync(0xffffffffu, l2b, off);
      dotb += __shfl_xor_sync(0xffffffffu, dotb, off);
      ddb += __shfl_xor_sync(0xffffffffu, ddb, off);
      qq += __shfl_xor_sync(0xffffffffu, qq, off);
    }
    if (lane < 2 && (lane == 0 || two)) {
      const float l2 = lane ? l2b : l2a, dot = lane ? dotb : dota, dd = lane ? ddb : dda;
      const uint32_t idx = lane ? idx1 : idx0;
      const int c = lane ? c1 : c0;
      float dist;
      if (p.metric == 0) dist = l2;
      else if (p.metric == 1) dist = 1.f - dot / fmaxf(sqrtf(qq) * sqrtf(dd), 1e-30f);
      else dist = 1.f - dot;
      sm.rr_keys[c] = (uint64_t(f32_to_ordered(dist)) << 32) | idx;
      sm.rr_score[c] = dot + (p.row_bias != nullptr ? p.row_bias[idx] : 0.f);
      sm.rr_row[c] = idx;
      if (c == 0) sm.q_norm = sqrtf(qq);
    }
  }
  if (n_rr == 0 && tid == 0) sm.q_norm = 0.f;  // empty shard: the merging side takes the max over ranks
  __syncthreads();
  K3_STAMP(6);
  // sort by (distance, row): <= 32 keys, one per lane of warp 0
  if (warp == 0) {
    const uint64_t sorted = warp_sort_u64(sm.rr_keys[lane], lane);
    __syncwarp();
    sm.rr_keys[lane] = sorted;
    __syncwarp();
  }
  if (x.world > 1) __syncthreads();  // (single table: only warp 0 goes on, it owns what it reads)
  K3_STAMP(7);

  // scan score of the weakest re-ranked candidate: every row that was NOT re-ranked scores <= it.
  // With fewer than `rerank` candidates no run was full, so every (eligible) row was a candidate
  // and was re-ranked: nothing is left outside (-inf).
  const float weakest = (n_rr == rerank && n_rr > 0) ? sim_key_score(sm.lists[n_rr - 1]) : -INFINITY;
  auto score_of_row = [&](uint32_t row) {
    float s = -INFINITY;
    for (int c = 0; c < n_rr; ++c)
      if (sm.rr_row[c] == row) s = sm.rr_score[c];
    return s;
  };

  auto entry = [&](int j, bool in_range) {
    Emit e;
    e.valid = false;
    e.dist = INFINITY;
    e.idx = -1;
    e.group = -1;
    e.score = -INFINITY;
    if (in_range && j < kMaxRerank) {
      const uint64_t key = sm.rr_keys[j];
      const uint32_t idx = uint32_t(key);
      if (key != kEmptyKey && idx < uint32_t(kInvalidIdx)) {
        e.valid = true;
        e.dist = ordered_to_f32(uint32_t(key >> 32));
        e.idx = p.index_base + int64_t(idx);
        e.group = (p.row_group != nullptr) ? p.row_group[idx] : -1;
        if (x.world > 1) e.score = score_of_row(idx);
      }
    }
    return e;
  };
  if (x.world <= 1) {
    // exactness certificate of the bf16 scan (see mrag_search_params.out_margin)
    if (out_margin != nullptr && tid == 0) {
      float margin = INFINITY;
      if (weakest > -INFINITY) {
        const int kth = min(k, n_rr) - 1;
        margin = (score_of_row(uint32_t(sm.rr_keys[kth])) - weakest) / fmaxf(sm.q_norm, 1e-30f);
      }
      *out_margin = margin;
    }
    if (warp == 0) {
      emit_filtered(entry, n_rr, k, p.filter_mode, exclude, out_dist, out_idx, out_group, nullptr, lane);
      if (p.done_flag != nullptr) {  // results (written by this warp) first, then the flag
        __threadfence_system();
        __syncwarp();
        if (lane == 0) st_release_sys(p.done_flag, __ldcg(p.done_seq_dev));
      }
    }
    return;
  }

  // ---- row-sharded: publish this shard's top-k to every rank ----
  // (the post-filter belongs after the GLOBAL top-k; pre-filtered lists hold eligible rows only)
  if (warp == 0)
    emit_filtered(entry, n_rr, k, 0, exclude, sm.rec_dist, sm.rec_idx, sm.rec_grp, sm.rec_score, lane);
  __syncthreads();
  const int slot = int(epoch & 1u);
  const size_t rec_bytes = xchg_rec_bytes(x.k_cap);
  const size_t cell = (size_t(slot) * x.world + x.rank) * x.nq_cap + q;  // (slot, src = me, query)
  const size_t flags_off = xchg_flags_offset(x.world, x.nq_cap, x.k_cap);
  for (int t = tid; t < x.world * k; t += NT) {
    const int r = t / k, j = t % k;
    char* rec = x.bufs[r] + cell * rec_bytes;
    reinterpret_cast<int64_t*>(rec)[j] = sm.rec_idx[j];
    reinterpret_cast<float*>(rec + size_t(x.k_cap) * 8)[j] = sm.rec_dist[j];
    reinterpret_cast<int32_t*>(rec + size_t(x.k_cap) * 12)[j] = sm.rec_grp[j];
    reinterpret_cast<float*>(rec + size_t(x.k_cap) * 16)[j] = sm.rec_score[j];
    if (j == 0) {
      float* hdr = reinterpret_cast<float*>(rec + size_t(x.k_cap) * 20);
      hdr[0] = weakest;
      hdr[1] = sm.q_norm;
    }
  }
  // The record stores above are ordered before the flag by the CTA barrier followed by a
  // system-scope RELEASE store (release is cumulative over what happened-before it in this
  // CTA); a separate __threadfence_system() per thread would only add a second fence round trip.
  __syncthreads();
  if (tid < x.world)
    st_release_sys(reinterpret_cast<uint32_t*>(x.bufs[tid] + flags_off) + cell, epoch);
  if (x.phase == 1) return;  // the wait + merge runs as its own kernel (large batches)
  k3_exchange_merge<NT>(x, epoch, q, k, p.filter_mode, exclude, out_dist, out_idx, out_group, out_margin, sm,
                        p.done_flag, p.done_seq_dev);
}

}  // namespace mrag
