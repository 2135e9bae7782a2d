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
#include "k3_body.cuh"
#include "kernels.h"

namespace mrag {

// ---- K0 -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k0_prepare_rows_kernel(float* __restrict__ rows, __nv_bfloat16* __restrict__ shadow, int64_t n,
                           int dim, int normalise, unsigned int* __restrict__ stats,
                           float* __restrict__ bias) {
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
  // l2 ranking score of a row = q.d - |d|^2 / 2 (sum (q-d)^2 = |q|^2 - 2 (q.d - |d|^2 / 2)): the
  // additive term is kept per row so that rows that are not exactly unit-norm (embeddings normalised
  // in bf16, zero-filled bad vectors) rank exactly as LanceDB's squared L2 ranks them
  if (lane == 0 && bias != nullptr) bias[row] = -0.5f * (normalise ? ss * scale * scale : ss);
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
                                bool normalise, unsigned int* stats, float* bias, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const int64_t threads = n * 32;
  const int64_t blocks = (threads + 255) / 256;
  k0_prepare_rows_kernel<<<unsigned(blocks), 256, 0, st>>>(
      rows_f32, static_cast<__nv_bfloat16*>(rows_bf16), n, dim, normalise ? 1 : 0, stats, bias);
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

// ---- K3 kernels (body in k3_body.cuh) ----------------------------------------------------------
// Block size by the number of candidate runs: 1 024 threads (32 warps: one re-rank candidate / one qualifying
// run per warp) when a query has hundreds of runs (K1 with 2-4 queries: one run per CTA), 256 threads for the
// tensor path's few dozen runs when the batch is several waves of blocks — it then keeps 8 blocks per SM in
// flight instead of 2, which is what the latency chain of a block (heads -> keys -> 16-32 master rows -> sort)
// needs to be hidden (4096 queries: p50 of the whole step 4.54 -> 4.44 ms). A batch that fits one wave keeps
// the wide blocks: more warps per query = shorter chain (64 queries: 70 vs 73 us per step).
constexpr int kK3Threads = 1024, kK3ThreadsSmall = 256, kK3SmallMaxRuns = 256, kK3SmallMinQueries = 600;

template <int NT>
__global__ void __launch_bounds__(NT) k3_merge_rerank_kernel(const K3Params p) {
  __shared__ K3Smem sm;
  // launched with programmatic stream serialization: wait here for the scan kernel's results
  asm volatile("griddepcontrol.wait;" ::: "memory");
  k3_body<NT>(p, blockIdx.x, sm);
}

// second phase of the exchange for batches that cannot all be resident at once: every block of the
// publishing kernel has finished before this one starts, so no wait can block a publish
__global__ void __launch_bounds__(256) k3_exchange_merge_kernel(const K3Params p) {
  __shared__ K3Smem sm;
  const int q = blockIdx.x;
  const XchgArgs& x = p.x;
  const uint32_t epoch = x.epoch_dev != nullptr ? __ldcg(x.epoch_dev) : x.epoch;
  const int exclude = (p.exclude_group != nullptr) ? p.exclude_group[q] : -1;
  k3_exchange_merge<256>(x, epoch, q, p.k, p.filter_mode, exclude, p.out_dist + int64_t(q) * p.k,
                         p.out_idx + int64_t(q) * p.k, p.out_group ? p.out_group + int64_t(q) * p.k : nullptr,
                         p.out_margin ? p.out_margin + q : nullptr, sm);
}

size_t exchange_bytes(int world, int nq_cap, int k_cap) { return xchg_total_bytes(world, nq_cap, k_cap); }

bool k3_params_ok(const K3Params& p, int nq) {
  if (p.x.world > 1 &&
      (nq > p.x.nq_cap || p.k > p.x.k_cap || p.x.k_cap > 32 || p.x.world * p.k > 256 || p.x.world > 8))
    return false;
  return p.n_runs >= 1 && p.n_runs <= kMaxRuns && p.run_len >= 1 && p.run_len <= 32 && p.rerank >= 1 &&
         p.rerank <= kMaxRerank && int64_t(p.rerank) * p.run_len <= kMaxSel && p.k >= 1 && p.k <= kMaxRerank;
}

cudaError_t launch_k3_merge_rerank(const K3Params& p_in, int nq, cudaStream_t st) {
  if (!k3_params_ok(p_in, nq)) return cudaErrorInvalidValue;
  K3Params p = p_in;
  // more queries than can be co-resident: publish in this kernel, wait + merge in a second one, so
  // that liveness never depends on the order in which the hardware dispatches blocks
  const bool two_phase = p.x.world > 1 && nq > kK3SinglePhaseMax;
  p.x.phase = two_phase ? 1 : 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(nq));
  const bool small = p.n_runs <= kK3SmallMaxRuns && nq >= kK3SmallMinQueries;
  cfg.blockDim = dim3(small ? kK3ThreadsSmall : kK3Threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = small ? cudaLaunchKernelEx(&cfg, k3_merge_rerank_kernel<kK3ThreadsSmall>, p)
                         : cudaLaunchKernelEx(&cfg, k3_merge_rerank_kernel<kK3Threads>, p);
  note_launch();
  if (le != cudaSuccess) return le;
  if (two_phase) {
    p.x.phase = 2;
    k3_exchange_merge_kernel<<<nq, 256, 0, st>>>(p);
    note_launch();
  }
  return cudaGetLastError();
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
      e.score = -INFINITY;
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
                  out_group ? out_group + int64_t(q) * k_out : nullptr, nullptr, lane);
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
