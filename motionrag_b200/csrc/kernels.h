// Internal launcher interface between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "k3_body.cuh"

namespace mrag {

// thread-local count of kernels this library launched (mrag_launch_count)
void note_launch(int n = 1);

// K0: store preparation -------------------------------------------------------------------
// rows (fp32, in place when normalise) -> bf16 shadow; one warp per row
// stats (device, 2 x u32, may be null): [0] max ||row|^2 - 1| as float bits, [1] zero-row count
// bias (device, [n] f32, may be null): -|row|^2 / 2 of each row as stored (l2 ranking term)
cudaError_t launch_prepare_rows(float* rows_f32, void* rows_bf16, int64_t n, int dim,
                                bool normalise, unsigned int* stats, float* bias, cudaStream_t st);
cudaError_t launch_cast_queries_bf16(const float* q, void* q_bf16, int nq, int dim,
                                     cudaStream_t st);

// K1: HBM-streaming scan + per-CTA top-KC -----------------------------------------------------
// db: fp32 (elt_bytes 4) or bf16 (elt_bytes 2) rows; queries fp32 [nq<=4][dim]
// cand: [nq][grid][kc] u64 similarity keys. Returns the grid it will use via k1_grid().
// Extras: row_bias (null = none) is added to every score before selection; row_group + exclude_group
// (both non-null) drop the rows of each query's excluded group before selection (pre-filter);
// ticket (single query only, null = off) points at a zero-initialised counter and makes the last
// CTA to finish run the K3 body `k3` itself (fused tail: the search is one launch).
// overlap (host-side flag, fused single-query form only): launch with the programmatic-stream-serialization
// attribute, so that the READ-ONLY scan phase of this search may run while the previous kernel in the stream
// (typically the previous search's single-CTA tail) is still finishing; the kernel waits for its predecessor
// (griddepcontrol.wait) before its first global write.
struct K1Extra {
  const float* row_bias;
  const int32_t* row_group;
  const int32_t* exclude_group;
  int* ticket;
  K3Params k3;
  int overlap;
};
// spare_sm: leave one SM free (grid <= sm_count - 1 in the one-CTA-per-SM single-query form) — the SM on
// which the previous search's tail CTA may still be running when this grid starts
int k1_grid(int64_t n_rows, int elt_bytes, int dim, int nq, int sm_count, bool spare_sm = false);
bool k1_supported(int dim, int nq);
cudaError_t launch_k1_stream(const void* db, int elt_bytes, int64_t n_rows, int dim,
                             const float* queries, int nq, uint64_t* cand, int kc, int grid,
                             const K1Extra& ex, cudaStream_t st);

// K2: tcgen05 batched scan + fused epilogue top-32 ------------------------------------------
struct K2Plan {
  int m_tiles;          // ceil(nq / 128)
  int n_tiles;          // ceil(n_rows / 256)
  int chunks;           // candidate runs per query (per epilogue set) = fixed_per_group + nb
  int tiles_per_chunk;  // longest tile range one list covers
  int grid;             // persistent CTAs
  // work assignment (k2_common.cuh: k2_segment)
  int m_groups, n_units, fixed_per_group, quota, lr_t0, bt, nb;
  int epi_sets;         // epilogue warp sets = candidate runs per (query, chunk)
  int kc;               // candidates kept per run: 16 or 32
};
constexpr int kK2CandMax = 32;  // candidates kept per run: 16 (k <= 12) or 32
// same extras as K1 (bias added in the epilogue, pre-filter applied on the insertion path)
struct K2Extra {
  const float* row_bias;
  const int32_t* row_group;
  const int32_t* exclude_group;
};
bool k2_supported(int dim);
K2Plan k2_plan(int64_t n_rows, int nq, int sm_count);  // caller sets .kc
// q_bf16 [q_rows_padded][dim] (rows >= nq zero), db_bf16 [db_rows_padded][dim] (rows >= n_rows
// zero), both padded to whole tiles; cand [nq][chunks][plan.epi_sets][32] u64 keys; gthr [nq] u32 zeroed
cudaError_t launch_k2_batch(const void* q_bf16, int q_rows_padded, const void* db_bf16,
                            int64_t db_rows_padded, int64_t n_rows, int dim, int nq,
                            const K2Plan& plan, uint64_t* cand, uint32_t* gthr, const K2Extra& ex,
                            cudaStream_t st);

// CTA-pair (cta_group::2) form of K2 for nq > 128: q_rows_padded must cover whole 256-row pairs
K2Plan k2_plan_pair(int64_t n_rows, int nq, int sm_count);
cudaError_t launch_k2_batch_pair(const void* q_bf16, int q_rows_padded, const void* db_bf16,
                                 int64_t db_rows_padded, int64_t n_rows, int dim, int nq,
                                 const K2Plan& plan, uint64_t* cand, uint32_t* gthr, const K2Extra& ex,
                                 cudaStream_t st);

// K3: candidate merge + exact fp32 re-score + filter (body and parameter block: k3_body.cuh) --
size_t exchange_bytes(int world, int nq_cap, int k_cap);
// batches up to this many queries run the cross-GPU exchange inside the K3 blocks (all of them are
// resident at once); larger ones publish in K3 and wait + merge in a second kernel
constexpr int kK3SinglePhaseMax = 128;
bool k3_params_ok(const K3Params& p, int nq);
// one block per query; with p.x.world > 1 the kernel also exchanges per-shard results over peer
// memory and emits the GLOBAL top-k
cudaError_t launch_k3_merge_rerank(const K3Params& p, int nq, cudaStream_t st);
cudaError_t launch_k3_merge_shards(const float* cand_dist, const int64_t* cand_idx,
                                   const int32_t* cand_group, int64_t shard_stride_bytes,
                                   int nshards, int nq, int k_in,
                                   int k_out, const int32_t* exclude_group, int filter_mode,
                                   float* out_dist, int64_t* out_idx, int32_t* out_group,
                                   cudaStream_t st);

// K4: feature gather into the CAMA context layout -------------------------------------------
cudaError_t launch_k4_gather(const void* const* shard_ptrs, int nshards, int64_t rows_per_shard,
                             int64_t n_rows, const int64_t* ref_idx, const void* sos, const void* uncond,
                             const void* pe, const void* cond, void* out, int b, int K, int L,
                             int C, int dtype, cudaStream_t st);

// exact fp32 distances of the given rows (second stage of text_image_search): cand_idx [nq][kc] (-1 = none),
// outputs [nq][k_out] sorted by (distance, candidate position)
cudaError_t launch_k3_rescore_rows(const float* db_f32, int64_t n_rows, int dim, const float* queries, int nq,
                                   const int64_t* cand_idx, int kc, int metric, int k_out, float* out_dist,
                                   int64_t* out_idx, cudaStream_t st);

// K5/K6/K7: CAMA transformer pieces (k5_cama.cu) ------------------------------------------------
// C[M,N] = A[M,K] W[N,K]^T (+bias)(gelu) -> bf16 `out` (splits > 1: the K splits are summed on chip
// inside a thread-block cluster), or fp32 partial sums [splits][M,N] when out is null.
// a_rows_alloc = rows the A buffer really has (TMA zero-fills tile rows beyond it).
cudaError_t launch_k5_linear(const void* a_bf16, int a_rows_alloc, const void* w_bf16, int M, int N, int K,
                             const void* bias, bool gelu, void* out_bf16, float* partial, int splits,
                             cudaStream_t st);
// only_group >= 0: just that group's query rows, output rows written compactly as [b][group_tokens]
cudaError_t launch_k6_attention(const void* qkv, void* out, int b, int T, int d_model, int heads,
                                int groups, int group_tokens, cudaStream_t st, int only_group = -1);
// resid_T > 0: the M rows are the last resid_L tokens of every resid_T-token sample (compact), the
// residual is read from the full-length stream
cudaError_t launch_k7_add_layernorm(const void* resid, const float* partial, int splits, const void* bias,
                                    const void* gamma, const void* beta, void* out, int M, int d, float eps,
                                    cudaStream_t st, int resid_T = 0, int resid_L = 0);

}  // namespace mrag
