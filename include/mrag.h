/*
 * mrag.h — C ABI of the B200-native motion-retrieval hot path (libmrag.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * The reference (MCG-NJU/MotionRAG) is pure Python and has no FFI of its own;
 * each entry point below names the reference call it replaces (file:line relative
 * to the reference root) — the binding a maintainer adds is the ctypes stub shown
 * in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success and a negative mrag_status on error;
 *     mrag_last_error() returns a thread-local human-readable message.
 *   - "dev" pointers are CUDA device pointers on the store's device; "host" pointers
 *     are ordinary (ideally pinned) host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *     No call synchronises the device unless documented (the *_host variants do).
 *   - nothing here ever falls back to a CPU implementation: without a CUDA device
 *     of compute capability 10.x the compute calls fail with MRAG_ERR_DEVICE.
 */
#ifndef MRAG_H_
#define MRAG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRAG_ABI_VERSION 4

#if defined(__GNUC__)
#define MRAG_API __attribute__((visibility("default")))
#else
#define MRAG_API
#endif

typedef enum mrag_status {
  MRAG_OK = 0,
  MRAG_ERR_ARG = -1,         /* bad argument (NULL, size, unsupported dim / k) */
  MRAG_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
  MRAG_ERR_DEVICE = -3,      /* no sm_100-class device: this library has no other path */
  MRAG_ERR_CAPACITY = -4,    /* store or workspace too small */
  MRAG_ERR_UNSUPPORTED = -5  /* valid request this build does not implement */
} mrag_status;

/* distance definitions of LanceDB 0.14 (reference: src/data/rag.py:54 — the
 * reference never sets a metric, so MRAG_METRIC_L2 is what it runs). */
typedef enum mrag_metric {
  MRAG_METRIC_L2 = 0,      /* _distance = sum (q-d)^2            */
  MRAG_METRIC_COSINE = 1,  /* _distance = 1 - q.d / (|q| |d|)    */
  MRAG_METRIC_DOT = 2      /* _distance = 1 - q.d                */
} mrag_metric;

/* which scan kernel serves the query batch */
typedef enum mrag_path {
  MRAG_PATH_AUTO = 0,        /* nq == 1 (and dim in {256,512,768,1024}): STREAM_BF16, else TENSOR_BF16 */
  MRAG_PATH_STREAM_F32 = 1,  /* K1 on the fp32 master rows (ranking exact in fp32, 4 B/elt streamed);
                                one pass over the table per 4 queries */
  MRAG_PATH_STREAM_BF16 = 2, /* K1 on the bf16 shadow rows + fp32 re-rank (2 B/elt streamed) */
  MRAG_PATH_TENSOR_BF16 = 3  /* K2 tcgen05 GEMM with fused epilogue top-k + fp32 re-rank */
} mrag_path;

/* `where video != "<own video>"` (reference: src/data/datamodule.py:235) */
typedef enum mrag_filter {
  MRAG_FILTER_NONE = 0,
  MRAG_FILTER_POST = 1, /* LanceDB 0.14 default: k nearest first, then drop excluded rows (<= k results) */
  MRAG_FILTER_PRE = 2   /* drop excluded rows first, then k nearest: applied inside the scan kernels (rows of
                           the excluded group never enter a candidate list), exact for any group size */
} mrag_filter;

typedef struct mrag_store mrag_store; /* opaque: HBM-resident, row-major, immutable between appends */

typedef struct mrag_store_info {
  int32_t dim;
  int32_t device;
  int64_t n_rows;
  int64_t capacity_rows;
  int32_t has_groups;
  int32_t sm_count;
  const void* rows_f32_dev;  /* [n_rows, dim] float32, L2-normalised when appended with normalise=1 */
  const void* rows_bf16_dev; /* [n_rows, dim] bfloat16 shadow of the same rows */
  const void* groups_dev;    /* [n_rows] int32 group id per row (video identity) or NULL */
  float max_norm_deviation;  /* max over non-zero rows of | |row|^2 - 1 | as stored. l2 searches rank by
                                q.d - |d|^2 / 2 (exact squared-L2 order for ANY rows: the per-row term is added
                                to the scan score whenever this exceeds fp32 normalisation noise or zero rows
                                exist); cosine searches are refused above 1e-3 (they rank by q.d) */
  int32_t reserved;
  int64_t zero_rows;         /* all-zero rows (LanceDB on_bad_vectors='fill'); they score q.d = 0 */
  const void* row_bias_dev;  /* [n_rows] float32: -|row|^2 / 2 of each row as stored */
} mrag_store_info;

typedef struct mrag_search_params {
  int32_t k;           /* results per query, 1..32 (reference: top_k, src/data/rag.py:36) */
  int32_t metric;      /* mrag_metric */
  int32_t path;        /* mrag_path */
  int32_t refine;      /* candidates re-scored in fp32 per query (k..64); 0 = default max(32, k).
                          Plays the role of the reference's refine_factor (src/data/rag.py:37). */
  int32_t filter_mode; /* mrag_filter; needs store groups + exclude_group */
  int32_t list_len;    /* candidates kept per scan run and re-scored in fp32: 0 = default (16 when k <= 12 on
                          the tensor / fp32-stream paths, else 32), or 16 / 32. Deeper lists give a larger
                          exactness margin (the k-th result is compared with the 32nd instead of the 16th best
                          scan score) at about twice the epilogue cost of the tensor path: callers re-issue only
                          the queries whose margin failed with list_len = 32 before falling back to the fp32 scan */
  int64_t index_base;  /* added to local row numbers in out_idx (row-sharded stores) */
  /* optional output, float[nq] (device memory for mrag_search*, host memory for the *_host
   * variants; NULL = not wanted): exactness certificate of the bf16 scan paths.
   *   margin = (exact ranking score of the k-th result - scan score of the weakest re-ranked candidate) / |q|
   * (ranking score = q.d, plus -|d|^2/2 for l2 on rows that are not unit-norm). Every row that was
   * NOT re-ranked has a scan score <= that weakest candidate, and |scan score - exact score| <=
   * eps |q| |d| with eps = 2^-9 (STREAM_BF16: rows rounded) or 2^-8 (TENSOR_BF16: rows and queries
   * rounded). Hence margin > eps * max|d| proves the returned top-k is the exact fp32 top-k;
   * otherwise re-issue that query with MRAG_PATH_STREAM_F32. +inf when every (eligible) row was
   * re-ranked. Row-sharded searches report the margin of the GLOBAL result: the k-th result after
   * the cross-GPU merge against the weakest re-ranked candidate of ANY shard (both travel in the
   * exchange records). NaN only for a query whose peer exchange timed out. */
  float* out_margin;
} mrag_search_params;

MRAG_API int mrag_abi_version(void);
MRAG_API const char* mrag_last_error(void);

/* ---- store: replaces lancedb.connect/open_table (src/data/rag.py:13-14) and the vector
 *      column written by tools/build_rag_database.py:35-50 ------------------------------ */
MRAG_API int mrag_store_create(int32_t dim, int64_t capacity_rows, int32_t device, mrag_store** out);
MRAG_API int mrag_store_destroy(mrag_store* s);
/* append n fp32 rows (host or device memory); normalise != 0 L2-normalises each row first
 * (tools/build_rag_database.py:31-37 stores normalised vectors); also refreshes the bf16 shadow */
MRAG_API int mrag_store_append(mrag_store* s, const float* rows, int64_t n, int32_t rows_on_device,
                      int32_t normalise, void* stream);
/* int32 group id per row for the `video != x` filter; n must equal the current row count */
MRAG_API int mrag_store_set_groups(mrag_store* s, const int32_t* groups, int64_t n, int32_t on_device,
                          void* stream);
MRAG_API int mrag_store_get_info(const mrag_store* s, mrag_store_info* out);

/* ---- search: replaces table.search(vec, col).limit(k)[.where(..)] (src/data/rag.py:54-59) -- */
typedef struct mrag_plan_info {
  int32_t path;            /* resolved mrag_path (never AUTO) */
  int32_t grid;            /* CTAs of the scan kernel (K1 or K2) */
  int32_t cands_per_query; /* candidate keys the scan leaves per query for K3 */
  int32_t rerank;          /* candidates K3 re-scores in fp32 */
  int32_t m_tiles, n_tiles, chunks, tiles_per_chunk; /* K2 tiling (0 for K1) */
  int64_t scan_bytes;      /* algorithmic bytes streamed by the scan kernel: n_rows*dim*elt */
  int64_t scan_flops;      /* algorithmic flops: 2*nq*n_rows*dim */
  size_t workspace_bytes;
  int32_t fused_tail;      /* 1: single-query streaming scan whose last CTA runs the K3 body (one launch) */
  int32_t row_bias;        /* 1: the per-row l2 term is added to the scan scores (rows not unit-norm) */
} mrag_plan_info;
/* validates (store, nq, params) and reports how the call would run, including the workspace
 * the caller must provide to mrag_search */
MRAG_API int mrag_search_plan(const mrag_store* s, int32_t nq, const mrag_search_params* p,
                              mrag_plan_info* out);
/* queries_dev [nq, dim] fp32 (not normalised, as in src/data/datamodule.py:300-302);
 * exclude_group_dev [nq] int32 or NULL; outputs [nq, k]: ascending distance, ties by lowest
 * row index, unused slots = (+inf, -1, -1). out_group_dev may be NULL.
 * Stream semantics: the call is ordered on `stream` like any kernel launch, with one refinement for
 * single-query searches (one launch: scan + select/re-rank tail in its last CTA). That kernel is launched
 * with programmatic stream serialization, so its READ-ONLY scan phase may start while the previous kernel
 * in the stream is still running — in practice the previous search's single-CTA tail, because only
 * libmrag's own kernels release their dependents early; it waits for that kernel to complete before it
 * writes anything (candidates, results, exchange records). Back-to-back searches therefore pipeline
 * (scan of query i+1 under the tail of query i) while every result is still produced in stream order.
 * MRAG_K1_OVERLAP=0 turns this off. */
MRAG_API int mrag_search(const mrag_store* s, const float* queries_dev, int32_t nq,
                const mrag_search_params* p, const int32_t* exclude_group_dev,
                float* out_dist_dev, int64_t* out_idx_dev, int32_t* out_group_dev,
                void* workspace_dev, size_t workspace_bytes, void* stream);
/* ---- row-sharded search with the cross-GPU merge fused into the last kernel -----------------
 * One process per GPU; every rank holds a contiguous row range (params.index_base = its first
 * global row) and an exchange buffer of mrag_exchange_bytes() zero-initialised bytes that all
 * peers have mapped (mrag_ipc_export/open). The last kernel of the search (the K3 block of a query,
 * or the last CTA of the single-query scan) stores its shard's top-k straight into every rank's
 * buffer over NVLink, raises per-query flags and merges world*k candidates as soon as the peers'
 * flags arrive — no collective call. Batches of more than 128 queries publish in K3 and wait +
 * merge in a second small kernel, so that progress never depends on the order in which the
 * hardware dispatches blocks. The flag wait is bounded (timeout_ms): a dead or desynchronised peer
 * yields empty results for the affected queries and an error readable with mrag_store_poll_error,
 * never a hung GPU. All ranks must issue the same sequence of calls (same nq, k) with the same
 * epoch = 1, 2, 3, ...; a rank whose shard is empty still takes part (it publishes no rows). */
typedef struct mrag_exchange {
  int32_t world, rank;     /* world <= 8 */
  int32_t nq_cap, k_cap;   /* capacity the buffers were sized for (k_cap <= 32) */
  uint32_t epoch;          /* call counter, identical on all ranks, starts at 1 */
  int32_t timeout_ms;      /* bound of the flag wait; 0 = default (10 s, env MRAG_XCHG_TIMEOUT_MS) */
  void* const* bufs_dev;   /* device array [world]: exchange-buffer base of every rank */
} mrag_exchange;
MRAG_API size_t mrag_exchange_bytes(int32_t world, int32_t nq_cap, int32_t k_cap);
MRAG_API int mrag_search_sharded(const mrag_store* s, const float* queries_dev, int32_t nq,
                const mrag_search_params* p, const int32_t* exclude_group_dev,
                float* out_dist_dev, int64_t* out_idx_dev, int32_t* out_group_dev,
                void* workspace_dev, size_t workspace_bytes, const mrag_exchange* xchg,
                void* stream);

/* device-side error word of the store (set by a peer exchange that timed out): *code_out = the
 * word (0 = none), which is cleared; returns MRAG_ERR_CUDA with a message when it was set. Only
 * meaningful once the stream the search ran on has been synchronised. */
MRAG_API int mrag_store_poll_error(const mrag_store* s, int32_t* code_out);

/* mrag_search bracketed by CUDA events on `stream`: *scan_ms_out = duration of the scan kernel
 * alone (K1 or K2), *total_ms_out = the whole call on the device. Synchronises the stream.
 * Used by bench.py for the roofline of the dominant kernel. */
MRAG_API int mrag_search_timed(const mrag_store* s, const float* queries_dev, int32_t nq,
                const mrag_search_params* p, const int32_t* exclude_group_dev,
                float* out_dist_dev, int64_t* out_idx_dev, int32_t* out_group_dev,
                void* workspace_dev, size_t workspace_bytes, void* stream,
                float* scan_ms_out, float* total_ms_out);
/* same call with HOST buffers: copies in, searches, copies out, synchronises the stream.
 * Device scratch and a pinned staging block live in the store and are reused; concurrent
 * host-buffer calls on one store serialise on an internal mutex. */
MRAG_API int mrag_search_host(const mrag_store* s, const float* queries_host, int32_t nq,
                     const mrag_search_params* p, const int32_t* exclude_group_host,
                     float* out_dist_host, int64_t* out_idx_host, int32_t* out_group_host,
                     void* stream);

/* the row-sharded search with HOST buffers (small calls replay a captured CUDA graph whose last kernel
 * contains the peer exchange; the epoch travels with the queries) */
MRAG_API int mrag_search_sharded_host(const mrag_store* s, const float* queries_host, int32_t nq,
                     const mrag_search_params* p, const int32_t* exclude_group_host,
                     float* out_dist_host, int64_t* out_idx_host, int32_t* out_group_host,
                     const mrag_exchange* xchg, void* stream);

/* ---- second stage of RAGDatabase.text_image_search (src/data/rag.py:118-128): the reference puts
 *      the text hits into a temporary table and runs the image search inside it. Here: exact
 *      distances (store metric formulas, fp32) of the candidate rows cand_idx_dev [nq, kc] (kc <= 64,
 *      -1 = no candidate) of `s` (the image-embedding store) against queries_dev [nq, dim]; outputs
 *      [nq, k_out] ascending by (distance, position in the candidate list), unused = (+inf, -1). */
MRAG_API int mrag_rescore_rows(const mrag_store* s, const float* queries_dev, int32_t nq,
                      const int64_t* cand_idx_dev, int32_t kc, int32_t metric, int32_t k_out,
                      float* out_dist_dev, int64_t* out_idx_dev, void* stream);

/* ---- cross-shard merge of per-shard results (after the all-gather of [nshards, nq, k]) ----
 * Shard g's [nq, k_in] block of every field starts shard_stride_bytes * g after the field's
 * base pointer (0 = dense [nshards, nq, k_in] arrays); this lets one packed per-rank record
 * {dist, idx, group} travel in a single all-gather. Each shard's list must be sorted by
 * (distance, index) and shards ordered by ascending row range. */
MRAG_API int mrag_merge_topk(const float* cand_dist_dev, const int64_t* cand_idx_dev,
                    const int32_t* cand_group_dev /* may be NULL */, int64_t shard_stride_bytes,
                    int32_t nshards, int32_t nq,
                    int32_t k_in, int32_t k_out, const int32_t* exclude_group_dev,
                    int32_t filter_mode, float* out_dist_dev, int64_t* out_idx_dev,
                    int32_t* out_group_dev /* may be NULL */, void* stream);

/* ---- context gather: replaces get_ref_videos + encode_vision for the K references and the
 *      context assembly x = cat([sos, feats[:, :-1]]) (+pe) (+cond)
 *      (src/data/dataset.py:285-312, src/projects/condition/module.py:264-268, 298-301) ------
 * shard_ptrs_dev: device array of nshards pointers to [rows_per_shard, L, C] feature blocks
 *                 (local or peer-mapped); global row r lives in shard r / rows_per_shard.
 * ref_idx_dev [b, K] int64 in similarity order (0 = most similar), -1 = missing/dropped; an index
 *                 >= n_rows_total (the rows the table really has) also selects the uncond row and is
 *                 never dereferenced.
 * out [b, (K+1)*L, C]; group 0 = sos, group g>=1 = feature of reference rank K-g.
 * dtype: 0 = bfloat16, 1 = float32 (all feature-like tensors share it).
 * pe_dev [(K+1)*L, C] and cond_dev [b, (K+1)*L, C] may be NULL; when given the adds happen in
 * the reference's order and rounding (x + pe, then += cond, each rounded to dtype). */
MRAG_API int mrag_gather_context(const void* const* shard_ptrs_dev, int32_t nshards,
                        int64_t rows_per_shard, const int64_t* ref_idx_dev, const void* sos_dev,
                        const void* uncond_row_dev, const void* pe_dev, const void* cond_dev,
                        void* out_dev, int32_t b, int32_t K, int32_t L, int32_t C, int32_t dtype,
                        int64_t n_rows_total, void* stream);

/* ---- CAMA causal motion transformer forward (consumer of the gathered context) ----
 * torch.nn.TransformerEncoder(num_layers, TransformerEncoderLayer(d_model, nhead, dim_feedforward,
 * dropout=0, activation="gelu", batch_first=True, norm_first=False, bias=True)) with the block-causal
 * mask of ActionTransformer.get_mask (reference configs/cogvideox/MotionRAG_open.yml:253-267,
 * src/projects/condition/module.py:131-135, 303-306). All weights are bf16 device pointers in
 * PyTorch's own layouts (Linear.weight = [out, in]); the caller keeps them alive. */
typedef struct mrag_cama_layer {
  const void *w_qkv, *b_qkv; /* self_attn.in_proj_weight [3d, d], in_proj_bias [3d] */
  const void *w_o, *b_o;     /* self_attn.out_proj.weight [d, d], .bias [d]          */
  const void *w_1, *b_1;     /* linear1.weight [d_ff, d], .bias [d_ff]               */
  const void *w_2, *b_2;     /* linear2.weight [d, d_ff], .bias [d]                  */
  const void *ln1_g, *ln1_b; /* norm1.weight / .bias [d]                             */
  const void *ln2_g, *ln2_b; /* norm2.weight / .bias [d]                             */
} mrag_cama_layer;
typedef struct mrag_cama mrag_cama;
/* head_dim must be 64, d_model in {256,512,768,1024}, d_ff a multiple of 512; tokens = groups *
 * group_tokens (10 * 25 in the reference); workspace for max_batch samples is owned by the handle */
MRAG_API int mrag_cama_create(int32_t n_layers, const mrag_cama_layer* layers, int32_t d_model,
                              int32_t n_heads, int32_t d_ff, int32_t groups, int32_t group_tokens,
                              int32_t max_batch, int32_t device, mrag_cama** out);
MRAG_API int mrag_cama_destroy(mrag_cama* c);
/* device pointers of the handle's input [max_batch, tokens, d_model] bf16 (write the context here —
 * mrag_gather_context can target it directly) and output [max_batch, tokens, d_model] bf16 buffers */
MRAG_API int mrag_cama_io(const mrag_cama* c, void** x_in_dev, void** y_out_dev);
/* runs the n_layers forward for the first b samples of the input buffer on `stream`; with use_graph
 * the 7*n_layers launch chain is captured once per b and replayed as one CUDA graph */
MRAG_API int mrag_cama_forward(mrag_cama* c, int32_t b, int32_t use_graph, void* stream);
/* ActionTransformer.predict (src/projects/condition/module.py:325-326 keeps `[:, -1]`, the last group of the
 * output): the same forward, but the last layer computes attention, out-projection, FFN and LayerNorms for the
 * last group's rows only. *y_last_dev (optional) receives the device pointer of the result
 * [b, group_tokens, d_model] bf16, owned by the handle and valid until the next call. Rows equal those of
 * mrag_cama_forward's output up to the fp32 summation order of the split-K GEMMs (the K split is chosen per
 * row count). */
MRAG_API int mrag_cama_predict(mrag_cama* c, int32_t b, int32_t use_graph, void* stream, void** y_last_dev);
/* the GEMM building block on its own (tests / profiling): C[M,N] = A[M,K] W[N,K]^T (+bias)(gelu) to
 * bf16 (splits > 1: split-K summed inside a thread-block cluster, splits a power of two <= 8), or fp32
 * partial sums [splits][M,N] when out_bf16_dev is NULL; N %% 128 == 0, K %% 64 == 0, splits | K/64 */
MRAG_API int mrag_linear(const void* a_dev, int32_t a_rows_alloc, const void* w_dev, int32_t M, int32_t N,
                         int32_t K, const void* bias_dev, int32_t gelu, void* out_bf16_dev,
                         float* partial_dev, int32_t splits, void* stream);

/* plain cudaMalloc / cudaFree on `device` (IPC-exportable blocks for sharded feature tables) */
MRAG_API int mrag_device_alloc(int32_t device, size_t bytes, void** dev_ptr_out);
MRAG_API int mrag_device_free(int32_t device, void* dev_ptr);

/* ---- peer memory plumbing for row-sharded feature tables (one process per GPU) ------------ */
MRAG_API int mrag_ipc_export(const void* dev_ptr, void* handle_out_64B);
MRAG_API int mrag_ipc_open(const void* handle_64B, void** dev_ptr_out);
MRAG_API int mrag_ipc_close(void* dev_ptr);

/* ---- introspection used by bench.py / tests -------------------------------------------- */
/* name and duration hooks: number of kernels launched by this library on this thread so far */
MRAG_API int64_t mrag_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MRAG_H_ */
