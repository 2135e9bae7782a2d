// C ABI of libmrag (see include/mrag.h). Host-side orchestration only: argument checks,
// workspace carving and kernel launches on the caller's stream. No CPU compute path exists.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/mrag.h"
#include "kernels.h"

namespace mrag {
static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;
void note_launch(int n) { g_launches += n; }

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  return fail(MRAG_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CK(call)                                          \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

int api_fail(int code, const char* fmt, ...) {  // shared with api_cama.cu
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
}  // namespace mrag

using namespace mrag;

struct mrag_store {
  int32_t dim = 0;
  int32_t device = 0;
  int32_t sm_count = 0;
  int64_t n_rows = 0;
  int64_t capacity = 0;  // rows, multiple of 256
  float* rows_f32 = nullptr;
  void* rows_bf16 = nullptr;
  int32_t* groups = nullptr;
  bool has_groups = false;
  unsigned int* norm_stats = nullptr;  // device: {max ||row|^2-1| (float bits), zero rows}
  float max_norm_dev = 0.f;            // host copy, refreshed by every append
  int64_t zero_rows = 0;
  float* row_bias = nullptr;           // device [capacity]: -|row|^2 / 2 (l2 ranking term of non-unit rows)
  // fused-tail tickets of the single-query scan: zero-initialised counters, re-armed by the kernel
  // itself; calls take them round robin so searches in flight on different streams never share one
  int* tickets = nullptr;
  mutable std::atomic<unsigned> next_ticket{0};
  // device-side errors (a peer exchange that timed out): one word in mapped pinned host memory
  int* err_host = nullptr;
  int* err_dev = nullptr;
  unsigned long long* stamps = nullptr;  // profiling only (MRAG_K3_STAMPS=1): 16 globaltimer stamps
  // scratch of the host-buffer entry point (mrag_search_host): grown on demand, reused
  mutable std::mutex host_mu;
  mutable char* host_dev = nullptr;
  mutable size_t host_dev_bytes = 0;
  mutable char* host_pin = nullptr;   // pinned staging for small transfers
  mutable size_t host_pin_bytes = 0;
  // small host-buffer calls replay a captured CUDA graph (H2D copy -> scan -> K3 -> D2H copy)
  // on a private stream: one launch instead of five submissions
  struct HostGraph {
    int32_t nq, k, metric, path, refine, filter_mode, has_ex, world, rank;
    int64_t index_base, n_rows;
    const void* xbufs;
    cudaGraphExec_t exec;
  };
  mutable std::vector<HostGraph> host_graphs;
  mutable cudaStream_t host_stream = nullptr;
  mutable cudaEvent_t host_event = nullptr;
  mutable bool host_dirty = true;     // work was enqueued on a caller stream that host searches must follow
  mutable uint32_t host_seq = 0;      // sequence number of the done-flag handshake
  void drop_host_graphs() const {
    for (auto& g : host_graphs) cudaGraphExecDestroy(g.exec);
    host_graphs.clear();
  }
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

constexpr int kTickets = 64;

// tuning / debugging knobs: read ONCE (first use), never on the per-call path
struct Knobs {
  bool k2_single;        // MRAG_K2_SINGLE=1: single-CTA K2 instead of the CTA-pair kernel
  bool k2_kc32;          // MRAG_K2_KC=32: long candidate lists in K2
  bool k1_fuse;          // MRAG_K1_FUSE=0: single-query scans launch K3 separately (A/B runs)
  long long xchg_timeout_ms;  // MRAG_XCHG_TIMEOUT_MS: bound of the peer-exchange flag wait
  bool k3_stamps;        // MRAG_K3_STAMPS=1: mrag_search_timed prints the phase timeline of query 0
  bool k1_overlap;       // MRAG_K1_OVERLAP=0: fused single-query scans are launched fully stream-ordered (A/B runs)
};
const Knobs& knobs() {
  static const Knobs k = [] {
    Knobs v;
    const char* e = getenv("MRAG_K2_SINGLE");
    v.k2_single = e && e[0] == '1';
    e = getenv("MRAG_K2_KC");
    v.k2_kc32 = e && atoi(e) == 32;
    e = getenv("MRAG_K1_FUSE");
    v.k1_fuse = !(e && e[0] == '0');
    e = getenv("MRAG_XCHG_TIMEOUT_MS");
    v.xchg_timeout_ms = e ? atoll(e) : 10000;
    if (v.xchg_timeout_ms < 1) v.xchg_timeout_ms = 1;
    e = getenv("MRAG_K3_STAMPS");
    v.k3_stamps = e && e[0] == '1';
    e = getenv("MRAG_K1_OVERLAP");
    v.k1_overlap = !(e && e[0] == '0');
    return v;
  }();
  return k;
}

// resolved execution plan of one search call
struct Plan {
  int path;       // MRAG_PATH_* (never AUTO)
  int kc;         // candidates per CTA (K1) / per chunk (K2)
  int rerank;     // entries re-scored by K3
  int k1_grid = 0;
  K2Plan k2{};
  bool k2_pair = false;  // cta_group::2 kernel
  bool use_bias = false; // l2 on rows that are not exactly unit-norm: rank by q.d - |d|^2 / 2
  bool fused = false;    // single-query streaming scan whose last CTA runs the K3 body
  int q_rows_padded = 0;
  int cands_per_query = 0;
  size_t off_cand = 0, off_qbf16 = 0, off_gthr = 0, total = 0;
};

int make_plan(const mrag_store* s, int32_t nq, const mrag_search_params* p, Plan* out, bool sharded = false) {
  if (!s || !p) return fail(MRAG_ERR_ARG, "null store or params");
  if (nq < 1) return fail(MRAG_ERR_ARG, "nq must be >= 1 (got %d)", nq);
  if (nq > 65536)
    return fail(MRAG_ERR_ARG, "nq must be <= 65536 per call (got %d): split the batch, the scan "
                "cost per query does not improve beyond a few thousand queries", nq);
  if (p->k < 1 || p->k > 32) return fail(MRAG_ERR_ARG, "k must be in 1..32 (got %d)", p->k);
  if (p->metric < 0 || p->metric > 2) return fail(MRAG_ERR_ARG, "unknown metric %d", p->metric);
  if (p->filter_mode < 0 || p->filter_mode > 2)
    return fail(MRAG_ERR_ARG, "unknown filter_mode %d", p->filter_mode);
  if (p->list_len != 0 && p->list_len != 16 && p->list_len != 32)
    return fail(MRAG_ERR_ARG, "list_len must be 0, 16 or 32 (got %d)", p->list_len);
  if (p->list_len == 16 && p->k > 16) return fail(MRAG_ERR_ARG, "list_len 16 cannot serve k = %d", p->k);
  // an empty shard of a row-sharded table still takes part in the exchange (it publishes nothing)
  if (s->n_rows < 1 && !sharded) return fail(MRAG_ERR_ARG, "store is empty");
  // The scan ranks by q.d. Squared L2 orders like q.d - |d|^2 / 2, so for l2 the per-row term is
  // simply added to the scan score (exact for any rows, incl. zero-filled ones). Cosine orders like
  // q.d / |d|: that needs unit rows (normalising on upload does not change a cosine distance).
  if (p->metric == MRAG_METRIC_COSINE && (s->max_norm_dev > 1e-3f || s->zero_rows > 0))
    return fail(MRAG_ERR_UNSUPPORTED,
                "cosine search needs unit-norm rows (max ||d|^2-1| = %.3g, %lld zero rows): append with "
                "normalise=1 (cosine distances do not change) or search with metric l2 / dot",
                s->max_norm_dev, (long long)s->zero_rows);
  int path = p->path;
  // AUTO: scan the bf16 shadow (half the bytes), re-rank the candidates in fp32 from the master
  // rows — the role refine_factor plays in the reference's own call (src/data/rag.py:54)
  // One query streams through K1; from two queries on the CUDA-core FMA rate would cap the bf16
  // stream (4 queries: 0.55 ms vs 0.24 ms measured), while one padded 128-query tensor tile
  // serves up to 128 queries at the HBM rate.
  if (path == MRAG_PATH_AUTO)
    path = (nq == 1 && k1_supported(s->dim, nq)) ? MRAG_PATH_STREAM_BF16 : MRAG_PATH_TENSOR_BF16;
  int refine = p->refine > 0 ? p->refine : 32;
  if (refine < p->k) refine = p->k;
  if (refine > 64) refine = 64;
  Plan pl;
  pl.path = path;
  // 2e-6: what fp32 normalisation leaves (|d|^2 = 1 +- a few ulp) — below the resolution of the fp32
  // distances themselves, so such tables skip the extra per-row load
  pl.use_bias = p->metric == MRAG_METRIC_L2 && (s->max_norm_dev > 2e-6f || s->zero_rows > 0);
  if (path == MRAG_PATH_STREAM_F32 || path == MRAG_PATH_STREAM_BF16) {
    // more than 4 queries are served by successive passes of <= 4 queries each
    if (!k1_supported(s->dim, nq < 4 ? nq : 4))
      return fail(MRAG_ERR_UNSUPPORTED,
                  "streaming path needs dim in {256,512,768,1024} (dim=%d)", s->dim);
    const bool f32 = (path == MRAG_PATH_STREAM_F32);
    if (f32) {
      pl.kc = p->list_len ? p->list_len : ((p->k <= 12) ? 16 : 32);
      pl.rerank = pl.kc;
    } else {
      pl.kc = 32;
      // never re-rank more than a run holds: only then is the re-ranked set exactly the global
      // top-`rerank` by scan score, which the exactness certificate (out_margin) relies on
      pl.rerank = refine > 32 ? 32 : refine;
    }
    pl.fused = nq == 1 && knobs().k1_fuse;
    pl.k1_grid = k1_grid(s->n_rows, f32 ? 4 : 2, s->dim, nq < 4 ? nq : 4, s->sm_count, pl.fused && knobs().k1_overlap);
    pl.cands_per_query = pl.k1_grid * pl.kc;
  } else if (path == MRAG_PATH_TENSOR_BF16) {
    if (!k2_supported(s->dim))
      return fail(MRAG_ERR_UNSUPPORTED, "tensor path needs dim %% 64 == 0 (dim=%d)", s->dim);
    // per-run list length: 16 entries when k <= 12 (the reference asks for K+3 = 12), else 32.
    // The shared per-query bound guarantees the global top-KC by bf16 score, not more, so the
    // fp32 re-rank covers at most KC candidates. MRAG_K2_KC=32 forces the long list.
    pl.kc = p->list_len ? p->list_len : ((p->k <= 12 && !knobs().k2_kc32) ? 16 : 32);
    pl.rerank = refine > pl.kc ? pl.kc : refine;
    // more than one query tile: the CTA-pair kernel (M = 256 per cluster); MRAG_K2_SINGLE=1
    // forces the single-CTA kernel for A/B measurements
    pl.k2_pair = nq > 128 && !knobs().k2_single;
    const int64_t rows = s->n_rows > 0 ? s->n_rows : 1;
    pl.k2 = pl.k2_pair ? k2_plan_pair(rows, nq, s->sm_count) : k2_plan(rows, nq, s->sm_count);
    pl.q_rows_padded = pl.k2_pair ? ((pl.k2.m_tiles + 1) / 2) * 256 : pl.k2.m_tiles * 128;
    pl.k2.kc = pl.kc;
    pl.cands_per_query = pl.k2.chunks * pl.k2.epi_sets * pl.kc;
  } else {
    return fail(MRAG_ERR_ARG, "unknown path %d", p->path);
  }
  if (pl.cands_per_query / pl.kc > 1024)
    return fail(MRAG_ERR_UNSUPPORTED, "too many candidate runs per query (%d)",
                pl.cands_per_query / pl.kc);
  size_t off = 0;
  pl.off_cand = off;
  off += align_up(size_t(nq) * pl.cands_per_query * sizeof(uint64_t), 256);
  pl.off_qbf16 = off;
  if (path == MRAG_PATH_TENSOR_BF16)
    off += align_up(size_t(pl.q_rows_padded) * s->dim * 2, 256);
  pl.off_gthr = off;
  if (path == MRAG_PATH_TENSOR_BF16) off += align_up(size_t(nq) * 4, 256);
  pl.total = off;
  *out = pl;
  return MRAG_OK;
}

}  // namespace

extern "C" {

int mrag_abi_version(void) { return MRAG_ABI_VERSION; }
const char* mrag_last_error(void) { return g_err; }
int64_t mrag_launch_count(void) { return g_launches; }

int mrag_store_create(int32_t dim, int64_t capacity_rows, int32_t device, mrag_store** out) {
  if (!out) return fail(MRAG_ERR_ARG, "out is null");
  *out = nullptr;
  if (dim < 64 || dim % 64 != 0 || dim > 4096)
    return fail(MRAG_ERR_ARG, "dim must be a multiple of 64 in 64..4096 (got %d)", dim);
  if (capacity_rows < 1 || capacity_rows >= 0x7fffff00ll)
    return fail(MRAG_ERR_ARG, "capacity_rows out of range (%lld)", (long long)capacity_rows);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(MRAG_ERR_DEVICE, "no CUDA device visible: libmrag has no CPU path");
  }
  if (device < 0 || device >= ndev) return fail(MRAG_ERR_ARG, "device %d out of range", device);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(MRAG_ERR_DEVICE, "device %d is sm_%d%d; libmrag is built for sm_100a only", device,
                prop.major, prop.minor);
  DeviceGuard g(device);
  if (!g.ok) return fail(MRAG_ERR_CUDA, "cannot select device %d", device);
  mrag_store* s = new (std::nothrow) mrag_store();
  if (!s) return fail(MRAG_ERR_CAPACITY, "host allocation failed");
  s->dim = dim;
  s->device = device;
  s->sm_count = prop.multiProcessorCount;
  s->capacity = (capacity_rows + 255) / 256 * 256;
  const size_t elems = size_t(s->capacity) * dim;
  cudaError_t e = cudaMalloc(&s->rows_f32, elems * 4);
  if (e == cudaSuccess) e = cudaMalloc(&s->rows_bf16, elems * 2);
  if (e == cudaSuccess) e = cudaMalloc(&s->groups, size_t(s->capacity) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&s->row_bias, size_t(s->capacity) * 4);
  if (e == cudaSuccess) e = cudaMemset(s->rows_f32, 0, elems * 4);
  if (e == cudaSuccess) e = cudaMemset(s->rows_bf16, 0, elems * 2);
  if (e == cudaSuccess) e = cudaMemset(s->groups, 0xff, size_t(s->capacity) * 4);
  if (e == cudaSuccess) e = cudaMemset(s->row_bias, 0, size_t(s->capacity) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&s->norm_stats, 8);
  if (e == cudaSuccess) e = cudaMemset(s->norm_stats, 0, 8);
  if (e == cudaSuccess) e = cudaMalloc(&s->tickets, kTickets * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(s->tickets, 0, kTickets * sizeof(int));
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&s->err_host), sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *s->err_host = 0;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&s->err_dev), s->err_host, 0);
  }
  if (e != cudaSuccess) {
    cudaFree(s->rows_f32);
    cudaFree(s->rows_bf16);
    cudaFree(s->groups);
    cudaFree(s->row_bias);
    cudaFree(s->norm_stats);
    cudaFree(s->tickets);
    if (s->err_host) cudaFreeHost(s->err_host);
    delete s;
    return cuda_fail(e, "store allocation");
  }
  *out = s;
  return MRAG_OK;
}

int mrag_store_destroy(mrag_store* s) {
  if (!s) return MRAG_OK;
  DeviceGuard g(s->device);
  cudaFree(s->rows_f32);
  cudaFree(s->rows_bf16);
  cudaFree(s->groups);
  cudaFree(s->row_bias);
  cudaFree(s->tickets);
  cudaFree(s->stamps);
  if (s->err_host) cudaFreeHost(s->err_host);
  s->drop_host_graphs();
  if (s->host_stream) cudaStreamDestroy(s->host_stream);
  if (s->host_event) cudaEventDestroy(s->host_event);
  cudaFree(s->host_dev);
  cudaFreeHost(s->host_pin);
  cudaFree(s->norm_stats);
  delete s;
  return MRAG_OK;
}

int mrag_store_append(mrag_store* s, const float* rows, int64_t n, int32_t rows_on_device,
                      int32_t normalise, void* stream) {
  if (!s || (!rows && n > 0)) return fail(MRAG_ERR_ARG, "null store or rows");
  if (n < 0) return fail(MRAG_ERR_ARG, "negative row count");
  if (s->n_rows + n > s->capacity)
    return fail(MRAG_ERR_CAPACITY, "append of %lld rows exceeds capacity %lld (have %lld)",
                (long long)n, (long long)s->capacity, (long long)s->n_rows);
  if (n == 0) return MRAG_OK;
  DeviceGuard g(s->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* dst = s->rows_f32 + size_t(s->n_rows) * s->dim;
  CK(cudaMemcpyAsync(dst, rows, size_t(n) * s->dim * 4,
                     rows_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  CK(launch_prepare_rows(dst, static_cast<char*>(s->rows_bf16) + size_t(s->n_rows) * s->dim * 2, n,
                         s->dim, normalise != 0, s->norm_stats, s->row_bias + s->n_rows, st));
  unsigned int host_stats[2] = {0, 0};
  CK(cudaMemcpyAsync(host_stats, s->norm_stats, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  memcpy(&s->max_norm_dev, &host_stats[0], 4);
  s->zero_rows = host_stats[1];
  s->n_rows += n;
  {
    std::lock_guard<std::mutex> lock(s->host_mu);
    s->drop_host_graphs();
    s->host_dirty = true;
  }
  return MRAG_OK;
}

int mrag_store_set_groups(mrag_store* s, const int32_t* groups, int64_t n, int32_t on_device,
                          void* stream) {
  if (!s || (!groups && n > 0)) return fail(MRAG_ERR_ARG, "null store or groups");
  if (n != s->n_rows)
    return fail(MRAG_ERR_ARG, "groups length %lld != row count %lld", (long long)n,
                (long long)s->n_rows);
  DeviceGuard g(s->device);
  if (n > 0)
    CK(cudaMemcpyAsync(s->groups, groups, size_t(n) * 4,
                       on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                       static_cast<cudaStream_t>(stream)));
  s->has_groups = true;
  {
    std::lock_guard<std::mutex> lock(s->host_mu);
    s->drop_host_graphs();
    s->host_dirty = true;
  }
  return MRAG_OK;
}

int mrag_store_get_info(const mrag_store* s, mrag_store_info* out) {
  if (!s || !out) return fail(MRAG_ERR_ARG, "null argument");
  out->dim = s->dim;
  out->device = s->device;
  out->n_rows = s->n_rows;
  out->capacity_rows = s->capacity;
  out->has_groups = s->has_groups ? 1 : 0;
  out->sm_count = s->sm_count;
  out->rows_f32_dev = s->rows_f32;
  out->rows_bf16_dev = s->rows_bf16;
  out->groups_dev = s->has_groups ? s->groups : nullptr;
  out->max_norm_deviation = s->max_norm_dev;
  out->zero_rows = s->zero_rows;
  out->row_bias_dev = s->row_bias;
  return MRAG_OK;
}

int mrag_store_poll_error(const mrag_store* s, int32_t* code_out) {
  if (!s || !code_out) return fail(MRAG_ERR_ARG, "null argument");
  // the word lives in mapped pinned memory: kernels that hit a device-side error (a peer exchange
  // whose flags never arrived) set it; valid once the stream that ran the search is synchronised
  const int v = __atomic_exchange_n(s->err_host, 0, __ATOMIC_ACQ_REL);
  *code_out = v;
  if (v != 0)
    return fail(MRAG_ERR_CUDA, "peer exchange timed out: a rank did not publish its shard results within "
                "%lld ms (dead or desynchronised peer); affected queries returned no rows", knobs().xchg_timeout_ms);
  return MRAG_OK;
}

int mrag_search_plan(const mrag_store* s, int32_t nq, const mrag_search_params* p,
                     mrag_plan_info* out) {
  if (!out) return fail(MRAG_ERR_ARG, "out is null");
  Plan pl;
  int rc = make_plan(s, nq, p, &pl);
  if (rc != MRAG_OK) return rc;
  memset(out, 0, sizeof(*out));
  out->path = pl.path;
  out->cands_per_query = pl.cands_per_query;
  out->rerank = pl.rerank;
  const bool tensor = (pl.path == MRAG_PATH_TENSOR_BF16);
  out->grid = tensor ? pl.k2.grid : pl.k1_grid;
  if (tensor) {
    out->m_tiles = pl.k2.m_tiles;
    out->n_tiles = pl.k2.n_tiles;
    out->chunks = pl.k2.chunks;
    out->tiles_per_chunk = pl.k2.tiles_per_chunk;
  }
  out->scan_bytes = s->n_rows * int64_t(s->dim) * (pl.path == MRAG_PATH_STREAM_F32 ? 4 : 2);
  out->scan_flops = 2ll * nq * s->n_rows * s->dim;
  out->workspace_bytes = pl.total;
  out->fused_tail = pl.fused ? 1 : 0;
  out->row_bias = pl.use_bias ? 1 : 0;
  return MRAG_OK;
}

// epoch_dev: device address holding the exchange epoch (graph replays), or null = xchg->epoch
static int search_impl(const mrag_store* s, const float* queries_dev, int32_t nq,
                       const mrag_search_params* p, const int32_t* exclude_group_dev,
                       float* out_dist_dev, int64_t* out_idx_dev, int32_t* out_group_dev,
                       void* workspace_dev, size_t workspace_bytes, void* stream,
                       cudaEvent_t before_scan, cudaEvent_t after_scan,
                       const mrag_exchange* xchg = nullptr, const uint32_t* epoch_dev = nullptr,
                       bool in_host_graph = false, uint32_t* done_flag = nullptr,
                       const uint32_t* done_seq_dev = nullptr) {
  const bool sharded = xchg != nullptr && xchg->world > 1;
  Plan pl;
  int rc = make_plan(s, nq, p, &pl, sharded);
  if (rc != MRAG_OK) return rc;
  if (!queries_dev || !out_dist_dev || !out_idx_dev || !workspace_dev)
    return fail(MRAG_ERR_ARG, "null device buffer");
  if (workspace_bytes < pl.total)
    return fail(MRAG_ERR_CAPACITY, "workspace too small: %zu < %zu", workspace_bytes, pl.total);
  if (p->filter_mode != MRAG_FILTER_NONE && exclude_group_dev && !s->has_groups)
    return fail(MRAG_ERR_ARG, "filter requested but the store has no group ids");
  DeviceGuard g(s->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace_dev);
  uint64_t* cand = reinterpret_cast<uint64_t*>(ws + pl.off_cand);
  const int32_t* groups = s->has_groups ? s->groups : nullptr;
  const int fm = (exclude_group_dev && groups) ? p->filter_mode : MRAG_FILTER_NONE;
  const float* bias = pl.use_bias ? s->row_bias : nullptr;

  // ---- parameter block of the K3 body (stand-alone kernel or fused tail of K1) ----
  K3Params kp{};
  kp.cand = cand;
  kp.n_runs = pl.cands_per_query / pl.kc;
  kp.run_len = pl.kc;
  kp.db = s->rows_f32;
  kp.row_bias = bias;
  kp.dim = s->dim;
  kp.queries = queries_dev;
  kp.row_group = groups;
  kp.exclude_group = exclude_group_dev;
  kp.filter_mode = fm;
  kp.metric = p->metric;
  kp.rerank = pl.rerank;
  kp.k = p->k;
  kp.index_base = p->index_base;
  kp.out_dist = out_dist_dev;
  kp.out_idx = out_idx_dev;
  kp.out_group = out_group_dev;
  kp.out_margin = p->out_margin;
  if (nq == 1) {  // one block finishes the search: it can tell the host directly
    kp.done_flag = done_flag;
    kp.done_seq_dev = done_seq_dev;
  }
  if (sharded) {
    if (!xchg->bufs_dev || xchg->rank < 0 || xchg->rank >= xchg->world || xchg->world > 8 ||
        nq > xchg->nq_cap || p->k > xchg->k_cap || xchg->k_cap > 32 || (xchg->epoch == 0 && !epoch_dev))
      return fail(MRAG_ERR_ARG, "bad exchange descriptor (world %d rank %d nq %d/%d k %d/%d epoch %u)",
                  xchg->world, xchg->rank, nq, xchg->nq_cap, p->k, xchg->k_cap, xchg->epoch);
    kp.x.world = xchg->world;
    kp.x.rank = xchg->rank;
    kp.x.nq_cap = xchg->nq_cap;
    kp.x.k_cap = xchg->k_cap;
    kp.x.epoch = xchg->epoch;
    kp.x.epoch_dev = epoch_dev;
    kp.x.bufs = reinterpret_cast<char* const*>(xchg->bufs_dev);
    const long long ms = xchg->timeout_ms > 0 ? xchg->timeout_ms : knobs().xchg_timeout_ms;
    kp.x.timeout_ns = static_cast<unsigned long long>(ms) * 1000000ull;
    kp.x.err_word = s->err_dev;
  }
  if (knobs().k3_stamps && before_scan != nullptr) {  // only the timed entry point profiles
    mrag_store* ms = const_cast<mrag_store*>(s);
    if (!ms->stamps) CK(cudaMalloc(&ms->stamps, 16 * 8));
    CK(cudaMemsetAsync(ms->stamps, 0, 16 * 8, st));
    CK(cudaMemsetAsync(ms->stamps, 0xff, 8, st));  // [0] is an atomicMin target
    kp.stamps = ms->stamps;
  }
  if (!k3_params_ok(kp, nq)) return fail(MRAG_ERR_ARG, "unsupported K3 shape (runs %d x %d, rerank %d, k %d)",
                                         kp.n_runs, kp.run_len, kp.rerank, kp.k);
  // pre-filter: applied inside the scan (rows of the excluded group never enter a list)
  const int32_t* pre_groups = fm == MRAG_FILTER_PRE ? groups : nullptr;
  const int32_t* pre_excl = fm == MRAG_FILTER_PRE ? exclude_group_dev : nullptr;

  if (pl.path != MRAG_PATH_TENSOR_BF16 && before_scan) CK(cudaEventRecord(before_scan, st));
  if (pl.path == MRAG_PATH_STREAM_F32 || pl.path == MRAG_PATH_STREAM_BF16) {
    const bool f32 = pl.path == MRAG_PATH_STREAM_F32;
    const void* rows = f32 ? static_cast<const void*>(s->rows_f32) : s->rows_bf16;
    for (int q0 = 0; q0 < nq; q0 += 4) {  // one pass over the table per group of <= 4 queries
      const int nqg = nq - q0 < 4 ? nq - q0 : 4;
      K1Extra ex{};
      ex.row_bias = bias;
      ex.row_group = pre_groups;
      ex.exclude_group = pre_excl ? pre_excl + q0 : nullptr;
      if (pl.fused) {
        // host-buffer calls (captured graphs) are serialised per store and share the last counter
        ex.ticket = s->tickets + (in_host_graph ? kTickets - 1
                                                : s->next_ticket.fetch_add(1, std::memory_order_relaxed) % (kTickets - 1));
        ex.k3 = kp;
        ex.overlap = knobs().k1_overlap ? 1 : 0;
      }
      // a 1..3-query tail uses the kernel instantiated for that count but is launched with the
      // plan's grid, so every query shares one candidate layout ([nq][grid][kc])
      CK(launch_k1_stream(rows, f32 ? 4 : 2, s->n_rows, s->dim, queries_dev + size_t(q0) * s->dim, nqg,
                          cand + size_t(q0) * pl.cands_per_query, pl.kc, pl.k1_grid, ex, st));
    }
  } else {
    void* qb = ws + pl.off_qbf16;
    const size_t qb_bytes = size_t(pl.q_rows_padded) * s->dim * 2;
    if (nq != pl.q_rows_padded) CK(cudaMemsetAsync(qb, 0, qb_bytes, st));
    CK(launch_cast_queries_bf16(queries_dev, qb, nq, s->dim, st));
    // rows beyond n_rows up to the 256-row tile edge are zero (store capacity is padded)
    const int64_t rows_padded = (s->n_rows + 255) / 256 * 256;
    uint32_t* gthr = reinterpret_cast<uint32_t*>(ws + pl.off_gthr);
    CK(cudaMemsetAsync(gthr, 0, size_t(nq) * 4, st));
    if (before_scan) CK(cudaEventRecord(before_scan, st));
    K2Extra ex{bias, pre_groups, pre_excl};
    cudaError_t e = pl.k2_pair
                        ? launch_k2_batch_pair(qb, pl.q_rows_padded, s->rows_bf16, rows_padded > 0 ? rows_padded : 256,
                                               s->n_rows, s->dim, nq, pl.k2, cand, gthr, ex, st)
                        : launch_k2_batch(qb, pl.q_rows_padded, s->rows_bf16, rows_padded > 0 ? rows_padded : 256,
                                          s->n_rows, s->dim, nq, pl.k2, cand, gthr, ex, st);
    if (e != cudaSuccess) return cuda_fail(e, "launch_k2_batch");
  }
  if (after_scan) CK(cudaEventRecord(after_scan, st));
  if (!pl.fused) CK(launch_k3_merge_rerank(kp, nq, st));
  return MRAG_OK;
}

int mrag_search(const mrag_store* s, const float* queries_dev, int32_t nq,
                const mrag_search_params* p, const int32_t* exclude_group_dev, float* out_dist_dev,
                int64_t* out_idx_dev, int32_t* out_group_dev, void* workspace_dev,
                size_t workspace_bytes, void* stream) {
  return search_impl(s, queries_dev, nq, p, exclude_group_dev, out_dist_dev, out_idx_dev,
                     out_group_dev, workspace_dev, workspace_bytes, stream, nullptr, nullptr);
}

int mrag_search_sharded(const mrag_store* s, const float* queries_dev, int32_t nq,
                        const mrag_search_params* p, const int32_t* exclude_group_dev,
                        float* out_dist_dev, int64_t* out_idx_dev, int32_t* out_group_dev,
                        void* workspace_dev, size_t workspace_bytes, const mrag_exchange* xchg,
                        void* stream) {
  if (!xchg) return fail(MRAG_ERR_ARG, "null exchange descriptor");
  return search_impl(s, queries_dev, nq, p, exclude_group_dev, out_dist_dev, out_idx_dev,
                     out_group_dev, workspace_dev, workspace_bytes, stream, nullptr, nullptr, xchg);
}

size_t mrag_exchange_bytes(int32_t world, int32_t nq_cap, int32_t k_cap) {
  if (world < 1 || world > 8 || nq_cap < 1 || k_cap < 1 || k_cap > 32) return 0;
  return exchange_bytes(world, nq_cap, k_cap);
}

int mrag_search_timed(const mrag_store* s, const float* queries_dev, int32_t nq,
                      const mrag_search_params* p, const int32_t* exclude_group_dev,
                      float* out_dist_dev, int64_t* out_idx_dev, int32_t* out_group_dev,
                      void* workspace_dev, size_t workspace_bytes, void* stream,
                      float* scan_ms_out, float* total_ms_out) {
  if (!s) return fail(MRAG_ERR_ARG, "null store");
  DeviceGuard g(s->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1, e2, e3;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventCreate(&e2));
  CK(cudaEventCreate(&e3));
  cudaEventRecord(e0, st);
  int rc = search_impl(s, queries_dev, nq, p, exclude_group_dev, out_dist_dev, out_idx_dev,
                       out_group_dev, workspace_dev, workspace_bytes, stream, e1, e2);
  cudaEventRecord(e3, st);
  cudaError_t e = cudaEventSynchronize(e3);
  if (rc == MRAG_OK && e != cudaSuccess) rc = cuda_fail(e, "event synchronize");
  if (rc == MRAG_OK) {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, e1, e2);
    cudaEventElapsedTime(&b, e0, e3);
    if (scan_ms_out) *scan_ms_out = a;
    if (total_ms_out) *total_ms_out = b;
    if (knobs().k3_stamps && s->stamps) {
      unsigned long long h[16];
      cudaMemcpy(h, s->stamps, sizeof(h), cudaMemcpyDeviceToHost);
      auto us = [&](int i) { return h[i] ? double(h[i] - h[0]) / 1e3 : -1.0; };
      fprintf(stderr, "[k3 stamps us from first CTA start] last CTA done streaming %.1f | tail enters %.1f | k3 start %.1f | "
              "heads+T %.1f | compact+sort %.1f | rerank %.1f | sorted %.1f | tail done %.1f | events: scan %.1f total %.1f\n",
              us(1), us(2), us(3), us(4), us(5), us(6), us(7), us(8), a * 1e3, b * 1e3);
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaEventDestroy(e2);
  cudaEventDestroy(e3);
  return rc;
}

// host-buffer search, optionally row-sharded (xchg != null: the peer exchange runs inside the graph)
static int search_host_impl(const mrag_store* s, const float* queries_host, int32_t nq,
                            const mrag_search_params* p, const int32_t* exclude_group_host,
                            float* out_dist_host, int64_t* out_idx_host, int32_t* out_group_host,
                            const mrag_exchange* xchg, void* stream) {
  const bool sharded = xchg != nullptr && xchg->world > 1;
  Plan pl;
  int rc = make_plan(s, nq, p, &pl, sharded);
  if (rc != MRAG_OK) return rc;
  if (!queries_host || !out_dist_host || !out_idx_host)
    return fail(MRAG_ERR_ARG, "null host buffer");
  DeviceGuard g(s->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(s->host_mu);  // one host-buffer call at a time per store
  // device block: [queries | epoch (16 B) | exclude] (one H2D) [dist | idx | group | margin] (one D2H) [workspace]
  const size_t q_raw = size_t(nq) * s->dim * 4, ex_raw = size_t(nq) * 4;
  const size_t ex_off = q_raw + 16;
  const size_t in_bytes = align_up(ex_off + ex_raw, 256);
  const size_t nk = size_t(nq) * p->k;
  const bool want_margin = p->out_margin != nullptr;
  const size_t out_raw = nk * 16 + size_t(nq) * 4;  // idx i64 | dist f32 | group i32 | margin f32
  const size_t out_bytes = align_up(out_raw, 256);
  const size_t total = in_bytes + out_bytes + pl.total;
  if (s->host_dev_bytes < total) {
    s->drop_host_graphs();
    cudaFree(s->host_dev);
    s->host_dev = nullptr;
    s->host_dev_bytes = 0;
    CK(cudaMalloc(reinterpret_cast<void**>(&s->host_dev), total));
    s->host_dev_bytes = total;
  }
  char* base = s->host_dev;
  float* q_d = reinterpret_cast<float*>(base);
  uint32_t* epoch_d = reinterpret_cast<uint32_t*>(base + q_raw);
  int32_t* ex_d = reinterpret_cast<int32_t*>(base + ex_off);
  int64_t* oi_d = reinterpret_cast<int64_t*>(base + in_bytes);
  float* od_d = reinterpret_cast<float*>(base + in_bytes + nk * 8);
  int32_t* og_d = reinterpret_cast<int32_t*>(base + in_bytes + nk * 12);
  float* om_d = reinterpret_cast<float*>(base + in_bytes + nk * 16);
  void* ws = base + in_bytes + out_bytes;
  mrag_search_params pd = *p;   // device-side view of the parameters (margin pointer swapped)
  // small transfers go through pinned staging (truly asynchronous, one copy each way)
  const bool staged = (ex_off + ex_raw) <= (256u << 10) && out_raw <= (256u << 10) - 64;
  if (staged && s->host_pin_bytes < (512u << 10)) {
    cudaFreeHost(s->host_pin);
    s->host_pin = nullptr;
    s->host_pin_bytes = 0;
    CK(cudaMallocHost(reinterpret_cast<void**>(&s->host_pin), 512u << 10));
    memset(s->host_pin, 0, 512u << 10);   // (the done flag starts at 0; sequence numbers start at 1)
    s->host_pin_bytes = 512u << 10;
  }
  cudaError_t e;
  if (staged) {
    // ---- graph path: fixed staging + fixed scratch make the whole call a static graph ----
    if (!s->host_stream) CK(cudaStreamCreateWithFlags(&s->host_stream, cudaStreamNonBlocking));
    if (!s->host_event) CK(cudaEventCreateWithFlags(&s->host_event, cudaEventDisableTiming));
    cudaStream_t hs = s->host_stream;
    memcpy(s->host_pin, queries_host, q_raw);
    // the exchange epoch changes with every call: it travels with the queries and the kernels read
    // it from device memory, so the captured graph stays valid
    uint32_t hdr[2] = {sharded ? xchg->epoch : 0u, ++s->host_seq};
    if (s->host_seq == 0) hdr[1] = ++s->host_seq;   // 0 is the flag's initial value
    memcpy(s->host_pin + q_raw, hdr, 8);
    if (exclude_group_host) memcpy(s->host_pin + ex_off, exclude_group_host, ex_raw);
    char* pin_out = s->host_pin + (256u << 10);
    // done flag: last 64 bytes of the staging block (results need at most 256 KB - 64: checked by `staged`)
    volatile uint32_t* done_flag = reinterpret_cast<volatile uint32_t*>(s->host_pin + (512u << 10) - 64);
    if (s->host_dirty) {   // order after work the caller enqueued on its stream (appends, group uploads)
      CK(cudaEventRecord(s->host_event, st));
      CK(cudaStreamWaitEvent(hs, s->host_event, 0));
      s->host_dirty = false;
    }
    const int has_ex = (exclude_group_host ? 1 : 0) | (want_margin ? 2 : 0);
    pd.out_margin = want_margin ? reinterpret_cast<float*>(pin_out + nk * 16) : nullptr;
    cudaGraphExec_t exec = nullptr;
    const int world = sharded ? xchg->world : 1, rank = sharded ? xchg->rank : 0;
    const void* xbufs = sharded ? static_cast<const void*>(xchg->bufs_dev) : nullptr;
    // (the graph bakes in nk-dependent result offsets: nq and k are part of the key)
    for (auto& hg : s->host_graphs)
      if (hg.nq == nq && hg.k == p->k && hg.metric == p->metric && hg.path == p->path &&
          hg.refine == p->refine + 1000 * p->list_len && hg.filter_mode == p->filter_mode && hg.has_ex == has_ex &&
          hg.index_base == p->index_base && hg.n_rows == s->n_rows && hg.world == world &&
          hg.rank == rank && hg.xbufs == xbufs)
        exec = hg.exec;
    if (!exec) {
      CK(cudaStreamBeginCapture(hs, cudaStreamCaptureModeThreadLocal));
      e = cudaMemcpyAsync(base, s->host_pin, exclude_group_host ? ex_off + ex_raw : q_raw + 16,
                          cudaMemcpyHostToDevice, hs);
      int crc = MRAG_OK;
      // results are written by K3 straight into the pinned block (zero-copy stores over PCIe):
      // no device->host copy node, the graph ends with the kernel
      if (e == cudaSuccess)
        crc = search_impl(s, q_d, nq, &pd, exclude_group_host ? ex_d : nullptr,
                          reinterpret_cast<float*>(pin_out + nk * 8),
                          reinterpret_cast<int64_t*>(pin_out),
                          reinterpret_cast<int32_t*>(pin_out + nk * 12), ws, pl.total, hs, nullptr,
                          nullptr, sharded ? xchg : nullptr, sharded ? epoch_d : nullptr, true,
                          const_cast<uint32_t*>(done_flag), epoch_d + 1);
      cudaGraph_t graph = nullptr;
      cudaError_t e2 = cudaStreamEndCapture(hs, &graph);
      if (crc != MRAG_OK) {
        if (graph) cudaGraphDestroy(graph);
        return crc;
      }
      if (e != cudaSuccess || e2 != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        return cuda_fail(e != cudaSuccess ? e : e2, "graph capture of the host search");
      }
      e = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
      if (s->host_graphs.size() >= 64) s->drop_host_graphs();
      s->host_graphs.push_back({nq, p->k, p->metric, p->path, p->refine + 1000 * p->list_len, p->filter_mode, has_ex, world, rank,
                                p->index_base, s->n_rows, xbufs, exec});
    }
    e = cudaGraphLaunch(exec, hs);
    if (e == cudaSuccess && nq == 1) {
      // single query: the finishing block stores the call's sequence number to the done flag right after
      // the results (system-scope release): return as soon as it shows up instead of waiting for the
      // kernel to retire and the driver to notice (~5 us); stream order still protects the next call
      const uint32_t want = hdr[1];
      uint32_t spins = 0;
      while (*done_flag != want) {
        if ((++spins & 0x3ffu) == 0u && cudaStreamQuery(hs) != cudaErrorNotReady) break;  // finished or failed
      }
      if (*done_flag != want) e = cudaStreamSynchronize(hs);
      else std::atomic_thread_fence(std::memory_order_acquire);
    } else if (e == cudaSuccess) {
      e = cudaStreamSynchronize(hs);
    }
    if (e != cudaSuccess) return cuda_fail(e, "graph launch of the host search");
    // kernels replayed by the graph
    note_launch(pl.path == MRAG_PATH_TENSOR_BF16 ? (sharded && nq > kK3SinglePhaseMax ? 4 : 3) : (pl.fused ? 1 : 2));
    memcpy(out_idx_host, pin_out, nk * 8);
    memcpy(out_dist_host, pin_out + nk * 8, nk * 4);
    if (out_group_host) memcpy(out_group_host, pin_out + nk * 12, nk * 4);
    if (want_margin) memcpy(p->out_margin, pin_out + nk * 16, size_t(nq) * 4);
    if (sharded && *s->err_host != 0) {
      int32_t code = 0;
      return mrag_store_poll_error(s, &code);
    }
    return MRAG_OK;
  }
  pd.out_margin = want_margin ? om_d : nullptr;
  e = cudaMemcpyAsync(q_d, queries_host, q_raw, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && exclude_group_host)
    e = cudaMemcpyAsync(ex_d, exclude_group_host, ex_raw, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(e, "host->device copy");
  rc = search_impl(s, q_d, nq, &pd, exclude_group_host ? ex_d : nullptr, od_d, oi_d, og_d, ws, pl.total,
                   stream, nullptr, nullptr, sharded ? xchg : nullptr, nullptr);
  if (rc != MRAG_OK) return rc;
  e = cudaMemcpyAsync(out_idx_host, oi_d, nk * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && want_margin)
    e = cudaMemcpyAsync(p->out_margin, om_d, size_t(nq) * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist_host, od_d, nk * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && out_group_host)
    e = cudaMemcpyAsync(out_group_host, og_d, nk * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "device->host copy");
  if (sharded && *s->err_host != 0) {
    int32_t code = 0;
    return mrag_store_poll_error(s, &code);
  }
  return MRAG_OK;
}

int mrag_search_host(const mrag_store* s, const float* queries_host, int32_t nq,
                     const mrag_search_params* p, const int32_t* exclude_group_host,
                     float* out_dist_host, int64_t* out_idx_host, int32_t* out_group_host,
                     void* stream) {
  return search_host_impl(s, queries_host, nq, p, exclude_group_host, out_dist_host, out_idx_host,
                          out_group_host, nullptr, stream);
}

int mrag_search_sharded_host(const mrag_store* s, const float* queries_host, int32_t nq,
                             const mrag_search_params* p, const int32_t* exclude_group_host,
                             float* out_dist_host, int64_t* out_idx_host, int32_t* out_group_host,
                             const mrag_exchange* xchg, void* stream) {
  if (!xchg) return fail(MRAG_ERR_ARG, "null exchange descriptor");
  return search_host_impl(s, queries_host, nq, p, exclude_group_host, out_dist_host, out_idx_host,
                          out_group_host, xchg, stream);
}

int mrag_rescore_rows(const mrag_store* s, const float* queries_dev, int32_t nq, const int64_t* cand_idx_dev,
                      int32_t kc, int32_t metric, int32_t k_out, float* out_dist_dev, int64_t* out_idx_dev,
                      void* stream) {
  if (!s || !queries_dev || !cand_idx_dev || !out_dist_dev || !out_idx_dev)
    return fail(MRAG_ERR_ARG, "null store or device buffer");
  if (nq < 1 || kc < 1 || kc > 64 || k_out < 1 || k_out > 64 || metric < 0 || metric > 2)
    return fail(MRAG_ERR_ARG, "need nq >= 1, 1 <= kc <= 64, 1 <= k_out <= 64 (got nq %d kc %d k_out %d)", nq, kc,
                k_out);
  if (s->n_rows < 1) return fail(MRAG_ERR_ARG, "store is empty");
  DeviceGuard guard(s->device);
  CK(launch_k3_rescore_rows(s->rows_f32, s->n_rows, s->dim, queries_dev, nq, cand_idx_dev, kc, metric, k_out,
                            out_dist_dev, out_idx_dev, static_cast<cudaStream_t>(stream)));
  return MRAG_OK;
}

int mrag_merge_topk(const float* cand_dist_dev, const int64_t* cand_idx_dev,
                    const int32_t* cand_group_dev, int64_t shard_stride_bytes, int32_t nshards,
                    int32_t nq, int32_t k_in,
                    int32_t k_out, const int32_t* exclude_group_dev, int32_t filter_mode,
                    float* out_dist_dev, int64_t* out_idx_dev, int32_t* out_group_dev,
                    void* stream) {
  if (!cand_dist_dev || !cand_idx_dev || !out_dist_dev || !out_idx_dev)
    return fail(MRAG_ERR_ARG, "null device buffer");
  if (nshards < 1 || nq < 1 || k_in < 1 || k_out < 1 || k_out > 32 || nshards * k_in > 256)
    return fail(MRAG_ERR_ARG, "need 1 <= k_out <= 32 and nshards*k_in <= 256 (got %d x %d -> %d)",
                nshards, k_in, k_out);
  if (filter_mode != MRAG_FILTER_NONE && exclude_group_dev && !cand_group_dev)
    return fail(MRAG_ERR_ARG, "filter requested without candidate group ids");
  if (shard_stride_bytes < 0 || shard_stride_bytes % 8 != 0)
    return fail(MRAG_ERR_ARG, "shard_stride_bytes must be a non-negative multiple of 8");
  const int fm = exclude_group_dev ? filter_mode : MRAG_FILTER_NONE;
  CK(launch_k3_merge_shards(cand_dist_dev, cand_idx_dev, cand_group_dev, shard_stride_bytes,
                            nshards, nq, k_in, k_out,
                            exclude_group_dev, fm, out_dist_dev, out_idx_dev, out_group_dev,
                            static_cast<cudaStream_t>(stream)));
  return MRAG_OK;
}

int mrag_gather_context(const void* const* shard_ptrs_dev, int32_t nshards, int64_t rows_per_shard,
                        const int64_t* ref_idx_dev, const void* sos_dev, const void* uncond_row_dev,
                        const void* pe_dev, const void* cond_dev, void* out_dev, int32_t b,
                        int32_t K, int32_t L, int32_t C, int32_t dtype, int64_t n_rows_total,
                        void* stream) {
  if (!shard_ptrs_dev || !ref_idx_dev || !sos_dev || !uncond_row_dev || !out_dev)
    return fail(MRAG_ERR_ARG, "null device buffer");
  if (nshards < 1 || rows_per_shard < 1 || b < 1 || K < 1 || L < 1 || C < 1)
    return fail(MRAG_ERR_ARG, "bad gather shape");
  if (n_rows_total < 0 || n_rows_total > int64_t(nshards) * rows_per_shard)
    return fail(MRAG_ERR_ARG, "n_rows_total %lld does not fit %d shards of %lld rows", (long long)n_rows_total,
                nshards, (long long)rows_per_shard);
  if (dtype != 0 && dtype != 1) return fail(MRAG_ERR_ARG, "dtype must be 0 (bf16) or 1 (f32)");
  if ((int64_t(L) * C * (dtype == 0 ? 2 : 4)) % 16 != 0)
    return fail(MRAG_ERR_ARG, "L*C*sizeof(elt) must be a multiple of 16 bytes");
  CK(launch_k4_gather(shard_ptrs_dev, nshards, rows_per_shard, n_rows_total, ref_idx_dev, sos_dev,
                      uncond_row_dev, pe_dev, cond_dev, out_dev, b, K, L, C, dtype,
                      static_cast<cudaStream_t>(stream)));
  return MRAG_OK;
}

int mrag_device_alloc(int32_t device, size_t bytes, void** dev_ptr_out) {
  if (!dev_ptr_out || bytes == 0) return fail(MRAG_ERR_ARG, "bad alloc request");
  DeviceGuard g(device);
  if (!g.ok) return fail(MRAG_ERR_DEVICE, "cannot select device %d", device);
  CK(cudaMalloc(dev_ptr_out, bytes));
  return MRAG_OK;
}

int mrag_device_free(int32_t device, void* dev_ptr) {
  if (!dev_ptr) return MRAG_OK;
  DeviceGuard g(device);
  CK(cudaFree(dev_ptr));
  return MRAG_OK;
}

int mrag_ipc_export(const void* dev_ptr, void* handle_out_64B) {
  if (!dev_ptr || !handle_out_64B) return fail(MRAG_ERR_ARG, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  memcpy(handle_out_64B, &h, 64);
  return MRAG_OK;
}

int mrag_ipc_open(const void* handle_64B, void** dev_ptr_out) {
  if (!handle_64B || !dev_ptr_out) return fail(MRAG_ERR_ARG, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64B, 64);
  CK(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return MRAG_OK;
}

int mrag_ipc_close(void* dev_ptr) {
  if (!dev_ptr) return MRAG_OK;
  CK(cudaIpcCloseMemHandle(dev_ptr));
  return MRAG_OK;
}

}  // extern "C"
