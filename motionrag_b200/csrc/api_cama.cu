// C ABI of the CAMA transformer forward (include/mrag.h, "next" row f-1). Host orchestration
// only: buffers, the launch chain of K5/K6/K7 and its CUDA-graph replay.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <cstdlib>
#include <vector>

#include "../../include/mrag.h"
#include "kernels.h"

namespace mrag {
int api_fail(int code, const char* fmt, ...);  // api.cu
}
using namespace mrag;

struct mrag_cama {
  int n_layers = 0, d = 0, heads = 0, dff = 0, groups = 0, gtok = 0, T = 0, max_b = 0, device = 0;
  int rows_alloc = 0;  // max_b * T rounded up to 128
  std::vector<mrag_cama_layer> layers;
  char* buf = nullptr;  // one allocation, carved below
  void *x_in = nullptr, *x_a = nullptr, *x_b = nullptr, *qkv = nullptr, *att = nullptr, *h = nullptr, *y_out = nullptr;
  void* y_last = nullptr;  // [max_b, group_tokens, d] bf16: the last group's rows of the prediction (mrag_cama_predict)
  float* partial = nullptr;
  int sm_count = 0;
  struct Graph {
    int b;
    bool last_only;
    cudaGraphExec_t exec;
  };
  std::vector<Graph> graphs;
  cudaStream_t capture_stream = nullptr;
};

namespace {
constexpr int kSplitsO = 4;   // out-proj: K = d      (16 k-blocks at d = 1024)
constexpr int kSplitsF = 8;   // ffn2    : K = d_ff   (64 k-blocks at d_ff = 4096)

// split K only as far as needed to give every SM a tile: the fp32 partial sums cost HBM traffic
int pick_splits(int M, int N, int max_splits) {
  const int tiles = ((M + 127) / 128) * (N / 128);
  int s = 1;
  while (s < max_splits && tiles * s < 148) s *= 2;
  return s;
}

// bf16-output GEMMs (QKV, FFN1) with less than a wave of 128 x 128 tiles: split K inside a cluster
// (k5_linear_kernel<.., REDUCE>) as far as the 2 x 148 co-resident CTAs allow
int pick_cluster_splits(int M, int N, int K) {
  static const int max_s = [] {
    const char* e = getenv("MRAG_K5_REDUCE_MAX");  // tuning knob: 1 disables
    const int v = e ? atoi(e) : 4;
    return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 4;
  }();
  const int tiles = ((M + 127) / 128) * (N / 128);
  // measured (scripts/cama_bench.py, MRAG_K5_REDUCE_MAX sweep): with the <= 64 tiles of one sample the
  // 128 x 64-tile form without clusters is as fast (186 vs 190 us per forward); from two samples on the
  // cluster form avoids a second wave of 128 x 64 tiles (b = 2: 258 -> 237 us)
  if (tiles >= 120 || tiles <= 64) return 1;
  int s = 1;
  while (s < max_s && tiles * s * 2 <= 296 && (K / 64) % (s * 2) == 0) s *= 2;
  return s;
}

struct Guard {
  int prev = -1;
  explicit Guard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~Guard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

// last_only: ActionTransformer.predict keeps only the last group of the output (module.py:326), so the
// LAST layer runs its attention, out-projection, FFN and both LayerNorms for that group's rows alone
// (b * group_tokens compact rows instead of b * T); K and V are still needed for every row, so its QKV
// GEMM stays whole. The result lands in y_last [b, group_tokens, d].
cudaError_t run_chain(const mrag_cama* c, int b, cudaStream_t st, bool last_only) {
  const int M = b * c->T, d = c->d, dff = c->dff;
  const void* xin = c->x_in;
  cudaError_t e = cudaSuccess;
  const int so = pick_splits(M, d, kSplitsO), sf = pick_splits(M, d, kSplitsF);
  const int sq = pick_cluster_splits(M, 3 * d, d), s1 = pick_cluster_splits(M, dff, d);
  for (int l = 0; l < c->n_layers && e == cudaSuccess; ++l) {
    const mrag_cama_layer& w = c->layers[l];
    void* x1 = c->x_a;
    const bool tail = last_only && l == c->n_layers - 1;
    e = launch_k5_linear(xin, c->rows_alloc, w.w_qkv, M, 3 * d, d, w.b_qkv, false, c->qkv, nullptr, sq, st);
    if (tail) {
      const int Ml = b * c->gtok;
      const int sol = pick_splits(Ml, d, kSplitsO), sfl = pick_splits(Ml, d, kSplitsF);
      const int s1l = pick_cluster_splits(Ml, dff, d);
      if (e == cudaSuccess)
        e = launch_k6_attention(c->qkv, c->att, b, c->T, d, c->heads, c->groups, c->gtok, st, c->groups - 1);
      if (e == cudaSuccess)
        e = launch_k5_linear(c->att, c->rows_alloc, w.w_o, Ml, d, d, nullptr, false, nullptr, c->partial, sol, st);
      if (e == cudaSuccess)
        e = launch_k7_add_layernorm(xin, c->partial, sol, w.b_o, w.ln1_g, w.ln1_b, x1, Ml, d, 1e-5f, st, c->T, c->gtok);
      if (e == cudaSuccess)
        e = launch_k5_linear(x1, c->rows_alloc, w.w_1, Ml, dff, d, w.b_1, true, c->h, nullptr, s1l, st);
      if (e == cudaSuccess)
        e = launch_k5_linear(c->h, c->rows_alloc, w.w_2, Ml, d, dff, nullptr, false, nullptr, c->partial, sfl, st);
      if (e == cudaSuccess)
        e = launch_k7_add_layernorm(x1, c->partial, sfl, w.b_2, w.ln2_g, w.ln2_b, c->y_last, Ml, d, 1e-5f, st);
      break;
    }
    void* x2 = (l == c->n_layers - 1) ? c->y_out : c->x_b;
    if (e == cudaSuccess) e = launch_k6_attention(c->qkv, c->att, b, c->T, d, c->heads, c->groups, c->gtok, st);
    if (e == cudaSuccess)
      e = launch_k5_linear(c->att, c->rows_alloc, w.w_o, M, d, d, nullptr, false, nullptr, c->partial, so, st);
    if (e == cudaSuccess)
      e = launch_k7_add_layernorm(xin, c->partial, so, w.b_o, w.ln1_g, w.ln1_b, x1, M, d, 1e-5f, st);
    if (e == cudaSuccess) e = launch_k5_linear(x1, c->rows_alloc, w.w_1, M, dff, d, w.b_1, true, c->h, nullptr, s1, st);
    if (e == cudaSuccess)
      e = launch_k5_linear(c->h, c->rows_alloc, w.w_2, M, d, dff, nullptr, false, nullptr, c->partial, sf, st);
    if (e == cudaSuccess)
      e = launch_k7_add_layernorm(x1, c->partial, sf, w.b_2, w.ln2_g, w.ln2_b, x2, M, d, 1e-5f, st);
    xin = x2;
  }
  return e;
}

int run_forward(mrag_cama* c, int32_t b, int32_t use_graph, void* stream, bool last_only) {
  if (!c) return api_fail(MRAG_ERR_ARG, "null handle");
  if (b < 1 || b > c->max_b) return api_fail(MRAG_ERR_ARG, "batch %d outside 1..%d", b, c->max_b);
  Guard g(c->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!use_graph) {
    cudaError_t e = run_chain(c, b, st, last_only);
    if (e != cudaSuccess) return api_fail(MRAG_ERR_CUDA, "transformer launch chain: %s", cudaGetErrorString(e));
    return MRAG_OK;
  }
  cudaGraphExec_t exec = nullptr;
  for (auto& gr : c->graphs)
    if (gr.b == b && gr.last_only == last_only) exec = gr.exec;
  if (!exec) {
    if (!c->capture_stream &&
        cudaStreamCreateWithFlags(&c->capture_stream, cudaStreamNonBlocking) != cudaSuccess)
      return api_fail(MRAG_ERR_CUDA, "cannot create the capture stream");
    cudaError_t e = cudaStreamBeginCapture(c->capture_stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return api_fail(MRAG_ERR_CUDA, "begin capture: %s", cudaGetErrorString(e));
    cudaError_t ce = run_chain(c, b, c->capture_stream, last_only);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(c->capture_stream, &graph);
    if (ce != cudaSuccess || e != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      return api_fail(MRAG_ERR_CUDA, "capture of the transformer chain: %s",
                      cudaGetErrorString(ce != cudaSuccess ? ce : e));
    }
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return api_fail(MRAG_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(e));
    c->graphs.push_back({b, last_only, exec});
  }
  cudaError_t e = cudaGraphLaunch(exec, st);
  if (e != cudaSuccess) return api_fail(MRAG_ERR_CUDA, "graph launch: %s", cudaGetErrorString(e));
  note_launch(7 * c->n_layers);
  return MRAG_OK;
}
}  // namespace

extern "C" {

int mrag_cama_create(int32_t n_layers, const mrag_cama_layer* layers, int32_t d_model, int32_t n_heads,
                     int32_t d_ff, int32_t groups, int32_t group_tokens, int32_t max_batch, int32_t device,
                     mrag_cama** out) {
  if (!out || !layers) return api_fail(MRAG_ERR_ARG, "null argument");
  *out = nullptr;
  if (n_layers < 1 || n_layers > 64 || max_batch < 1 || groups < 1 || group_tokens < 1)
    return api_fail(MRAG_ERR_ARG, "bad transformer shape");
  if (d_model != n_heads * 64 || d_model % 256 != 0 || d_model > 1024 || d_ff % 128 != 0 ||
      (d_model / 64) % kSplitsO != 0 || (d_ff / 64) % kSplitsF != 0)
    return api_fail(MRAG_ERR_UNSUPPORTED,
                    "kernels are built for head_dim 64, d_model in {256,512,768,1024} and d_ff %% 512 == 0 "
                    "(got d_model %d, heads %d, d_ff %d)", d_model, n_heads, d_ff);
  if (int64_t(groups) * group_tokens > 704)
    return api_fail(MRAG_ERR_UNSUPPORTED, "sequence of %lld tokens: the attention kernel stages K and V of the whole "
                    "sequence in shared memory (at most 704 tokens)", (long long)(int64_t(groups) * group_tokens));
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return api_fail(MRAG_ERR_DEVICE, "no such CUDA device %d: libmrag has no CPU path", device);
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return api_fail(MRAG_ERR_DEVICE, "device %d is not sm_100-class", device);
  Guard g(device);
  mrag_cama* c = new (std::nothrow) mrag_cama();
  if (!c) return api_fail(MRAG_ERR_CAPACITY, "host allocation failed");
  c->n_layers = n_layers;
  c->d = d_model;
  c->heads = n_heads;
  c->dff = d_ff;
  c->groups = groups;
  c->gtok = group_tokens;
  c->T = groups * group_tokens;
  c->max_b = max_batch;
  c->device = device;
  c->rows_alloc = (max_batch * c->T + 127) / 128 * 128;
  c->layers.assign(layers, layers + n_layers);
  const size_t R = size_t(c->rows_alloc);
  const size_t sz_x = R * d_model * 2, sz_qkv = R * 3 * d_model * 2, sz_h = R * d_ff * 2;
  const size_t sz_p = size_t(kSplitsF) * R * d_model * 4;
  const size_t total = 6 * sz_x + sz_qkv + sz_h + sz_p + 1024;
  c->sm_count = prop.multiProcessorCount;
  if (cudaMalloc(reinterpret_cast<void**>(&c->buf), total) != cudaSuccess) {
    cudaGetLastError();
    delete c;
    return api_fail(MRAG_ERR_CUDA, "cannot allocate %zu bytes of transformer workspace", total);
  }
  cudaMemset(c->buf, 0, total);
  char* p = c->buf;
  c->x_in = p; p += sz_x;
  c->x_a = p; p += sz_x;
  c->x_b = p; p += sz_x;
  c->att = p; p += sz_x;
  c->y_out = p; p += sz_x;
  c->y_last = p; p += sz_x;
  c->qkv = p; p += sz_qkv;
  c->h = p; p += sz_h;
  c->partial = reinterpret_cast<float*>(p); p += sz_p;
  *out = c;
  return MRAG_OK;
}

int mrag_cama_destroy(mrag_cama* c) {
  if (!c) return MRAG_OK;
  Guard g(c->device);
  for (auto& gr : c->graphs) cudaGraphExecDestroy(gr.exec);
  if (c->capture_stream) cudaStreamDestroy(c->capture_stream);
  cudaFree(c->buf);
  delete c;
  return MRAG_OK;
}

int mrag_cama_io(const mrag_cama* c, void** x_in_dev, void** y_out_dev) {
  if (!c) return api_fail(MRAG_ERR_ARG, "null handle");
  if (x_in_dev) *x_in_dev = c->x_in;
  if (y_out_dev) *y_out_dev = c->y_out;
  return MRAG_OK;
}

int mrag_cama_forward(mrag_cama* c, int32_t b, int32_t use_graph, void* stream) {
  return run_forward(c, b, use_graph, stream, false);
}

int mrag_cama_predict(mrag_cama* c, int32_t b, int32_t use_graph, void* stream, void** y_last_dev) {
  const int rc = run_forward(c, b, use_graph, stream, true);
  if (rc == MRAG_OK && y_last_dev) *y_last_dev = c->y_last;
  return rc;
}

int mrag_linear(const void* a_dev, int32_t a_rows_alloc, const void* w_dev, int32_t M, int32_t N, int32_t K,
                const void* bias_dev, int32_t gelu, void* out_bf16_dev, float* partial_dev, int32_t splits,
                void* stream) {
  if (!a_dev || !w_dev || (!out_bf16_dev && !partial_dev)) return api_fail(MRAG_ERR_ARG, "null device buffer");
  if (splits < 1 || (K / 64) % splits != 0) return api_fail(MRAG_ERR_ARG, "splits must divide K/64");
  cudaError_t e = launch_k5_linear(a_dev, a_rows_alloc, w_dev, M, N, K, bias_dev, gelu != 0, out_bf16_dev,
                                   partial_dev, splits, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return api_fail(MRAG_ERR_CUDA, "linear: %s", cudaGetErrorString(e));
  return MRAG_OK;
}

}  // extern "C"
