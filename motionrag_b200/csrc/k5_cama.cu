// K5/K6/K7 — the CAMA causal motion transformer forward, the immediate consumer
// of the gathered context tensor: torch.nn.TransformerEncoder(num_layers=4,
// TransformerEncoderLayer(d_model=1024, nhead=16, dim_feedforward=4096, dropout=0, activation=gelu,
// batch_first, norm_first=False, bias=True)) with the block-causal mask of
// ActionTransformer.get_mask (reference configs/cogvideox/MotionRAG_open.yml:253-267,
// src/projects/condition/module.py:131-135, 303-306).
//
// Per layer (post-norm):   qkv = x Wqkv^T + b          K5 (bf16 out)
//                          a   = blockcausal_attn(qkv) K6
//                          p   = a Wo^T                K5 (split-K, fp32 partial sums)
//                          x1  = LN1(x + p + bo)       K7
//                          h   = gelu(x1 W1^T + b1)    K5 (bf16 out)
//                          p   = h W2^T                K5 (split-K, fp32 partial sums)
//                          x2  = LN2(x1 + p + b2)      K7
// K5 is a tcgen05 GEMM built from the same parts as the retrieval scan K2 (TMA ring of
// SWIZZLE_128B k-blocks -> single-thread tcgen05.mma with the accumulator in TMEM ->
// tcgen05.ld epilogue), tile 128 x 128, one tile (and one K split) per CTA. The residual path
// stays in fp32 until the LayerNorm (torch rounds the projection to bf16 first), so results are
// at least as close to an fp32 evaluation as torch's own bf16 path.
#include "k2_common.cuh"

namespace mrag {

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmStages = 6, kGemmThreads = 192;
constexpr uint32_t kGemmABytes = kGemmBM * kBK * 2, kGemmBBytes = kGemmBN * kBK * 2;
constexpr uint32_t kGemmStageBytes = kGemmABytes + kGemmBBytes;
constexpr uint32_t kGemmSmem = kGemmStages * kGemmStageBytes + 256 + 1024;

struct GemmArgs {
  int M, N, K;          // C[M,N] = A[M,K] W[N,K]^T ; K multiple of 64, N multiple of 128
  int kb_per_split;     // k-blocks (of 64) handled by one CTA along grid.z
  const __nv_bfloat16* bias;  // [N] or null (bf16 mode only)
  int gelu;             // bf16 mode: apply exact (erf) GELU after the bias
  __nv_bfloat16* out;   // bf16 mode: [M,N]
  float* partial;       // split mode: [splits][M,N] fp32 (out == null)
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__global__ void __launch_bounds__(kGemmThreads, 1)
    k5_linear_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                     const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGemmStages * kGemmStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kGemmStages;
  uint64_t* tfull_bar = bars + 2 * kGemmStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGemmStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // let the next kernel of the chain become resident as soon as every CTA of this grid has started;
  // it parks in its own griddepcontrol.wait until this grid has completed
  asm volatile("griddepcontrol.launch_dependents;");
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
  const int kb0 = split * g.kb_per_split;
  const int kb1 = min(g.K / kBK, kb0 + g.kb_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // inputs of this GEMM are produced by the previous kernel in the stream (PDL-safe either way)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_a = policy_evict_last();    // activations: re-read by every N tile
      const uint64_t pol_w = policy_evict_normal();  // weights: 100 MB total, L2-resident across calls
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kGemmStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], kGemmStageBytes);
        tma_load_2d(sa, &tm_a, &full_bar[stage], kb * kBK, m_tile * kGemmBM, pol_a);
        tma_load_2d(sa + kGemmABytes, &tm_w, &full_bar[stage], kb * kBK, n_tile * kGemmBN, pol_w);
        if (++stage == kGemmStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, kGemmBN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kGemmStageBytes);
        const uint64_t da = umma_desc_k_sw128(sa);
        const uint64_t db = umma_desc_k_sw128(sa + kGemmABytes);
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k)
          umma_bf16(tmem_base, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (kb == kb1 - 1) umma_commit(tfull_bar);
        if (++stage == kGemmStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // epilogue: thread = output row, 32 columns per tcgen05.ld
    const int quarter = warp & 3;
    const int m = m_tile * kGemmBM + quarter * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < kGemmBN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(t_addr + c * 32, v);
      tmem_ld_wait();
      const int n0 = n_tile * kGemmBN + c * 32;
      if (m < g.M) {
        if (g.out != nullptr) {
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float a = __uint_as_float(v[j]), b = __uint_as_float(v[j + 1]);
            if (g.bias != nullptr) {
              const __nv_bfloat162 bb = *reinterpret_cast<const __nv_bfloat162*>(g.bias + n0 + j);
              a += __bfloat162float(bb.x);
              b += __bfloat162float(bb.y);
            }
            if (g.gelu) {
              a = gelu_erf(a);
              b = gelu_erf(b);
            }
            const __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
            packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&r);
          }
          uint4* dst = reinterpret_cast<uint4*>(g.out + size_t(m) * g.N + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        } else {
          float4* dst = reinterpret_cast<float4*>(g.partial + (size_t(split) * g.M + m) * g.N + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// ---- K6: block-causal attention, head_dim 64 ---------------------------------------------------
// grid (b * heads, groups); block 128 threads. Group g attends to keys of groups 0..g.
// qkv [M, 3*d] bf16 (q | k | v, head h at columns h*64 inside each third), out [M, d] bf16.
constexpr int kAttnThreads = 128;
constexpr int kHeadDim = 64;
constexpr int kKeyStride = 66;  // bf16 elements per staged key/value row: 33 words -> conflict-free

__global__ void __launch_bounds__(kAttnThreads)
    k6_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T,
                        int d_model, int heads, int group_tokens, float scale) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int bh = blockIdx.x, g = blockIdx.y;
  const int b = bh / heads, h = bh % heads;
  const int nk = (g + 1) * group_tokens;  // visible keys
  __nv_bfloat16* ks = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* vs = ks + size_t(nk) * kKeyStride;
  float* probs = reinterpret_cast<float*>(vs + size_t(nk) * kKeyStride);  // [4 warps][nk]
  float* qs = probs + 4 * nk;                                              // [4 warps][64]
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t row_stride = size_t(3) * d_model;
  const __nv_bfloat16* base = qkv + size_t(b) * T * row_stride;
  // stage K and V rows of this head: 64 bf16 = 128 B per row, 4-byte words
  for (int i = tid; i < nk * 32; i += kAttnThreads) {
    const int r = i >> 5, w = i & 31;
    const uint32_t kw = *reinterpret_cast<const uint32_t*>(base + size_t(r) * row_stride + d_model + h * kHeadDim + 2 * w);
    const uint32_t vw = *reinterpret_cast<const uint32_t*>(base + size_t(r) * row_stride + 2 * d_model + h * kHeadDim + 2 * w);
    *reinterpret_cast<uint32_t*>(ks + size_t(r) * kKeyStride + 2 * w) = kw;
    *reinterpret_cast<uint32_t*>(vs + size_t(r) * kKeyStride + 2 * w) = vw;
  }
  __syncthreads();
  float* my_p = probs + warp * nk;
  float* my_q = qs + warp * kHeadDim;
  for (int qi = warp; qi < group_tokens; qi += 4) {
    const int t = g * group_tokens + qi;
    const uint32_t qw = *reinterpret_cast<const uint32_t*>(base + size_t(t) * row_stride + h * kHeadDim + 2 * lane);
    my_q[2 * lane] = bf16lo_to_f32(qw) * scale;
    my_q[2 * lane + 1] = bf16hi_to_f32(qw) * scale;
    __syncwarp();
    // scores: lane owns keys lane, lane+32, ...
    float mx = -INFINITY;
    for (int k0 = lane; k0 < nk; k0 += 32) {
      const __nv_bfloat16* kr = ks + size_t(k0) * kKeyStride;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 32; ++w) {
        const uint32_t kw = *reinterpret_cast<const uint32_t*>(kr + 2 * w);
        s = fmaf(my_q[2 * w], bf16lo_to_f32(kw), s);
        s = fmaf(my_q[2 * w + 1], bf16hi_to_f32(kw), s);
      }
      my_p[k0] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
    for (int k0 = lane; k0 < nk; k0 += 32) {
      const float p = __expf(my_p[k0] - mx);
      my_p[k0] = p;
      sum += p;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    __syncwarp();
    // output: lane owns dims 2*lane, 2*lane+1
    float o0 = 0.f, o1 = 0.f;
    for (int k0 = 0; k0 < nk; ++k0) {
      const float p = my_p[k0];
      const uint32_t vw = *reinterpret_cast<const uint32_t*>(vs + size_t(k0) * kKeyStride + 2 * lane);
      o0 = fmaf(p, bf16lo_to_f32(vw), o0);
      o1 = fmaf(p, bf16hi_to_f32(vw), o1);
    }
    const float inv = 1.f / sum;
    const __nv_bfloat162 r = __floats2bfloat162_rn(o0 * inv, o1 * inv);
    *reinterpret_cast<__nv_bfloat162*>(out + (size_t(b) * T + t) * d_model + h * kHeadDim + 2 * lane) = r;
    __syncwarp();
  }
}

// ---- K7: y = LayerNorm(resid + sum_s partial[s] + bias) * gamma + beta ---------------------------
// one warp per row; d multiple of 256 (8 elements per lane per pass)
__global__ void __launch_bounds__(256)
    k7_add_layernorm_kernel(const __nv_bfloat16* __restrict__ resid, const float* __restrict__ partial,
                            int splits, const __nv_bfloat16* __restrict__ bias,
                            const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                            __nv_bfloat16* __restrict__ out, int M, int d, float eps) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  constexpr int MAXV = 4;  // d <= 1024: 4 passes of 8 elements per lane
  float x[MAXV][8];
  const int passes = d / 256;
  float s1 = 0.f;
#pragma unroll
  for (int p = 0; p < MAXV; ++p) {
    if (p >= passes) break;
    const int c = p * 256 + lane * 8;
    const uint4 rv = *reinterpret_cast<const uint4*>(resid + size_t(row) * d + c);
    const uint4 bv = *reinterpret_cast<const uint4*>(bias + c);
    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x[p][2 * j] = bf16lo_to_f32(rw[j]) + bf16lo_to_f32(bw[j]);
      x[p][2 * j + 1] = bf16hi_to_f32(rw[j]) + bf16hi_to_f32(bw[j]);
    }
    for (int s = 0; s < splits; ++s) {
      const float4* pp = reinterpret_cast<const float4*>(partial + (size_t(s) * M + row) * d + c);
      const float4 a = pp[0], b = pp[1];
      x[p][0] += a.x; x[p][1] += a.y; x[p][2] += a.z; x[p][3] += a.w;
      x[p][4] += b.x; x[p][5] += b.y; x[p][6] += b.z; x[p][7] += b.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s1 += x[p][j];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, off);
  const float mean = s1 / d;
  float s2 = 0.f;
#pragma unroll
  for (int p = 0; p < MAXV; ++p) {
    if (p >= passes) break;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = x[p][j] - mean;
      s2 = fmaf(t, t, s2);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  const float rstd = rsqrtf(s2 / d + eps);
#pragma unroll
  for (int p = 0; p < MAXV; ++p) {
    if (p >= passes) break;
    const int c = p * 256 + lane * 8;
    const uint4 gv = *reinterpret_cast<const uint4*>(gamma + c);
    const uint4 ev = *reinterpret_cast<const uint4*>(beta + c);
    const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, ew[4] = {ev.x, ev.y, ev.z, ev.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = (x[p][2 * j] - mean) * rstd * bf16lo_to_f32(gw[j]) + bf16lo_to_f32(ew[j]);
      const float b = (x[p][2 * j + 1] - mean) * rstd * bf16hi_to_f32(gw[j]) + bf16hi_to_f32(ew[j]);
      const __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
      o[j] = *reinterpret_cast<const uint32_t*>(&r);
    }
    *reinterpret_cast<uint4*>(out + size_t(row) * d + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---- host side ---------------------------------------------------------------------------------
static cudaError_t launch_pdl(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelExC(&cfg, fn, args);
}

cudaError_t launch_k5_linear(const void* a_bf16, int a_rows_alloc, const void* w_bf16, int M, int N, int K,
                             const void* bias, bool gelu, void* out_bf16, float* partial, int splits,
                             cudaStream_t st) {
  if (N % kGemmBN != 0 || K % kBK != 0 || M < 1 || splits < 1) return cudaErrorInvalidValue;
  CUtensorMap tm_a, tm_w;
  if (!make_tmap(&tm_a, a_bf16, a_rows_alloc, K, kGemmBM) || !make_tmap(&tm_w, w_bf16, N, K, kGemmBN))
    return cudaErrorInvalidValue;
  GemmArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  const int kblocks = K / kBK;
  g.kb_per_split = (kblocks + splits - 1) / splits;
  g.bias = static_cast<const __nv_bfloat16*>(bias);
  g.gelu = gelu ? 1 : 0;
  g.out = static_cast<__nv_bfloat16*>(out_bf16);
  g.partial = partial;
  cudaError_t e = cudaFuncSetAttribute(k5_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGemmSmem));
  if (e != cudaSuccess) return e;
  void* args[] = {&tm_a, &tm_w, &g};
  e = launch_pdl(reinterpret_cast<const void*>(k5_linear_kernel),
                 dim3(unsigned(N / kGemmBN), unsigned((M + kGemmBM - 1) / kGemmBM), unsigned(splits)),
                 dim3(kGemmThreads), kGemmSmem, st, args);
  note_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_k6_attention(const void* qkv, void* out, int b, int T, int d_model, int heads,
                                int groups, int group_tokens, cudaStream_t st) {
  if (d_model != heads * kHeadDim || groups * group_tokens != T) return cudaErrorInvalidValue;
  const size_t smem = size_t(2) * T * kKeyStride * 2 + size_t(4) * T * 4 + 4 * kHeadDim * 4;
  cudaError_t e = cudaFuncSetAttribute(k6_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  float scale = 1.f / sqrtf(float(kHeadDim));
  void* args[] = {&q, &o, &T, &d_model, &heads, &group_tokens, &scale};
  e = launch_pdl(reinterpret_cast<const void*>(k6_attention_kernel), dim3(unsigned(b * heads), unsigned(groups)),
                 dim3(kAttnThreads), smem, st, args);
  note_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_k7_add_layernorm(const void* resid, const float* partial, int splits, const void* bias,
                                    const void* gamma, const void* beta, void* out, int M, int d, float eps,
                                    cudaStream_t st) {
  if (d % 256 != 0 || d > 1024) return cudaErrorInvalidValue;
  const __nv_bfloat16* r = static_cast<const __nv_bfloat16*>(resid);
  const __nv_bfloat16* bi = static_cast<const __nv_bfloat16*>(bias);
  const __nv_bfloat16* ga = static_cast<const __nv_bfloat16*>(gamma);
  const __nv_bfloat16* be = static_cast<const __nv_bfloat16*>(beta);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  void* args[] = {&r, &partial, &splits, &bi, &ga, &be, &o, &M, &d, &eps};
  cudaError_t e = launch_pdl(reinterpret_cast<const void*>(k7_add_layernorm_kernel), dim3(unsigned((M * 32 + 255) / 256)),
                             dim3(256), 0, st, args);
  note_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mrag
