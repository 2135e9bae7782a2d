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
// tcgen05.ld epilogue). Four forms, chosen per launch by launch_k5_linear from the tile count:
//   k5_linear_kernel<BN,STAGES,false>   one 128 x BN tile and one K split per CTA (less than ~2 rounds of tiles)
//   k5_linear_kernel<128,3,true>        K splits of a tile = one cluster, partials summed through DSMEM
//   k5_linear_persistent_kernel<BN>     one CTA per SM walks the tiles, two TMEM accumulators (N % 256 != 0)
//   k5_linear_pair_kernel               the same on a CTA pair: cta_group::2, 256 x 256 tiles (N % 256 == 0)
// The residual path stays in fp32 until the LayerNorm (torch rounds the projection to bf16 first), so
// results are at least as close to an fp32 evaluation as torch's own bf16 path.
#include <cstdio>
#include <cstdlib>

#include "k2_common.cuh"

namespace mrag {

constexpr int kGemmBM = 128, kGemmThreads = 192;
constexpr uint32_t kGemmABytes = kGemmBM * kBK * 2;
// BN = 128 by default; BN = 64 doubles the CTA count when a GEMM would otherwise leave most SMs idle.
// STAGES = 6 (one CTA per SM) for grids of at most one wave; STAGES = 3 lets two CTAs share an SM so
// that one tile's epilogue overlaps the other's main loop when there are several waves of tiles.
template <int BN, int STAGES>
struct GemmCfg {
  static constexpr uint32_t kBBytes = BN * kBK * 2;
  static constexpr uint32_t kStageBytes = kGemmABytes + kBBytes;
  static constexpr uint32_t kSmem = STAGES * kStageBytes + 256 + 1024;
};

struct GemmArgs {
  int M, N, K;          // C[M,N] = A[M,K] W[N,K]^T ; K multiple of 64, N multiple of 128
  int kb_per_split;     // k-blocks (of 64) handled by one CTA along grid.z
  const __nv_bfloat16* bias;  // [N] or null (bf16 mode only)
  int gelu;             // bf16 mode: apply exact (erf) GELU after the bias
  __nv_bfloat16* out;   // bf16 mode: [M,N]
  float* partial;       // split mode: [splits][M,N] fp32 (out == null)
};

// erf GELU, 0.5 x (1 + erf(x / sqrt 2)), evaluated through erfc(z) = (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z),
// z = |x| / sqrt 2 (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 on erf): at most 3.4e-7 absolute off the float64
// GELU over [-8, 8] (tests/test_host_logic.py restates it), three orders below the bf16 rounding of the output.
// 14 instructions, two of them MUFU, against ~30 for erff(): the FFN1 epilogue of the multi-tile kernels was
// issue-bound on it (profiles/r2_k5_pair_ffn1_b16.txt). The negative side is h * erfc(z) directly, no cancellation.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = fabsf(0.5f * x) * (p * t * e);  // |x| / 2 * erfc(z)
  return x > 0.f ? x - r : -r;
}

// ---- epilogue pieces shared by the multi-tile kernels (thread = output row, warp = 32 rows x CW columns) ----------
// The bias slice of the warp's columns is parked in a warp-private strip of shared memory before the accumulator is
// waited for (one coalesced load instead of 16 scalar loads per 32 columns on the critical path; broadcast LDS.128).
template <int CW>
__device__ __forceinline__ void k5_stage_bias(const GemmArgs& g, int n0, float* bias_s, int lane) {
  __syncwarp();  // the previous tile's reads of the strip are done
  if constexpr (CW == 64) {
    float2 b = make_float2(0.f, 0.f);
    if (g.bias != nullptr) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(g.bias + n0 + 2 * lane);
      b = make_float2(bf16lo_to_f32(w), bf16hi_to_f32(w));
    }
    *reinterpret_cast<float2*>(bias_s + 2 * lane) = b;
  } else {
    static_assert(CW == 32, "32 or 64 columns per epilogue warp");
    bias_s[lane] = g.bias != nullptr ? __bfloat162float(g.bias[n0 + lane]) : 0.f;
  }
  __syncwarp();
}

// 32 accumulator columns of one row: + bias (+ GELU) -> bf16, or the raw fp32 partial sums
__device__ __forceinline__ void k5_write_cols32(const GemmArgs& g, const uint32_t (&v)[32], const float* bias_s, int m,
                                                int n0) {
  if (g.out != nullptr) {
    uint4* dst = reinterpret_cast<uint4*>(g.out + size_t(m) * g.N + n0);
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c), b1 = *reinterpret_cast<const float4*>(bias_s + c + 4);
      float x[8] = {__uint_as_float(v[c]) + b0.x,     __uint_as_float(v[c + 1]) + b0.y, __uint_as_float(v[c + 2]) + b0.z,
                    __uint_as_float(v[c + 3]) + b0.w, __uint_as_float(v[c + 4]) + b1.x, __uint_as_float(v[c + 5]) + b1.y,
                    __uint_as_float(v[c + 6]) + b1.z, __uint_as_float(v[c + 7]) + b1.w};
      if (g.gelu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = gelu_erf(x[e]);
      }
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 r = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
        w[e] = *reinterpret_cast<const uint32_t*>(&r);
      }
      dst[c >> 3] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
    float4* dst = reinterpret_cast<float4*>(g.partial + size_t(m) * g.N + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                           __uint_as_float(v[4 * j + 3]));
  }
}

// REDUCE: the grid.z K splits of a tile form one thread-block cluster and their fp32 partial tiles are
// summed on chip: every CTA parks its TMEM accumulator in its own shared memory, then CTA r of the
// cluster sums rows [r*128/S, (r+1)*128/S) over the S tiles through distributed shared memory (fixed
// order, so results are deterministic) and writes the final bf16 (+bias, +GELU) — no fp32 round trip
// through HBM. Used when M is small: each CTA then walks only K/S of the reduction dimension, which
// shortens the TMA -> MMA latency chain that bounds a GEMM of less than one wave of tiles.
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta_rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(remote)
               : "memory");
  return v;
}

template <int BN, int STAGES, bool REDUCE>
__global__ void __launch_bounds__(kGemmThreads, STAGES <= 3 ? 2 : 1)
    k5_linear_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                     const GemmArgs g) {
  constexpr uint32_t kGemmStageBytes = GemmCfg<BN, STAGES>::kStageBytes;
  constexpr int kGemmBN = BN;
  constexpr int kGemmStages = STAGES;
  constexpr int kStgStride = BN + 4;  // floats per parked accumulator row (bank-conflict-free float4 rows)
  static_assert(!REDUCE || kGemmBM * kStgStride * 4 <= STAGES * GemmCfg<BN, STAGES>::kStageBytes,
                "parked accumulator tile must fit in the pipeline stages");
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
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
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
      if constexpr (REDUCE) {
        // the pipeline stages are idle once tfull has fired (every MMA has read its operands)
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(smem) +
                                                size_t(quarter * 32 + lane) * kStgStride + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      } else if (m < g.M) {
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
  if constexpr (REDUCE) {
    __syncwarp();        // the producer / MMA lanes rejoin their warps before the aligned cluster barrier
    cluster_sync_all();  // every CTA of the cluster has parked its partial tile
    const uint32_t S = gridDim.z, rank = cluster_ctarank();
    const int rows_per = kGemmBM / int(S);
    const uint32_t stg_u32 = smem_u32(smem);
    for (int u = threadIdx.x; u < rows_per * (BN / 4); u += kGemmThreads) {
      const int r = int(rank) * rows_per + u / (BN / 4), c4 = u % (BN / 4);
      const uint32_t addr = stg_u32 + uint32_t(r * kStgStride + 4 * c4) * 4u;
      float4 acc = ld_dsmem_v4(addr, 0);
      for (uint32_t sidx = 1; sidx < S; ++sidx) {
        const float4 t = ld_dsmem_v4(addr, sidx);
        acc.x += t.x;
        acc.y += t.y;
        acc.z += t.z;
        acc.w += t.w;
      }
      const int m = m_tile * kGemmBM + r;
      if (m < g.M) {
        const int n0 = n_tile * BN + 4 * c4;
        if (g.out != nullptr) {
          if (g.bias != nullptr) {
            const uint2 bw = *reinterpret_cast<const uint2*>(g.bias + n0);
            acc.x += bf16lo_to_f32(bw.x);
            acc.y += bf16hi_to_f32(bw.x);
            acc.z += bf16lo_to_f32(bw.y);
            acc.w += bf16hi_to_f32(bw.y);
          }
          if (g.gelu) {
            acc.x = gelu_erf(acc.x);
            acc.y = gelu_erf(acc.y);
            acc.z = gelu_erf(acc.z);
            acc.w = gelu_erf(acc.w);
          }
          const __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
          *reinterpret_cast<uint2*>(g.out + size_t(m) * g.N + n0) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        } else {
          *reinterpret_cast<float4*>(g.partial + size_t(m) * g.N + n0) = acc;  // one reduced partial
        }
      }
    }
    cluster_sync_all();  // nobody leaves while a peer may still read its tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

// ---- K5p: persistent form of K5 for GEMMs of several waves of tiles -----------------------------
// One CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest, so the CTAs of one
// round share A tiles in L2). Two 128-column TMEM accumulators: the MMA thread starts the k-loop of
// tile i+1 while the sixteen epilogue warps (four per TMEM lane quarter, a quarter of the columns each) drain
// tile i — bias / erf-GELU / bf16 packing no longer sits on the tensor pipe's critical path. (Eight warps, two
// per scheduler, could not hide their own latencies: issue slots 36 % busy while a GELU tile took 2.8x its MMA
// time.) The TMA producer runs ahead across tile boundaries through the same ring. bf16 output or one fp32 partial.
constexpr int kPersEpiWarps = 16;
constexpr int kPersThreads = 64 + 32 * kPersEpiWarps;
constexpr uint32_t kPersBiasBytes = kPersEpiWarps * 64 * 4;  // warp-private bias strips
// BN = 256 halves the L2 -> shared-memory bytes per flop (the bound of 128 x 128 tiles at ~0.65 PF/s):
// 48 KB per k-block for a 128 x 256 x 64 MMA block, 4 stages, both accumulators fill the 512 TMEM columns
template <int BN>
struct PersCfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr uint32_t kStageBytes = kGemmABytes + BN * kBK * 2;
  static constexpr uint32_t kSmem = kStages * kStageBytes + 256 + kPersBiasBytes + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kPersThreads, 1)
    k5_linear_persistent_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                                const GemmArgs g, int m_tiles, int n_tiles) {
  constexpr int kPersBN = BN, kPersStages = PersCfg<BN>::kStages;
  constexpr uint32_t kPersStageBytes = PersCfg<BN>::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPersStages * kPersStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kPersStages;
  uint64_t* tfull_bar = bars + 2 * kPersStages;       // [2] accumulator ready
  uint64_t* tempty_bar = bars + 2 * kPersStages + 2;  // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPersStages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;");
  const int total = m_tiles * n_tiles, kblocks = g.K / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < kPersStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], kPersEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * kPersBN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_a = policy_evict_last(), pol_w = policy_evict_normal();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int m_tile = t / n_tiles, n_tile = t % n_tiles;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kPersStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], kPersStageBytes);
          tma_load_2d(sa, &tm_a, &full_bar[stage], kb * kBK, m_tile * kGemmBM, pol_a);
          tma_load_2d(sa + kGemmABytes, &tm_w, &full_bar[stage], kb * kBK, n_tile * kPersBN, pol_w);
          if (++stage == kPersStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, kPersBN);
      int stage = 0;
      uint32_t phase = 0;
      int i = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
        const int buf = i & 1;
        mbar_wait(&tempty_bar[buf], ((i >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(buf * kPersBN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kPersStageBytes);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + kGemmABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_bf16(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == kPersStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // epilogue warp: TMEM lane quarter = warp % 4 (hardware rule), column quarter by warp group
    constexpr int kCW = kPersBN / 4;
    const int quarter = warp & 3, part = (warp - 2) >> 2;
    float* bias_s = reinterpret_cast<float*>(smem + kPersStages * kPersStageBytes + 256) + (warp - 2) * 64;
    int i = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
      const int m_tile = t / n_tiles, n_tile = t % n_tiles, buf = i & 1;
      const int n0 = n_tile * kPersBN + part * kCW;
      if (g.out != nullptr) k5_stage_bias<kCW>(g, n0, bias_s, lane);
      mbar_wait(&tfull_bar[buf], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf * kPersBN + part * kCW);
      const int m = m_tile * kGemmBM + quarter * 32 + lane;
      uint32_t v[kCW / 32][32];
#pragma unroll
      for (int c = 0; c < kCW / 32; ++c) tmem_ld_32x32(t_addr + c * 32, v[c]);
      tmem_ld_wait();
      tc_fence_before();  // this warp's share of the accumulator is in registers: release it now
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      if (m < g.M) {
#pragma unroll
        for (int c = 0; c < kCW / 32; ++c) k5_write_cols32(g, v[c], bias_s + c * 32, m, n0 + c * 32);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * kPersBN>(tmem_base);
  }
}

// ---- K5 pair form: 256 x 256 tiles on a CTA pair (cta_group::2) ------------------------------------
// The persistent kernel above is bound by the L2 -> shared-memory path (128 x 256 tiles: 48 KB per
// 128 x 256 x 64 MMA block). Two CTAs of a cluster share one M = 256 x N = 256 MMA stream instead: each
// stages its own 128 rows of A and 128 of the 256 rows of the W tile (32 KB per k-block and SM for the
// same flops, 6 stages), the leader's single thread issues tcgen05.mma.cta_group::2, commits are
// multicast to both CTAs, and each CTA drains the 128 x 256 accumulator of its own rows with sixteen
// epilogue warps while the next tile accumulates in the other half of TMEM. Same structure as the
// CTA-pair scan kernel (k2_batch2.cu). Tiles: (256-row pair, 256-column block), n fastest.
constexpr int kPairBN = 256, kPairStages = 6;
constexpr uint32_t kPairStageBytes = kGemmABytes + (kPairBN / 2) * kBK * 2;  // 16 KB A + 16 KB half of W
constexpr uint32_t kPairSmem = kPairStages * kPairStageBytes + 256 + kPersBiasBytes + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPersThreads, 1)
    k5_linear_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                          const GemmArgs g, int m_pairs, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPairStages * kPairStageBytes);
  uint64_t* full_bar = bars;                          // [stages] used in the leader only
  uint64_t* empty_bar = bars + kPairStages;           // [stages] each CTA its own (multicast commit)
  uint64_t* tfull_bar = bars + 2 * kPairStages;       // [2]      each CTA its own (multicast commit)
  uint64_t* tempty_bar = bars + 2 * kPairStages + 2;  // [2]      used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;");
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int total = m_pairs * n_tiles, kblocks = g.K / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full_bar[s], 1);   // leader producer's arrive.expect_tx (+ the bytes of both CTAs)
      mbar_init(&empty_bar[s], 1);  // one multicast commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 2 * kPersEpiWarps);  // epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<2 * kPairBN>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers exist before anything is signalled remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_a = policy_evict_last(), pol_w = policy_evict_normal();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair_id; t < total; t += n_pairs) {
        const int m_pair = t / n_tiles, n_tile = t % n_tiles;
        const int a_row0 = (m_pair * 2 + int(cta_rank)) * kGemmBM;
        const int w_row0 = n_tile * kPairBN + int(cta_rank) * (kPairBN / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kPairStageBytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
          tma_load_2d_pair(sa, &tm_a, &full_bar[stage], kb * kBK, a_row0, pol_a);
          tma_load_2d_pair(sa + kGemmABytes, &tm_w, &full_bar[stage], kb * kBK, w_row0, pol_w);
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kGemmBM, kPairBN);
      int stage = 0;
      uint32_t phase = 0;
      int i = 0;
      for (int t = pair_id; t < total; t += n_pairs, ++i) {
        const int buf = i & 1;
        mbar_wait(&tempty_bar[buf], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(buf * kPairBN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kPairStageBytes);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + kGemmABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_bf16_pair(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair(&tfull_bar[buf]);
      }
    }
  } else {
    constexpr int kCW = kPairBN / 4;  // 64 columns per epilogue warp
    const int quarter = warp & 3, part = (warp - 2) >> 2;
    float* bias_s = reinterpret_cast<float*>(smem + kPairStages * kPairStageBytes + 256) + (warp - 2) * 64;
    int i = 0;
    for (int t = pair_id; t < total; t += n_pairs, ++i) {
      const int m_pair = t / n_tiles, n_tile = t % n_tiles, buf = i & 1;
      const int n0 = n_tile * kPairBN + part * kCW;
      if (g.out != nullptr) k5_stage_bias<kCW>(g, n0, bias_s, lane);
      mbar_wait(&tfull_bar[buf], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf * kPairBN + part * kCW);
      const int m = (m_pair * 2 + int(cta_rank)) * kGemmBM + quarter * 32 + lane;
      uint32_t v[2][32];
      tmem_ld_32x32(t_addr, v[0]);
      tmem_ld_32x32(t_addr + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before();  // accumulator share in registers: hand it back to the leader's MMA thread
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[buf], 0);
      if (m < g.M) {
        k5_write_cols32(g, v[0], bias_s, m, n0);
        k5_write_cols32(g, v[1], bias_s + 32, m, n0 + 32);
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the pair's MMAs / remote arrives may still land
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<2 * kPairBN>(tmem_base);
  }
}

// ---- K6: block-causal attention, head_dim 64 ---------------------------------------------------
// grid (b * heads, groups, 32-row blocks of a group). Group g attends to the keys of groups 0..g
// (get_mask), so there is no mask inside a CTA apart from the padding of the last 64-key chunk.
// qkv [M, 3*d] bf16 (q | k | v, head h at columns h*64 inside each third), out [M, d] bf16.
// The shapes (25 query rows x <= 250 keys x 64 dims per CTA) are far below a tcgen05 tile (M >= 64
// per instruction, one issuing thread, TMEM round trips), so this kernel uses warp-level
// mma.sync.m16n8k16 bf16 with fp32 accumulation: 4 warps = 2 query tiles of 16 rows x 2 key-chunk
// parities; each warp walks its 64-key chunks with an online softmax (scores and the running
// output never leave registers; P is re-used as the A operand of P V straight from the score
// accumulators), the two parities are merged through shared memory at the end. K and V rows of the
// visible prefix are staged once per CTA with cp.async as bf16, rows padded to 144 B so the
// B-fragment loads (32-bit for K, ldmatrix.trans for V) are bank-conflict free.
constexpr int kAttnThreads = 128;
constexpr int kHeadDim = 64;
constexpr int kKVStride = 72;   // bf16 elements per staged K / V row
constexpr int kKeyChunk = 64;
constexpr int kAttnMaxT = 704;  // staged K + V of the longest prefix must fit in shared memory

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// one work item: query rows [32*zblk, 32*zblk+32) of group g of (sample b, head h). Called by every
// thread of the CTA (it contains CTA barriers); the first four warps compute, all threads help staging.
__device__ __forceinline__ void attention_item(uint8_t* smem_attn, const __nv_bfloat16* __restrict__ qkv,
                                               __nv_bfloat16* __restrict__ out, int T, int d_model, int group_tokens,
                                               float scale_log2e, int b, int h, int g, int zblk, int n_threads,
                                               bool compact_out = false) {
  const int nk = (g + 1) * group_tokens;               // visible keys
  const int n_chunks = (nk + kKeyChunk - 1) / kKeyChunk;
  const int rows_pad = n_chunks * kKeyChunk;
  __nv_bfloat16* ks = reinterpret_cast<__nv_bfloat16*>(smem_attn);  // [rows_pad][72]
  __nv_bfloat16* vs = ks + size_t(rows_pad) * kKVStride;            // [rows_pad][72]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool worker = warp < 4;
  const int quad = lane >> 2, qlane = lane & 3;
  const size_t row_stride = size_t(3) * d_model;
  const __nv_bfloat16* base = qkv + size_t(b) * T * row_stride;

  // stage K and V rows of this head (16-byte pieces); rows of the last chunk beyond nk are zeroed
  // (their scores are masked, but 0 * garbage must not produce NaN in P V)
  for (int i = tid; i < rows_pad * 8; i += n_threads) {
    const int r = i >> 3, c = i & 7;
    __nv_bfloat16* dk = ks + size_t(r) * kKVStride + 8 * c;
    __nv_bfloat16* dv = vs + size_t(r) * kKVStride + 8 * c;
    if (r < nk) {
      const __nv_bfloat16* src = base + size_t(r) * row_stride + h * kHeadDim + 8 * c;
      cp_async_16(dk, src + d_model);
      cp_async_16(dv, src + 2 * d_model);
    } else {
      *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // query fragments (A operand, 16 rows x 64 dims) straight from global while the copies fly
  const int mt = warp & 1, par = warp >> 1;
  const int r0 = zblk * 32 + mt * 16 + quad, r1 = r0 + 8;
  const bool valid0 = worker && r0 < group_tokens, valid1 = worker && r1 < group_tokens;
  const __nv_bfloat16* q0 = base + size_t(g * group_tokens + (valid0 ? r0 : 0)) * row_stride + h * kHeadDim;
  const __nv_bfloat16* q1 = base + size_t(g * group_tokens + (valid1 ? r1 : 0)) * row_stride + h * kHeadDim;
  uint32_t qa[4][4];
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    const int col = 16 * kt + 2 * qlane;
    // plain loads: inside the fused kernel qkv is rewritten every layer, the read-only path could go stale
    qa[kt][0] = valid0 ? *reinterpret_cast<const uint32_t*>(q0 + col) : 0u;
    qa[kt][1] = valid1 ? *reinterpret_cast<const uint32_t*>(q1 + col) : 0u;
    qa[kt][2] = valid0 ? *reinterpret_cast<const uint32_t*>(q0 + col + 8) : 0u;
    qa[kt][3] = valid1 ? *reinterpret_cast<const uint32_t*>(q1 + col + 8) : 0u;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // running max / per-lane partial sums of rows r0, r1
  const uint32_t vs_u32 = smem_u32(vs);
  for (int ch = worker ? par : n_chunks; ch < n_chunks; ch += 2) {
    const int key0 = ch * kKeyChunk;
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      const __nv_bfloat16* kp = ks + size_t(key0 + 8 * j + quad) * kKVStride + 2 * qlane;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt)
        mma_bf16_16816(s[j], qa[kt], *reinterpret_cast<const uint32_t*>(kp + 16 * kt),
                       *reinterpret_cast<const uint32_t*>(kp + 16 * kt + 8));
    }
    if (key0 + kKeyChunk > nk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = key0 + 8 * j + 2 * qlane;
        if (key >= nk) s[j][0] = s[j][2] = -INFINITY;
        if (key + 1 >= nk) s[j][1] = s[j][3] = -INFINITY;
      }
    }
    float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c0 = fmaxf(c0, fmaxf(s[j][0], s[j][1]));
      c1 = fmaxf(c1, fmaxf(s[j][2], s[j][3]));
    }
    c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1));
    c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2));
    c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1));
    c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
    const float n0 = fmaxf(m0, c0), n1 = fmaxf(m1, c1);  // finite: every chunk holds a visible key
    const float a0 = exp2f((m0 - n0) * scale_log2e), a1 = exp2f((m1 - n1) * scale_log2e);
    m0 = n0;
    m1 = n1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f((s[j][0] - m0) * scale_log2e);
      s[j][1] = exp2f((s[j][1] - m0) * scale_log2e);
      s[j][2] = exp2f((s[j][2] - m1) * scale_log2e);
      s[j][3] = exp2f((s[j][3] - m1) * scale_log2e);
      rs0 += s[j][0] + s[j][1];
      rs1 += s[j][2] + s[j][3];
    }
    l0 = l0 * a0 + rs0;
    l1 = l1 * a1 + rs1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= a0;
      o[j][1] *= a0;
      o[j][2] *= a1;
      o[j][3] *= a1;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      const uint32_t vrow = vs_u32 + uint32_t((key0 + 16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)) * kKVStride +
                                              8 * (lane >> 4)) * 2u;
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // 16 output dims per step
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, vrow + uint32_t(16 * jp) * 2u);
        mma_bf16_16816(o[2 * jp], pa, vb[0], vb[1]);
        mma_bf16_16816(o[2 * jp + 1], pa, vb[2], vb[3]);
      }
    }
  }

  // merge the two key parities: warps 2, 3 park (o, m, l) in shared memory (K is no longer needed)
  __syncthreads();
  float* xch = reinterpret_cast<float*>(smem_attn) + mt * (36 * 32);
  if (worker && par == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) xch[(4 * j + e) * 32 + lane] = o[j][e];
    xch[32 * 32 + lane] = m0;
    xch[33 * 32 + lane] = m1;
    xch[34 * 32 + lane] = l0;
    xch[35 * 32 + lane] = l1;
  }
  __syncthreads();
  if (worker && par == 0) {
    const float pm0 = xch[32 * 32 + lane], pm1 = xch[33 * 32 + lane];
    const float t0 = fmaxf(m0, pm0), t1 = fmaxf(m1, pm1);
    const float f0 = exp2f((m0 - t0) * scale_log2e), g0 = exp2f((pm0 - t0) * scale_log2e);
    const float f1 = exp2f((m1 - t1) * scale_log2e), g1 = exp2f((pm1 - t1) * scale_log2e);
    l0 = l0 * f0 + xch[34 * 32 + lane] * g0;
    l1 = l1 * f1 + xch[35 * 32 + lane] * g1;
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    // compact_out: only one group is computed and its rows are written densely as [b][group_tokens]
    const size_t orow = compact_out ? size_t(b) * group_tokens : size_t(b) * T + g * group_tokens;
    __nv_bfloat16* o0 = out + (orow + r0) * d_model + h * kHeadDim + 2 * qlane;
    __nv_bfloat16* o1 = out + (orow + r1) * d_model + h * kHeadDim + 2 * qlane;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x0 = (o[j][0] * f0 + xch[(4 * j + 0) * 32 + lane] * g0) * i0;
      const float x1 = (o[j][1] * f0 + xch[(4 * j + 1) * 32 + lane] * g0) * i0;
      const float y0 = (o[j][2] * f1 + xch[(4 * j + 2) * 32 + lane] * g1) * i1;
      const float y1 = (o[j][3] * f1 + xch[(4 * j + 3) * 32 + lane] * g1) * i1;
      if (valid0) *reinterpret_cast<uint32_t*>(o0 + 8 * j) = pack_bf16x2(x0, x1);
      if (valid1) *reinterpret_cast<uint32_t*>(o1 + 8 * j) = pack_bf16x2(y0, y1);
    }
  }
}

__global__ void __launch_bounds__(kAttnThreads, 3)
    k6_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T,
                        int d_model, int heads, int group_tokens, float scale_log2e, int only_group) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int bh = blockIdx.x;
  // only_group >= 0 (grid.y == 1): just that group's query rows, written compactly (last layer of predict)
  const int g = only_group >= 0 ? only_group : int(gridDim.y) - 1 - int(blockIdx.y) /* longest key prefixes first */;
  attention_item(smem_attn, qkv, out, T, d_model, group_tokens, scale_log2e, bh / heads, bh % heads, g,
                 int(blockIdx.z), kAttnThreads, only_group >= 0);
}

// ---- K6w: the same attention for batches that fill the GPU with one CTA per (sample, head) ------------
// k6_attention_kernel re-stages the visible K / V prefix in every (sample, head, group) CTA: 5.5x the bytes of
// one head per (sample, head) and a stage -> wait -> compute -> merge sequence per 25 query rows with three small
// CTAs per SM to hide it (b = 16: 2 560 CTAs, 32 us per layer, tensor pipe 21 % active, issue slots 42 % busy,
// profiles/r2_k6_attention_b16.txt). Here a 256-thread CTA owns ALL T query rows of one (sample, head): K and V are
// staged once, the T rows are cut into 16-row MMA tiles that ignore group borders, and warp w takes tiles
// w, 15 - w, 16 + w, ... (short and long key prefixes paired, 5 or 6 key chunks per warp at T = 250). A tile walks
// the 64-key chunks its LAST row may see; rows of the earlier group inside a tile are masked per element in the
// chunks beyond their own prefix (every row sees chunk 0, so the running maximum is finite from the first chunk
// on). No parity split, no merge through shared memory, one barrier per CTA.
constexpr int kAttnWideThreads = 256;

__global__ void __launch_bounds__(kAttnWideThreads, 2)
    k6_attention_wide_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int d_model,
                             int heads, int group_tokens, float scale_log2e) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int rows_pad = (T + kKeyChunk - 1) / kKeyChunk * kKeyChunk;
  __nv_bfloat16* ks = reinterpret_cast<__nv_bfloat16*>(smem_attn);  // [rows_pad][72]
  __nv_bfloat16* vs = ks + size_t(rows_pad) * kKVStride;            // [rows_pad][72]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int quad = lane >> 2, qlane = lane & 3;
  const size_t row_stride = size_t(3) * d_model;
  const __nv_bfloat16* base = qkv + size_t(b) * T * row_stride + h * kHeadDim;

  for (int i = tid; i < rows_pad * 8; i += kAttnWideThreads) {
    const int r = i >> 3, c = i & 7;
    __nv_bfloat16* dk = ks + size_t(r) * kKVStride + 8 * c;
    __nv_bfloat16* dv = vs + size_t(r) * kKVStride + 8 * c;
    if (r < T) {
      const __nv_bfloat16* src = base + size_t(r) * row_stride + 8 * c;
      cp_async_16(dk, src + d_model);
      cp_async_16(dv, src + 2 * d_model);
    } else {  // padding rows of the last chunk: masked scores, but 0 * garbage must not produce NaN in P V
      *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int n_tiles = (T + 15) / 16;
  const uint32_t vs_u32 = smem_u32(vs);
  constexpr int kWarps = kAttnWideThreads / 32;
  for (int k = 0;; ++k) {
    const int t = (k & 1) ? kWarps * (k + 1) - 1 - warp : kWarps * k + warp;
    if (kWarps * (k & ~1) >= n_tiles) break;   // both tiles of this pair of rounds are past the end for every warp
    if (t >= n_tiles) continue;
    const int r0 = 16 * t + quad, r1 = r0 + 8;
    const bool valid0 = r0 < T, valid1 = r1 < T;
    const int last_row = min(16 * t + 15, T - 1);
    const int vis_min = (16 * t / group_tokens + 1) * group_tokens;     // keys the tile's first row sees
    const int vis_max = (last_row / group_tokens + 1) * group_tokens;   // keys its last row sees
    const int vis0 = ((valid0 ? r0 : last_row) / group_tokens + 1) * group_tokens;
    const int vis1 = ((valid1 ? r1 : last_row) / group_tokens + 1) * group_tokens;
    const int n_chunks = (vis_max + kKeyChunk - 1) / kKeyChunk;
    const __nv_bfloat16* q0 = base + size_t(valid0 ? r0 : 0) * row_stride;
    const __nv_bfloat16* q1 = base + size_t(valid1 ? r1 : 0) * row_stride;
    uint32_t qa[4][4];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const int col = 16 * kt + 2 * qlane;
      qa[kt][0] = valid0 ? *reinterpret_cast<const uint32_t*>(q0 + col) : 0u;
      qa[kt][1] = valid1 ? *reinterpret_cast<const uint32_t*>(q1 + col) : 0u;
      qa[kt][2] = valid0 ? *reinterpret_cast<const uint32_t*>(q0 + col + 8) : 0u;
      qa[kt][3] = valid1 ? *reinterpret_cast<const uint32_t*>(q1 + col + 8) : 0u;
    }
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int key0 = ch * kKeyChunk;
      float sc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
        const __nv_bfloat16* kp = ks + size_t(key0 + 8 * j + quad) * kKVStride + 2 * qlane;
#pragma unroll
        for (int kt = 0; kt < 4; ++kt)
          mma_bf16_16816(sc[j], qa[kt], *reinterpret_cast<const uint32_t*>(kp + 16 * kt),
                         *reinterpret_cast<const uint32_t*>(kp + 16 * kt + 8));
      }
      if (key0 + kKeyChunk > vis_min) {  // warp-uniform
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int key = key0 + 8 * j + 2 * qlane;
          if (key >= vis0) sc[j][0] = -INFINITY;
          if (key + 1 >= vis0) sc[j][1] = -INFINITY;
          if (key >= vis1) sc[j][2] = -INFINITY;
          if (key + 1 >= vis1) sc[j][3] = -INFINITY;
        }
      }
      float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c0 = fmaxf(c0, fmaxf(sc[j][0], sc[j][1]));
        c1 = fmaxf(c1, fmaxf(sc[j][2], sc[j][3]));
      }
      c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1));
      c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2));
      c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1));
      c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
      const float n0 = fmaxf(m0, c0), n1 = fmaxf(m1, c1);  // finite: chunk 0 holds a visible key for every row
      const float a0 = exp2f((m0 - n0) * scale_log2e), a1 = exp2f((m1 - n1) * scale_log2e);
      m0 = n0;
      m1 = n1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j][0] = exp2f((sc[j][0] - m0) * scale_log2e);
        sc[j][1] = exp2f((sc[j][1] - m0) * scale_log2e);
        sc[j][2] = exp2f((sc[j][2] - m1) * scale_log2e);
        sc[j][3] = exp2f((sc[j][3] - m1) * scale_log2e);
        rs0 += sc[j][0] + sc[j][1];
        rs1 += sc[j][2] + sc[j][3];
      }
      l0 = l0 * a0 + rs0;
      l1 = l1 * a1 + rs1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j][0] *= a0;
        o[j][1] *= a0;
        o[j][2] *= a1;
        o[j][3] *= a1;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
        uint32_t pa[4];
        pa[0] = pack_bf16x2(sc[2 * kk][0], sc[2 * kk][1]);
        pa[1] = pack_bf16x2(sc[2 * kk][2], sc[2 * kk][3]);
        pa[2] = pack_bf16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
        const uint32_t vrow = vs_u32 + uint32_t((key0 + 16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1)) * kKVStride +
                                                8 * (lane >> 4)) * 2u;
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {  // 16 output dims per step
          uint32_t vb[4];
          ldmatrix_x4_trans(vb, vrow + uint32_t(16 * jp) * 2u);
          mma_bf16_16816(o[2 * jp], pa, vb[0], vb[1]);
          mma_bf16_16816(o[2 * jp + 1], pa, vb[2], vb[3]);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    __nv_bfloat16* o0 = out + (size_t(b) * T + r0) * d_model + h * kHeadDim + 2 * qlane;
    __nv_bfloat16* o1 = out + (size_t(b) * T + r1) * d_model + h * kHeadDim + 2 * qlane;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (valid0) *reinterpret_cast<uint32_t*>(o0 + 8 * j) = pack_bf16x2(o[j][0] * i0, o[j][1] * i0);
      if (valid1) *reinterpret_cast<uint32_t*>(o1 + 8 * j) = pack_bf16x2(o[j][2] * i1, o[j][3] * i1);
    }
  }
}

// ---- K7: y = LayerNorm(resid + sum_s partial[s] + bias) * gamma + beta ---------------------------
// One WARP per row, 8 rows per CTA: lane l owns the 8-element chunks l, l+32, ... of the row (16-byte bf16 /
// 32-byte fp32 pieces, coalesced across the warp), mean and variance are two xor-shuffle reductions — no shared
// memory, no CTA barrier (the first version ran one 128-thread CTA per row with two barrier-separated block
// reductions: 13 us for 4 000 rows, most of it CTA launch and barrier latency).
constexpr int kMaxSplits = 8;
constexpr int kLnRowsPerCta = 8;
constexpr int kLnWarpRowsMin = 2000;  // rows from which the warp-per-row form is used (see the CTA-per-row form below)

__global__ void __launch_bounds__(32 * kLnRowsPerCta)
    k7_add_layernorm_kernel(const __nv_bfloat16* __restrict__ resid, const float* __restrict__ partial,
                            int splits, const __nv_bfloat16* __restrict__ bias,
                            const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                            __nv_bfloat16* __restrict__ out, int M, int d, float eps, int resid_T, int resid_L) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * kLnRowsPerCta + warp;
  if (row >= M) return;  // warp-uniform
  // resid_T > 0: the rows of this launch are the LAST resid_L tokens of each resid_T-token sample, stored
  // compactly, while the residual stream still has all tokens (last layer of predict)
  const size_t rrow = resid_T > 0 ? size_t(row / resid_L) * resid_T + (resid_T - resid_L) + row % resid_L : size_t(row);
  const int chunks = d >> 8;  // 8-element chunks per lane: d / (32 * 8), 1..4
  float x[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < chunks) {
      const int c = (j * 32 + lane) * 8;
      const uint4 rv = *reinterpret_cast<const uint4*>(resid + rrow * d + c);
      const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bias + c));
      const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        x[j][2 * e] = bf16lo_to_f32(rw[e]) + bf16lo_to_f32(bw[e]);
        x[j][2 * e + 1] = bf16hi_to_f32(rw[e]) + bf16hi_to_f32(bw[e]);
      }
#pragma unroll
      for (int sp = 0; sp < kMaxSplits; ++sp) {
        if (sp < splits) {
          const float4* pp = reinterpret_cast<const float4*>(partial + (size_t(sp) * M + row) * d + c);
          const float4 a = pp[0], b = pp[1];
          x[j][0] += a.x; x[j][1] += a.y; x[j][2] += a.z; x[j][3] += a.w;
          x[j][4] += b.x; x[j][5] += b.y; x[j][6] += b.z; x[j][7] += b.w;
        }
      }
    }
  }
  float s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < chunks)
#pragma unroll
      for (int e = 0; e < 8; ++e) s1 += x[j][e];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, off);
  const float mean = s1 / d;
  float s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < chunks)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float t = x[j][e] - mean;
        s2 = fmaf(t, t, s2);
      }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  const float rstd = rsqrtf(s2 / d + eps);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j < chunks) {
      const int c = (j * 32 + lane) * 8;
      const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gamma + c));
      const uint4 ev = __ldg(reinterpret_cast<const uint4*>(beta + c));
      const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, ew[4] = {ev.x, ev.y, ev.z, ev.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = (x[j][2 * e] - mean) * rstd * bf16lo_to_f32(gw[e]) + bf16lo_to_f32(ew[e]);
        const float b = (x[j][2 * e + 1] - mean) * rstd * bf16hi_to_f32(gw[e]) + bf16hi_to_f32(ew[e]);
        const __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
        o[e] = *reinterpret_cast<const uint32_t*>(&r);
      }
      *reinterpret_cast<uint4*>(out + size_t(row) * d + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// CTA-per-row form (d/8 threads, 8 elements each, two barrier-separated block reductions): four times the
// parallelism inside a row and one CTA per row — the shorter latency chain when there are only a few hundred
// rows (one sample: 250 rows; 4.4 us against 9 us for the warp-per-row form, which wins from ~2 000 rows on).
__device__ __forceinline__ float block_sum(float v, float* red, int warp, int lane, int nwarps) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nwarps; ++w) t += red[w];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(128)
    k7_add_layernorm_cta_kernel(const __nv_bfloat16* __restrict__ resid, const float* __restrict__ partial,
                            int splits, const __nv_bfloat16* __restrict__ bias,
                            const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                            __nv_bfloat16* __restrict__ out, int M, int d, float eps, int resid_T, int resid_L) {
  __shared__ float red[4];
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int row = blockIdx.x;
  const int c = tid * 8;
  // resid_T > 0: the rows of this launch are the LAST resid_L tokens of each resid_T-token sample, stored
  // compactly, while the residual stream still has all tokens (last layer of predict)
  const size_t rrow = resid_T > 0 ? size_t(row / resid_L) * resid_T + (resid_T - resid_L) + row % resid_L : size_t(row);
  const uint4 rv = *reinterpret_cast<const uint4*>(resid + rrow * d + c);
  const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bias + c));
  const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gamma + c));
  const uint4 ev = __ldg(reinterpret_cast<const uint4*>(beta + c));
  float4 pa[kMaxSplits], pb[kMaxSplits];
#pragma unroll
  for (int s = 0; s < kMaxSplits; ++s) {
    if (s < splits) {
      const float4* pp = reinterpret_cast<const float4*>(partial + (size_t(s) * M + row) * d + c);
      pa[s] = pp[0];
      pb[s] = pp[1];
    }
  }
  const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
  float x[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    x[2 * j] = bf16lo_to_f32(rw[j]) + bf16lo_to_f32(bw[j]);
    x[2 * j + 1] = bf16hi_to_f32(rw[j]) + bf16hi_to_f32(bw[j]);
  }
#pragma unroll
  for (int s = 0; s < kMaxSplits; ++s) {
    if (s < splits) {
      x[0] += pa[s].x; x[1] += pa[s].y; x[2] += pa[s].z; x[3] += pa[s].w;
      x[4] += pb[s].x; x[5] += pb[s].y; x[6] += pb[s].z; x[7] += pb[s].w;
    }
  }
  float s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s1 += x[j];
  const float mean = block_sum(s1, red, warp, lane, nwarps) / d;
  float s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float t = x[j] - mean;
    s2 = fmaf(t, t, s2);
  }
  const float rstd = rsqrtf(block_sum(s2, red, warp, lane, nwarps) / d + eps);
  const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, ew[4] = {ev.x, ev.y, ev.z, ev.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = (x[2 * j] - mean) * rstd * bf16lo_to_f32(gw[j]) + bf16lo_to_f32(ew[j]);
    const float b = (x[2 * j + 1] - mean) * rstd * bf16hi_to_f32(gw[j]) + bf16hi_to_f32(ew[j]);
    const __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
    o[j] = *reinterpret_cast<const uint32_t*>(&r);
  }
  *reinterpret_cast<uint4*>(out + size_t(row) * d + c) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---- host side ---------------------------------------------------------------------------------
static cudaError_t launch_pdl(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args,
                              unsigned cluster_z = 1) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cluster_z > 1) {  // the K splits of one tile (grid.z) share a cluster
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 1;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = cluster_z;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelExC(&cfg, fn, args);
}

template <int BN, int STAGES, bool REDUCE = false>
static cudaError_t launch_k5_bn(const CUtensorMap& tm_a, const CUtensorMap& tm_w, const GemmArgs& g, int splits,
                                cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k5_linear_kernel<BN, STAGES, REDUCE>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, int(GemmCfg<BN, STAGES>::kSmem));
  if (e != cudaSuccess) return e;
  void* args[] = {const_cast<CUtensorMap*>(&tm_a), const_cast<CUtensorMap*>(&tm_w), const_cast<GemmArgs*>(&g)};
  return launch_pdl(reinterpret_cast<const void*>(k5_linear_kernel<BN, STAGES, REDUCE>),
                    dim3(unsigned(g.N / BN), unsigned((g.M + kGemmBM - 1) / kGemmBM), unsigned(splits)),
                    dim3(kGemmThreads), GemmCfg<BN, STAGES>::kSmem, st, args, REDUCE ? unsigned(splits) : 1u);
}

cudaError_t launch_k5_linear(const void* a_bf16, int a_rows_alloc, const void* w_bf16, int M, int N, int K,
                             const void* bias, bool gelu, void* out_bf16, float* partial, int splits,
                             cudaStream_t st) {
  if (N % 128 != 0 || K % kBK != 0 || M < 1 || splits < 1) return cudaErrorInvalidValue;
  const int m_tiles = (M + kGemmBM - 1) / kGemmBM;
  const int ctas128 = (N / 128) * m_tiles * splits;
  const bool reduce = out_bf16 != nullptr && splits > 1;
  int bn = (ctas128 < 120 && !reduce) ? 64 : 128;  // fill the 148 SMs when tiles are few
  static const int bn_knob = [] {  // MRAG_K5_BN: tuning knob (scripts/cama_gemm_bench.py), read once
    const char* v = getenv("MRAG_K5_BN");
    return v ? atoi(v) : 0;
  }();
  if ((bn_knob == 64 || bn_knob == 128) && !reduce) bn = bn_knob;
  CUtensorMap tm_a, tm_w;
  if (!make_tmap(&tm_a, a_bf16, a_rows_alloc, K, kGemmBM) || !make_tmap(&tm_w, w_bf16, N, K, bn))
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
  cudaError_t e;
  static const bool persistent_ok = [] {  // MRAG_K5_PERSISTENT=0 falls back to one tile per CTA (A/B runs)
    const char* v = getenv("MRAG_K5_PERSISTENT");
    return !(v && atoi(v) == 0);
  }();
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // persistent kernel when the tile count makes >= 2 full rounds over the SMs (wave quantisation would
  // otherwise cost more than the overlap gains): 128 x 256 tiles when they still do, else 128 x 128
  const int tiles256 = (N % 256 == 0) ? (N / 256) * m_tiles : 0;
  static const bool pair_ok = [] {  // MRAG_K5_PAIR=0 keeps the single-CTA persistent kernel (A/B runs)
    const char* v = getenv("MRAG_K5_PAIR");
    return !(v && atoi(v) == 0);
  }();
  const int m_pairs = (M + 2 * kGemmBM - 1) / (2 * kGemmBM);
  const int pair_tiles = (N % kPairBN == 0) ? (N / kPairBN) * m_pairs : 0;
  // The pair kernel also takes GEMMs of ONE round of 256 x 256 tiles when they occupy half of the pairs (N = 1024
  // from b = 10 on: out-proj, FFN2; b = 16: 616 -> 590 us per forward): 128 x 128 tiles move twice the bytes per flop from L2 into shared memory, and that
  // path, not the tensor pipe, bounds them (FFN2 at b = 16: 537 MB through it in 37 us).
  static const int pair_min = [] {  // MRAG_K5_PAIR_MIN: smallest tile count for the pair kernel (A/B runs)
    const char* v = getenv("MRAG_K5_PAIR_MIN");
    return v ? atoi(v) : 0;
  }();
  const int pair_floor = pair_min > 0 ? pair_min : sms / 4;  // half of the pairs busy (b = 10: 520 -> 503 us per forward)
  if (!reduce && splits == 1 && persistent_ok && pair_ok && pair_tiles >= pair_floor) {
    if (!make_tmap(&tm_w, w_bf16, N, K, kPairBN / 2)) return cudaErrorInvalidValue;
    e = cudaFuncSetAttribute(k5_linear_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kPairSmem));
    if (e != cudaSuccess) return e;
    int mp = m_pairs, nt = N / kPairBN;
    const int pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
    void* args[] = {&tm_a, &tm_w, &g, &mp, &nt};
    e = launch_pdl(reinterpret_cast<const void*>(k5_linear_pair_kernel), dim3(unsigned(2 * pairs)),
                   dim3(kPersThreads), kPairSmem, st, args);
  } else if (!reduce && splits == 1 && persistent_ok && (tiles256 >= 2 * sms || ctas128 >= 2 * sms)) {
    const bool wide = tiles256 >= 2 * sms;
    const int pbn = wide ? 256 : 128;
    if (!make_tmap(&tm_w, w_bf16, N, K, pbn)) return cudaErrorInvalidValue;
    const void* fn = wide ? reinterpret_cast<const void*>(k5_linear_persistent_kernel<256>)
                          : reinterpret_cast<const void*>(k5_linear_persistent_kernel<128>);
    const uint32_t smem = wide ? PersCfg<256>::kSmem : PersCfg<128>::kSmem;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int mt = m_tiles, nt = N / pbn;
    const int total = mt * nt;
    void* args[] = {&tm_a, &tm_w, &g, &mt, &nt};
    e = launch_pdl(fn, dim3(unsigned(total < sms ? total : sms)), dim3(kPersThreads), smem, st, args);
  } else if (reduce) {  // bf16 output of a split-K GEMM: sum the splits inside a cluster (128-wide tiles, two CTAs per SM)
    if (splits > 8 || (splits & (splits - 1)) != 0) return cudaErrorInvalidValue;
    e = launch_k5_bn<128, 3, true>(tm_a, tm_w, g, splits, st);
  } else {
    // ring depth measured at b = 1 (whole forward): 3 stages 227 us, 4: 192 us, 6: 186 us, 8: 188 us
    e = bn == 64        ? launch_k5_bn<64, 6>(tm_a, tm_w, g, splits, st)
        : ctas128 > 148 ? launch_k5_bn<128, 3>(tm_a, tm_w, g, splits, st)
                        : launch_k5_bn<128, 6>(tm_a, tm_w, g, splits, st);
  }
  note_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_k6_attention(const void* qkv, void* out, int b, int T, int d_model, int heads,
                                int groups, int group_tokens, cudaStream_t st, int only_group) {
  if (only_group >= groups) return cudaErrorInvalidValue;
  if (d_model != heads * kHeadDim || groups * group_tokens != T || T > kAttnMaxT) return cudaErrorInvalidValue;
  const int rows_pad = (T + kKeyChunk - 1) / kKeyChunk * kKeyChunk;
  const size_t smem = size_t(rows_pad) * kKVStride * 2 * 2;
  const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  float scale_log2e = 1.4426950408889634f / sqrtf(float(kHeadDim));
  // one CTA per (sample, head) once those alone load every SM; the (sample, head, group) grid below that. Measured
  // per forward with the threshold moved: 128 pairs 422 vs 415 us, 160 pairs 520 vs 524 us, 256 pairs 590 vs 619 us.
  static const int wide_min = [] {  // MRAG_K6_WIDE_MIN: (sample, head) count from which K6w runs; 0 = never (A/B runs)
    const char* v = getenv("MRAG_K6_WIDE_MIN");
    return v ? atoi(v) : 176;
  }();
  cudaError_t e;
  if (only_group < 0 && wide_min > 0 && b * heads >= wide_min) {
    e = cudaFuncSetAttribute(k6_attention_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    void* args[] = {&q, &o, &T, &d_model, &heads, &group_tokens, &scale_log2e};
    e = launch_pdl(reinterpret_cast<const void*>(k6_attention_wide_kernel), dim3(unsigned(b * heads)),
                   dim3(kAttnWideThreads), smem, st, args);
  } else {
    e = cudaFuncSetAttribute(k6_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    void* args[] = {&q, &o, &T, &d_model, &heads, &group_tokens, &scale_log2e, &only_group};
    e = launch_pdl(reinterpret_cast<const void*>(k6_attention_kernel),
                   dim3(unsigned(b * heads), unsigned(only_group >= 0 ? 1 : groups), unsigned((group_tokens + 31) / 32)),
                   dim3(kAttnThreads), smem, st, args);
  }
  note_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_k7_add_layernorm(const void* resid, const float* partial, int splits, const void* bias,
                                    const void* gamma, const void* beta, void* out, int M, int d, float eps,
                                    cudaStream_t st, int resid_T, int resid_L) {
  if (d % 256 != 0 || d > 1024 || splits > kMaxSplits) return cudaErrorInvalidValue;
  const __nv_bfloat16* r = static_cast<const __nv_bfloat16*>(resid);
  const __nv_bfloat16* bi = static_cast<const __nv_bfloat16*>(bias);
  const __nv_bfloat16* ga = static_cast<const __nv_bfloat16*>(gamma);
  const __nv_bfloat16* be = static_cast<const __nv_bfloat16*>(beta);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  void* args[] = {&r, &partial, &splits, &bi, &ga, &be, &o, &M, &d, &eps, &resid_T, &resid_L};
  cudaError_t e = M >= kLnWarpRowsMin
                      ? launch_pdl(reinterpret_cast<const void*>(k7_add_layernorm_kernel),
                                   dim3(unsigned((M + kLnRowsPerCta - 1) / kLnRowsPerCta)), dim3(32 * kLnRowsPerCta), 0, st, args)
                      : launch_pdl(reinterpret_cast<const void*>(k7_add_layernorm_cta_kernel), dim3(unsigned(M)),
                                   dim3(unsigned(d / 8)), 0, st, args);
  note_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mrag
