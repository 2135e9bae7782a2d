// Shared device helpers for libmrag (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ != 1000)
#error "libmrag kernels are written for sm_100a (B200) only"
#endif
#endif

namespace mrag {

constexpr int kInvalidIdx = 0x7fffffff;  // empty candidate slot

// ---- order-preserving float <-> uint ---------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
// candidate key: ascending u64 order == (score descending, row index ascending)
__host__ __device__ __forceinline__ uint64_t make_sim_key(float score, int idx) {
  return (uint64_t(~f32_to_ordered(score)) << 32) | uint32_t(idx);
}
__host__ __device__ __forceinline__ float sim_key_score(uint64_t key) {
  return ordered_to_f32(~uint32_t(key >> 32));
}
__host__ __device__ __forceinline__ int sim_key_idx(uint64_t key) { return int(uint32_t(key)); }
constexpr uint64_t kEmptyKey = ~0ull;

// ---- streaming loads -------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ float bf16lo_to_f32(uint32_t packed) {
  return __uint_as_float(packed << 16);
}
__device__ __forceinline__ float bf16hi_to_f32(uint32_t packed) {
  return __uint_as_float(packed & 0xffff0000u);
}

// ---- block-wide bitonic sort of u64 keys in shared memory (n power of two) ---------------
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* keys, int n, int tid, int nthreads) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = tid; t < (n >> 1); t += nthreads) {
        int lo = 2 * t - (t & (stride - 1));  // index of the lower element of the pair
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        uint64_t a = keys[lo], b = keys[hi];
        if ((a > b) == up) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

// ---- barrier-light sort of a small key array: rank by counting -----------------------------------
// Thread i (i < n) counts the keys that precede its own (ties by position) and writes its key
// to that rank in `out`. n iterations of a broadcast smem read, two block barriers in total —
// for n <= a few hundred this beats the ~log^2(n) barriers of the bitonic network. `out` must
// not alias `keys`; entries >= n of `out` are left untouched.
__device__ __forceinline__ void rank_sort_smem(const uint64_t* keys, uint64_t* out, int n, int tid,
                                               int nthreads) {
  __syncthreads();
  for (int i = tid; i < n; i += nthreads) {
    const uint64_t mine = keys[i];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const uint64_t o = keys[j];
      r += (o < mine) || (o == mine && j < i);
    }
    out[r] = mine;
  }
  __syncthreads();
}

// ---- warp-level sorted lists of u64 keys (one key per lane, registers + shuffles only) ----------
// ascending bitonic sort across the warp
__device__ __forceinline__ uint64_t warp_sort_u64(uint64_t key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, key, j);
      const bool up = (lane & k) == 0;     // this k-block sorts ascending
      const bool lower = (lane & j) == 0;  // this lane keeps the smaller of the pair when ascending
      const uint64_t lo = key < other ? key : other;
      const uint64_t hi = key < other ? other : key;
      key = (lower == up) ? lo : hi;
    }
  }
  return key;
}
// a: sorted ascending across lanes; b_rev: lane l holds element 31-l of another ascending list.
// min(a[l], b[31-l]) is a bitonic sequence holding the 32 smallest keys of the union; five
// compare-exchange steps sort it: the result is the sorted list of the 32 best keys of both lists.
__device__ __forceinline__ uint64_t warp_merge_keep32(uint64_t a, uint64_t b_rev, int lane) {
  uint64_t c = a < b_rev ? a : b_rev;
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint64_t other = __shfl_xor_sync(0xffffffffu, c, j);
    const uint64_t lo = c < other ? c : other;
    const uint64_t hi = c < other ? other : c;
    c = ((lane & j) == 0) ? lo : hi;
  }
  return c;
}
// The 32 smallest of n <= 1024 keys, sorted, for a block of NT threads: every warp sorts the groups
// of 32 keys it owns and folds them into one list, warp 0 folds the per-warp lists. `lists` is
// shared scratch of NT keys; on return lists[0..32) holds the result (kEmptyKey-padded). Two block
// barriers; key_at(i) is called once per key (it may load from global memory: a warp's loads are
// all issued before its first sort).
template <int NT, typename KeyFn>
__device__ __forceinline__ void block_top32(KeyFn key_at, int n, uint64_t* lists) {
  constexpr int NW = NT / 32;
  constexpr int MAXG = (1024 / 32 + NW - 1) / NW;  // groups per warp at n = 1024
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t kk[MAXG];
#pragma unroll
  for (int u = 0; u < MAXG; ++u) {
    const int i = (warp + u * NW) * 32 + lane;
    kk[u] = (i < n) ? key_at(i) : kEmptyKey;
  }
  uint64_t acc = kEmptyKey;
#pragma unroll
  for (int u = 0; u < MAXG; ++u) {
    if ((warp + u * NW) * 32 < n) {  // warp-uniform
      const uint64_t sorted = warp_sort_u64(kk[u], lane);
      acc = (u == 0) ? sorted : warp_merge_keep32(acc, __shfl_sync(0xffffffffu, sorted, 31 - lane), lane);
    }
  }
  lists[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) {
    const int used = min(NW, (n + 31) / 32);
    for (int w = 1; w < used; ++w) acc = warp_merge_keep32(acc, lists[w * 32 + 31 - lane], lane);
    lists[lane] = acc;
  }
  __syncthreads();
}

// ---- mbarrier / TMA / tcgen05 wrappers ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 2-D tiled TMA load, completion signalled on an mbarrier of this CTA
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1, uint64_t cache_policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(cache_policy)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- CTA-pair (cta_group::2) variants ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `target_rank` of this cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t target_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(target_rank));
  // relaxed: the only data handed over is TMEM contents, ordered by tcgen05.fence::before_thread_sync;
  // a release at cluster scope would drain this thread's outstanding memory traffic (~1300 cycles)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are credited to the mbarrier of the
// even (leader) CTA: clearing bit 24 of the shared::cluster address selects the pair's CTA 0
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint64_t* bar,
                                                 int c0, int c1, uint64_t cache_policy) {
  const uint32_t leader_bar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1), "l"(cache_policy)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
// M=256 MMA across the CTA pair: each CTA supplies its 128 rows of A and its N/2 rows of B;
// issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive (once the MMAs retire) on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(uint16_t(3))
      : "memory");
}

// shared-memory matrix descriptor: K-major tile, 128-byte swizzle, rows of 64 bf16 (128 B),
// 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor: start[0,14) LBO[16,30) SBO[32,46)
// version[46,48)=1 layout_type[61,64)=2 (SWIZZLE_128B)).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3ffff) >> 4);
  d |= uint64_t(1) << 16;            // leading byte offset (ignored for swizzled K-major)
  d |= uint64_t(1024 >> 4) << 32;    // stride byte offset: next 8-row group
  d |= uint64_t(1) << 46;            // descriptor version (Blackwell)
  d |= uint64_t(2) << 61;            // SWIZZLE_128B
  return d;
}
// instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

}  // namespace mrag
