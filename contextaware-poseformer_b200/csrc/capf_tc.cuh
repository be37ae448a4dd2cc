// Device-side building blocks shared by the tcgen05 kernels of libcapf_b200 (capf_tc.cu, capf_tc_halo.cu):
// PTX wrappers (mbarrier, TMA, TMEM, tcgen05.mma/ld/commit), smem matrix descriptors, and the fused epilogue.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

#include <type_traits>

#include "capf_common.cuh"
#include "capf_internal.h"

namespace capf {

// =======================================================================================================
// device-side PTX wrappers
// =======================================================================================================
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of the (converged) warp: lets the compiler treat the guarded region as single-threaded, which is what
// keeps descriptor / coordinate arithmetic in uniform registers without per-value "waterfall" loops.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0x3ff) == 0 && clock64() - t0 > 4000000000ll) __trap();
  }
}

// L2 eviction-priority policies: activations are produced by one kernel and consumed by the next, so outputs are
// written evict_last (stay in the 126 MB L2 for the consumer) and inputs are read evict_first (dead after this read).
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
__device__ __forceinline__ void tma_load_4d_hint(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(dst),
      "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void cp_async16_hint(uint32_t dst_smem, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_global_v4_hint(void* p, uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// shared memory (a dense box in the map's swizzled layout) -> global; out-of-bounds elements of the box are not written
__device__ __forceinline__ void tma_store_4d_hint(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;" ::"l"((uint64_t)map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, uint32_t src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"((uint64_t)map), "r"(src), "r"(c0),
               "r"(c1), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the two 64-bit shared-memory descriptors given as (lo, hi) words: the hi words are kernel constants and
// the lo words advance by plain 32-bit adds, which keeps the single issuing thread's instruction count down.
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every tcgen05 op previously issued by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp receives row (lane base + t).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace ptx

// cta_group::2 (CTA pair) variants, shared by capf_tc2.cu and capf_tc_block64.cu
namespace ptx2 {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same variable in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Same without the release: `mbarrier.arrive.release.cluster` is a cluster-scope fence in front of the arrive and was measured
// at ~800 clk per call (tools/block_trace.py, 64-channel block).  Enough where what the arrive publishes is already complete
// when it is issued: TMEM reads after tcgen05.wait::ld, shared-memory writes after fence.proxy.async.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t bar_cluster, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the previously issued MMAs of this thread have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm_hint(const CUtensorMap* map, uint32_t bar_cluster, uint32_t dst, int c0, int c1, int c2, int c3,
                                                     uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(dst),
      "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}
// (lo, hi) descriptor words as in ptx::umma_f16_lohi
__device__ __forceinline__ void umma2_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
}  // namespace ptx2

constexpr int TC_SMEM_LIMIT = 232448;       // 227 KB opt-in maximum per CTA

// High word of a K-major shared-memory matrix descriptor: stride byte offset (distance between 8-row groups) >> 4 in
// bits [32,46), descriptor version 1 in [46,48), swizzle mode in [61,64) (0 none, 2 = 128 B, 4 = 64 B, 6 = 32 B).
__host__ __device__ inline uint32_t tc_desc_hi(int swizzle_bytes, int sbo_bytes) {
  const uint32_t layout = swizzle_bytes == 128 ? 2u : swizzle_bytes == 64 ? 4u : swizzle_bytes == 32 ? 6u : 0u;
  return ((uint32_t)sbo_bytes >> 4) | (1u << 14) | (layout << 29);
}
// Low word: start address >> 4 in [0,14), leading byte offset >> 4 in [16,30) (distance between the two 16-byte
// K slices of one MMA in the un-swizzled layout; ignored, canonical value 1, in the swizzled K-major layouts).
__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t smem_addr, uint32_t lbo_field) {
  return ((smem_addr >> 4) & 0x3fffu) | (lbo_field << 16);
}
__device__ __forceinline__ uint64_t tc_make_desc(uint32_t smem_addr, uint32_t lbo_field, uint32_t hi) {
  return ((uint64_t)hi << 32) | (uint64_t)(((smem_addr >> 4) & 0x3fffu) | (lbo_field << 16));
}
// tcgen05 instruction descriptor: kind::f16, fp32 accumulate, A and B K-major, M = 128, N = n.
__host__ __device__ inline uint32_t tc_idesc(bool bf16, int n) {
  const uint32_t fmt = bf16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// 16 consecutive output elements as raw 16-byte vectors: loads are issued early and converted late so that the
// global-memory latency of the residual overlaps the TMEM read of the accumulator.
template <typename TO> struct Vec16;
template <> struct Vec16<float> {
  float4 v[4];
  __device__ __forceinline__ void load(const float* p) {
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = *reinterpret_cast<const float4*>(p + 4 * q);
  }
  __device__ __forceinline__ void add_to(float (&r)[16]) const {
#pragma unroll
    for (int q = 0; q < 4; ++q) { r[4 * q] += v[q].x; r[4 * q + 1] += v[q].y; r[4 * q + 2] += v[q].z; r[4 * q + 3] += v[q].w; }
  }
  static __device__ __forceinline__ void store(float* p, const float (&r)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
  }
};
template <> struct Vec16<__half> {
  uint4 v[2];
  __device__ __forceinline__ void load(const __half* p) {
#pragma unroll
    for (int q = 0; q < 2; ++q) v[q] = *reinterpret_cast<const uint4*>(p + 8 * q);
  }
  __device__ __forceinline__ void add_to(float (&r)[16]) const {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __half22float2(h[e]);
        r[8 * q + 2 * e] += f.x; r[8 * q + 2 * e + 1] += f.y;
      }
    }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&r)[16]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint4 o;
      __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(r[8 * q + 2 * e], r[8 * q + 2 * e + 1]);
      *reinterpret_cast<uint4*>(p + 8 * q) = o;
    }
  }
};
template <> struct Vec16<__nv_bfloat16> {
  uint4 v[2];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
#pragma unroll
    for (int q = 0; q < 2; ++q) v[q] = *reinterpret_cast<const uint4*>(p + 8 * q);
  }
  __device__ __forceinline__ void add_to(float (&r)[16]) const {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __bfloat1622float2(h[e]);
        r[8 * q + 2 * e] += f.x; r[8 * q + 2 * e + 1] += f.y;
      }
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&r)[16]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint4 o;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(r[8 * q + 2 * e], r[8 * q + 2 * e + 1]);
      *reinterpret_cast<uint4*>(p + 8 * q) = o;
    }
  }
};

// 16-byte global store with an L2 eviction-priority policy
__device__ __forceinline__ void st16_hint(void* p, uint4 v, uint64_t pol) { ptx::st_global_v4_hint(p, v, pol); }

// 16 bias values (broadcast loads through the read-only path); zeros when the op has no bias.  Issued BEFORE the
// tcgen05.wait::ld so the L1 latency overlaps the TMEM read.
struct Bias16 {
  float4 b[4];
  __device__ __forceinline__ void load(const float* __restrict__ bias, int n) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4)
      b[q4] = bias ? __ldg(reinterpret_cast<const float4*>(bias + n) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
};

// halo-band 3x3 convolution (capf_tc_halo.cu)
struct TcHaloState;
int tc_halo_supported(const capf_op& op);
int tc_halo_prepare(const capf_op& op, TcHaloState** out);
int tc_halo_launch(const TcHaloState* s, cudaStream_t st);
void tc_halo_release(TcHaloState* s);
void tc_halo_describe(const TcHaloState* s, char* buf, int cap);

// halo band + streamed weights for C = Cout = 128 (capf_tc_halo128.cu)
struct TcHalo128State;
int tc_halo128_supported(const capf_op& op);
int tc_halo128_prepare(const capf_op& op, TcHalo128State** out);
int tc_halo128_launch(const TcHalo128State* s, cudaStream_t st);
void tc_halo128_release(TcHalo128State* s);
void tc_halo128_describe(const TcHalo128State* s, char* buf, int cap);

// halo tile in four 64-channel planes + streamed weights: 3x3 / stride 1 with C = 256 -> Cout = 32 (capf_tc_halo256.cu)
struct TcHalo256State;
int tc_halo256_supported(const capf_op& op);
int tc_halo256_prepare(const capf_op& op, TcHalo256State** out);
int tc_halo256_launch(const TcHalo256State* s, cudaStream_t st);
void tc_halo256_release(TcHalo256State* s);
void tc_halo256_describe(const TcHalo256State* s, char* buf, int cap);

// bias + GELU + residual + ReLU on 16 accumulator columns of a 16-bit output row staged in shared memory: the residual (if
// any) is read from, and the result written back to, the two 16-byte chunks at smem addresses s0 / s1 (runtime-flag
// variant used by the stem kernel; the GEMM / halo kernels use the compile-time epi16 below).
template <typename TO>
__device__ __forceinline__ void finish16_smem(const Bias16& bs, int act, const uint32_t (&raw)[16], bool has_res, uint32_t s0, uint32_t s1) {
  static_assert(sizeof(TO) == 2, "staged epilogue is for 16-bit outputs");
  float v[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(raw[e]);
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    v[4 * q4] += bs.b[q4].x; v[4 * q4 + 1] += bs.b[q4].y; v[4 * q4 + 2] += bs.b[q4].z; v[4 * q4 + 3] += bs.b[q4].w;
  }
  if (act == CAPF_ACT_GELU) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = gelu_erf(v[e]);
  }
  if (has_res) {
    Vec16<TO> rv;
    rv.v[0] = ptx::ld_shared_v4(s0);
    rv.v[1] = ptx::ld_shared_v4(s1);
    rv.add_to(v);
  }
  if (act == CAPF_ACT_RELU) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
  }
  uint4 o[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t* po = reinterpret_cast<uint32_t*>(&o[h]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if constexpr (std::is_same<TO, __half>::value) {
        __half2 t = __floats2half2_rn(v[8 * h + 2 * e], v[8 * h + 2 * e + 1]);
        po[e] = *reinterpret_cast<uint32_t*>(&t);
      } else {
        __nv_bfloat162 t = __floats2bfloat162_rn(v[8 * h + 2 * e], v[8 * h + 2 * e + 1]);
        po[e] = *reinterpret_cast<uint32_t*>(&t);
      }
    }
  }
  ptx::st_shared_v4(s0, o[0]);
  ptx::st_shared_v4(s1, o[1]);
}

// Epilogue of 16 accumulator columns of one output row, staged in shared memory: y = max(floor, gelu?(acc + bias) +
// residual?) -> TO, written to (and, with a residual, first read from) the CPG consecutive 16-byte chunks of the row's
// slot in the XOR-swizzled staging tile (chunk k of row r lives at chunk position (c0 + k) ^ (r & 7)).  Compile-time
// MODE (bit 0 residual, bit 1 GELU) and plain shared-memory pointers keep the instruction stream short and let the
// compiler schedule it; `floor` is 0 for ReLU and -inf otherwise.
template <typename TO, int MODE>
__device__ __forceinline__ void epi16(const uint32_t (&raw)[16], const float* __restrict__ sbias, float floor, uint8_t* row_ptr, uint32_t chunk0,
                                      uint32_t rx) {
  constexpr int CPG = (int)sizeof(TO);           // 16-byte chunks per 16 columns
  float v[16];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const float4 b = *reinterpret_cast<const float4*>(sbias + 4 * q4);
    v[4 * q4] = __uint_as_float(raw[4 * q4]) + b.x;
    v[4 * q4 + 1] = __uint_as_float(raw[4 * q4 + 1]) + b.y;
    v[4 * q4 + 2] = __uint_as_float(raw[4 * q4 + 2]) + b.z;
    v[4 * q4 + 3] = __uint_as_float(raw[4 * q4 + 3]) + b.w;
  }
  if constexpr ((MODE & 2) != 0) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = gelu_erf(v[e]);
  }
  // row_ptr = this thread's 128-byte row of the staging tile, rx = row & 7 (the swizzle key), chunk0 = first chunk
  uint4* slot[CPG];
#pragma unroll
  for (int k = 0; k < CPG; ++k) slot[k] = reinterpret_cast<uint4*>(row_ptr + (((chunk0 + (uint32_t)k) ^ rx) << 4));
  if constexpr ((MODE & 1) != 0) {
    if constexpr (sizeof(TO) == 2) {
      Vec16<TO> rv;
      rv.v[0] = *slot[0];
      rv.v[1] = *slot[1];
      rv.add_to(v);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 t = *slot[k];
        v[4 * k] += __uint_as_float(t.x); v[4 * k + 1] += __uint_as_float(t.y);
        v[4 * k + 2] += __uint_as_float(t.z); v[4 * k + 3] += __uint_as_float(t.w);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], floor);
  if constexpr (sizeof(TO) == 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 o;
      uint32_t* po = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if constexpr (std::is_same<TO, __half>::value) {
          __half2 t = __floats2half2_rn(v[8 * h + 2 * e], v[8 * h + 2 * e + 1]);
          po[e] = *reinterpret_cast<uint32_t*>(&t);
        } else {
          __nv_bfloat162 t = __floats2bfloat162_rn(v[8 * h + 2 * e], v[8 * h + 2 * e + 1]);
          po[e] = *reinterpret_cast<uint32_t*>(&t);
        }
      }
      *slot[h] = o;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *slot[k] = make_uint4(__float_as_uint(v[4 * k]), __float_as_uint(v[4 * k + 1]), __float_as_uint(v[4 * k + 2]), __float_as_uint(v[4 * k + 3]));
  }
}

// fused 32-channel BasicBlock (capf_tc_block.cu)
struct TcBlockState;
int tc_block_supported(const capf_op& op);
int tc_block_prepare(const capf_op& op, TcBlockState** out);
int tc_block_launch(const TcBlockState* s, cudaStream_t st);
void tc_block_release(TcBlockState* s);
void tc_block_describe(const TcBlockState* s, char* buf, int cap);

// fused 64-channel BasicBlock on CTA pairs (capf_tc_block64.cu); reached through tc_block_* by channel count
struct TcBlock64State;
int tc_block64_supported(const capf_op& op);
int tc_block64_prepare(const capf_op& op, TcBlock64State** out);
int tc_block64_launch(const TcBlock64State* s, cudaStream_t st);
void tc_block64_release(TcBlock64State* s);
void tc_block64_describe(const TcBlock64State* s, char* buf, int cap);

// Bottleneck conv3 + residual -> conv1 of the next block (capf_tc_chain.cu, CAPF_OP_EXPAND_REDUCE)
struct TcChainState;
int tc_chain_supported(const capf_op& op);
int tc_chain_prepare(const capf_op& op, TcChainState** out);
int tc_chain_launch(const TcChainState* s, cudaStream_t st);
void tc_chain_release(TcChainState* s);
void tc_chain_describe(const TcChainState* s, char* buf, int cap);

// 2-CTA (cta_group::2) GEMM for the wide lifter Linears (capf_tc2.cu)
// fused Mlp of a 128-wide transformer block: fc1 + GELU + fc2 + residual (capf_tc_mlp.cu)
struct TcMlpState;
int tc_mlp_supported(const capf_op& op);
int tc_mlp_prepare(const capf_op& op, TcMlpState** out);
int tc_mlp_launch(const TcMlpState* s, cudaStream_t st);
void tc_mlp_release(TcMlpState* s);
void tc_mlp_describe(const TcMlpState* s, char* buf, int cap);

struct Tc2State;
int tc2_supported(const capf_op& op);
int tc2_prepare(const capf_op& op, Tc2State** out);
int tc2_launch(const Tc2State* s, cudaStream_t st);
void tc2_release(Tc2State* s);
void tc2_describe(const Tc2State* s, char* buf, int cap);

// host helpers (capf_tc.cu)
int tc_get_encoder();
int tc_encode_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* ptr, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, const cuuint32_t* estr, int swz_bytes, const char* what);

}  // namespace capf
