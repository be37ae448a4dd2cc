// Micro-benchmark: sustained rate of tcgen05.mma (kind::f16, SS operands, 128-byte-swizzled K-major, K = 16 per instruction)
// from resident shared-memory operands, by CTA group (1: M = 128, 2: CTA pair, M = 256), N, and number of issuing warps --
// the structural bound of the halo / fused-block kernels, whose MMAs have N = 32..128 and read A nine times from one band.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../contextaware-poseformer_b200/csrc mma_rate.cu -o mma_rate
// ./mma_rate            (prints clk per MMA for a table of configurations)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "capf_tc.cuh"
using namespace capf;
namespace capf { int g_use_pdl = 0; }

struct P { int cg, N, sets, issuers, shift_px, same_a; };

// every issuer runs `sets` times the 36-MMA pattern of one 3x3 / 64-channel sub-tile (9 taps x 4 K steps) into its own accumulator
template <int CG>
__global__ void __launch_bounds__(256, 1) k(P p, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base, slot = base + 64, a0 = base + 1024, b0 = a0 + 96 * 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (CG == 2) rank = ptx2::cluster_ctarank();
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, p.issuers);
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < (96 + 64) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (a0 - ptx::smem_u32(raw)))[i] = 0x3c003c00u;
  ptx::fence_proxy_async();
  if (warp == 7) {
    if constexpr (CG == 2) { ptx2::tmem_alloc2(slot, 512); ptx2::tmem_relinquish2(); }
    else { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) ptx2::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(raw + (slot - ptx::smem_u32(raw)));
  const uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)((p.cg == 2 ? 256 : 128) >> 4) << 24);
  const uint32_t hi = tc_desc_hi(128, 1024);
  long long t0 = 0, t1 = 0;
  if (warp < p.issuers && rank == 0) {
    const int Wp = 33;
    const uint32_t a_lo = tc_desc_lo(a0 + (uint32_t)p.shift_px * 128u + (p.same_a ? 0u : (uint32_t)warp * 128u * 128u), 1u), b_lo = tc_desc_lo(b0, 1u);
    const uint32_t d = tmem + (uint32_t)(warp * p.N);
    const uint32_t brows16 = (uint32_t)((p.cg == 2 ? p.N / 2 : p.N) * 128) >> 4;      // one tap of B
    t0 = clock64();
    for (int s = 0; s < p.sets; ++s) {
      if (ptx::elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t ao = (uint32_t)((tap / 3) * Wp + tap % 3) * 8u + 2u * kk, bo = (uint32_t)(tap % 2) * brows16 + 2u * kk;
            if constexpr (CG == 2) ptx2::umma2_f16_lohi(d, a_lo + ao, hi, b_lo + bo, hi, idesc, (tap | kk | s) ? 1u : 0u);
            else ptx::umma_f16_lohi(d, a_lo + ao, hi, b_lo + bo, hi, idesc, (tap | kk | s) ? 1u : 0u);
          }
        }
      }
      __syncwarp();
    }
    if (ptx::elect_one()) {
      if constexpr (CG == 2) ptx2::umma2_commit_mc(bar); else ptx::umma_commit(bar);
    }
    __syncwarp();
    t1 = clock64();
  }
  ptx::mbar_wait(bar, 0);
  const long long t2 = clock64();
  if (warp < p.issuers && rank == 0 && lane == 0 && blockIdx.x == 0) { out[2 * warp] = t1 - t0; out[2 * warp + 1] = t2 - t0; }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) ptx2::cluster_sync();
  if (warp == 7) {
    if constexpr (CG == 2) ptx2::tmem_dealloc2(tmem, 512); else ptx::tmem_dealloc(tmem, 512);
  }
}

static void run(int cg, int N, int issuers, int shift, int same_a, long long* dout) {
  P p{cg, N, 40, issuers, shift, same_a};
  const int smem = 1024 + 1024 + (96 + 64) * 1024;
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = cg == 2 ? 1 : 0;
  long long h[8];
  double best = 1e30, best_issue = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(dout, 0, 64);
    cudaError_t le = cg == 2 ? cudaLaunchKernelEx(&cfg, k<2>, p, dout) : cudaLaunchKernelEx(&cfg, k<1>, p, dout);
    if (le != cudaSuccess) { printf("launch cg %d N %d: %s\n", cg, N, cudaGetErrorString(le)); exit(1); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cg %d N %d: %s\n", cg, N, cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, dout, 64, cudaMemcpyDeviceToHost);
    double done = 0, iss = 0;
    for (int w = 0; w < issuers; ++w) { if (h[2 * w + 1] > done) done = (double)h[2 * w + 1]; if (h[2 * w] > iss) iss = (double)h[2 * w]; }
    if (done < best) { best = done; best_issue = iss; }
  }
  const double n = 36.0 * p.sets * issuers;
  const double flop = 2.0 * (cg == 2 ? 256 : 128) * N * 16;
  printf("cta_group %d  M %3d  N %3d  issuers %d  shift %2d px  %s: %6.1f clk/MMA (issue loop %6.1f), %5.0f FLOP/clk/SM\n", cg, cg == 2 ? 256 : 128, N, issuers, shift,
         same_a ? "same A " : "own A  ", best / n, best_issue / n, flop / cg / (best / n));
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  for (int cg = 1; cg <= 2; ++cg)
    for (int N : {16, 32, 48, 64, 96, 128, 192, 256})
      for (int issuers : {1, 2}) {
        if (issuers * N > 512) continue;
        run(cg, N, issuers, 0, 0, dout);
      }
  run(1, 64, 2, 5, 0, dout);
  run(1, 64, 2, 0, 1, dout);
  return 0;
}
