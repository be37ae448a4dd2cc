// Micro-benchmark: how fast can ONE CTA per SM pull boxes through TMA, by box shape / OOB / boxes in flight?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../contextaware-poseformer_b200/csrc tma_bench.cu -o tma_bench -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "capf_tc.cuh"
using namespace capf;
namespace capf { int g_use_pdl = 0; int g_num_sms = 148; int set_error(int c, const char*) { return c; } int set_errorf(int c, const char*, ...) { return c; }
int check_launch(const char*) { return 0; } }

struct P { int box_bytes, n_slots, iters, c1, c2, c3, dy, img_stride_rows, H; };

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, P p, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bars = base, data = base + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_slots; ++i) ptx::mbar_init(bars + 8 * i, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    int y = 0, img = blockIdx.x;
    // keep n_slots boxes in flight: slot s is re-armed as soon as its previous box has landed
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.n_slots;
      if (it >= p.n_slots) ptx::mbar_wait(bars + 8 * s, ((it / p.n_slots) - 1) & 1);
      ptx::mbar_arrive_expect_tx(bars + 8 * s, p.box_bytes);
      ptx::tma_load_4d(&map, bars + 8 * s, data + s * ((p.box_bytes + 1023) & ~1023), 0, p.c1, y, img);
      y += p.dy;
      if (y + p.dy > p.H) { y = 0; img += gridDim.x; }
    }
    for (int it = p.iters; it < p.iters + p.n_slots; ++it) {
      const int s = it % p.n_slots;
      ptx::mbar_wait(bars + 8 * s, ((it / p.n_slots) - 1) & 1);
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main(int argc, char** argv) {
  // args: C W H N  box_w box_h  x0  n_slots iters swizzle
  int C = atoi(argv[1]), W = atoi(argv[2]), H = atoi(argv[3]), N = atoi(argv[4]);
  int bw = atoi(argv[5]), bh = atoi(argv[6]), x0 = atoi(argv[7]), slots = atoi(argv[8]), iters = atoi(argv[9]), swz = atoi(argv[10]);
  void* d; size_t bytes = (size_t)N * H * W * C * 2;
  cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
  tc_get_encoder();
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh, 1}, es[4] = {1, 1, 1, 1};
  if (tc_encode_map(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, str, box, es, swz, "bench")) { printf("encode failed\n"); return 1; }
  P p; p.box_bytes = C * 2 * bw * bh; p.n_slots = slots; p.iters = iters; p.c1 = x0; p.dy = bh; p.H = H;
  long long* out; cudaMalloc(&out, 148 * 8);
  int smem = 2048 + slots * ((p.box_bytes + 1023) & ~1023);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 3; ++rep) {
    k<<<148, 128, smem>>>(m, p, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
  }
  std::vector<long long> h(148); cudaMemcpy(h.data(), out, 148 * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (auto v : h) avg += v; avg /= 148;
  printf("C=%d box %dx%d (%d B, %d-B rows) x0=%d slots=%d swz=%d: %.1f cycles/box, %.2f B/clk/SM, ~%.2f TB/s chip @1.9GHz\n", C, bw, bh, p.box_bytes, C * 2, x0,
         slots, swz, avg / iters, (double)p.box_bytes * iters / avg, (double)p.box_bytes * iters / avg * 148 * 1.9e9 / 1e12);
  return 0;
}
