// Fused HRNet BasicBlock for the 32-channel branch (pose_hrnet.py:66-95):  y = relu(bn2(conv2(relu(bn1(conv1(x))))) + x),
// both convolutions 3x3 / stride 1 / pad 1, 32 -> 32 channels, 16-bit NHWC in and out.
//
// As two halo-band launches the block is HBM-bound: conv1 reads x and writes u, conv2 reads u and x (residual) and writes
// y -- 5 passes of a 67 MB tensor at bs = 256 (64 such convolutions = 22 % of the step).  Here one CTA keeps, per band of
// bh output rows,
//     X   the input band (bh + 4 rows + 1, TMA, double buffered)             -> conv1 operand AND conv2's residual
//     MID relu(bn1(conv1(x))) on bh + 2 rows, written by the epilogue warps straight into shared memory in the swizzled
//         K-major layout the tensor pipe reads                           -> conv2 operand; never leaves the SM
// so HBM sees one read of x and one write of y.  Both convolutions use the shifted-window trick of capf_tc_halo.cu (the
// nine taps read the same buffer through descriptors moved by (r * Wp + s) pixels; one zero column shared by the left /
// right padding).  conv2's zero padding needs MID to be ZERO outside the image (not conv1 of the padding), which the
// first epilogue enforces; its one-pixel border of MID costs (bh + 2) / bh more conv1 work.
//
// Roles (640 threads): warp 0 TMA producer, warp 1 issues conv1 and warp 3 conv2 (each sub-tile by sub-tile, in order), warp 2
// TMEM allocator, warps 4..19 four 4-warp epilogue groups (two for epilogue 1, two for epilogue 2, each pair splitting a
// sub-tile's channels).  Every 128-pixel sub-tile of a band owns one TMEM accumulator (phase A: slots [0, n1max), phase B:
// [n1max, n1max + n2max)), so a band needs no accumulator recycling; phase B sub-tile j starts as soon as the MID rows it
// reads have been written (per-sub-tile "mid ready" barriers), which pipelines conv1 -> epilogue 1 -> conv2 inside a band,
// and conv1 of the next band is issued while conv2 of this one runs.  Barriers that are used once per band keep their phase
// parity in a per-role bit mask (the ragged last band of an image uses fewer sub-tiles).
//
// Bound (tools/microbench/mma_rate.cu): a 128 x 32 x 16 SS-MMA takes 40 clk -- the 4 KB of A it reads from shared memory at
// 128 B/clk (+ 1 KB of B), not its 16 clk of tensor work; a band is 234 such MMAs.  HBM traffic: 77 MB per block instead of 2 x 143 MB.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <new>

#include "capf_tc.cuh"

namespace capf {

constexpr int BLK_THREADS = 640;            // warps 0-3: TMA / MMA / TMEM / MMA; warps 4-19: four 4-warp epilogue groups
constexpr int BLK_C = 32;                   // channels in = out
constexpr int BLK_PIX = BLK_C * 2;          // bytes per pixel row (64-byte swizzle span)
constexpr int BLK_W_BYTES = 9 * BLK_C * BLK_C * 2;   // folded weights of one convolution: 18 KB
constexpr int BLK_GROUPS = 4;
constexpr int BLK_EPI_WARPS = 4 * BLK_GROUPS;
constexpr int BLK_MAX_ACC = 16;             // 16 x 32 TMEM columns
// header layout (bytes from the 1024-aligned base)
constexpr int BH_W = 0, BH_XFULL = 8, BH_XEMPTY = 24, BH_MIDFREE = 40, BH_C1DONE = 48, BH_TFULL = 64, BH_TEMPTY = 192, BH_MIDRDY = 320, BH_TMEM = 448;
constexpr int BLK_HEADER = 1024;

struct BlockP {
  int H, W, Nimg, Wp;
  uint32_t wp_magic;
  int bh, bands_per_img, num_bands;
  int n1max, n2max;
  int x_bytes, mid_bytes, x_tx_bytes, tmem_cols;
  uint32_t idesc, desc_hi;
  const float* bias1;
  const float* bias2;
  void* out;
  long long* trace;      // optional (debug, op.in[5]): per-band wait cycles of CTA 0's issuers, see tools/block_trace.py
};

__device__ __forceinline__ int blk_div_wp(int v, uint32_t magic) { return (int)__umulhi((uint32_t)v, magic); }

// 16-byte chunk c of pixel h in a 64-byte-swizzled pixel-row buffer whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t blk_chunk(uint32_t h, uint32_t c) { return h * (uint32_t)BLK_PIX + ((c ^ ((h >> 1) & 3u)) << 4); }

template <typename T>
__global__ void __launch_bounds__(BLK_THREADS, 1)
tc_block32_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                  const BlockP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_w = base + BH_W, bar_xfull = base + BH_XFULL, bar_xempty = base + BH_XEMPTY;
  const uint32_t bar_c1done = base + BH_C1DONE;
  const uint32_t bar_tfull = base + BH_TFULL, bar_tempty = base + BH_TEMPTY, bar_midrdy = base + BH_MIDRDY, tmem_slot = base + BH_TMEM;
  const uint32_t smem_w1 = base + BLK_HEADER, smem_w2 = smem_w1 + BLK_W_BYTES;
  const uint32_t smem_x = smem_w2 + BLK_W_BYTES;
  const uint32_t smem_mid = smem_x + 2u * (uint32_t)p.x_bytes;
  uint8_t* const gen = smem_raw + (base - raw);                 // generic pointer to `base`
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + BH_TMEM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapX);
    ptx::prefetch_tmap(&mapW1);
    ptx::prefetch_tmap(&mapW2);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(bar_w, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_xfull + 8 * b, 1);
      ptx::mbar_init(bar_xempty + 8 * b, 1 + 8);                // conv1's commit + the 8 epilogue-2 warps (residuals, in-place results, band store)
      ptx::mbar_init(bar_c1done + 8 * b, 1);                    // conv1's commit: epilogue 2 may overwrite X[b]
    }
    for (int a = 0; a < BLK_MAX_ACC; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);
      ptx::mbar_init(bar_tempty + 8 * a, 8);                    // one arrival per epilogue warp of the phase (two groups)
      ptx::mbar_init(bar_midrdy + 8 * a, 8);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp == 3) {                   // the zero pixel in front of MID (image column -1 of its first row)
    if (lane < 4) *reinterpret_cast<uint4*>(gen + (smem_mid - base) + 16 * lane) = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();
  if (warp != 0) pdl_wait();

  const int band0 = (int)(((long long)p.num_bands * blockIdx.x) / gridDim.x);
  const int band1 = (int)(((long long)p.num_bands * (blockIdx.x + 1)) / gridDim.x);
  const int img0 = band0 / p.bands_per_img;

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(bar_w, 2u * BLK_W_BYTES);
      for (int c = 0; c < 9; ++c) {
        ptx::tma_load_2d(&mapW1, bar_w, smem_w1 + c * 2048, c * BLK_C, 0);
        ptx::tma_load_2d(&mapW2, bar_w, smem_w2 + c * 2048, c * BLK_C, 0);
      }
      pdl_wait();
      const uint64_t pol_in = ptx::policy_evict_first();
      int img = img0, bin = band0 - img0 * p.bands_per_img;
      uint32_t k = 0;
      for (int band = band0; band < band1; ++band, ++k) {
        const uint32_t buf = k & 1u, ph = (k >> 1) & 1u;
        ptx::mbar_wait(bar_xempty + 8 * buf, ph ^ 1u);
        ptx::mbar_arrive_expect_tx(bar_xfull + 8 * buf, (uint32_t)p.x_tx_bytes);
        ptx::tma_load_4d_hint(&mapX, bar_xfull + 8 * buf, smem_x + buf * (uint32_t)p.x_bytes, 0, -1, bin * p.bh - 2, img, pol_in);
        if (++bin == p.bands_per_img) { bin = 0; ++img; }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer, phase A ===============================
    // One warp issues conv1 of every band, another (warp 3) conv2, each sub-tile by sub-tile IN ORDER: the tensor pipe
    // finishes sub-tile j while epilogue 1 of sub-tile j - 1 runs, and this warp already issues conv1 of the next band while
    // conv2 of the current one is in the pipe.  (Earlier: three issuers taking sub-tiles round-robin -- their MMAs interleave,
    // all sub-tiles of a phase complete together, and 10-35 % of a band was the wait for MID rows.)  One thread sustains the
    // pipe's 40 clk per 128 x 32 x 16 MMA (tools/microbench/mma_rate.cu: the shared-memory read of A, 4 KB per MMA, is the bound).
    ptx::mbar_wait(bar_w, 0);
    ptx::tc_fence_after();
    uint32_t tap_off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) tap_off[tap] = (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 4u;     // 4 x 16 B per pixel
    const uint32_t w1_lo = tc_desc_lo(smem_w1, 1u);
    int bin = band0 - img0 * p.bands_per_img;
    uint32_t k = 0, pm = 0;           // pm: phase parity of the per-accumulator barriers (bit a flips when slot a is used)
    for (int band = band0; band < band1; ++band, ++k) {
      const uint32_t buf = k & 1u, ph = (k >> 1) & 1u;
      const int bh_eff = min(p.bh, p.H - bin * p.bh);
      const int n1 = ((bh_eff + 2) * p.Wp + 127) >> 7;
      const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && k < 32;
      long long tw = 0, w_x = 0, w_ta = 0;
      if (tr) tw = clock64();
      ptx::mbar_wait(bar_xfull + 8 * buf, ph);
      if (tr) { w_x = clock64() - tw; p.trace[0 * 32 + k] = clock64(); }
      ptx::tc_fence_after();
      const uint32_t x_lo = tc_desc_lo(smem_x + buf * (uint32_t)p.x_bytes, 1u);
      for (int j = 0; j < n1; ++j) {
        if (tr) tw = clock64();
        ptx::mbar_wait(bar_tempty + 8 * j, ((pm >> j) & 1u) ^ 1u);
        if (tr) w_ta += clock64() - tw;
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t d = tmem_base + (uint32_t)(j * BLK_C), a_sub = x_lo + (uint32_t)(j * 128) * 4u;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              ptx::umma_f16_lohi(d, a_sub + tap_off[tap] + 2u * kk, p.desc_hi, w1_lo + (uint32_t)tap * 128u + 2u * kk, p.desc_hi, p.idesc, (tap | kk) ? 1u : 0u);
          }
          ptx::umma_commit(bar_tfull + 8 * j);
        }
        __syncwarp();
      }
      if (ptx::elect_one()) {                                            // conv1's reads of X[buf] complete
        ptx::umma_commit(bar_xempty + 8 * buf);
        ptx::umma_commit(bar_c1done + 8 * buf);
      }
      __syncwarp();
      if (tr) { p.trace[1 * 32 + k] = w_x; p.trace[2 * 32 + k] = w_ta; p.trace[5 * 32 + k] = clock64(); }
      pm ^= (1u << n1) - 1u;
      if (++bin == p.bands_per_img) bin = 0;
    }
  } else if (warp == 3) {
    // ===================================== MMA issuer, phase B ===============================
    ptx::mbar_wait(bar_w, 0);
    ptx::tc_fence_after();
    uint32_t tap_off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) tap_off[tap] = (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 4u;
    const uint32_t w2_lo = tc_desc_lo(smem_w2, 1u), mid_lo = tc_desc_lo(smem_mid, 1u);
    int bin = band0 - img0 * p.bands_per_img;
    uint32_t k = 0, pm = 0;
    for (int band = band0; band < band1; ++band, ++k) {
      const int bh_eff = min(p.bh, p.H - bin * p.bh);
      const int n1 = ((bh_eff + 2) * p.Wp + 127) >> 7, n2 = (bh_eff * p.Wp + 127) >> 7;
      const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && k < 32;
      long long tw = 0, w_mid = 0, w_tb = 0;
      if (tr) p.trace[8 * 32 + k] = clock64();
      // conv2 over bh_eff rows, sub-tile j as soon as the MID pixels it reads exist
      int ready_upto = -1;
      for (int j = 0; j < n2; ++j) {
        const int need = min(n1 - 1, (j * 128 + 128 + 2 * p.Wp) >> 7);
        if (tr) tw = clock64();
        while (ready_upto < need) {
          ++ready_upto;
          ptx::mbar_wait(bar_midrdy + 8 * ready_upto, (pm >> ready_upto) & 1u);
        }
        if (tr) { w_mid += clock64() - tw; tw = clock64(); }
        const int a = p.n1max + j;
        ptx::mbar_wait(bar_tempty + 8 * a, ((pm >> a) & 1u) ^ 1u);
        if (tr) w_tb += clock64() - tw;
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t d = tmem_base + (uint32_t)(a * BLK_C), a_sub = mid_lo + (uint32_t)(j * 128) * 4u;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              ptx::umma_f16_lohi(d, a_sub + tap_off[tap] + 2u * kk, p.desc_hi, w2_lo + (uint32_t)tap * 128u + 2u * kk, p.desc_hi, p.idesc, (tap | kk) ? 1u : 0u);
          }
          ptx::umma_commit(bar_tfull + 8 * a);
        }
        __syncwarp();
      }
      if (tr) { p.trace[(8 + 3) * 32 + k] = w_mid; p.trace[(8 + 4) * 32 + k] = w_tb; p.trace[(8 + 5) * 32 + k] = clock64(); }
      pm ^= ((1u << n1) - 1u) | (((1u << n2) - 1u) << p.n1max);
      if (++bin == p.bands_per_img) bin = 0;
    }
  } else if (warp >= 4) {
    // ===================================== epilogues =========================================
    // Groups 0 and 1 both work on every phase-A sub-tile, groups 2 and 3 on every phase-B sub-tile, 16 of the 32 channels
    // each (half the latency per sub-tile; a thread's 16 folded-BN shifts live in registers).  Epilogue 2 writes its result
    // over the pixel's residual IN the X band; at the end of the band its 8 warps copy the band out, 4 lanes per pixel and 8
    // pixels (four 128-byte lines) per instruction -- per-thread stores of a pixel's 64 bytes touch 16 lines per instruction
    // and the load/store unit, which also carries the MID writes, paced the epilogues (measured on the 64-channel kernel).
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const int ch0 = 16 * (grp & 1);                                     // this group's 16 channels = 2 chunks of the pixel row
    float breg[16];
    {
      const float* bsrc = grp < 2 ? p.bias1 : p.bias2;
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        const float4 v = bsrc ? __ldg(reinterpret_cast<const float4*>(bsrc + ch0 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
        breg[e] = v.x; breg[e + 1] = v.y; breg[e + 2] = v.z; breg[e + 3] = v.w;
      }
    }
    int img = img0, bin = band0 - img0 * p.bands_per_img, n2_prev = 0;
    uint32_t k = 0, pm = 0;
    for (int band = band0; band < band1; ++band, ++k) {
      const uint32_t buf = k & 1u;
      const int y0 = bin * p.bh;
      const int bh_eff = min(p.bh, p.H - y0);
      const int n1 = ((bh_eff + 2) * p.Wp + 127) >> 7, n2 = (bh_eff * p.Wp + 127) >> 7;
      if (grp < 2) {
        // ---- epilogue 1: relu(acc + b1), zero outside the image, 16-bit, into MID (shifted by one pixel) ----
        uint8_t* const mg = gen + (smem_mid - base);
        int freed_upto = -1;
        for (int j = 0; j < n1; ++j) {
          // MID pixels of sub-tile j were last read by conv2 sub-tiles <= j + 1 of the previous band (in-order issue: the
          // completion of sub-tile jj implies all earlier ones); their accumulator-full barriers double as "MID rows free"
          if (k > 0) {
            const int jj = min(j + 1, n2_prev - 1);
            if (jj > freed_upto) {
              const int a = p.n1max + jj;
              ptx::mbar_wait(bar_tfull + 8 * a, ((pm >> a) & 1u) ^ 1u);      // bit a of pm has flipped since the previous band
              freed_upto = jj;
            }
          }
          ptx::mbar_wait(bar_tfull + 8 * j, (pm >> j) & 1u);
          ptx::tc_fence_after();
          uint32_t a0[16];
          ptx::tmem_ld16(tmem_base + (uint32_t)(j * BLK_C + ch0) + ((uint32_t)(q * 32) << 16), a0);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * j);
          const int mp = j * 128 + q * 32 + lane;
          const int iy = blk_div_wp(mp, p.wp_magic), ix = mp - iy * p.Wp;
          const int yi = y0 - 1 + iy;
          const bool valid = ix < p.W && yi >= 0 && yi < p.H;
          if (mp < (bh_eff + 2) * p.Wp) {
            const uint32_t h = (uint32_t)mp + 1u;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = valid ? fmaxf(__uint_as_float(a0[8 * c + e]) + breg[8 * c + e], 0.f) : 0.f;
              *reinterpret_cast<uint4*>(mg + blk_chunk(h, (uint32_t)(2 * grp + c))) = pack8<T>(f);
            }
          }
          ptx::fence_proxy_async();                                     // generic-proxy writes of MID -> tensor-pipe reads
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_midrdy + 8 * j);
        }
      } else {
        // ---- epilogue 2: relu(acc + b2 + x) in place over x in the X band; coalesced store of the band ----------------
        uint8_t* const xg = gen + (smem_x - base) + buf * (uint32_t)p.x_bytes;
        ptx::mbar_wait(bar_c1done + 8 * buf, (k >> 1) & 1u);           // conv1 of this band no longer reads X[buf]
        for (int j = 0; j < n2; ++j) {
          const int a = p.n1max + j;
          ptx::mbar_wait(bar_tfull + 8 * a, (pm >> a) & 1u);
          ptx::tc_fence_after();
          uint32_t a0[16];
          ptx::tmem_ld16(tmem_base + (uint32_t)(a * BLK_C + ch0) + ((uint32_t)(q * 32) << 16), a0);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * a);
          const int mp = j * 128 + q * 32 + lane;
          const int iy = blk_div_wp(mp, p.wp_magic), ix = mp - iy * p.Wp;
          if (mp < bh_eff * p.Wp && ix < p.W) {
            const uint32_t hx = (uint32_t)((iy + 2) * p.Wp + ix + 1);  // this pixel in the X band (2 halo rows, 1 zero column)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint4* const slot_p = reinterpret_cast<uint4*>(xg + blk_chunk(hx, (uint32_t)(2 * (grp - 2) + c)));
              float f[8], r[8];
              unpack8<T>(*slot_p, r);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(a0[8 * c + e]) + breg[8 * c + e] + r[e], 0.f);
              *slot_p = pack8<T>(f);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");                  // the band is complete in X[buf]
        {
          T* const out = reinterpret_cast<T*>(p.out);
          const uint64_t pol_out = ptx::policy_evict_last();
          const int npix = bh_eff * p.W;
          const int c = lane & 3;
          for (int px = (warp - 12) * 8 + (lane >> 2); px < npix; px += 64) {
            const int oy = px / p.W, ox = px - oy * p.W;
            const uint32_t hx = (uint32_t)((oy + 2) * p.Wp + ox + 1);
            const uint4 v = *reinterpret_cast<const uint4*>(xg + blk_chunk(hx, (uint32_t)c));
            ptx::st_global_v4_hint(out + ((size_t)((img * p.H + y0 + oy) * p.W + ox)) * BLK_C + c * 8, v, pol_out);
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_xempty + 8 * buf);          // this warp no longer reads X[buf]
      }
      pm ^= ((1u << n1) - 1u) | (((1u << n2) - 1u) << p.n1max);
      n2_prev = n2;
      if (++bin == p.bands_per_img) { bin = 0; ++img; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// =======================================================================================================
// host side
// =======================================================================================================
struct TcBlockState {
  CUtensorMap mapX, mapW1, mapW2;
  BlockP p;
  int grid, smem_bytes, dtype;
  TcBlock64State* b64 = nullptr;   // non-null: a 64-channel block on CTA pairs (capf_tc_block64.cu)
};

static int block_plan(const capf_op& op, BlockP& p, int& smem_bytes) {
  const int N = op.i[0], H = op.i[1], W = op.i[2], C = op.i[3];
  if (C != BLK_C || N <= 0 || H <= 0 || W <= 0 || W + 1 > 256) return 0;
  if (op.dtype_in != op.dtype_out || (op.dtype_in != CAPF_F16 && op.dtype_in != CAPF_BF16)) return 0;
  if ((long long)N * H * W * C >= (1ll << 31)) return 0;
  memset(&p, 0, sizeof(p));
  p.H = H; p.W = W; p.Nimg = N; p.Wp = W + 1;
  p.wp_magic = (uint32_t)(((1ull << 32) + p.Wp - 1) / p.Wp);
  const int fixed = 1024 + BLK_HEADER + 2 * BLK_W_BYTES;
  int best_bh = 0;
  double best_cost = 1e300;
  for (int bh = 1; bh <= H && bh + 5 <= 256; ++bh) {
    const int n1 = ((bh + 2) * p.Wp + 127) / 128, n2 = (bh * p.Wp + 127) / 128;
    if (n1 + n2 > BLK_MAX_ACC) break;
    const int px = ((std::max((bh + 5) * p.Wp, n1 * 128 + 2 * p.Wp + 2) + 15) & ~15);
    const int pm = ((std::max((bh + 2) * p.Wp + 1, n2 * 128 + 2 * p.Wp + 2) + 15) & ~15);
    const int xb = (px * BLK_PIX + 1023) & ~1023, mb = (pm * BLK_PIX + 1023) & ~1023;
    if (fixed + 2 * xb + mb > TC_SMEM_LIMIT) break;
    const int full = H / bh, rem = H - full * bh;
    double tiles = (double)full * (n1 + n2);
    if (rem) tiles += ((rem + 2) * p.Wp + 127) / 128 + (rem * p.Wp + 127) / 128;
    if (tiles < best_cost - 1e-9 || (tiles < best_cost + 1e-9 && bh > best_bh)) { best_cost = tiles; best_bh = bh; }
  }
  if (!best_bh) return 0;
  p.bh = best_bh;
  p.bands_per_img = (H + p.bh - 1) / p.bh;
  if ((long long)N * p.bands_per_img >= (1ll << 31)) return 0;
  p.num_bands = N * p.bands_per_img;
  p.n1max = ((p.bh + 2) * p.Wp + 127) / 128;
  p.n2max = (p.bh * p.Wp + 127) / 128;
  const int px = ((std::max((p.bh + 5) * p.Wp, p.n1max * 128 + 2 * p.Wp + 2) + 15) & ~15);
  const int pm = ((std::max((p.bh + 2) * p.Wp + 1, p.n2max * 128 + 2 * p.Wp + 2) + 15) & ~15);
  p.x_bytes = (px * BLK_PIX + 1023) & ~1023;
  p.mid_bytes = (pm * BLK_PIX + 1023) & ~1023;
  p.x_tx_bytes = (p.bh + 5) * p.Wp * BLK_PIX;      // bh + 4 halo rows + one more: the rightmost tap of the last row reads the next row's zero column
  int cols = 32;
  while (cols < (p.n1max + p.n2max) * BLK_C) cols <<= 1;
  p.tmem_cols = cols;
  smem_bytes = fixed + 2 * p.x_bytes + p.mid_bytes;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
  return 1;
}

int tc_block_supported(const capf_op& op) {
  const char* ev = getenv("CAPF_FUSE_BLOCKS");
  if (ev && ev[0] == '0') return 0;
  if (op.kind != CAPF_OP_BASICBLOCK) return 0;
  if (op.i[3] == 64) return tc_block64_supported(op);
  if (!op.in[0] || !op.in[1] || !op.in[3] || !op.out[0]) return 0;
  if (((uintptr_t)op.in[0] | (uintptr_t)op.in[1] | (uintptr_t)op.in[3] | (uintptr_t)op.out[0]) & 15) return 0;
  BlockP p;
  int smem;
  return block_plan(op, p, smem);
}

int tc_block_prepare(const capf_op& op, TcBlockState** out) {
  *out = nullptr;
  int e = tc_get_encoder();
  if (e) return e;
  TcBlockState* s = new (std::nothrow) TcBlockState();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_block_prepare: out of host memory");
  if (op.i[3] == 64) {
    e = tc_block64_prepare(op, &s->b64);
    if (e) { delete s; return e; }
    *out = s;
    return CAPF_OK;
  }
  if (!block_plan(op, s->p, s->smem_bytes)) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "fused BasicBlock: shape not supported"); }
  BlockP& p = s->p;
  const bool bf16 = op.dtype_in == CAPF_BF16;
  p.idesc = tc_idesc(bf16, BLK_C);
  p.desc_hi = tc_desc_hi(BLK_PIX, 8 * BLK_PIX);
  p.bias1 = (const float*)op.in[2];
  p.bias2 = (const float*)op.in[4];
  p.out = op.out[0];
  p.trace = (long long*)op.in[5];      // debug only (NULL in every program the host layer builds)
  s->grid = p.num_bands < num_sms() ? p.num_bands : num_sms();
  s->dtype = op.dtype_in;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  for (int w = 0; w < 2 && !e; ++w) {
    cuuint64_t dims[2] = {(cuuint64_t)(9 * BLK_C), (cuuint64_t)BLK_C};
    cuuint64_t strides[1] = {(cuuint64_t)(9 * BLK_C) * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLK_C, (cuuint32_t)BLK_C};
    cuuint32_t es[2] = {1, 1};
    e = tc_encode_map(w ? &s->mapW2 : &s->mapW1, dt, 2, w ? op.in[3] : op.in[1], dims, strides, box, es, BLK_PIX, "B weights (fused block)");
  }
  if (!e) {
    cuuint64_t adims[4] = {(cuuint64_t)BLK_C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.Nimg};
    cuuint64_t astr[3] = {(cuuint64_t)BLK_C * 2, (cuuint64_t)p.W * BLK_C * 2, (cuuint64_t)p.H * p.W * BLK_C * 2};
    cuuint32_t abox[4] = {(cuuint32_t)BLK_C, (cuuint32_t)p.Wp, (cuuint32_t)(p.bh + 5), 1};
    cuuint32_t aes[4] = {1, 1, 1, 1};
    e = tc_encode_map(&s->mapX, dt, 4, op.in[0], adims, astr, abox, aes, BLK_PIX, "X band (fused block)");
  }
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

template <typename T>
static int block_launch_typed(const TcBlockState* s, cudaStream_t st) {
  static PerDevice<bool> opted_;
  std::atomic<bool>& opted = opted_.get();
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(tc_block32_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_block32_kernel smem opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  launch_k(tc_block32_kernel<T>, dim3(s->grid), dim3(BLK_THREADS), s->smem_bytes, st, s->mapX, s->mapW1, s->mapW2, s->p);
  return check_launch("tc_block32_kernel");
}

int tc_block_launch(const TcBlockState* s, cudaStream_t st) {
  if (s->b64) return tc_block64_launch(s->b64, st);
  return s->dtype == CAPF_F16 ? block_launch_typed<__half>(s, st) : block_launch_typed<__nv_bfloat16>(s, st);
}

void tc_block_release(TcBlockState* s) {
  if (s && s->b64) tc_block64_release(s->b64);
  delete s;
}

void tc_block_describe(const TcBlockState* s, char* buf, int cap) {
  if (s->b64) { tc_block64_describe(s->b64, buf, cap); return; }
  snprintf(buf, cap, "tc_block32_kernel[fused BasicBlock, band %d rows, %d+%d sub-tiles]", s->p.bh, s->p.n1max, s->p.n2max);
}

}  // namespace capf
