// Fused HRNet BasicBlock for the 64-channel branch (pose_hrnet.py:66-95) on CTA PAIRS (cluster of 2, tcgen05 cta_group::2):
//     y = relu(bn2(conv2(relu(bn1(conv1(x))))) + x),   both convolutions 3x3 / stride 1 / pad 1, 64 -> 64, 16-bit NHWC.
//
// Same scheme as the 32-channel kernel (capf_tc_block.cu): per band of bh output rows a CTA keeps the input band X (bh + 4
// halo rows, TMA, double buffered: conv1 operand AND conv2's residual) and MID = relu(bn1(conv1(x))) on bh + 2 rows (written
// by the first epilogue straight into the swizzled K-major operand layout) in shared memory, and both convolutions read
// their operand through shifted-window descriptors (tap (r, s) = the same buffer moved by r * Wp + s pixels).
//
// What is different at 64 channels: the folded weights of the two convolutions are 2 x 72 KB -- with X and MID they do not
// fit one SM.  So two CTAs (the two SMs of a TPC) work as a pair: CTA r keeps only output channels [32 r, 32 r + 32) of
// both weight sets (2 x 36 KB), each CTA holds the band of ITS OWN image rows, and every MMA is a cta_group::2 instruction
// of M = 256 (rows 0-127: the leader's sub-tile, rows 128-255: the peer's), N = 64: each tensor core reads its local A and
// the two B halves over the pair datapath, and each CTA's TMEM receives the 128 x 64 accumulator of its own sub-tile.
// Per MMA an SM reads 4 KB of A + 1 KB of B instead of 4 + 2 KB, the intermediate tensor never leaves the SM, and the block is
// one launch instead of two.
//
// Protocol (barriers live at the same offset in both CTAs; "leader" = cluster rank 0, which issues every MMA):
//   w          leader: the weight halves of BOTH CTAs have landed (the peer's TMA completes on the leader's barrier)
//   xfull[b]   leader: X[b] of both CTAs landed;   xempty[b]  both: the issuers' multicast commits + the CTA's own 8 epilogue-2 warps
//   tfull[a]   both (multicast commit);            tempty[a]  leader: the epilogue warps of both CTAs (the peer's arrive remotely)
//   midrdy[a]  leader: epilogue-1 warps of both CTAs wrote the MID pixels of sub-tile a (MID rows are released to the next band's
//              epilogue 1 by the tfull barriers of the conv2 sub-tiles that read them);   c1done[b]  both: conv1 finished with X[b]
// Shared memory: header | X0 | X1 | MID | W1 | W2.  The shifted windows of a band's last sub-tile read past the band (rows that
// are dropped); the order of the buffers makes those reads land in the next buffer of the same allocation.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <new>

#include "capf_tc.cuh"

namespace capf {

constexpr int B64_THREADS = 640;            // warps 0-3: TMA / conv1 MMA / TMEM / conv2 MMA; warps 4-19: four 4-warp epilogue groups
constexpr int B64_C = 64;
constexpr int B64_PIX = B64_C * 2;          // bytes per pixel row = the 128-byte swizzle span
constexpr int B64_TAP_BYTES = 32 * B64_PIX; // one tap of this CTA's weight half: 32 output channels x 64 input channels
constexpr int B64_WHALF_BYTES = 9 * B64_TAP_BYTES;   // 36 KB
constexpr int B64_MAX_ACC = 8;              // 8 x 64 TMEM columns
constexpr int B6_W = 0, B6_XFULL = 8, B6_XEMPTY = 24, B6_MIDFREE = 40, B6_C1DONE = 48, B6_TFULL = 64, B6_TEMPTY = 128, B6_MIDRDY = 192, B6_TMEM = 256;
constexpr int B64_HEADER = 1024;

struct Block64P {
  int H, W, Nimg, Wp;
  uint32_t wp_magic;
  int bh, bands_per_img, num_bands, num_slots;
  int n1max, n2max;
  int x_bytes, mid_bytes, x_tx_bytes, tmem_cols;
  uint32_t idesc, desc_hi;
  const float* bias1;
  const float* bias2;
  void* out;
  long long* trace;      // optional (debug, op.in[5]): per-band wait cycles of pair 0's issuers, see tools/block_trace.py
};

__device__ __forceinline__ int b64_div_wp(int v, uint32_t magic) { return (int)__umulhi((uint32_t)v, magic); }

// 16-byte chunk c of pixel h in a 128-byte-swizzled pixel-row buffer whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t b64_chunk(uint32_t h, uint32_t c) { return h * (uint32_t)B64_PIX + ((c ^ (h & 7u)) << 4); }

// the band a CTA works on in pair slot `slot` (clamped: the peer of an odd last band repeats it and stores nothing)
struct B64Band {
  int img, bin, y0, bh_eff;
  bool live;
  __device__ __forceinline__ void set(const Block64P& p, int slot, int rank) {
    int band = 2 * slot + rank;
    live = band < p.num_bands;
    if (!live) band = p.num_bands - 1;
    img = band / p.bands_per_img;
    bin = band - img * p.bands_per_img;
    y0 = bin * p.bh;
    bh_eff = min(p.bh, p.H - y0);
  }
};
// sub-tile counts of a slot: the larger of its two bands (both CTAs step through the same MMAs)
__device__ __forceinline__ void b64_slot_tiles(const Block64P& p, int slot, int& n1, int& n2) {
  const int b0 = 2 * slot, b1 = min(2 * slot + 1, p.num_bands - 1);
  const int bin0 = b0 % p.bands_per_img, bin1 = b1 % p.bands_per_img;
  const int bh = max(min(p.bh, p.H - bin0 * p.bh), min(p.bh, p.H - bin1 * p.bh));
  n1 = ((bh + 2) * p.Wp + 127) >> 7;
  n2 = (bh * p.Wp + 127) >> 7;
}

template <typename T>
__global__ void __launch_bounds__(B64_THREADS, 1)
tc_block64_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                  const Block64P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_w = base + B6_W, bar_xfull = base + B6_XFULL, bar_xempty = base + B6_XEMPTY;
  const uint32_t bar_c1done = base + B6_C1DONE;
  const uint32_t bar_tfull = base + B6_TFULL, bar_tempty = base + B6_TEMPTY, bar_midrdy = base + B6_MIDRDY, tmem_slot = base + B6_TMEM;
  const uint32_t smem_x = base + B64_HEADER;
  const uint32_t smem_mid = smem_x + 2u * (uint32_t)p.x_bytes;
  const uint32_t smem_w1 = smem_mid + (uint32_t)p.mid_bytes, smem_w2 = smem_w1 + B64_WHALF_BYTES;
  uint8_t* const gen = smem_raw + (base - raw);                 // generic pointer to `base`
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + B6_TMEM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx2::cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapX);
    ptx::prefetch_tmap(&mapW1);
    ptx::prefetch_tmap(&mapW2);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(bar_w, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_xfull + 8 * b, 1);
      ptx::mbar_init(bar_xempty + 8 * b, 1 + 8);                // conv1's multicast commit + this CTA's 8 epilogue-2 warps (residuals, in-place results, band store)
      ptx::mbar_init(bar_c1done + 8 * b, 1);                    // conv1's multicast commit: epilogue 2 may overwrite X[b]
    }
    for (int a = 0; a < B64_MAX_ACC; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);
      ptx::mbar_init(bar_tempty + 8 * a, 16);                    // one arrival per epilogue warp of the phase, both CTAs: 2 x 8
      ptx::mbar_init(bar_midrdy + 8 * a, 16);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {                   // one warp of EACH CTA of the pair executes the cta_group::2 allocation
    ptx2::tmem_alloc2(tmem_slot, (uint32_t)p.tmem_cols);
    ptx2::tmem_relinquish2();
  }
  if (warp == 3) {                   // the zero pixel in front of MID (image column -1 of its first row)
    if (lane < 8) *reinterpret_cast<uint4*>(gen + (smem_mid - base) + 16 * lane) = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx2::cluster_sync();              // barriers of both CTAs initialised, TMEM of both allocated
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();
  if (warp != 0) pdl_wait();

  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int slot0 = (int)(((long long)p.num_slots * pair) / n_pairs);
  const int slot1 = (int)(((long long)p.num_slots * (pair + 1)) / n_pairs);

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ==========================
    if (ptx::elect_one()) {
      const uint32_t w_leader = ptx2::mapa(bar_w, 0u);
      if (leader) ptx::mbar_arrive_expect_tx(bar_w, 4u * B64_WHALF_BYTES);
      for (int c = 0; c < 9; ++c) {
        ptx2::tma_load_2d_2sm(&mapW1, w_leader, smem_w1 + c * B64_TAP_BYTES, c * B64_C, (int)rank * 32);
        ptx2::tma_load_2d_2sm(&mapW2, w_leader, smem_w2 + c * B64_TAP_BYTES, c * B64_C, (int)rank * 32);
      }
      pdl_wait();
      const uint64_t pol_in = ptx::policy_evict_first();
      uint32_t k = 0;
      for (int slot = slot0; slot < slot1; ++slot, ++k) {
        const uint32_t buf = k & 1u, ph = (k >> 1) & 1u;
        B64Band bd;
        bd.set(p, slot, (int)rank);
        ptx::mbar_wait(bar_xempty + 8 * buf, ph ^ 1u);
        if (leader) ptx::mbar_arrive_expect_tx(bar_xfull + 8 * buf, 2u * (uint32_t)p.x_tx_bytes);
        ptx2::tma_load_4d_2sm_hint(&mapX, ptx2::mapa(bar_xfull + 8 * buf, 0u), smem_x + buf * (uint32_t)p.x_bytes, 0, -1, bd.y0 - 2, bd.img, pol_in);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer, phase A (leader CTA only) ==============
    // One warp issues conv1 of every band, another (warp 3) conv2, each sub-tile by sub-tile IN ORDER: the tensor pipe then
    // finishes sub-tile j while epilogue 1 of sub-tile j - 1 runs, and this warp is already issuing conv1 of the next band
    // while conv2 of the current one is in the pipe.  (Three issuers taking sub-tiles round-robin interleave their MMAs, all
    // sub-tiles of a phase complete together and the pipe idles for the whole first epilogue: measured 15 k clk per band
    // against 9.4 k of MMAs.)  A single thread sustains the pipe's 43 clk per M = 256 / N = 64 MMA (tools/microbench/mma_rate.cu).
    if (leader) {
      ptx::mbar_wait(bar_w, 0);
      ptx::tc_fence_after();
      uint32_t tap_off[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) tap_off[tap] = (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 8u;     // 8 x 16 B per pixel
      const uint32_t w1_lo = tc_desc_lo(smem_w1, 1u);
      uint32_t k = 0, pm = 0;         // pm: phase parity of the per-accumulator barriers (bit a flips when slot a is used)
      for (int slot = slot0; slot < slot1; ++slot, ++k) {
        const uint32_t buf = k & 1u, ph = (k >> 1) & 1u;
        int n1, n2;
        b64_slot_tiles(p, slot, n1, n2);
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && k < 32;
        long long tw = 0, w_x = 0, w_ta = 0, w_is = 0;
        if (tr) tw = clock64();
        ptx::mbar_wait(bar_xfull + 8 * buf, ph);
        if (tr) { w_x = clock64() - tw; p.trace[0 * 32 + k] = clock64(); }
        ptx::tc_fence_after();
        const uint32_t x_lo = tc_desc_lo(smem_x + buf * (uint32_t)p.x_bytes, 1u);
        for (int j = 0; j < n1; ++j) {
          if (tr) tw = clock64();
          ptx::mbar_wait(bar_tempty + 8 * j, ((pm >> j) & 1u) ^ 1u);
          if (tr) { w_ta += clock64() - tw; tw = clock64(); }
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t d = tmem_base + (uint32_t)(j * B64_C), a_sub = x_lo + (uint32_t)(j * 128) * 8u;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                ptx2::umma2_f16_lohi(d, a_sub + tap_off[tap] + 2u * kk, p.desc_hi, w1_lo + (uint32_t)tap * (B64_TAP_BYTES >> 4) + 2u * kk, p.desc_hi,
                                     p.idesc, (tap | kk) ? 1u : 0u);
            }
            ptx2::umma2_commit_mc(bar_tfull + 8 * j);
          }
          __syncwarp();
          if (tr) w_is += clock64() - tw;
        }
        if (ptx::elect_one()) {                                                  // conv1's reads of X[buf] complete (both CTAs)
          ptx2::umma2_commit_mc(bar_xempty + 8 * buf);
          ptx2::umma2_commit_mc(bar_c1done + 8 * buf);
        }
        __syncwarp();
        if (tr) {
          p.trace[1 * 32 + k] = w_x; p.trace[2 * 32 + k] = w_ta; p.trace[5 * 32 + k] = clock64(); p.trace[6 * 32 + k] = w_is;
        }
        pm ^= (1u << n1) - 1u;
      }
    }
  } else if (warp == 3) {
    // ===================================== MMA issuer, phase B (leader CTA only) ==============
    if (leader) {
      ptx::mbar_wait(bar_w, 0);
      ptx::tc_fence_after();
      uint32_t tap_off[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) tap_off[tap] = (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 8u;
      const uint32_t w2_lo = tc_desc_lo(smem_w2, 1u), mid_lo = tc_desc_lo(smem_mid, 1u);
      uint32_t k = 0, pm = 0;
      for (int slot = slot0; slot < slot1; ++slot, ++k) {
        int n1, n2;
        b64_slot_tiles(p, slot, n1, n2);
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && k < 32;
        long long tw = 0, w_mid = 0, w_tb = 0, w_is = 0;
        if (tr) p.trace[8 * 32 + k] = clock64();
        // conv2 over bh rows, sub-tile j as soon as the MID pixels it reads exist in both CTAs
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
          if (tr) { w_tb += clock64() - tw; tw = clock64(); }
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t d = tmem_base + (uint32_t)(a * B64_C), a_sub = mid_lo + (uint32_t)(j * 128) * 8u;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                ptx2::umma2_f16_lohi(d, a_sub + tap_off[tap] + 2u * kk, p.desc_hi, w2_lo + (uint32_t)tap * (B64_TAP_BYTES >> 4) + 2u * kk, p.desc_hi,
                                     p.idesc, (tap | kk) ? 1u : 0u);
            }
            ptx2::umma2_commit_mc(bar_tfull + 8 * a);
          }
          __syncwarp();
          if (tr) w_is += clock64() - tw;
        }
        while (ready_upto < n1 - 1) {            // keep this warp's view of every midrdy barrier in step (ragged bands)
          ++ready_upto;
          ptx::mbar_wait(bar_midrdy + 8 * ready_upto, (pm >> ready_upto) & 1u);
        }
        if (tr) {
          p.trace[(8 + 3) * 32 + k] = w_mid; p.trace[(8 + 4) * 32 + k] = w_tb; p.trace[(8 + 5) * 32 + k] = clock64(); p.trace[(8 + 6) * 32 + k] = w_is;
        }
        pm ^= ((1u << n1) - 1u) | (((1u << n2) - 1u) << p.n1max);
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogues (both CTAs: own band) ====================
    // Groups 0 and 1 both work on every phase-A sub-tile, groups 2 and 3 on every phase-B sub-tile, 32 of the 64 channels each
    // (half the latency per sub-tile, and a thread's 32 folded-BN shifts live in registers).  Measured with the first version
    // (shifts read from shared memory, epilogue 2 storing each pixel's 128 bytes with eight per-lane 16-byte st.global = 32
    // cache lines per instruction): ~100 instructions of epilogue 1 took 1.4-3.7 k clk and one epilogue-2 sub-tile 3 k clk --
    // the load/store unit, not the tensor pipe, paced the kernel.  Now the epilogues issue only the shared-memory vectors
    // they must (4 STS / 4 LDS + 4 STS per thread and sub-tile): epilogue 2 writes the result over the pixel's residual IN
    // the X band, and at the end of the band its 8 warps copy the band out with whole 128-byte lines per 8 lanes.
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const uint32_t tempty_leader = ptx2::mapa(bar_tempty, 0u), midrdy_leader = ptx2::mapa(bar_midrdy, 0u);
    const int ch0 = 32 * (grp & 1);                                     // this group's 32 channels = 4 chunks of the pixel row
    float breg[32];
    {
      const float* bsrc = grp < 2 ? p.bias1 : p.bias2;
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 v = bsrc ? __ldg(reinterpret_cast<const float4*>(bsrc + ch0 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
        breg[e] = v.x; breg[e + 1] = v.y; breg[e + 2] = v.z; breg[e + 3] = v.w;
      }
    }
    int n2_prev = 0;
    uint32_t k = 0, pm = 0;
    for (int slot = slot0; slot < slot1; ++slot, ++k) {
      const uint32_t buf = k & 1u;
      B64Band bd;
      bd.set(p, slot, (int)rank);
      int n1, n2;
      b64_slot_tiles(p, slot, n1, n2);
      if (grp < 2) {
        // ---- epilogue 1: relu(acc + b1), zero outside the image, 16-bit, into MID (shifted by one pixel) ----
        uint8_t* const mg = gen + (smem_mid - base);
        const bool tre = p.trace != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0;
        int freed_upto = -1;
        for (int j = 0; j < n1; ++j) {
          const int ev = (int)k * n1 + j;
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
          if (tre && ev < 32) p.trace[16 * 32 + ev] = clock64();
          ptx::mbar_wait(bar_tfull + 8 * j, (pm >> j) & 1u);
          ptx::tc_fence_after();
          if (tre && ev < 32) p.trace[17 * 32 + ev] = clock64();
          uint32_t a0[16], a1[16];
          const uint32_t taddr = tmem_base + (uint32_t)(j * B64_C + ch0) + ((uint32_t)(q * 32) << 16);
          ptx::tmem_ld16(taddr, a0);
          ptx::tmem_ld16(taddr + 16u, a1);
          ptx::tmem_ld_wait();
          if (tre && ev < 32) p.trace[18 * 32 + ev] = clock64();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx2::mbar_arrive_cluster_relaxed(tempty_leader + 8 * j);
          const int mp = j * 128 + q * 32 + lane;
          const int iy = b64_div_wp(mp, p.wp_magic), ix = mp - iy * p.Wp;
          const int yi = bd.y0 - 1 + iy;
          const bool valid = ix < p.W && yi >= 0 && yi < p.H;
          if (mp < (bd.bh_eff + 2) * p.Wp) {
            const uint32_t h = (uint32_t)mp + 1u;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float acc = __uint_as_float(c < 2 ? a0[8 * c + e] : a1[8 * (c - 2) + e]);
                f[e] = valid ? fmaxf(acc + breg[8 * c + e], 0.f) : 0.f;
              }
              *reinterpret_cast<uint4*>(mg + b64_chunk(h, (uint32_t)(4 * grp + c))) = pack8<T>(f);
            }
          }
          if (tre && ev < 32) p.trace[19 * 32 + ev] = clock64();
          ptx::fence_proxy_async();                                     // generic-proxy writes of MID -> tensor-pipe reads
          __syncwarp();
          if (lane == 0) ptx2::mbar_arrive_cluster_relaxed(midrdy_leader + 8 * j);
          if (tre && ev < 32) p.trace[20 * 32 + ev] = clock64();
        }
      } else {
        // ---- epilogue 2: relu(acc + b2 + x) in place over x in the X band; coalesced store of the band ---------------------
        uint8_t* const xg = gen + (smem_x - base) + buf * (uint32_t)p.x_bytes;
        ptx::mbar_wait(bar_c1done + 8 * buf, (k >> 1) & 1u);           // conv1 of this band no longer reads X[buf] (either CTA)
        const bool tre = p.trace != nullptr && blockIdx.x == 0 && warp == 12 && lane == 0;
        for (int j = 0; j < n2; ++j) {
          const int a = p.n1max + j;
          const int ev = (int)k * n2 + j;
          if (tre && ev < 32) p.trace[21 * 32 + ev] = clock64();
          ptx::mbar_wait(bar_tfull + 8 * a, (pm >> a) & 1u);
          ptx::tc_fence_after();
          if (tre && ev < 32) p.trace[22 * 32 + ev] = clock64();
          uint32_t a0[16], a1[16];
          const uint32_t taddr = tmem_base + (uint32_t)(a * B64_C + ch0) + ((uint32_t)(q * 32) << 16);
          ptx::tmem_ld16(taddr, a0);
          ptx::tmem_ld16(taddr + 16u, a1);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx2::mbar_arrive_cluster_relaxed(tempty_leader + 8 * a);
          const int mp = j * 128 + q * 32 + lane;
          const int iy = b64_div_wp(mp, p.wp_magic), ix = mp - iy * p.Wp;
          if (mp < bd.bh_eff * p.Wp && ix < p.W) {
            const uint32_t hx = (uint32_t)((iy + 2) * p.Wp + ix + 1);  // this pixel in the X band (2 halo rows, 1 zero column)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4* const slot_p = reinterpret_cast<uint4*>(xg + b64_chunk(hx, (uint32_t)(4 * (grp - 2) + c)));
              float f[8], r[8];
              unpack8<T>(*slot_p, r);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float acc = __uint_as_float(c < 2 ? a0[8 * c + e] : a1[8 * (c - 2) + e]);
                f[e] = fmaxf(acc + breg[8 * c + e] + r[e], 0.f);
              }
              *slot_p = pack8<T>(f);
            }
          }
          if (tre && ev < 32) p.trace[23 * 32 + ev] = clock64();
        }
        // the band is complete in X[buf] once all 8 epilogue-2 warps are here; they store it together, 8 lanes per pixel
        // (one 128-byte line) and 4 pixels per instruction.  (A TMA tensor store of the band would need its first row at a
        // 1024-byte-aligned shared-memory address with the 128-byte swizzle -- 2 * Wp pixels into the band it is not: the
        // instruction faults.)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (bd.live) {
          T* const out = reinterpret_cast<T*>(p.out);
          const uint64_t pol_out = ptx::policy_evict_last();
          const int npix = bd.bh_eff * p.W;
          const int c = lane & 7;
          for (int px = (warp - 12) * 4 + (lane >> 3); px < npix; px += 32) {
            const int oy = px / p.W, ox = px - oy * p.W;
            const uint32_t hx = (uint32_t)((oy + 2) * p.Wp + ox + 1);
            const uint4 v = *reinterpret_cast<const uint4*>(xg + b64_chunk(hx, (uint32_t)c));
            ptx::st_global_v4_hint(out + ((size_t)((bd.img * p.H + bd.y0 + oy) * p.W + ox)) * B64_C + c * 8, v, pol_out);
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_xempty + 8 * buf);          // this warp no longer reads X[buf]
      }
      pm ^= ((1u << n1) - 1u) | (((1u << n2) - 1u) << p.n1max);
      n2_prev = n2;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx2::cluster_sync();              // the peer may still be reading operands of / arriving at this CTA
  if (warp == 2) ptx2::tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
}

// =======================================================================================================
// host side
// =======================================================================================================
struct TcBlock64State {
  CUtensorMap mapX, mapW1, mapW2;
  Block64P p;
  int grid, smem_bytes, dtype;
};

static int round1k(int v) { return (v + 1023) & ~1023; }

static int block64_plan(const capf_op& op, Block64P& p, int& smem_bytes) {
  const int N = op.i[0], H = op.i[1], W = op.i[2], C = op.i[3];
  if (C != B64_C || N <= 0 || H <= 0 || W <= 0 || W + 1 > 256) return 0;
  if (num_sms() < 2) return 0;
  if (op.dtype_in != op.dtype_out || (op.dtype_in != CAPF_F16 && op.dtype_in != CAPF_BF16)) return 0;
  if ((long long)N * H * W * C >= (1ll << 31)) return 0;
  memset(&p, 0, sizeof(p));
  p.H = H; p.W = W; p.Nimg = N; p.Wp = W + 1;
  p.wp_magic = (uint32_t)(((1ull << 32) + p.Wp - 1) / p.Wp);
  const int fixed = 1024 + B64_HEADER + 2 * B64_WHALF_BYTES;
  int best_bh = 0;
  double best_cost = 1e300;
  for (int bh = 1; bh <= H && bh + 5 <= 256; ++bh) {
    const int n1 = ((bh + 2) * p.Wp + 127) / 128, n2 = (bh * p.Wp + 127) / 128;
    if (n1 + n2 > B64_MAX_ACC) break;
    const int xb = round1k((bh + 5) * p.Wp * B64_PIX), mb = round1k(((bh + 2) * p.Wp + 1) * B64_PIX);
    if (fixed + 2 * xb + mb > TC_SMEM_LIMIT) break;
    // the windows of the last sub-tiles read up to n * 128 + 2 Wp + 2 pixels from the start of X1 / MID: inside the allocation
    if ((n1 * 128 + 2 * p.Wp + 2) * B64_PIX > xb + mb + 2 * B64_WHALF_BYTES) continue;
    if ((n2 * 128 + 2 * p.Wp + 2) * B64_PIX > mb + 2 * B64_WHALF_BYTES) continue;
    const int full = H / bh, rem = H - full * bh;
    double tiles = (double)full * (n1 + n2);
    if (rem) tiles += ((rem + 2) * p.Wp + 127) / 128 + (rem * p.Wp + 127) / 128;
    if (tiles < best_cost - 1e-9 || (tiles < best_cost + 1e-9 && bh > best_bh)) { best_cost = tiles; best_bh = bh; }
  }
  if (!best_bh) return 0;
  p.bh = best_bh;
  p.bands_per_img = (H + p.bh - 1) / p.bh;
  if ((long long)N * p.bands_per_img >= (1ll << 30)) return 0;
  p.num_bands = N * p.bands_per_img;
  p.num_slots = (p.num_bands + 1) / 2;
  p.n1max = ((p.bh + 2) * p.Wp + 127) / 128;
  p.n2max = (p.bh * p.Wp + 127) / 128;
  p.x_bytes = round1k((p.bh + 5) * p.Wp * B64_PIX);
  p.mid_bytes = round1k(((p.bh + 2) * p.Wp + 1) * B64_PIX);
  p.x_tx_bytes = (p.bh + 5) * p.Wp * B64_PIX;      // bh + 4 halo rows + one more: the rightmost tap of the last row reads the next row's zero column
  int cols = 32;
  while (cols < (p.n1max + p.n2max) * B64_C) cols <<= 1;
  p.tmem_cols = cols;
  smem_bytes = fixed + 2 * p.x_bytes + p.mid_bytes;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;     // one CTA per SM: the pair owns both TMEMs
  return 1;
}

int tc_block64_supported(const capf_op& op) {
  const char* ev = getenv("CAPF_FUSE_BLOCKS64");
  if (ev && ev[0] == '0') return 0;
  if (op.kind != CAPF_OP_BASICBLOCK) return 0;
  if (!op.in[0] || !op.in[1] || !op.in[3] || !op.out[0]) return 0;
  if (((uintptr_t)op.in[0] | (uintptr_t)op.in[1] | (uintptr_t)op.in[3] | (uintptr_t)op.out[0]) & 15) return 0;
  Block64P p;
  int smem;
  return block64_plan(op, p, smem);
}

int tc_block64_prepare(const capf_op& op, TcBlock64State** out) {
  *out = nullptr;
  int e = tc_get_encoder();
  if (e) return e;
  TcBlock64State* s = new (std::nothrow) TcBlock64State();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_block64_prepare: out of host memory");
  if (!block64_plan(op, s->p, s->smem_bytes)) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "fused BasicBlock (64 channels): shape not supported"); }
  Block64P& p = s->p;
  const bool bf16 = op.dtype_in == CAPF_BF16;
  const uint32_t fmt = bf16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(B64_C >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // M = 256 (pair), N = 64
  p.desc_hi = tc_desc_hi(B64_PIX, 8 * B64_PIX);
  p.bias1 = (const float*)op.in[2];
  p.bias2 = (const float*)op.in[4];
  p.out = op.out[0];
  p.trace = (long long*)op.in[5];      // debug only (NULL in every program the host layer builds)
  const int pairs = num_sms() / 2;
  s->grid = 2 * (p.num_slots < pairs ? p.num_slots : pairs);
  s->dtype = op.dtype_in;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  for (int w = 0; w < 2 && !e; ++w) {
    cuuint64_t dims[2] = {(cuuint64_t)(9 * B64_C), (cuuint64_t)B64_C};
    cuuint64_t strides[1] = {(cuuint64_t)(9 * B64_C) * 2};
    cuuint32_t box[2] = {(cuuint32_t)B64_C, 32u};
    cuuint32_t es[2] = {1, 1};
    e = tc_encode_map(w ? &s->mapW2 : &s->mapW1, dt, 2, w ? op.in[3] : op.in[1], dims, strides, box, es, B64_PIX, "B weight halves (fused block, pair)");
  }
  if (!e) {
    cuuint64_t adims[4] = {(cuuint64_t)B64_C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.Nimg};
    cuuint64_t astr[3] = {(cuuint64_t)B64_C * 2, (cuuint64_t)p.W * B64_C * 2, (cuuint64_t)p.H * p.W * B64_C * 2};
    cuuint32_t abox[4] = {(cuuint32_t)B64_C, (cuuint32_t)p.Wp, (cuuint32_t)(p.bh + 5), 1};
    cuuint32_t aes[4] = {1, 1, 1, 1};
    e = tc_encode_map(&s->mapX, dt, 4, op.in[0], adims, astr, abox, aes, B64_PIX, "X band (fused block, pair)");
  }
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

template <typename T>
static int block64_launch_typed(const TcBlock64State* s, cudaStream_t st) {
  static PerDevice<bool> opted_;
  std::atomic<bool>& opted = opted_.get();
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(tc_block64_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_block64_kernel smem opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(s->grid);
  cfg.blockDim = dim3(B64_THREADS);
  cfg.dynamicSmemBytes = s->smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_block64_kernel<T>, s->mapX, s->mapW1, s->mapW2, s->p);
  if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_block64_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("tc_block64_kernel");
}

int tc_block64_launch(const TcBlock64State* s, cudaStream_t st) {
  return s->dtype == CAPF_F16 ? block64_launch_typed<__half>(s, st) : block64_launch_typed<__nv_bfloat16>(s, st);
}

void tc_block64_release(TcBlock64State* s) { delete s; }

void tc_block64_describe(const TcBlock64State* s, char* buf, int cap) {
  snprintf(buf, cap, "tc_block64_kernel[fused BasicBlock on CTA pairs, band %d rows, %d+%d sub-tiles]", s->p.bh, s->p.n1max, s->p.n2max);
}

}  // namespace capf
