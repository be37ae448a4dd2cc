// tcgen05 3x3 / stride-1 / pad-1 convolution with a shared-memory halo band ("shifted-window" implicit GEMM).
//
// The per-tap TMA kernel (capf_tc.cu) re-reads every input pixel nine times from L2, which is what bounds the
// C = 32 / 64 BasicBlock convolutions of HRNet (pose_hrnet.py:66-95) -- 47 % of the backbone FLOPs.  Here a CTA
// loads a band of (bh + 2) input rows ONCE into shared memory and all nine filter taps read it in place:
//
//   halo[chunk][pixel][8 ch]   un-swizzled K-major "core matrix" layout: 16 bytes per (pixel, 8-channel chunk),
//                              pixels contiguous, one zero column shared by the left / right padding (Wp = W + 1)
//
// so the A operand of tap (r, s) for the 128 consecutive padded pixels m'..m'+127 is the SAME buffer with its start
// address moved by (r * Wp + s) * 16 bytes -- a descriptor change, not a copy.  Rows whose padded x lands on the
// zero column are computed and dropped (1 / Wp of the tensor work).  The folded weights ([Cout][9 * Cin], swizzled
// K-major chunks exactly as in capf_tc.cu) are fetched once per CTA by TMA and stay resident.
//
// Roles (512 threads): warp 0 = TMA producer (weights once, then the halo of each band; out-of-image elements are
// zero-filled by TMA), warps 1 and 3 = tcgen05.mma issuers, warp 2 = TMEM allocator, warps 4..15 = three 4-warp
// epilogue groups taking 128-row sub-tiles round-robin.  Halo bands are double buffered; four TMEM accumulators.
#include <cstdio>
#include <cstdlib>
#include <new>

#include "capf_tc.cuh"

namespace capf {

constexpr int HALO_THREADS = 512;
constexpr int HALO_HEADER_BYTES = 1024;
constexpr int HALO_BIAS_OFF = 512;            // 64 fp32 bias values inside the header
constexpr int HALO_EPI_GROUPS = 3;          // warps 4..15 (measured: a 4th group costs band height, 11.97 vs 11.6 ms/step)
constexpr int HALO_MAX_ACC = 2 * HALO_EPI_GROUPS;   // two TMEM accumulator stages per epilogue group

struct HaloP {
  int C, Cout;              // Cout == UMMA N (single column tile)
  int H, W, Nimg;           // stride 1, pad 1: output size == input size
  int Wp;                   // padded row pitch in pixels = W + 1
  uint32_t wp_magic;        // ceil(2^32 / Wp)
  int bh;                   // band height (output rows per band)
  int bands_per_img, num_bands;
  int P_alloc;              // pixels per chunk plane of one halo buffer
  int chunks;               // C / 8   (16-byte channel chunks per pixel)
  int kb, cpt;              // weight chunk width (elements) and chunks per tap, as in capf_tc.cu
  int b_chunk_bytes, b_bytes;
  int halo_bytes;           // one halo buffer
  int plane_tx_bytes;       // bytes per 8-channel plane of a band (n_boxes slabs of box_rows halo rows)
  int box_rows, n_boxes;    // the band is fetched as n_boxes TMA boxes of box_rows rows each
  int acc_stages, tmem_cols, acc_stride;
  uint32_t idesc, b_desc_hi, a_desc_hi;
  int a_rows;               // 0: halo stored as un-swizzled 8-channel planes; 1: swizzled pixel rows of C channels (C = 16|32|64)
  int act;
  const void* x;
  const float* bias;
  const void* res;
  void* out;
  int l2_hints;             // 1: evict_first input / residual reads, evict_last output writes (env CAPF_L2_HINTS)
  long long* trace;         // optional (debug): per-role clock64 timeline of CTA 0, see tools/halo_trace.py
};

// v / Wp without a divide: magic = ceil(2^32 / Wp), exact for the pixel counts of one band (< 2^16)
__device__ __forceinline__ int div_wp(int v, uint32_t magic) { return (int)__umulhi((uint32_t)v, magic); }

// Walks the (band, 128-row sub-tile) sequence of one CTA in issue order.  Every role steps through the same sequence,
// which is what keeps the accumulator-stage / phase bookkeeping implicit.
struct HaloWalk {
  int band, band_end, img, bin, j, n_sub, bh_eff, y0;
  uint32_t it;      // sub-tile counter of this CTA
  uint32_t g, m;    // it % HALO_EPI_GROUPS (the epilogue group that drains it) and it / HALO_EPI_GROUPS
  __device__ __forceinline__ void load_band(const HaloP& p) {
    y0 = bin * p.bh;
    bh_eff = min(p.bh, p.H - y0);
    n_sub = (bh_eff * p.Wp + 127) >> 7;
  }
  __device__ __forceinline__ void init(const HaloP& p, int b0, int b1) {
    band = b0; band_end = b1;
    img = b0 / p.bands_per_img;
    bin = b0 - img * p.bands_per_img;
    j = 0; it = 0; g = 0; m = 0;
    load_band(p);
  }
  __device__ __forceinline__ bool valid() const { return band < band_end; }
  // Accumulator stage of the current sub-tile and the parity of its use count.  Every epilogue group owns two
  // stages, so a stage always has ONE producer (the issuer warp of that sub-tile parity: uses of a stage are 6
  // sub-tiles apart) and ONE consumer group -- the ordering the parity-based mbarrier waits rely on.
  __device__ __forceinline__ uint32_t acc() const { return 2u * g + (m & 1u); }
  __device__ __forceinline__ uint32_t use_parity() const { return (m >> 1) & 1u; }
  __device__ __forceinline__ void step(const HaloP& p) {
    ++it;
    if (++g == (uint32_t)HALO_EPI_GROUPS) { g = 0; ++m; }
    if (++j == n_sub) {
      j = 0;
      ++band;
      if (++bin == p.bands_per_img) { bin = 0; ++img; }
      load_band(p);
    }
  }
};

template <int C, int NV, typename TI, typename TO, bool RES>
__global__ void __launch_bounds__(HALO_THREADS, 1)
tc_conv3_halo_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const HaloP p) {
  constexpr int KSTEPS = C / 16;                                   // 16-channel MMA steps per filter tap
  constexpr int KB = C % 64 == 0 ? 64 : C % 32 == 0 ? 32 : 16;     // weight chunk width (elements)
  constexpr int KPC = KB / 16, CPT = C / KB, CHUNKS = C / 8;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_b = base;                       // weights landed
  const uint32_t bar_hfull = base + 8;               // [2] halo buffer filled   (TMA transaction bytes)
  const uint32_t bar_hempty = base + 24;             // [2] halo buffer consumed (tcgen05.commit)
  const uint32_t bar_tfull = base + 40;              // [HALO_MAX_ACC] accumulator complete
  const uint32_t bar_tempty = base + 40 + 8 * HALO_MAX_ACC;  // [HALO_MAX_ACC] accumulator drained (128 arrivals)
  const uint32_t tmem_slot = base + 40 + 16 * HALO_MAX_ACC;
  const uint32_t smem_b = base + HALO_HEADER_BYTES;
  const uint32_t smem_halo = smem_b + ((p.b_bytes + 1023) & ~1023);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapA);
    ptx::prefetch_tmap(&mapB);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(bar_b, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_hfull + 8 * b, 1);
      ptx::mbar_init(bar_hempty + 8 * b, 2);     // both MMA issuer warps commit
    }
    for (int a = 0; a < HALO_MAX_ACC; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);
      ptx::mbar_init(bar_tempty + 8 * a, 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp == 3 && lane < 16) {     // folded-BN shift (constant data, not produced by the predecessor kernel)
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias && 4 * lane < p.Cout) b4 = __ldg(reinterpret_cast<const float4*>(p.bias) + lane);
    *reinterpret_cast<float4*>(smem_raw + (base + HALO_BIAS_OFF - raw) + 16 * lane) = b4;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[11 * 256] = clock64();
  pdl_trigger();                                      // one-wave persistent grid: successor may be scheduled as SMs drain
  if (warp != 0) pdl_wait();                          // warp 0 first starts the (constant) weight fetch, then waits

  // contiguous band range of this CTA
  const int band0 = (int)(((long long)p.num_bands * blockIdx.x) / gridDim.x);
  const int band1 = (int)(((long long)p.num_bands * (blockIdx.x + 1)) / gridDim.x);

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    // weights once; then per band one box per 8-channel plane: {8 ch, Wp pixels from x = -1, bh + 3 rows from y0 - 1}.
    // Out-of-image elements arrive as zeros = the convolution padding (and the shared zero column).
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(bar_b, (uint32_t)p.b_bytes);
      for (int c = 0; c < 9 * CPT; ++c) ptx::tma_load_2d(&mapB, bar_b, smem_b + c * p.b_chunk_bytes, c * KB, 0);
      pdl_wait();                                     // activations of the predecessor are read from here on
      const int img0 = band0 / p.bands_per_img;
      int img = img0, bin = band0 - img0 * p.bands_per_img;
      const uint64_t pol_in = p.l2_hints ? ptx::policy_evict_first() : 0;
      uint32_t k = 0;
      for (int band = band0; band < band1; ++band, ++k) {
        const uint32_t buf = k & 1u, hph = (k >> 1) & 1u;
        ptx::mbar_wait(bar_hempty + 8 * buf, hph ^ 1u);
        if (p.trace && blockIdx.x == 0) p.trace[0 * 256 + (k & 255)] = clock64();
        const uint32_t full = bar_hfull + 8 * buf;
        ptx::mbar_arrive_expect_tx(full, (uint32_t)(CHUNKS * p.plane_tx_bytes));
        const uint32_t halo = smem_halo + buf * p.halo_bytes;
        const int y_top = bin * p.bh - 1;
        for (int b = 0; b < p.n_boxes; ++b) {
          const uint32_t slab = (uint32_t)(b * p.box_rows * p.Wp);          // first halo pixel of this slab
          if (p.a_rows) {
            if (p.l2_hints) ptx::tma_load_4d_hint(&mapA, full, halo + slab * (uint32_t)(C * 2), 0, -1, y_top + b * p.box_rows, img, pol_in);
            else ptx::tma_load_4d(&mapA, full, halo + slab * (uint32_t)(C * 2), 0, -1, y_top + b * p.box_rows, img);
          } else {
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c)
              ptx::tma_load_4d(&mapA, full, halo + ((uint32_t)c * (uint32_t)p.P_alloc + slab) * 16u, 8 * c, -1, y_top + b * p.box_rows, img);
          }
        }
        if (++bin == p.bands_per_img) { bin = 0; ++img; }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================================== MMA issuers ======================================
    // Two issuing warps (even / odd sub-tiles, disjoint accumulators): with N = Cout <= 64 one tcgen05.mma retires in
    // 16-64 cycles, about what a single thread needs to set one up, so one issuer alone would pace the tensor pipe.
    const uint32_t parity = warp == 1 ? 0u : 1u;
    ptx::mbar_wait(bar_b, 0);
    ptx::tc_fence_after();
    const uint32_t pix_units = p.a_rows ? (uint32_t)CHUNKS : 1u;      // descriptor address units (16 B) per halo pixel
    const uint32_t k_units = p.a_rows ? 2u : 2u * (uint32_t)p.P_alloc;  // ... per 16-channel K step
    const uint32_t a_lbo = p.a_rows ? 1u : (uint32_t)p.P_alloc;
    const uint32_t b_lo0 = tc_desc_lo(smem_b, 1u);
    const uint32_t b_chunk16 = (uint32_t)p.b_chunk_bytes >> 4;
    uint32_t tap_off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) tap_off[tap] = (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * pix_units;
    HaloWalk w;
    w.init(p, band0, band1);
    uint32_t k = 0;
    while (w.valid()) {
      const uint32_t buf = k & 1u, hph = (k >> 1) & 1u;
      long long tw0 = 0;
      if (p.trace && blockIdx.x == 0 && lane == 0) tw0 = clock64();
      ptx::mbar_wait(bar_hfull + 8 * buf, hph);
      ptx::tc_fence_after();
      if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[(1 + parity) * 256 + (k & 255)] = clock64() - tw0;   // halo wait
      const uint32_t a_lo0 = tc_desc_lo(smem_halo + buf * p.halo_bytes, a_lbo);
      const int n_sub = w.n_sub;
      for (int j = 0; j < n_sub; ++j) {
        if ((w.it & 1u) == parity) {
          const uint32_t acc = w.acc(), aph = w.use_parity();
          long long tw1 = 0;
          if (p.trace && blockIdx.x == 0 && lane == 0) tw1 = clock64();
          ptx::mbar_wait(bar_tempty + 8 * acc, aph ^ 1u);
          ptx::tc_fence_after();
          if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[(12 + parity) * 256 + ((w.it >> 1) & 255)] = clock64() - tw1;  // tempty wait
          if (ptx::elect_one()) {
            const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
            const uint32_t a_sub = a_lo0 + (uint32_t)(j * 128) * pix_units;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int kk = 0; kk < KSTEPS; ++kk) {
                const uint32_t a_lo = a_sub + tap_off[tap] + (uint32_t)kk * k_units;
                const uint32_t b_lo = b_lo0 + (uint32_t)(tap * CPT + kk / KPC) * b_chunk16 + (uint32_t)((kk % KPC) * 2);
                ptx::umma_f16_lohi(d_tmem, a_lo, p.a_desc_hi, b_lo, p.b_desc_hi, p.idesc, (tap | kk) ? 1u : 0u);
              }
            }
            ptx::umma_commit(bar_tfull + 8 * acc);
            if (p.trace && blockIdx.x == 0) p.trace[(3 + parity) * 256 + ((w.it >> 1) & 255)] = clock64();
          }
          __syncwarp();
        }
        w.step(p);
      }
      // this warp's MMAs on the band are all issued: its share of "band consumed" (2 arrivals free the buffer)
      if (ptx::elect_one()) ptx::umma_commit(bar_hempty + 8 * buf);
      __syncwarp();
      ++k;
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    // Three 4-warp groups take sub-tiles round-robin; a warp owns 32 accumulator rows (= 32 consecutive padded pixels).
    // All global traffic of the epilogue goes through a warp-private, XOR-swizzled staging tile of 32 pixels x Cout:
    //   * the residual of the warp's NEXT sub-tile is requested with cp.async right after the current one has been
    //     written out (a whole group period ahead of its use), 16 bytes per lane, consecutive lanes -> consecutive
    //     addresses (a per-row-per-thread LDG pattern would touch 16-32 cache lines per instruction and made the LSU,
    //     not the tensor pipe, the limiter of the C = 64 layers);
    //   * each thread adds its row in place (conflict-free row-wise access thanks to the swizzle);
    //   * the finished tile leaves with the same coalesced lane mapping.
    // The epilogue of a sub-tile is ONE warp's dependent instruction stream per 32 pixels, and with N <= 64 it, not the
    // tensor pipe, paces the kernel (the issuers were measured waiting 1-2.5 k cycles per sub-tile for a free
    // accumulator): the residual add is a compile-time variant, the bias sits in shared memory, and every lane
    // computes the image offset of ITS pixel once per sub-tile -- the coalesced phases fetch it with one shuffle.
    constexpr int NG = HALO_EPI_GROUPS;
    constexpr int CH = NV * 2;                        // 16-byte chunks per pixel row of the staging tile
    constexpr int ROWB = NV * 32;                     // bytes per pixel row
    constexpr int MODE = RES ? 1 : 0;
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const TO* res = reinterpret_cast<const TO*>(p.res);
    TO* out = reinterpret_cast<TO*>(p.out);
    const uint32_t stage = smem_halo + 2u * (uint32_t)p.halo_bytes + (uint32_t)(warp - 4) * (32u * ROWB);
    uint8_t* const stage_ptr = smem_raw + (stage - raw);
    const float* const sbias = reinterpret_cast<const float*>(smem_raw + (base + HALO_BIAS_OFF - raw));
    // chunk c of pixel r lives at r * ROWB + ((c ^ swz(r)) * 16); swz keeps both access patterns conflict-free
    // (CH = 6, i.e. Cout = 48, is not a power of two: stored plain, a few bank conflicts on that rare shape)
    auto swz = [](int r) { return CH == 8 ? (r & 7) : CH == 4 ? ((r >> 1) & 3) : CH == 2 ? ((r >> 2) & 1) : 0; };
    auto slot_off = [&](int r, int c) { return (uint32_t)(r * ROWB + ((c ^ swz(r)) * 16)); };
    const float floor_v = p.act == CAPF_ACT_RELU ? 0.f : -__int_as_float(0x7f800000);

    HaloWalk w;
    w.init(p, band0, band1);
    for (int i = 0; i < grp && w.valid(); ++i) w.step(p);
    uint32_t trace_n = 0;
    const uint64_t pol_in = ptx::policy_evict_first(), pol_out = ptx::policy_evict_last();

    // global element offset of this lane's padded pixel in sub-tile t, or -1 when it is padding / past the band
    auto my_pixel_off = [&](const HaloWalk& t) -> int {
      const int mp = t.j * 128 + q * 32 + lane;
      const int iy = div_wp(mp, p.wp_magic), ix = mp - iy * p.Wp;
      const bool live = t.valid() && ix < p.W && iy < t.bh_eff;
      return live ? (((t.img * p.H + t.y0 + iy) * p.W + ix) * p.Cout) : -1;
    };
    // coalesced lane mapping: item = i * 32 + lane -> (pixel row r = item / CH, chunk c = item % CH)
    auto prefetch_residual = [&](int myoff) {
      if (!RES) return;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int item = i * 32 + lane, r = item / CH, c = item % CH;
        const int off = __shfl_sync(0xffffffffu, myoff, r);
        if (off >= 0) {
          if (p.l2_hints) ptx::cp_async16_hint(stage + slot_off(r, c), res + off + c * 8, pol_in);
          else ptx::cp_async16(stage + slot_off(r, c), res + off + c * 8);
        }
      }
      ptx::cp_async_commit();
    };
    int myoff = my_pixel_off(w);
    prefetch_residual(myoff);
    while (w.valid()) {
      const uint32_t acc = w.acc(), aph = w.use_parity();
      for (int i = 0; i < NG && w.valid(); ++i) w.step(p);
      const int nextoff = my_pixel_off(w);           // -1 everywhere once the walk has ended
      const uint32_t taddr = tmem_base + acc * p.acc_stride + ((uint32_t)(q * 32) << 16);
      ptx::mbar_wait(bar_tfull + 8 * acc, aph);
      ptx::tc_fence_after();
      if (p.trace && blockIdx.x == 0 && q == 0 && lane == 0) p.trace[(5 + grp) * 256 + (trace_n & 255)] = clock64();
      if (RES) {
        ptx::cp_async_wait_all();                    // this tile's residual (requested one group period ago) has landed
        __syncwarp();
      }
#pragma unroll
      for (int v = 0; v < NV; v += 2) {
        const bool two = v + 1 < NV;
        uint32_t a0[16], a1[16];
        ptx::tmem_ld16(taddr + (uint32_t)(16 * v), a0);
        if (two) ptx::tmem_ld16(taddr + (uint32_t)(16 * v + 16), a1);
        ptx::tmem_ld_wait();
        epi16<TO, MODE>(a0, sbias + 16 * v, floor_v, stage_ptr + lane * ROWB, (uint32_t)(2 * v), (uint32_t)swz(lane));
        if (two) epi16<TO, MODE>(a1, sbias + 16 * v + 16, floor_v, stage_ptr + lane * ROWB, (uint32_t)(2 * v + 2), (uint32_t)swz(lane));
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_tempty + 8 * acc);          // accumulator drained: the issuer may reuse the stage
      __syncwarp();
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int item = i * 32 + lane, r = item / CH, c = item % CH;
        const int off = __shfl_sync(0xffffffffu, myoff, r);
        if (off >= 0) {
          const uint4 val = *reinterpret_cast<const uint4*>(stage_ptr + slot_off(r, c));
          if (p.l2_hints) ptx::st_global_v4_hint(out + off + c * 8, val, pol_out);
          else *reinterpret_cast<uint4*>(out + off + c * 8) = val;
        }
      }
      __syncwarp();
      myoff = nextoff;
      prefetch_residual(myoff);
      if (p.trace && blockIdx.x == 0 && q == 0 && lane == 0) p.trace[(8 + grp) * 256 + (trace_n & 255)] = clock64();
      ++trace_n;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// =======================================================================================================
// host side
// =======================================================================================================
struct TcHaloState {
  CUtensorMap mapA, mapB;
  HaloP p;
  int grid, smem_bytes, dtype_in, dtype_out;
};

// pixels of one 8-channel plane: the TMA box ((bh + 3) rows of Wp) and the furthest tap read of the last sub-tile
// Halo rows per TMA box.  Measured with tools/microbench/tma_bench.cu on B200: a box costs ~350 cycles of TMA-engine
// time plus ~1 cycle per 64 bytes, whatever its shape (64- or 128-byte rows, with or without zero-filled columns), so
// a band is fetched as ONE box (79 KB in 1.9k cycles = 42 B/clk/SM) unless it exceeds the 256-row box limit.
static int halo_box_rows(int bh) { return bh + 3 <= 256 ? bh + 3 : 128; }
static int halo_plane_pixels(int bh, int Wp, int box_rows) {
  const int n_sub = (bh * Wp + 127) / 128;
  const int n_boxes = (bh + 3 + box_rows - 1) / box_rows;
  const int reach = n_sub * 128 + 2 * Wp + 2, box = n_boxes * box_rows * Wp;
  return ((reach > box ? reach : box) + 7) & ~7;
}

static int halo_plan(const capf_op& op, HaloP& p, int& smem_bytes) {
  const int N = op.i[0], H = op.i[1], W = op.i[2], C = op.i[3], Cout = op.i[4];
  if (op.i[5] != 3 || op.i[6] != 3 || op.i[7] != 1 || op.i[8] != 1) return 0;
  if (op.dtype_out != op.dtype_in) return 0;      // 16-bit activations in and out (the backbone case)
  if ((C != 16 && C != 32 && C != 48 && C != 64) || Cout % 16 || Cout > 64) return 0;   // instantiated widths; <= 4 residual vectors
  if (W + 1 > 256 || N <= 0 || H <= 0 || W <= 0) return 0;
  if (op.i[11] == CAPF_ACT_GELU) return 0;                         // backbone convs only: identity / ReLU epilogues
  if ((long long)N * H * W * Cout >= (1ll << 31)) return 0;        // 32-bit element offsets in the epilogue
  memset(&p, 0, sizeof(p));
  p.C = C; p.Cout = Cout; p.H = H; p.W = W; p.Nimg = N; p.Wp = W + 1;
  p.wp_magic = (uint32_t)(((1ull << 32) + p.Wp - 1) / p.Wp);
  p.chunks = C / 8;
  p.kb = C % 64 == 0 ? 64 : C % 32 == 0 ? 32 : 16;
  p.cpt = C / p.kb;
  p.b_chunk_bytes = Cout * p.kb * 2;
  p.b_bytes = 9 * p.cpt * p.b_chunk_bytes;
  // un-swizzled 8-channel planes work for every C; swizzled whole-pixel rows need a power-of-two row (and move 4-8x
  // fewer TMA elements).  UMMA applies the swizzle XOR to the absolute shared-memory address, so the shifted-window
  // starts need no descriptor base offset (measured on B200: base_offset = 0 is exact, (start >> 7) & 7 is wrong).
  p.a_rows = op.i[13] != 2 && (C == 16 || C == 32 || C == 64);
  const int b_region = (p.b_bytes + 1023) & ~1023;
  const int stage_bytes = 4 * HALO_EPI_GROUPS * 32 * Cout * 2;   // warp-private epilogue staging tiles
  const int budget = TC_SMEM_LIMIT - 1024 - HALO_HEADER_BYTES - b_region - stage_bytes;
  // band height: fewest 128-row sub-tiles per image, then the tallest band (fewest halo re-reads)
  int best_bh = 0;
  long long best_tiles = 1ll << 60;
  for (int bh = 1; bh <= H; ++bh) {
    const int n_sub_full = (bh * p.Wp + 127) / 128;
    const int P_alloc = halo_plane_pixels(bh, p.Wp, halo_box_rows(bh));
    const long long halo_bytes = ((long long)P_alloc * C * 2 + 1023) & ~1023ll;
    if (P_alloc > 16383 || 2 * halo_bytes > budget) break;
    const int full = H / bh, rem = H - full * bh;
    long long tiles = (long long)full * n_sub_full + (rem ? (rem * p.Wp + 127) / 128 : 0);
    if (tiles < best_tiles || (tiles == best_tiles && bh > best_bh)) { best_tiles = tiles; best_bh = bh; }
  }
  if (!best_bh) return 0;
  p.bh = best_bh;
  p.bands_per_img = (H + p.bh - 1) / p.bh;
  const long long nb = (long long)N * p.bands_per_img;
  if (nb >= (1ll << 31)) return 0;
  p.num_bands = (int)nb;
  const int n_sub_full = (p.bh * p.Wp + 127) / 128;
  (void)n_sub_full;
  p.box_rows = halo_box_rows(p.bh);
  p.P_alloc = halo_plane_pixels(p.bh, p.Wp, p.box_rows);
  p.halo_bytes = (p.P_alloc * C * 2 + 1023) & ~1023;
  p.n_boxes = (p.bh + 3 + p.box_rows - 1) / p.box_rows;
  p.plane_tx_bytes = p.n_boxes * p.box_rows * p.Wp * 16;
  p.acc_stages = HALO_MAX_ACC;
  int cols = 32;
  while (cols < p.acc_stages * Cout) cols <<= 1;
  p.tmem_cols = cols;
  p.acc_stride = Cout;
  smem_bytes = 1024 + HALO_HEADER_BYTES + b_region + 2 * p.halo_bytes + stage_bytes;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;   // one CTA per SM
  return 1;
}

int tc_halo_supported(const capf_op& op) {
  HaloP p;
  int smem;
  return halo_plan(op, p, smem);
}

int tc_halo_prepare(const capf_op& op, TcHaloState** out) {
  *out = nullptr;
  int e = tc_get_encoder();
  if (e) return e;
  TcHaloState* s = new (std::nothrow) TcHaloState();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_halo_prepare: out of host memory");
  if (!halo_plan(op, s->p, s->smem_bytes)) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "halo conv: shape not supported"); }
  HaloP& p = s->p;
  const bool bf16 = op.dtype_in == CAPF_BF16;
  p.idesc = tc_idesc(bf16, p.Cout);
  p.b_desc_hi = tc_desc_hi(p.kb * 2, 8 * p.kb * 2);
  if (p.a_rows) {
    p.a_desc_hi = tc_desc_hi(p.C * 2, 8 * p.C * 2);   // swizzle span == pixel row; 8-row groups contiguous
  } else {
    p.a_desc_hi = tc_desc_hi(0, 128);                 // un-swizzled: 8-row groups are 128 contiguous bytes
  }
  p.act = op.i[11];
  p.x = op.in[0];
  p.bias = (const float*)op.in[2];
  p.res = op.in[3];
  p.out = op.out[0];
  { const char* e = getenv("CAPF_L2_HINTS"); p.l2_hints = (e && e[0] == '0') ? 0 : 1; }
  p.trace = (long long*)op.in[4];     // debug only (NULL in every program the host layer builds)
  s->grid = p.num_bands < num_sms() ? p.num_bands : num_sms();
  s->dtype_in = op.dtype_in;
  s->dtype_out = op.dtype_out;
  const int K = 9 * p.C;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)p.Cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)p.kb, (cuuint32_t)p.Cout};
  cuuint32_t es[2] = {1, 1};
  e = tc_encode_map(&s->mapB, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, op.in[1], dims, strides, box,
                    es, p.kb * 2, "B weights (halo)");
  if (!e) {
    // 8-channel planes of the NHWC input: box {8, Wp, bh + 3, 1}, no swizzle (16-byte rows = UMMA core-matrix rows)
    cuuint64_t adims[4] = {(cuuint64_t)p.C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.Nimg};
    cuuint64_t astr[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.W * p.C * 2, (cuuint64_t)p.H * p.W * p.C * 2};
    cuuint32_t abox[4] = {(cuuint32_t)(p.a_rows ? p.C : 8), (cuuint32_t)p.Wp, (cuuint32_t)p.box_rows, 1};
    cuuint32_t aes[4] = {1, 1, 1, 1};
    e = tc_encode_map(&s->mapA, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, op.in[0], adims, astr, abox,
                      aes, p.a_rows ? p.C * 2 : 0, "A halo");
  }
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

template <int C, int NV, typename TI, typename TO, bool RES>
static int halo_launch_cnr(const TcHaloState* s, cudaStream_t st) {
  static PerDevice<bool> opted_;
  std::atomic<bool>& opted = opted_.get();
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv3_halo_kernel<C, NV, TI, TO, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_conv3_halo_kernel smem opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  launch_k(tc_conv3_halo_kernel<C, NV, TI, TO, RES>, dim3(s->grid), dim3(HALO_THREADS), s->smem_bytes, st, s->mapA, s->mapB, s->p);
  return check_launch("tc_conv3_halo_kernel");
}

template <int C, int NV, typename TI, typename TO>
static int halo_launch_cn(const TcHaloState* s, cudaStream_t st) {
  return s->p.res ? halo_launch_cnr<C, NV, TI, TO, true>(s, st) : halo_launch_cnr<C, NV, TI, TO, false>(s, st);
}

template <int C, typename TI, typename TO>
static int halo_launch_c(const TcHaloState* s, cudaStream_t st) {
  switch (s->p.Cout >> 4) {
    case 1: return halo_launch_cn<C, 1, TI, TO>(s, st);
    case 2: return halo_launch_cn<C, 2, TI, TO>(s, st);
    case 3: return halo_launch_cn<C, 3, TI, TO>(s, st);
    case 4: return halo_launch_cn<C, 4, TI, TO>(s, st);
    default: return set_error(CAPF_ERR_UNSUPPORTED, "halo conv: Cout not instantiated");
  }
}

template <typename TI, typename TO>
static int halo_launch_typed(const TcHaloState* s, cudaStream_t st) {
  switch (s->p.C) {
    case 16: return halo_launch_c<16, TI, TO>(s, st);
    case 32: return halo_launch_c<32, TI, TO>(s, st);
    case 48: return halo_launch_c<48, TI, TO>(s, st);
    case 64: return halo_launch_c<64, TI, TO>(s, st);
    default: return set_error(CAPF_ERR_UNSUPPORTED, "halo conv: channel count not instantiated");
  }
}

int tc_halo_launch(const TcHaloState* s, cudaStream_t st) {
  if (s->dtype_in == CAPF_F16 && s->dtype_out == CAPF_F16) return halo_launch_typed<__half, __half>(s, st);
  if (s->dtype_in == CAPF_BF16 && s->dtype_out == CAPF_BF16) return halo_launch_typed<__nv_bfloat16, __nv_bfloat16>(s, st);
  return set_error(CAPF_ERR_UNSUPPORTED, "halo conv: dtype combination");
}

void tc_halo_release(TcHaloState* s) { delete s; }

void tc_halo_describe(const TcHaloState* s, char* buf, int cap) {
  snprintf(buf, cap, "tc_conv3_halo_kernel<C=%d,Cout=%d>[band %d rows]", s->p.C, s->p.Cout, s->p.bh);
}

}  // namespace capf
