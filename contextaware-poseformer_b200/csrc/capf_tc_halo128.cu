// tcgen05 3x3 / stride-1 / pad-1 convolution for C = Cout = 128: shared-memory halo band + STREAMED weights.
//
// The 128-channel BasicBlocks of HRNet (pose_hrnet.py:66-95; branch 2: 56 launches per forward at 16 x 16) ran on the
// per-tap TMA kernel (capf_tc.cu), which fetches every input pixel nine times through L2: 885 KB per 256-pixel tile
// against 9.2 k tensor clocks of work -- bound by the chip-wide L2 -> SM feed (~42 B/clk/SM) at 0.44 of the tensor peak.
// The halo-band kernel (capf_tc_halo.cu) removes the nine-fold re-read but keeps all weights resident, which stops at
// C = 64 (9 * 128 * 128 * 2 B = 295 KB do not fit).  Here:
//
//   A   a band of (bh + 3) input rows, loaded ONCE per band as two 64-channel planes of 128-byte-swizzled pixel rows
//       (pitch Wp = W + 1: one zero column shared by the left / right padding); the A operand of tap (r, s), K step kk
//       is the same buffer read through a descriptor whose start moves by (r * Wp + s) pixels -- exactly the shifted
//       window of capf_tc_halo.cu -- in plane kk / 4;
//   B   the folded weights [128][9 * 128] stream through a ring of 16 KB (tap, plane) chunks; every chunk feeds ALL
//       128-row sub-tiles of the band (up to 4, one 128-column TMEM accumulator each) before its slot is released,
//       so a band costs one pass over the 295 KB of weights: 378 KB per 16 x 16 image instead of 885 KB;
//   D   n_sub accumulators complete together; three 4-warp epilogue groups drain one sub-tile each, 64 columns at a
//       time through warp-private swizzled staging tiles (residual by cp.async, 16-byte coalesced stores).
//
// Roles (512 threads): warp 0 = weight-ring producer (independent of the predecessor kernel: starts before the PDL
// wait), warp 2 = TMEM allocator, then band producer, warps 1 and 3 = tcgen05.mma issuers (even / odd sub-tiles: one
// thread needs ~80 clk to set up an MMA that retires in 64, measured with tools/halo128_trace.py), warps 4..15 = epilogue.
#include <cstdio>
#include <cstdlib>
#include <new>

#include "capf_tc.cuh"

namespace capf {

constexpr int H128_THREADS = 512;
constexpr int H128_HEADER = 2048;            // barriers (first 512 B) + 128 fp32 bias values at +1024
constexpr int H128_BIAS_OFF = 1024;
constexpr int H128_C = 128;                  // input = output channels
constexpr int H128_CHUNK_BYTES = 128 * 128;  // one weight chunk: 128 output channels x 64 input channels x 2 B
constexpr int H128_NCHUNK = 18;              // 9 taps x 2 planes
constexpr int H128_MAX_B = 6;                // ring stages (upper bound)
constexpr int H128_EPI_WARPS = 12;
constexpr int H128_STG_BYTES = 32 * 128;     // staging tile of one warp: 32 pixels x 64 channels x 2 B

struct H128P {
  int H, W, Nimg, Wp;
  uint32_t wp_magic;
  int bh, bands_per_img, num_bands;
  int P_alloc;              // pixels per plane of the band buffer
  int plane_bytes;          // P_alloc * 128
  int box_rows, n_boxes;
  int a_tx_bytes;           // bytes the TMA boxes of one band deliver (both planes)
  int nb;                   // weight ring stages
  int n_sub_max, tmem_cols;
  uint32_t idesc, desc_hi;
  int act;
  const float* bias;
  const void* res;
  void* out;
  long long* trace;         // optional (debug, op.in[4]): clock64 timeline of CTA 0, tools/halo128_trace.py
};

__device__ __forceinline__ int h128_div(int v, uint32_t magic) { return (int)__umulhi((uint32_t)v, magic); }

template <typename TO, bool RES>
__global__ void __launch_bounds__(H128_THREADS, 1)
tc_conv3_halo128_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const H128P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_bfull = base;                          // [H128_MAX_B]
  const uint32_t bar_bempty = base + 8 * H128_MAX_B;        // [H128_MAX_B]
  const uint32_t bar_hfull = base + 16 * H128_MAX_B;        // band landed
  const uint32_t bar_hempty = bar_hfull + 8;                // band consumed (all MMAs of the band complete)
  const uint32_t bar_tfull = bar_hfull + 16;                // accumulators complete
  const uint32_t bar_tempty = bar_hfull + 24;               // accumulators drained (384 arrivals)
  const uint32_t tmem_slot = bar_hfull + 32;
  const uint32_t smem_a = base + H128_HEADER;               // 2 planes
  const uint32_t smem_b = smem_a + 2u * (uint32_t)p.plane_bytes;
  const uint32_t smem_stg = smem_b + (uint32_t)(p.nb * H128_CHUNK_BYTES);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapA);
    ptx::prefetch_tmap(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.nb; ++s) {
      ptx::mbar_init(bar_bfull + 8 * s, 1);
      ptx::mbar_init(bar_bempty + 8 * s, 2);        // both issuer warps commit
    }
    ptx::mbar_init(bar_hfull, 1);
    ptx::mbar_init(bar_hempty, 2);
    ptx::mbar_init(bar_tfull, 2);
    ptx::mbar_init(bar_tempty, H128_EPI_WARPS * 32);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp == 3) {                  // folded-BN shift: constant data
    float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(smem_raw + (base + H128_BIAS_OFF - raw) + 16 * lane) = b4;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const bool tr = p.trace != nullptr && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) p.trace[0] = clock64();
  pdl_trigger();
  if (warp != 0) pdl_wait();        // warp 0 only ever touches the (constant) weights

  const int band0 = (int)(((long long)p.num_bands * blockIdx.x) / gridDim.x);
  const int band1 = (int)(((long long)p.num_bands * (blockIdx.x + 1)) / gridDim.x);
  const int nbands = band1 - band0;

  if (warp == 0) {
    // ===================================== weight ring producer ==============================
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int b = 0; b < nbands; ++b) {
        for (int c = 0; c < H128_NCHUNK; ++c) {
          ptx::mbar_wait(bar_bempty + 8 * stage, phase ^ 1u);
          const uint32_t full = bar_bfull + 8 * stage;
          ptx::mbar_arrive_expect_tx(full, (uint32_t)H128_CHUNK_BYTES);
          ptx::tma_load_2d(&mapB, full, smem_b + stage * H128_CHUNK_BYTES, c * 64, 0);
          if (++stage == (uint32_t)p.nb) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================== band producer ====================================
    if (ptx::elect_one()) {
      const uint64_t pol_in = ptx::policy_evict_first();
      int img = band0 / p.bands_per_img, bin = band0 - img * p.bands_per_img;
      for (int b = 0; b < nbands; ++b) {
        ptx::mbar_wait(bar_hempty, (uint32_t)(b & 1) ^ 1u);
        ptx::mbar_arrive_expect_tx(bar_hfull, (uint32_t)p.a_tx_bytes);
        const int y_top = bin * p.bh - 1;
        for (int bx = 0; bx < p.n_boxes; ++bx) {
          const uint32_t slab = (uint32_t)(bx * p.box_rows * p.Wp) * 128u;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
            ptx::tma_load_4d_hint(&mapA, bar_hfull, smem_a + (uint32_t)pl * (uint32_t)p.plane_bytes + slab, 64 * pl, -1, y_top + bx * p.box_rows, img, pol_in);
        }
        if (++bin == p.bands_per_img) { bin = 0; ++img; }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================================== MMA issuers (even / odd sub-tiles) ================
    const int parity = warp == 1 ? 0 : 1;
    uint32_t tap_off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) tap_off[tap] = (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 8u;      // 16-byte units (128 B per pixel)
    const uint32_t a_lo0 = tc_desc_lo(smem_a, 1u), plane16 = (uint32_t)p.plane_bytes >> 4;
    const uint32_t b_lo0 = tc_desc_lo(smem_b, 1u);
    uint32_t stage = 0, phase = 0;
    int bin = band0 % p.bands_per_img;
    for (int b = 0; b < nbands; ++b) {
      const int bh_eff = min(p.bh, p.H - bin * p.bh);
      const int n_sub = (bh_eff * p.Wp + 127) >> 7;
      ptx::mbar_wait(bar_tempty, (uint32_t)(b & 1) ^ 1u);       // previous band's accumulators drained
      if (tr && lane == 0 && b < 8 && parity == 0) p.trace[16 + 4 * b] = clock64();
      ptx::mbar_wait(bar_hfull, (uint32_t)(b & 1));
      ptx::tc_fence_after();
      if (tr && lane == 0 && b < 8 && parity == 0) p.trace[17 + 4 * b] = clock64();
      for (int c = 0; c < H128_NCHUNK; ++c) {
        ptx::mbar_wait(bar_bfull + 8 * stage, phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t a_c = a_lo0 + (uint32_t)(c & 1) * plane16 + tap_off[c >> 1];
          const uint32_t b_c = b_lo0 + stage * (uint32_t)(H128_CHUNK_BYTES >> 4);
          for (int j = parity; j < n_sub; j += 2) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(j * H128_C);
            const uint32_t a_j = a_c + (uint32_t)(j * 128) * 8u;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16_lohi(d_tmem, a_j + 2u * kk, p.desc_hi, b_c + 2u * kk, p.desc_hi, p.idesc, (c | kk) ? 1u : 0u);
          }
          ptx::umma_commit(bar_bempty + 8 * stage);
          if (c == H128_NCHUNK - 1) {
            ptx::umma_commit(bar_tfull);
            ptx::umma_commit(bar_hempty);
          }
          if (tr && b < 8 && parity == 0) { if (c == 0) p.trace[18 + 4 * b] = clock64(); if (c == H128_NCHUNK - 1) p.trace[19 + 4 * b] = clock64(); }
          if (tr && b == 0 && c < 18) p.trace[96 + 18 * parity + c] = clock64();
        }
        __syncwarp();
        if (++stage == (uint32_t)p.nb) { stage = 0; phase ^= 1u; }
      }
      if (++bin == p.bands_per_img) bin = 0;
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    constexpr int MODE = RES ? 1 : 0;
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const TO* res = reinterpret_cast<const TO*>(p.res);
    TO* out = reinterpret_cast<TO*>(p.out);
    const uint32_t stg = smem_stg + (uint32_t)(warp - 4) * H128_STG_BYTES;
    uint8_t* const stg_ptr = smem_raw + (stg - raw);
    const float* const sbias = reinterpret_cast<const float*>(smem_raw + (base + H128_BIAS_OFF - raw));
    auto slot_off = [](int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); };
    const float floor_v = p.act == CAPF_ACT_RELU ? 0.f : -__int_as_float(0x7f800000);
    const uint64_t pol_in = ptx::policy_evict_first(), pol_out = ptx::policy_evict_last();
    int img = band0 / p.bands_per_img, bin = band0 - img * p.bands_per_img;
    for (int b = 0; b < nbands; ++b) {
      const int y0 = bin * p.bh, bh_eff = min(p.bh, p.H - y0);
      const int n_sub = (bh_eff * p.Wp + 127) >> 7;
      bool waited = false;
      for (int j = grp; j < n_sub; j += 3) {
        // element offset of this lane's padded pixel, -1 for the padding column / rows past the band
        const int mp = j * 128 + q * 32 + lane;
        const int iy = h128_div(mp, p.wp_magic), ix = mp - iy * p.Wp;
        const int myoff = (ix < p.W && iy < bh_eff) ? (((img * p.H + y0 + iy) * p.W + ix) * H128_C) : -1;
        const uint32_t taddr = tmem_base + (uint32_t)(j * H128_C) + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (RES) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int item = i * 32 + lane, r = item >> 3, c = item & 7;
              const int off = __shfl_sync(0xffffffffu, myoff, r);
              if (off >= 0) ptx::cp_async16_hint(stg + slot_off(r, c), res + off + half * 64 + c * 8, pol_in);
            }
            ptx::cp_async_commit();
          }
          if (!waited) {
            ptx::mbar_wait(bar_tfull, (uint32_t)(b & 1));
            ptx::tc_fence_after();
            waited = true;
            if (tr && warp == 4 && lane == 0 && b < 8) p.trace[48 + 4 * b] = clock64();
          }
          if (RES) {
            ptx::cp_async_wait_all();
            __syncwarp();
          }
          if (tr && warp == 4 && lane == 0 && b == 1) p.trace[80 + 4 * half] = clock64();
#pragma unroll
          for (int v = 0; v < 4; v += 2) {
            uint32_t a0[16], a1[16];
            ptx::tmem_ld16(taddr + (uint32_t)(half * 64 + 16 * v), a0);
            ptx::tmem_ld16(taddr + (uint32_t)(half * 64 + 16 * v + 16), a1);
            ptx::tmem_ld_wait();
            epi16<TO, MODE>(a0, sbias + half * 64 + 16 * v, floor_v, stg_ptr + lane * 128, (uint32_t)(2 * v), (uint32_t)lane & 7u);
            epi16<TO, MODE>(a1, sbias + half * 64 + 16 * v + 16, floor_v, stg_ptr + lane * 128, (uint32_t)(2 * v + 2), (uint32_t)lane & 7u);
          }
          __syncwarp();
          if (tr && warp == 4 && lane == 0 && b == 1) p.trace[81 + 4 * half] = clock64();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int item = i * 32 + lane, r = item >> 3, c = item & 7;
            const int off = __shfl_sync(0xffffffffu, myoff, r);
            if (off >= 0) ptx::st_global_v4_hint(out + off + half * 64 + c * 8, *reinterpret_cast<const uint4*>(stg_ptr + slot_off(r, c)), pol_out);
          }
          __syncwarp();
          if (tr && warp == 4 && lane == 0 && b == 1) p.trace[82 + 4 * half] = clock64();
        }
      }
      if (!waited) {                                   // a group without a sub-tile in this band still takes part in the hand-over
        ptx::mbar_wait(bar_tfull, (uint32_t)(b & 1));
        ptx::tc_fence_after();
      }
      if (tr && warp == 4 && lane == 0 && b < 8) p.trace[49 + 4 * b] = clock64();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_tempty);                    // 384 arrivals: every accumulator of the band has been read
      if (++bin == p.bands_per_img) { bin = 0; ++img; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (tr && threadIdx.x == 0) { p.trace[1] = clock64(); p.trace[2] = nbands; }
}

// =======================================================================================================
// host side
// =======================================================================================================
struct TcHalo128State {
  CUtensorMap mapA, mapB;
  H128P p;
  int grid, smem_bytes, dtype;
};

static int h128_box_rows(int bh) { return bh + 3 <= 256 ? bh + 3 : 128; }
static int h128_plane_pixels(int bh, int Wp, int box_rows) {
  const int n_sub = (bh * Wp + 127) / 128;
  const int n_boxes = (bh + 3 + box_rows - 1) / box_rows;
  const int reach = n_sub * 128 + 2 * Wp + 2, box = n_boxes * box_rows * Wp;
  return ((reach > box ? reach : box) + 7) & ~7;
}

static int h128_plan(const capf_op& op, H128P& p, int& smem_bytes) {
  const char* ev = getenv("CAPF_HALO128");
  if (ev && ev[0] == '0') return 0;
  const int N = op.i[0], H = op.i[1], W = op.i[2], C = op.i[3], Cout = op.i[4];
  if (op.kind != CAPF_OP_CONV2D || op.i[5] != 3 || op.i[6] != 3 || op.i[7] != 1 || op.i[8] != 1) return 0;
  if (C != H128_C || Cout != H128_C) return 0;
  if (op.dtype_out != op.dtype_in || (op.dtype_in != CAPF_F16 && op.dtype_in != CAPF_BF16)) return 0;
  if (op.i[18] != 0 || op.i[11] == CAPF_ACT_GELU) return 0;
  if (W + 1 > 256 || N <= 0 || H <= 0 || W <= 0) return 0;
  if ((long long)N * H * W * Cout >= (1ll << 31)) return 0;
  memset(&p, 0, sizeof(p));
  p.H = H; p.W = W; p.Nimg = N; p.Wp = W + 1;
  p.wp_magic = (uint32_t)(((1ull << 32) + p.Wp - 1) / p.Wp);
  const int fixed = 1024 + H128_HEADER + H128_EPI_WARPS * H128_STG_BYTES;
  // band height: at most 4 sub-tiles (4 x 128 TMEM columns), at least 3 weight-ring stages; fewest sub-tiles per image
  // (padding rows of the last sub-tile of a band are wasted tensor work), then the tallest band
  int best_bh = 0;
  double best_cost = 1e300;
  for (int bh = 1; bh <= H; ++bh) {
    const int n_sub = (bh * p.Wp + 127) / 128;
    if (n_sub > 4) break;
    const int P_alloc = h128_plane_pixels(bh, p.Wp, h128_box_rows(bh));
    if (P_alloc > 16383) break;
    const long long a_bytes = 2ll * (((long long)P_alloc * 128 + 1023) & ~1023ll);
    if (fixed + a_bytes + 3 * H128_CHUNK_BYTES > TC_SMEM_LIMIT) break;
    const int full = H / bh, rem = H - full * bh;
    // cost per image: MMA sub-tiles + one weight pass (~ 0.6 sub-tile of feed time) per band
    const double cost = (double)full * (n_sub + 0.6) + (rem ? ((rem * p.Wp + 127) / 128 + 0.6) : 0.0);
    if (cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && bh > best_bh)) { best_cost = cost; best_bh = bh; }
  }
  if (!best_bh) return 0;
  p.bh = best_bh;
  p.bands_per_img = (H + p.bh - 1) / p.bh;
  const long long nbands = (long long)N * p.bands_per_img;
  if (nbands >= (1ll << 31)) return 0;
  p.num_bands = (int)nbands;
  p.box_rows = h128_box_rows(p.bh);
  p.n_boxes = (p.bh + 3 + p.box_rows - 1) / p.box_rows;
  p.P_alloc = h128_plane_pixels(p.bh, p.Wp, p.box_rows);
  p.plane_bytes = (p.P_alloc * 128 + 1023) & ~1023;
  p.a_tx_bytes = 2 * p.n_boxes * p.box_rows * p.Wp * 128;
  int nb = (TC_SMEM_LIMIT - fixed - 2 * p.plane_bytes) / H128_CHUNK_BYTES;
  if (nb > H128_MAX_B) nb = H128_MAX_B;
  if (nb < 3) return 0;
  p.nb = nb;
  p.n_sub_max = (p.bh * p.Wp + 127) / 128;
  int cols = 128;
  while (cols < p.n_sub_max * H128_C) cols <<= 1;
  p.tmem_cols = cols;
  smem_bytes = fixed + 2 * p.plane_bytes + nb * H128_CHUNK_BYTES;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;     // one CTA per SM
  return 1;
}

int tc_halo128_supported(const capf_op& op) {
  H128P p;
  int smem;
  return h128_plan(op, p, smem);
}

int tc_halo128_prepare(const capf_op& op, TcHalo128State** out) {
  *out = nullptr;
  int e = tc_get_encoder();
  if (e) return e;
  TcHalo128State* s = new (std::nothrow) TcHalo128State();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_halo128_prepare: out of host memory");
  if (!h128_plan(op, s->p, s->smem_bytes)) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "halo128 conv: shape not supported"); }
  H128P& p = s->p;
  const bool bf16 = op.dtype_in == CAPF_BF16;
  p.idesc = tc_idesc(bf16, H128_C);
  p.desc_hi = tc_desc_hi(128, 1024);
  p.act = op.i[11];
  p.bias = (const float*)op.in[2];
  p.res = op.in[3];
  p.out = op.out[0];
  p.trace = (long long*)op.in[4];       // debug only (NULL in every program the host layer builds)
  s->grid = p.num_bands < num_sms() ? p.num_bands : num_sms();
  s->dtype = op.dtype_in;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  {
    const int K = 9 * H128_C;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)H128_C};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)H128_C};
    cuuint32_t es[2] = {1, 1};
    e = tc_encode_map(&s->mapB, dt, 2, op.in[1], dims, strides, box, es, 128, "B weights (halo128)");
  }
  if (!e) {
    cuuint64_t adims[4] = {(cuuint64_t)H128_C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.Nimg};
    cuuint64_t astr[3] = {(cuuint64_t)H128_C * 2, (cuuint64_t)p.W * H128_C * 2, (cuuint64_t)p.H * p.W * H128_C * 2};
    cuuint32_t abox[4] = {64, (cuuint32_t)p.Wp, (cuuint32_t)p.box_rows, 1};
    cuuint32_t aes[4] = {1, 1, 1, 1};
    e = tc_encode_map(&s->mapA, dt, 4, op.in[0], adims, astr, abox, aes, 128, "A halo128");
  }
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

template <typename T, bool RES>
static int h128_launch_r(const TcHalo128State* s, cudaStream_t st) {
  static PerDevice<bool> opted_;
  std::atomic<bool>& opted = opted_.get();
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv3_halo128_kernel<T, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_conv3_halo128_kernel smem opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  launch_k(tc_conv3_halo128_kernel<T, RES>, dim3(s->grid), dim3(H128_THREADS), s->smem_bytes, st, s->mapA, s->mapB, s->p);
  return check_launch("tc_conv3_halo128_kernel");
}

int tc_halo128_launch(const TcHalo128State* s, cudaStream_t st) {
  if (s->dtype == CAPF_F16) return s->p.res ? h128_launch_r<__half, true>(s, st) : h128_launch_r<__half, false>(s, st);
  return s->p.res ? h128_launch_r<__nv_bfloat16, true>(s, st) : h128_launch_r<__nv_bfloat16, false>(s, st);
}

void tc_halo128_release(TcHalo128State* s) { delete s; }

void tc_halo128_describe(const TcHalo128State* s, char* buf, int cap) {
  snprintf(buf, cap, "tc_conv3_halo128_kernel[band %d rows, %d sub-tiles, %d weight stages]", s->p.bh, s->p.n_sub_max, s->p.nb);
}

}  // namespace capf
