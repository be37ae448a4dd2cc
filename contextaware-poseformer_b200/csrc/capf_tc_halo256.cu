// tcgen05 3x3 / stride-1 / pad-1 convolution for C = 256 -> Cout = 32: shared-memory halo TILE, plane by plane, STREAMED weights.
//
// HRNet's transition1.0 (pose_hrnet.py:372-411: Conv2d(256, 32, 3, 1, 1) + BN + ReLU on the 64 x 64 output of layer1) is the
// single most expensive convolution of the forward: on the per-tap TMA kernel (capf_tc.cu) every input pixel -- 512 bytes -- is
// fetched nine times through L2 (4.8 GB per launch at bs = 256; the chip-wide L2 -> SM feed makes that 460 us) for 155 GFLOP of
// N = 32 tensor work.  The halo kernels keep a band in shared memory instead, but a 64-pixel-wide band of 256 channels holds
// two rows at most.  Here:
//
//   A   a TILE of (bh + 2) x (tw + 2) input pixels (tw = a fraction of the row: explicit halo columns on both sides, zero-filled
//       by TMA outside the image) as FOUR 64-channel planes of 128-byte-swizzled pixel rows (pitch Wp = tw + 2); tap (r, s),
//       K step kk of plane pl is the plane's buffer read through a descriptor whose start moves by (r * Wp + s) pixels
//       (the shifted window of capf_tc_halo.cu);
//   K   the loop runs PLANE-major (plane, filter row, tap, k) and the planes of consecutive tiles rotate through THREE plane
//       buffers, each with its own full / empty barrier pair: the next planes stream in while the current one is being
//       multiplied, and the fourth buffer's 42 KB go to the weight ring instead;
//   B   the folded weights [32][9 * 256] stream through a ring of 12 KB chunks (the three taps of one filter row of one plane;
//       7 stages = 84 KB in flight); every chunk feeds all 128-row sub-tiles of the tile: 12 MMAs per issuer and barrier round trip;
//   D   two accumulator stages (n_sub x 32 TMEM columns each): the epilogue of tile b overlaps the MMAs of tile b + 1.
//
// Measured with the wait-cycle counters of CTA 0 (ONE_TRACE=1 tools/one_conv.py): with 4 KB (plane, tap) chunks and an 8-deep ring
// the even issuer spent 62 % of the kernel blocked on MMA issue (82 clk per MMA = the N = 32 operand-port time of its own and the
// other issuer's MMA), 13 % waiting for weights and 21 % in per-chunk overhead -- hence the larger chunks and the deeper ring.
//
// L2 -> SM bytes per launch: 1.5 x the input (tile halo) + one pass over the 147 KB of weights per tile = 1.6 GB instead of
// 4.8 GB; what remains is the shared-memory operand port of N = 32 MMAs (4 KB of A + 1 KB of B per 16 tensor clocks of work).
//
// Roles (384 threads): warp 0 = weight-ring producer (constant data: starts before the PDL wait), warp 2 = TMEM allocator, then
// tile producer, warps 1 and 3 = tcgen05.mma issuers (even / odd sub-tiles), warps 4..11 = two 4-warp epilogue groups.
#include <cstdio>
#include <cstdlib>
#include <new>

#include "capf_tc.cuh"

namespace capf {

constexpr int H256_THREADS = 384;
constexpr int H256_HEADER = 2048;            // barriers (first 512 B) + 32 fp32 bias values at +1024
constexpr int H256_BIAS_OFF = 1024;
constexpr int H256_C = 256, H256_NPL = 4;    // input channels = 4 planes of 64
constexpr int H256_COUT = 32;
constexpr int H256_NPB = 3;                  // plane BUFFERS: the 4 planes of consecutive tiles rotate through 3 buffers
constexpr int H256_TAP_BYTES = H256_COUT * 128;     // weights of one (plane, tap): 32 output channels x 64 input channels x 2 B
constexpr int H256_CHUNK_BYTES = 3 * H256_TAP_BYTES;   // one weight chunk = the three taps of a filter row of one plane
constexpr int H256_MAX_B = 8;                // weight ring stages
constexpr int H256_MIN_B = 4;
constexpr int H256_EPI_WARPS = 8;
constexpr int H256_STG_BYTES = 32 * 64;      // staging tile of one warp: 32 pixels x 32 channels x 2 B
constexpr int H256_ACC_STRIDE = 128;         // TMEM columns per accumulator stage (up to 4 sub-tiles x 32)

struct H256P {
  int H, W, Nimg;
  int tw, tiles_x, Wp;
  uint32_t wp_magic;
  int bh, bands_y, num_bands;
  int plane_bytes;          // bytes of one plane buffer (1024-aligned)
  int plane_tx_bytes;       // bytes the TMA box of one plane delivers
  int nb;                   // weight ring stages
  int n_sub_max;
  uint32_t idesc, desc_hi;
  int act;
  const float* bias;
  void* out;
  long long* trace;         // optional (debug, op.in[4]): wait-cycle counters of CTA 0's even issuer, see tools/one_conv.py
};

// linear band index -> (image, band row, tile column); tile column fastest
struct H256Walk {
  int img, by, tx;
  __device__ __forceinline__ void init(const H256P& p, int band) {
    tx = band % p.tiles_x;
    const int t = band / p.tiles_x;
    by = t % p.bands_y;
    img = t / p.bands_y;
  }
  __device__ __forceinline__ void next(const H256P& p) {
    if (++tx == p.tiles_x) {
      tx = 0;
      if (++by == p.bands_y) { by = 0; ++img; }
    }
  }
};

template <typename TO>
__global__ void __launch_bounds__(H256_THREADS, 1)
tc_conv3_halo256_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const H256P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_bfull = base;                          // [H256_MAX_B] weight chunk landed
  const uint32_t bar_bempty = base + 64;                    // [H256_MAX_B] weight chunk consumed (both issuers)
  const uint32_t bar_pfull = base + 128;                    // [H256_NPB] plane buffer filled
  const uint32_t bar_pempty = base + 160;                   // [H256_NPB] plane buffer consumed (both issuers)
  const uint32_t bar_tfull = base + 192;                    // [2] accumulator stage complete (both issuers)
  const uint32_t bar_tempty = base + 208;                   // [2] accumulator stage drained (256 arrivals)
  const uint32_t tmem_slot = base + 224;
  const uint32_t smem_a = base + H256_HEADER;               // H256_NPB plane buffers
  const uint32_t smem_b = smem_a + (uint32_t)H256_NPB * (uint32_t)p.plane_bytes;
  const uint32_t smem_stg = smem_b + (uint32_t)(p.nb * H256_CHUNK_BYTES);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapA);
    ptx::prefetch_tmap(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.nb; ++s) {
      ptx::mbar_init(bar_bfull + 8 * s, 1);
      ptx::mbar_init(bar_bempty + 8 * s, 2);
    }
    for (int pb = 0; pb < H256_NPB; ++pb) {
      ptx::mbar_init(bar_pfull + 8 * pb, 1);
      ptx::mbar_init(bar_pempty + 8 * pb, 2);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 2);
      ptx::mbar_init(bar_tempty + 8 * a, H256_EPI_WARPS * 32);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 2u * H256_ACC_STRIDE);
    ptx::tmem_relinquish();
  }
  if (warp == 3 && lane < H256_COUT / 4) {      // folded-BN shift: constant data
    float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(smem_raw + (base + H256_BIAS_OFF - raw) + 16 * lane) = b4;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
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
        for (int pl = 0; pl < H256_NPL; ++pl) {
          for (int trow = 0; trow < 3; ++trow) {
            ptx::mbar_wait(bar_bempty + 8 * stage, phase ^ 1u);
            const uint32_t full = bar_bfull + 8 * stage;
            ptx::mbar_arrive_expect_tx(full, (uint32_t)H256_CHUNK_BYTES);
#pragma unroll
            for (int ts = 0; ts < 3; ++ts)
              ptx::tma_load_2d(&mapB, full, smem_b + stage * H256_CHUNK_BYTES + ts * H256_TAP_BYTES, (trow * 3 + ts) * H256_C + pl * 64, 0);
            if (++stage == (uint32_t)p.nb) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================== tile producer (one TMA box per plane) =============
    if (ptx::elect_one()) {
      H256Walk w;
      w.init(p, band0);
      uint32_t pb = 0, pphase = 0;                     // plane buffer of the running (tile, plane) sequence and its fill parity
      for (int b = 0; b < nbands; ++b, w.next(p)) {
        const int x_left = w.tx * p.tw - 1, y_top = w.by * p.bh - 1;
        for (int pl = 0; pl < H256_NPL; ++pl) {
          ptx::mbar_wait(bar_pempty + 8 * pb, pphase ^ 1u);      // the MMAs that read this buffer's previous plane are complete
          ptx::mbar_arrive_expect_tx(bar_pfull + 8 * pb, (uint32_t)p.plane_tx_bytes);
          ptx::tma_load_4d(&mapA, bar_pfull + 8 * pb, smem_a + pb * (uint32_t)p.plane_bytes, 64 * pl, x_left, y_top, w.img);
          if (++pb == (uint32_t)H256_NPB) { pb = 0; pphase ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================================== MMA issuers (even / odd sub-tiles) ================
    const int parity = warp == 1 ? 0 : 1;
    const uint32_t a_lo0 = tc_desc_lo(smem_a, 1u), plane16 = (uint32_t)p.plane_bytes >> 4;
    const uint32_t b_lo0 = tc_desc_lo(smem_b, 1u);
    uint32_t stage = 0, phase = 0, pb = 0, pphase = 0;
    H256Walk w;
    w.init(p, band0);
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && parity == 0;
    long long t_acc = 0, t_plane = 0, t_w = 0, t_issue = 0, t_all = tr ? clock64() : 0, t0 = 0;
    for (int b = 0; b < nbands; ++b, w.next(p)) {
      const int bh_eff = min(p.bh, p.H - w.by * p.bh);
      const int n_sub = (bh_eff * p.Wp + 127) >> 7;
      const uint32_t acc = (uint32_t)(b & 1);
      if (tr) t0 = clock64();
      ptx::mbar_wait(bar_tempty + 8 * acc, (uint32_t)((b >> 1) & 1) ^ 1u);      // this stage's previous accumulators drained
      ptx::tc_fence_after();
      if (tr) t_acc += clock64() - t0;
      const uint32_t d_base = tmem_base + acc * (uint32_t)H256_ACC_STRIDE;
      for (int pl = 0; pl < H256_NPL; ++pl) {
        if (tr) t0 = clock64();
        ptx::mbar_wait(bar_pfull + 8 * pb, pphase);
        ptx::tc_fence_after();
        if (tr) t_plane += clock64() - t0;
        const uint32_t a_pl = a_lo0 + pb * plane16;
#pragma unroll 1
        for (int trow = 0; trow < 3; ++trow) {
          if (tr) t0 = clock64();
          ptx::mbar_wait(bar_bfull + 8 * stage, phase);
          ptx::tc_fence_after();
          if (tr) { const long long t1 = clock64(); t_w += t1 - t0; t0 = t1; }
          if (ptx::elect_one()) {
            const uint32_t a_r = a_pl + (uint32_t)(trow * p.Wp) * 8u;      // shifted window: 16-byte units, 128 B per pixel
            const uint32_t b_r = b_lo0 + stage * (uint32_t)(H256_CHUNK_BYTES >> 4);
            for (int j = parity; j < n_sub; j += 2) {
              const uint32_t d_tmem = d_base + (uint32_t)(j * H256_COUT);
              const uint32_t a_j = a_r + (uint32_t)(j * 128) * 8u;
#pragma unroll
              for (int ts = 0; ts < 3; ++ts) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  ptx::umma_f16_lohi(d_tmem, a_j + 8u * ts + 2u * kk, p.desc_hi, b_r + (uint32_t)(ts * (H256_TAP_BYTES >> 4)) + 2u * kk, p.desc_hi, p.idesc,
                                     (pl | trow | ts | kk) ? 1u : 0u);
              }
            }
            ptx::umma_commit(bar_bempty + 8 * stage);
            if (trow == 2) {
              ptx::umma_commit(bar_pempty + 8 * pb);
              if (pl == H256_NPL - 1) ptx::umma_commit(bar_tfull + 8 * acc);
            }
          }
          __syncwarp();
          if (tr) t_issue += clock64() - t0;
          if (++stage == (uint32_t)p.nb) { stage = 0; phase ^= 1u; }
        }
        if (++pb == (uint32_t)H256_NPB) { pb = 0; pphase ^= 1u; }
      }
    }
    if (tr && lane == 0) {
      p.trace[0] = clock64() - t_all; p.trace[1] = nbands; p.trace[2] = t_acc; p.trace[3] = t_plane; p.trace[4] = t_w; p.trace[5] = t_issue;
      p.trace[6] = p.bh; p.trace[7] = p.tw; p.trace[8] = p.nb;
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    const int q = warp & 3, grp = (warp - 4) >> 2;
    TO* out = reinterpret_cast<TO*>(p.out);
    uint8_t* const stg_ptr = smem_raw + (smem_stg - raw) + (uint32_t)(warp - 4) * H256_STG_BYTES;
    const float* const sbias = reinterpret_cast<const float*>(smem_raw + (base + H256_BIAS_OFF - raw));
    auto slot_off = [](int r, int c) { return (uint32_t)(r * 64 + ((c ^ (r & 3)) << 4)); };      // 64-byte rows, 4 chunks, XOR-swizzled
    const float floor_v = p.act == CAPF_ACT_RELU ? 0.f : -__int_as_float(0x7f800000);
    const uint64_t pol_out = ptx::policy_evict_last();
    H256Walk w;
    w.init(p, band0);
    for (int b = 0; b < nbands; ++b, w.next(p)) {
      const int y0 = w.by * p.bh, bh_eff = min(p.bh, p.H - y0);
      const int x0 = w.tx * p.tw, tw_eff = min(p.tw, p.W - x0);
      const int n_sub = (bh_eff * p.Wp + 127) >> 7;
      const uint32_t acc = (uint32_t)(b & 1);
      ptx::mbar_wait(bar_tfull + 8 * acc, (uint32_t)((b >> 1) & 1));
      ptx::tc_fence_after();
      for (int j = grp; j < n_sub; j += 2) {
        // element offset of this lane's pixel of the padded tile, -1 for the halo columns / rows past the tile
        const int mp = j * 128 + q * 32 + lane;
        const int iy = (int)__umulhi((uint32_t)mp, p.wp_magic), ix = mp - iy * p.Wp;
        const int myoff = (ix < tw_eff && iy < bh_eff) ? (((w.img * p.H + y0 + iy) * p.W + x0 + ix) * H256_COUT) : -1;
        const uint32_t taddr = tmem_base + acc * (uint32_t)H256_ACC_STRIDE + (uint32_t)(j * H256_COUT) + ((uint32_t)(q * 32) << 16);
        uint32_t a0[16], a1[16];
        ptx::tmem_ld16(taddr, a0);
        ptx::tmem_ld16(taddr + 16u, a1);
        ptx::tmem_ld_wait();
        epi16<TO, 0>(a0, sbias, floor_v, stg_ptr + lane * 64, 0u, (uint32_t)lane & 3u);
        epi16<TO, 0>(a1, sbias + 16, floor_v, stg_ptr + lane * 64, 2u, (uint32_t)lane & 3u);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int item = i * 32 + lane, r = item >> 2, c = item & 3;
          const int off = __shfl_sync(0xffffffffu, myoff, r);
          if (off >= 0) ptx::st_global_v4_hint(out + off + c * 8, *reinterpret_cast<const uint4*>(stg_ptr + slot_off(r, c)), pol_out);
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_tempty + 8 * acc);          // 256 arrivals: every accumulator of the stage has been read
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 2u * H256_ACC_STRIDE);
}

// =======================================================================================================
// host side
// =======================================================================================================
struct TcHalo256State {
  CUtensorMap mapA, mapB;
  H256P p;
  int grid, smem_bytes, dtype;
};

static int h256_plane_bytes(int bh, int Wp) {
  const int n_sub = (bh * Wp + 127) / 128;
  const int reach = n_sub * 128 + 2 * Wp + 2, box = (bh + 2) * Wp;      // the shifted windows of the last sub-tile read past the box
  const int pixels = ((reach > box ? reach : box) + 7) & ~7;
  return (pixels * 128 + 1023) & ~1023;
}

static int h256_plan(const capf_op& op, H256P& p, int& smem_bytes) {
  const char* ev = getenv("CAPF_HALO256");
  if (ev && ev[0] == '0') return 0;
  const int N = op.i[0], H = op.i[1], W = op.i[2], C = op.i[3], Cout = op.i[4];
  if (op.kind != CAPF_OP_CONV2D || op.i[5] != 3 || op.i[6] != 3 || op.i[7] != 1 || op.i[8] != 1) return 0;
  if (C != H256_C || Cout != H256_COUT || op.in[3]) return 0;                       // no residual variant (transition convs have none)
  if (op.dtype_out != op.dtype_in || (op.dtype_in != CAPF_F16 && op.dtype_in != CAPF_BF16)) return 0;
  if (op.i[18] != 0 || op.i[19] != 0 || op.i[20] != 0 || op.i[11] == CAPF_ACT_GELU) return 0;
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  if ((long long)N * H * W * C >= (1ll << 31)) return 0;
  memset(&p, 0, sizeof(p));
  p.H = H; p.W = W; p.Nimg = N;
  const int fixed = 1024 + H256_HEADER + H256_EPI_WARPS * H256_STG_BYTES;
  // tile = (bh rows) x (tw columns): at most 4 sub-tiles (TMEM stage), the four planes + >= 6 weight stages in shared memory;
  // cost per image in sub-tile units: MMA sub-tiles + ~0.3 per tile for the pipeline hand-over
  double best_cost = 1e300;
  int best_bh = 0, best_tx = 0;
  for (int tiles_x = 1; tiles_x <= 8 && tiles_x <= W; ++tiles_x) {
    const int tw = (W + tiles_x - 1) / tiles_x, Wp = tw + 2;
    if (Wp > 256 || (tiles_x > 1 && tw * (tiles_x - 1) >= W)) continue;
    for (int bh = 1; bh <= H && bh + 2 <= 256; ++bh) {
      const int n_sub = (bh * Wp + 127) / 128;
      if (n_sub > 4) break;
      if (fixed + H256_NPB * (long long)h256_plane_bytes(bh, Wp) + H256_MIN_B * H256_CHUNK_BYTES > TC_SMEM_LIMIT) break;
      const int full = H / bh, rem = H - full * bh;
      const double cost = tiles_x * ((double)full * (n_sub + 0.3) + (rem ? ((rem * Wp + 127) / 128 + 0.3) : 0.0));
      if (cost < best_cost - 1e-9) { best_cost = cost; best_bh = bh; best_tx = tiles_x; }
    }
  }
  if (!best_bh) return 0;
  p.bh = best_bh;
  p.tiles_x = best_tx;
  p.tw = (W + best_tx - 1) / best_tx;
  p.Wp = p.tw + 2;
  p.wp_magic = (uint32_t)(((1ull << 32) + p.Wp - 1) / p.Wp);
  p.bands_y = (H + p.bh - 1) / p.bh;
  const long long nbands = (long long)N * p.bands_y * p.tiles_x;
  if (nbands >= (1ll << 31)) return 0;
  p.num_bands = (int)nbands;
  p.plane_bytes = h256_plane_bytes(p.bh, p.Wp);
  p.plane_tx_bytes = (p.bh + 2) * p.Wp * 128;
  int nb = (TC_SMEM_LIMIT - fixed - H256_NPB * p.plane_bytes) / H256_CHUNK_BYTES;
  if (nb > H256_MAX_B) nb = H256_MAX_B;
  if (nb < H256_MIN_B) return 0;
  p.nb = nb;
  p.n_sub_max = (p.bh * p.Wp + 127) / 128;
  smem_bytes = fixed + H256_NPB * p.plane_bytes + nb * H256_CHUNK_BYTES;
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;     // one CTA per SM
  return 1;
}

int tc_halo256_supported(const capf_op& op) {
  H256P p;
  int smem;
  return h256_plan(op, p, smem);
}

int tc_halo256_prepare(const capf_op& op, TcHalo256State** out) {
  *out = nullptr;
  int e = tc_get_encoder();
  if (e) return e;
  TcHalo256State* s = new (std::nothrow) TcHalo256State();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_halo256_prepare: out of host memory");
  if (!h256_plan(op, s->p, s->smem_bytes)) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "halo256 conv: shape not supported"); }
  H256P& p = s->p;
  const bool bf16 = op.dtype_in == CAPF_BF16;
  p.idesc = tc_idesc(bf16, H256_COUT);
  p.desc_hi = tc_desc_hi(128, 1024);
  p.act = op.i[11];
  p.bias = (const float*)op.in[2];
  p.out = op.out[0];
  p.trace = (long long*)op.in[4];       // debug only (NULL in every program the host layer builds)
  s->grid = p.num_bands < num_sms() ? p.num_bands : num_sms();
  s->dtype = op.dtype_in;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  {
    const int K = 9 * H256_C;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)H256_COUT};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)H256_COUT};
    cuuint32_t es[2] = {1, 1};
    e = tc_encode_map(&s->mapB, dt, 2, op.in[1], dims, strides, box, es, 128, "B weights (halo256)");
  }
  if (!e) {
    cuuint64_t adims[4] = {(cuuint64_t)H256_C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.Nimg};
    cuuint64_t astr[3] = {(cuuint64_t)H256_C * 2, (cuuint64_t)p.W * H256_C * 2, (cuuint64_t)p.H * p.W * H256_C * 2};
    cuuint32_t abox[4] = {64, (cuuint32_t)p.Wp, (cuuint32_t)(p.bh + 2), 1};
    cuuint32_t aes[4] = {1, 1, 1, 1};
    e = tc_encode_map(&s->mapA, dt, 4, op.in[0], adims, astr, abox, aes, 128, "A halo256");
  }
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

template <typename T>
static int h256_launch_t(const TcHalo256State* s, cudaStream_t st) {
  static PerDevice<bool> opted_;
  std::atomic<bool>& opted = opted_.get();
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv3_halo256_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_conv3_halo256_kernel smem opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  launch_k(tc_conv3_halo256_kernel<T>, dim3(s->grid), dim3(H256_THREADS), s->smem_bytes, st, s->mapA, s->mapB, s->p);
  return check_launch("tc_conv3_halo256_kernel");
}

int tc_halo256_launch(const TcHalo256State* s, cudaStream_t st) {
  return s->dtype == CAPF_F16 ? h256_launch_t<__half>(s, st) : h256_launch_t<__nv_bfloat16>(s, st);
}

void tc_halo256_release(TcHalo256State* s) { delete s; }

void tc_halo256_describe(const TcHalo256State* s, char* buf, int cap) {
  snprintf(buf, cap, "tc_conv3_halo256_kernel[tile %d rows x %d cols, %d sub-tiles, %d weight stages]", s->p.bh, s->p.tw, s->p.n_sub_max, s->p.nb);
}

}  // namespace capf
