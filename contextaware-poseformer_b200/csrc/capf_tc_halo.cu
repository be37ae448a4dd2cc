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
// Roles (512 threads): warp 0 = weight TMA, warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator,
// warps 4..11 = two 4-warp epilogue groups (alternate 128-row sub-tiles), warps 12..15 = halo loaders (cp.async,
// zero fill outside the image).  Halo bands are double buffered; up to four TMEM accumulators are in flight.
#include <new>

#include "capf_tc.cuh"

namespace capf {

constexpr int HALO_THREADS = 512;
constexpr int HALO_HEADER_BYTES = 1024;
constexpr int HALO_MAX_ACC = 4;

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
  int acc_stages, acc_shift, tmem_cols, acc_stride;   // acc_stages = 1 << acc_shift accumulators in flight
  uint32_t idesc, b_desc_hi, a_desc_hi;
  int act;
  const void* x;
  const float* bias;
  const void* res;
  void* out;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// pix / Wp without a divide: magic = ceil(2^32 / Wp), exact for the pixel counts of one band (< 2^16)
__device__ __forceinline__ int div_wp(int v, uint32_t magic) { return (int)__umulhi((uint32_t)v, magic); }

template <int C, typename TI, typename TO>
__global__ void __launch_bounds__(HALO_THREADS, 1)
tc_conv3_halo_kernel(const __grid_constant__ CUtensorMap mapB, const HaloP p) {
  constexpr int KSTEPS = C / 16;                                   // 16-channel MMA steps per filter tap
  constexpr int KB = C % 64 == 0 ? 64 : C % 32 == 0 ? 32 : 16;     // weight chunk width (elements)
  constexpr int KPC = KB / 16, CPT = C / KB, CHUNKS = C / 8;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_b = base;                       // weights landed
  const uint32_t bar_hfull = base + 8;               // [2] halo buffer filled   (128 loader arrivals)
  const uint32_t bar_hempty = base + 24;             // [2] halo buffer consumed (tcgen05.commit)
  const uint32_t bar_tfull = base + 40;              // [HALO_MAX_ACC] accumulator complete
  const uint32_t bar_tempty = base + 40 + 8 * HALO_MAX_ACC;  // [HALO_MAX_ACC] accumulator drained (128 arrivals)
  const uint32_t tmem_slot = base + 40 + 16 * HALO_MAX_ACC;
  const uint32_t smem_b = base + HALO_HEADER_BYTES;
  const uint32_t smem_halo = smem_b + ((p.b_bytes + 1023) & ~1023);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&mapB);
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(bar_b, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_hfull + 8 * b, 128);
      ptx::mbar_init(bar_hempty + 8 * b, 1);
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
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // contiguous band range of this CTA; (image, band-in-image) advance by carry, never by division
  const int band0 = (int)(((long long)p.num_bands * blockIdx.x) / gridDim.x);
  const int band1 = (int)(((long long)p.num_bands * (blockIdx.x + 1)) / gridDim.x);
  const int img0 = band0 / p.bands_per_img, bin0 = band0 - img0 * p.bands_per_img;

  if (warp == 0) {
    // ===================================== resident weights ==================================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(bar_b, (uint32_t)p.b_bytes);
      for (int c = 0; c < 9 * CPT; ++c) ptx::tma_load_2d(&mapB, bar_b, smem_b + c * p.b_chunk_bytes, c * KB, 0);
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    ptx::mbar_wait(bar_b, 0);
    ptx::tc_fence_after();
    const uint32_t lbo_field = (uint32_t)p.P_alloc;      // chunk-plane pitch = P_alloc * 16 bytes, >> 4
    const uint64_t b_desc0 = tc_make_desc(smem_b, 1u, p.b_desc_hi);
    const uint32_t b_chunk16 = (uint32_t)p.b_chunk_bytes >> 4;
    uint32_t it = 0, k = 0;
    int bin = bin0;
    for (int band = band0; band < band1; ++band, ++k) {
      const uint32_t buf = k & 1u, hph = (k >> 1) & 1u;
      const int bh_eff = min(p.bh, p.H - bin * p.bh);
      const int n_sub = (bh_eff * p.Wp + 127) >> 7;
      if (++bin == p.bands_per_img) bin = 0;
      ptx::mbar_wait(bar_hfull + 8 * buf, hph);
      ptx::tc_fence_after();
      const uint64_t a_desc0 = tc_make_desc(smem_halo + buf * p.halo_bytes, lbo_field, p.a_desc_hi);
      for (int j = 0; j < n_sub; ++j, ++it) {
        const uint32_t acc = it & (uint32_t)(p.acc_stages - 1), aph = (it >> p.acc_shift) & 1u;
        ptx::mbar_wait(bar_tempty + 8 * acc, aph ^ 1u);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
          // descriptor start-address field counts 16-byte units == halo pixels
          const uint64_t a_sub = a_desc0 + (uint32_t)(j * 128);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint64_t a_tap = a_sub + (uint32_t)((tap / 3) * p.Wp + (tap % 3));
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) {
              const uint64_t a_d = a_tap + (uint32_t)(2 * kk) * (uint32_t)p.P_alloc;
              const uint64_t b_d = b_desc0 + (uint32_t)(tap * CPT + kk / KPC) * b_chunk16 + (uint32_t)((kk % KPC) * 2);
              ptx::umma_f16(d_tmem, a_d, b_d, p.idesc, (tap | kk) ? 1u : 0u);
            }
          }
          ptx::umma_commit(bar_tfull + 8 * acc);
          if (j == n_sub - 1) ptx::umma_commit(bar_hempty + 8 * buf);   // band fully read -> loaders may refill
        }
        __syncwarp();
      }
    }
  } else if (warp >= 12) {
    // ===================================== halo loaders =====================================
    const int tl = threadIdx.x - 12 * 32;            // 0..127
    const TI* x = reinterpret_cast<const TI*>(p.x);
    uint32_t k = 0;
    int img = img0, bin = bin0;
    for (int band = band0; band < band1; ++band, ++k) {
      const uint32_t buf = k & 1u, hph = (k >> 1) & 1u;
      const int y0 = bin * p.bh;
      const int bh_eff = min(p.bh, p.H - y0);
      const int total = ((bh_eff + 2) * p.Wp + 1) * CHUNKS;     // + the zero pixel right of the last row
      ptx::mbar_wait(bar_hempty + 8 * buf, hph ^ 1u);
      const uint32_t halo = smem_halo + buf * p.halo_bytes;
      const TI* imgp = x + (size_t)img * p.H * p.W * C;
#pragma unroll 4
      for (int t = tl; t < total; t += 128) {
        const int pix = t / CHUNKS, c = t % CHUNKS;              // CHUNKS is a compile-time constant
        const int hy = div_wp(pix, p.wp_magic), hx = pix - hy * p.Wp;
        const int iy = y0 - 1 + hy, ix = hx - 1;
        const bool ok = hx > 0 && iy >= 0 && iy < p.H && hy < bh_eff + 2;
        const TI* src = ok ? imgp + ((size_t)iy * p.W + ix) * C + c * 8 : x;
        cp_async16_zfill(halo + ((uint32_t)c * (uint32_t)p.P_alloc + (uint32_t)pix) * 16u, src, ok ? 16u : 0u);
      }
      cp_async_wait_all();
      ptx::fence_proxy_async();                       // generic-proxy writes -> visible to the tensor (async) proxy
      ptx::mbar_arrive(bar_hfull + 8 * buf);
      if (++bin == p.bands_per_img) { bin = 0; ++img; }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;                 // 0 | 1: sub-tiles alternate between the two groups
    const int row = q * 32 + lane;
    const int ngroups = (p.Cout + 31) / 32;
    const TO* res = reinterpret_cast<const TO*>(p.res);
    TO* out = reinterpret_cast<TO*>(p.out);
    uint32_t it = 0;
    int img = img0, bin = bin0;
    for (int band = band0; band < band1; ++band) {
      const int y0 = bin * p.bh;
      const int bh_eff = min(p.bh, p.H - y0);
      const int n_sub = (bh_eff * p.Wp + 127) >> 7;
      const size_t img_row0 = (size_t)img * p.H + y0;
      if (++bin == p.bands_per_img) { bin = 0; ++img; }
      for (int j = 0; j < n_sub; ++j, ++it) {
        if ((int)(it & 1u) != grp) continue;
        const uint32_t acc = it & (uint32_t)(p.acc_stages - 1), aph = (it >> p.acc_shift) & 1u;
        const int mp = j * 128 + row;
        const int iy = div_wp(mp, p.wp_magic), ix = mp - iy * p.Wp;
        const bool live = ix < p.W && iy < bh_eff;
        const bool has_res = live && res != nullptr;
        const size_t off0 = live ? ((img_row0 + iy) * p.W + ix) * p.Cout : 0;
        const uint32_t taddr = tmem_base + acc * p.acc_stride + ((uint32_t)(q * 32) << 16);

        Vec16<TO> r0[2], r1[2];
        auto fetch = [&](int g, Vec16<TO> (&r)[2]) {
          if (has_res) {
            const int c = 32 * g;
            r[0].load(res + off0 + c);
            if (c + 16 < p.Cout) r[1].load(res + off0 + c + 16);
          }
        };
        auto group = [&](int g, const Vec16<TO> (&r)[2], Vec16<TO> (&rnext)[2]) {
          const int c = 32 * g;
          const bool two = c + 16 < p.Cout;
          uint32_t a0[16], a1[16];
          ptx::tmem_ld16(taddr + (uint32_t)c, a0);
          if (two) ptx::tmem_ld16(taddr + (uint32_t)(c + 16), a1);
          if (g + 1 < ngroups) fetch(g + 1, rnext);
          ptx::tmem_ld_wait();
          if (live) {
            finish16<TO>(p.bias, p.act, a0, r[0], has_res, c, out + off0 + c);
            if (two) finish16<TO>(p.bias, p.act, a1, r[1], has_res, c + 16, out + off0 + c + 16);
          }
        };
        fetch(0, r0);
        ptx::mbar_wait(bar_tfull + 8 * acc, aph);
        ptx::tc_fence_after();
        for (int g = 0; g < ngroups; g += 2) {
          group(g, r0, r1);
          if (g + 1 < ngroups) group(g + 1, r1, r0);
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_tempty + 8 * acc);
      }
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
  CUtensorMap mapB;
  HaloP p;
  int grid, smem_bytes, dtype_in, dtype_out;
};

static int halo_plan(const capf_op& op, HaloP& p, int& smem_bytes) {
  const int N = op.i[0], H = op.i[1], W = op.i[2], C = op.i[3], Cout = op.i[4];
  if (op.i[5] != 3 || op.i[6] != 3 || op.i[7] != 1 || op.i[8] != 1) return 0;
  if ((C != 16 && C != 32 && C != 48 && C != 64) || Cout % 16 || Cout > 256) return 0;   // instantiated widths
  memset(&p, 0, sizeof(p));
  p.C = C; p.Cout = Cout; p.H = H; p.W = W; p.Nimg = N; p.Wp = W + 1;
  p.wp_magic = (uint32_t)(((1ull << 32) + p.Wp - 1) / p.Wp);
  p.chunks = C / 8;
  p.kb = C % 64 == 0 ? 64 : C % 32 == 0 ? 32 : 16;
  p.cpt = C / p.kb;
  p.b_chunk_bytes = Cout * p.kb * 2;
  p.b_bytes = 9 * p.cpt * p.b_chunk_bytes;
  const int b_region = (p.b_bytes + 1023) & ~1023;
  const int budget = TC_SMEM_LIMIT - 1024 - HALO_HEADER_BYTES - b_region;
  // band height: fewest 128-row sub-tiles per image, then the tallest band (fewest halo re-reads)
  int best_bh = 0;
  long long best_tiles = 1ll << 60;
  for (int bh = 1; bh <= H; ++bh) {
    const int n_sub_full = (bh * p.Wp + 127) / 128;
    const int P_alloc = (n_sub_full * 128 + 2 * p.Wp + 2 + 7) & ~7;
    const long long halo_bytes = (long long)P_alloc * C * 2;
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
  p.P_alloc = (n_sub_full * 128 + 2 * p.Wp + 2 + 7) & ~7;
  p.halo_bytes = p.P_alloc * C * 2;
  p.acc_shift = 4 * Cout <= 512 ? 2 : 1;
  p.acc_stages = 1 << p.acc_shift;
  int cols = 32;
  while (cols < p.acc_stages * Cout) cols <<= 1;
  p.tmem_cols = cols;
  p.acc_stride = Cout;
  smem_bytes = 1024 + HALO_HEADER_BYTES + b_region + 2 * p.halo_bytes;
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
  p.a_desc_hi = tc_desc_hi(0, 128);               // un-swizzled: 8-row groups are 128 contiguous bytes
  p.act = op.i[11];
  p.x = op.in[0];
  p.bias = (const float*)op.in[2];
  p.res = op.in[3];
  p.out = op.out[0];
  s->grid = p.num_bands < g_num_sms ? p.num_bands : g_num_sms;
  s->dtype_in = op.dtype_in;
  s->dtype_out = op.dtype_out;
  const int K = 9 * p.C;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)p.Cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)p.kb, (cuuint32_t)p.Cout};
  cuuint32_t es[2] = {1, 1};
  e = tc_encode_map(&s->mapB, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, op.in[1], dims, strides, box,
                    es, p.kb * 2, "B weights (halo)");
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

template <int C, typename TI, typename TO>
static int halo_launch_c(const TcHaloState* s, cudaStream_t st) {
  static bool opted = false;
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv3_halo_kernel<C, TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_conv3_halo_kernel smem opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  tc_conv3_halo_kernel<C, TI, TO><<<s->grid, HALO_THREADS, s->smem_bytes, st>>>(s->mapB, s->p);
  return check_launch("tc_conv3_halo_kernel");
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
  if (s->dtype_in == CAPF_F16 && s->dtype_out == CAPF_F32) return halo_launch_typed<__half, float>(s, st);
  if (s->dtype_in == CAPF_BF16 && s->dtype_out == CAPF_F32) return halo_launch_typed<__nv_bfloat16, float>(s, st);
  return set_error(CAPF_ERR_UNSUPPORTED, "halo conv: dtype combination");
}

void tc_halo_release(TcHaloState* s) { delete s; }

}  // namespace capf
