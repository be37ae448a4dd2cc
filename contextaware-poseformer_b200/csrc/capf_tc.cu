// tcgen05 / TMA implicit-GEMM kernel of libcapf_b200 (sm_100a).
//
// One persistent, warp-specialised kernel serves every 16-bit GEMM-shaped operator on the CA_PF.forward path:
//   * nn.Conv2d(bias=False)+BatchNorm2d(eval)(+residual)(+ReLU) of the backbones (pose_hrnet.py:79-136, 235-277,
//     382-408; networks/resnet.py:62-85; networks/refineNet.py:26-45)  -- "conv" mode, A fetched by 4-D TMA boxes
//     of the NHWC activation, one box per filter tap (zero fill outside the image = the conv padding; a box
//     traversal stride of 2 = the conv stride);
//   * nn.Linear(+GELU)(+residual) of the lifter (pose_dformer.py:25-31,49,56,132,221)  -- "rows" mode, A is a plain
//     row-major [M][K] matrix.
// D[128 x BN] (fp32, TMEM) += A[128 x 16] (smem, K-major, swizzled) * B[BN x 16]^T (smem, K-major, swizzled).
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator,
// warps 4..11 = epilogue (tcgen05.ld -> bias/GELU/residual/ReLU -> global).  Two TMEM accumulator stages let the
// epilogue of tile i overlap the MMAs of tile i+1; a ring of smem stages decouples TMA from the tensor pipe.
#include <new>

#include "capf_tc.cuh"

namespace capf {

// =======================================================================================================
// kernel parameters
// =======================================================================================================
constexpr int TC_THREADS = 384;          // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_HEADER_BYTES = 1024;       // barriers + TMEM base pointer
constexpr int TC_A_STAGE_BYTES = 128 * 128; // 128 rows x 64 elements x 2 B

struct TcP {
  int mode;                 // 0 = rows ([M][K] matrix), 1 = conv (NHWC, 4-D boxes)
  int M;                    // rows mode: valid rows
  int Cout, BN, n_tiles_n;
  int num_chunks;           // K chunks = taps * (Cin / kb)
  int kb;                   // elements per chunk: 16 | 32 | 64  (32 | 64 | 128-byte swizzled rows)
  int cpt;                  // chunks per filter tap = Cin / kb
  int cps;                  // chunks per pipeline stage = 64 / kb
  int KW, stride, pad;
  int bw, bh, bn;           // conv mode: output-pixel box (x, y, image) of one 128-row tile, bw*bh*bn <= 128
  int tiles_x, tiles_y;
  int Ho, Wo, Nimg;
  int num_tiles;            // m_tiles * n_tiles_n
  int num_stages;
  int a_chunk_bytes, b_chunk_bytes, stage_bytes;
  int tx_bytes_per_chunk;   // bytes the two TMA boxes of one chunk deliver
  uint32_t idesc;           // tcgen05 instruction descriptor (kind::f16, fp32 accumulate, M=128, N=BN)
  uint32_t desc_hi;         // high word of the smem matrix descriptors (SBO, version, swizzle mode)
  int tmem_cols;            // allocated TMEM columns (power of two >= 2*BN)
  int act;
  const float* bias;
  const void* res;
  void* out;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t hi) { return tc_make_desc(smem_addr, 1u, hi); }

// Tiles are handed out as one contiguous range per CTA, so every role walks its range with carry-propagating
// counters instead of integer divisions (the producer and the MMA issuer are single threads: a division costs them
// more than a TMA or MMA issue).  Order: column tile fastest, then x, y, image box.
struct TileWalk {
  int n_tile, tx, ty, tn;
  __device__ __forceinline__ void init(const TcP& p, int tile) {
    n_tile = tile % p.n_tiles_n;
    int mt = tile / p.n_tiles_n;
    tx = mt % p.tiles_x;
    int r = mt / p.tiles_x;
    ty = r % p.tiles_y;
    tn = r / p.tiles_y;
  }
  __device__ __forceinline__ void next(const TcP& p) {
    if (++n_tile == p.n_tiles_n) {
      n_tile = 0;
      if (++tx == p.tiles_x) {
        tx = 0;
        if (++ty == p.tiles_y) { ty = 0; ++tn; }
      }
    }
  }
};

__device__ __forceinline__ void tile_range(const TcP& p, int& t0, int& t1) {
  t0 = (int)(((long long)p.num_tiles * blockIdx.x) / gridDim.x);
  t1 = (int)(((long long)p.num_tiles * (blockIdx.x + 1)) / gridDim.x);
}

// =======================================================================================================
// the kernel
// =======================================================================================================
template <typename TO>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzled tiles need 1024-byte alignment
  const uint32_t bar_full = base;                            // [TC_MAX_STAGES] x 8 B
  const uint32_t bar_empty = base + 8 * TC_MAX_STAGES;       // [TC_MAX_STAGES]
  const uint32_t bar_tfull = base + 16 * TC_MAX_STAGES;      // [2]
  const uint32_t bar_tempty = bar_tfull + 16;                // [2]
  const uint32_t tmem_slot = bar_tempty + 16;                // uint32
  const uint32_t stage0 = base + TC_HEADER_BYTES;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapA);
    ptx::prefetch_tmap(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);
      ptx::mbar_init(bar_tempty + 8 * a, TC_THREADS - 128);
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
  const uint32_t acc_stride = (uint32_t)p.tmem_cols >> 1;
  // PDL: this one-wave persistent grid is fully resident -> let the successor be scheduled as SMs drain; nothing
  // produced by the predecessor (activations, residual) has been touched before this point.
  pdl_trigger();
  pdl_wait();
  int t0, t1;
  tile_range(p, t0, t1);

  if (warp == 0) {
    // ===================================== TMA producer (one elected thread) ================
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      TileWalk w;
      w.init(p, t0);
      for (int tile = t0; tile < t1; ++tile, w.next(p)) {
        const int m0 = (w.tx + p.tiles_x * (w.ty + p.tiles_y * w.tn)) * 128;            // rows mode (tiles_y == 1)
        const int ix_base = w.tx * p.bw * p.stride - p.pad, iy_base = w.ty * p.bh * p.stride - p.pad, n0 = w.tn * p.bn;
        const int nb0 = w.n_tile * p.BN;
        int r = 0, sx = 0, cc = 0, kcol = 0;       // filter tap (r, sx), channel chunk inside the tap, B column
        for (int c0 = 0; c0 < p.num_chunks; c0 += p.cps) {
          ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const int nc = min(p.cps, p.num_chunks - c0);
          const uint32_t full = bar_full + 8 * stage;
          ptx::mbar_arrive_expect_tx(full, (uint32_t)(nc * p.tx_bytes_per_chunk));
          uint32_t a_dst = stage0 + stage * p.stage_bytes;
          uint32_t b_dst = a_dst + TC_A_STAGE_BYTES;
          for (int j = 0; j < nc; ++j) {
            if (p.mode == 1) ptx::tma_load_4d(&mapA, full, a_dst, cc * p.kb, ix_base + sx, iy_base + r, n0);
            else ptx::tma_load_2d(&mapA, full, a_dst, kcol, m0);
            ptx::tma_load_2d(&mapB, full, b_dst, kcol, nb0);
            a_dst += p.a_chunk_bytes;
            b_dst += p.b_chunk_bytes;
            kcol += p.kb;
            if (++cc == p.cpt) {
              cc = 0;
              if (++sx == p.KW) { sx = 0; ++r; }
            }
          }
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    // The whole warp walks the pipeline (uniform control flow); one elected lane issues the tcgen05 ops.
    // A full stage is always 4 MMAs of K = 16: (64 / kb) chunks x (kb / 16) steps.
    uint32_t a_off[4], b_off[4];
    {
      const int ksteps = p.kb >> 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = i / ksteps, k = i - j * ksteps;
        a_off[i] = (uint32_t)(j * p.a_chunk_bytes + 32 * k) >> 4;
        b_off[i] = (uint32_t)(j * p.b_chunk_bytes + 32 * k) >> 4;
      }
    }
    const int mma_per_chunk = p.kb >> 4;
    uint32_t stage = 0, phase = 0, it = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * acc_stride;
      uint32_t accumulate = 0;
      for (int c0 = 0; c0 < p.num_chunks; c0 += p.cps) {
        ptx::mbar_wait(bar_full + 8 * stage, phase);
        ptx::tc_fence_after();
        const int nmma = min(p.cps, p.num_chunks - c0) * mma_per_chunk;
        const bool last = c0 + p.cps >= p.num_chunks;
        if (ptx::elect_one()) {
          const uint32_t a_src = stage0 + stage * p.stage_bytes;
          const uint64_t a_desc = make_desc(a_src, p.desc_hi), b_desc = make_desc(a_src + TC_A_STAGE_BYTES, p.desc_hi);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < nmma) {
              ptx::umma_f16(d_tmem, a_desc + a_off[i], b_desc + b_off[i], p.idesc, accumulate);
              accumulate = 1;
            }
          }
          ptx::umma_commit(bar_empty + 8 * stage);      // smem slot reusable once these MMAs have read it
          if (last) ptx::umma_commit(bar_tfull + 8 * acc);  // accumulator complete -> epilogue
        }
        __syncwarp();
        accumulate = 1;
        if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    // 8 warps: quadrant q = warp & 3 owns TMEM lanes [32q, 32q+32) (= tile rows); the two warps of a quadrant split
    // the BN accumulator columns.  Per group of 32 columns: two tcgen05.ld in flight, the residual of the NEXT group
    // already requested, one wait, then bias/GELU/residual/ReLU and 16-byte stores.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int split = ((p.BN / 16 + 1) / 2) * 16;
    const int cbeg = half ? split : 0, cend = half ? p.BN : split;
    const int ngroups = (cend - cbeg + 31) / 32;
    const TO* res = reinterpret_cast<const TO*>(p.res);
    TO* out = reinterpret_cast<TO*>(p.out);
    // tile-invariant position of this thread's row inside the output-pixel box
    const int bx = row % p.bw;
    const int by = (row / p.bw) % p.bh, bi = row / (p.bw * p.bh);
    uint32_t it = 0;
    TileWalk w;
    w.init(p, t0);
    for (int tile = t0; tile < t1; ++tile, ++it, w.next(p)) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      long long orow;  // output row (pixel) index, -1 when this tile row is padding
      if (p.mode == 1) {
        const int ox = w.tx * p.bw + bx, oy = w.ty * p.bh + by, n = w.tn * p.bn + bi;
        const bool ok = bi < p.bn && ox < p.Wo && oy < p.Ho && n < p.Nimg;
        orow = ok ? ((long long)n * p.Ho + oy) * p.Wo + ox : -1;
      } else {
        const int m = w.tx * 128 + row;
        orow = m < p.M ? (long long)m : -1;
      }
      const bool live = orow >= 0;
      const bool has_res = live && res != nullptr;
      const int ncol0 = w.n_tile * p.BN;
      const size_t off0 = (size_t)(live ? orow : 0) * p.Cout + ncol0;
      const uint32_t taddr = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);

      Vec16<TO> r0[2], r1[2];
      auto fetch = [&](int g, Vec16<TO> (&r)[2]) {
        if (has_res) {
          const int c = cbeg + 32 * g;
          r[0].load(res + off0 + c);
          if (c + 16 < cend) r[1].load(res + off0 + c + 16);
        }
      };
      auto group = [&](int g, const Vec16<TO> (&r)[2], Vec16<TO> (&rnext)[2]) {
        const int c = cbeg + 32 * g;
        const bool two = c + 16 < cend;
        uint32_t a0[16], a1[16];
        ptx::tmem_ld16(taddr + (uint32_t)c, a0);
        if (two) ptx::tmem_ld16(taddr + (uint32_t)(c + 16), a1);
        if (g + 1 < ngroups) fetch(g + 1, rnext);
        Bias16 b0, b1;
        b0.load(p.bias, ncol0 + c);
        b1.load(p.bias, ncol0 + (two ? c + 16 : c));
        ptx::tmem_ld_wait();
        if (live) {
          finish16<TO>(b0, p.act, a0, r[0], has_res, out + off0 + c);
          if (two) finish16<TO>(b1, p.act, a1, r[1], has_res, out + off0 + c + 16);
        }
      };
      fetch(0, r0);                                  // independent of the MMAs: issue before waiting for them
      ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
      ptx::tc_fence_after();
      for (int g = 0; g < ngroups; g += 2) {
        group(g, r0, r1);
        if (g + 1 < ngroups) group(g + 1, r1, r0);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_tempty + 8 * acc);   // 256 arrivals free this accumulator stage
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// =======================================================================================================
// host side: tensor maps + launch geometry
// =======================================================================================================
struct TcConvState {
  TcHaloState* halo = nullptr;   // non-null: the op runs on the halo-band kernel instead of tc_gemm_kernel
  CUtensorMap mapA, mapB;
  TcP p;
  int grid;
  int smem_bytes;
  int dtype_out;
};

static PFN_cuTensorMapEncodeTiled g_encode = nullptr;

int tc_get_encoder() {
  if (g_encode) return CAPF_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return set_errorf(CAPF_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  return CAPF_OK;
}

struct ConvGeo {
  int N, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo, act;
};

static ConvGeo geo_of(const capf_op& op) {
  ConvGeo g;
  g.N = op.i[0]; g.H = op.i[1]; g.W = op.i[2]; g.Cin = op.i[3]; g.Cout = op.i[4];
  g.KH = op.i[5]; g.KW = op.i[6]; g.stride = op.i[7]; g.pad = op.i[8]; g.Ho = op.i[9]; g.Wo = op.i[10];
  g.act = op.i[11];
  return g;
}

int tc_conv_supported(const capf_op& op) {
  if (op.kind != CAPF_OP_CONV2D) return 0;
  if (op.dtype_in != CAPF_F16 && op.dtype_in != CAPF_BF16) return 0;
  if (op.dtype_out != CAPF_F16 && op.dtype_out != CAPF_BF16 && op.dtype_out != CAPF_F32) return 0;
  const ConvGeo g = geo_of(op);
  if (g.Cin <= 0 || g.Cin % 16 || g.Cout <= 0 || g.Cout % 16) return 0;
  if (g.KH < 1 || g.KW < 1 || g.KH > 7 || g.KW > 7 || g.stride < 1 || g.stride > 2 || g.pad < 0) return 0;
  if (g.N <= 0 || g.H <= 0 || g.W <= 0) return 0;
  if (g.Ho != (g.H + 2 * g.pad - g.KH) / g.stride + 1 || g.Wo != (g.W + 2 * g.pad - g.KW) / g.stride + 1) return 0;
  if ((long long)g.N * g.Ho * g.Wo >= (1ll << 31)) return 0;
  if (!op.in[0] || !op.in[1] || !op.out[0]) return 0;
  if (((uintptr_t)op.in[0] | (uintptr_t)op.in[1]) & 15) return 0;       // TMA base alignment
  if (((uintptr_t)op.out[0] | (uintptr_t)op.in[3]) & 15) return 0;      // 16-byte epilogue vectors
  if (op.in[2] && ((uintptr_t)op.in[2] & 15)) return 0;
  return 1;
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Output-pixel box (bw x bh x bn <= 128 rows) that wastes the fewest accumulator rows.
static void choose_box(const ConvGeo& g, int max_w, int& bw, int& bh, int& bn) {
  double best = -1.0;
  bw = bh = bn = 1;
  for (int w = 1; w <= g.Wo && w <= 128 && w <= max_w; ++w) {
    for (int h = 1; h <= g.Ho && w * h <= 128; ++h) {
      int n = 128 / (w * h);
      if (n > g.N) n = g.N;
      if (n < 1) continue;
      double tiles = (double)ceil_div(g.Wo, w) * ceil_div(g.Ho, h) * ceil_div(g.N, n);
      double util = ((double)g.Wo * g.Ho * g.N) / (tiles * 128.0);
      // prefer wide boxes (longer contiguous runs for TMA and for the epilogue stores) on ties
      double score = util + 1e-6 * w + 1e-9 * h;
      if (score > best) { best = score; bw = w; bh = h; bn = n; }
    }
  }
}

int tc_encode_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* ptr, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, int swz_bytes, const char* what) {
  CUtensorMapSwizzle sw = swz_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swz_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(m, dt, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_errorf(CAPF_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return CAPF_OK;
}

int tc_conv_prepare(const capf_op& op, TcConvState** out) {
  *out = nullptr;
  int e = tc_get_encoder();
  if (e) return e;
  const ConvGeo g = geo_of(op);
  TcConvState* s = new (std::nothrow) TcConvState();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_conv_prepare: out of host memory");
  // i[13]: kernel variant hint (0 = automatic, 1 = per-tap TMA GEMM, 2 = halo band); tests use it for A/B parity
  if (op.i[13] != 1 && tc_halo_supported(op)) {
    e = tc_halo_prepare(op, &s->halo);
    if (e) { delete s; return e; }
    *out = s;
    return CAPF_OK;
  }
  if (op.i[13] >= 2) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "tc conv: halo variant requested but not applicable"); }
  TcP& p = s->p;
  memset(&p, 0, sizeof(p));
  const bool rows = (g.KH == 1 && g.KW == 1 && g.stride == 1 && g.pad == 0);
  p.mode = rows ? 0 : 1;
  p.kb = g.Cin % 64 == 0 ? 64 : g.Cin % 32 == 0 ? 32 : 16;
  p.cpt = g.Cin / p.kb;
  p.cps = 64 / p.kb;
  p.num_chunks = g.KH * g.KW * p.cpt;
  p.KW = g.KW; p.stride = g.stride; p.pad = g.pad;
  p.Cout = g.Cout; p.Ho = g.Ho; p.Wo = g.Wo; p.Nimg = g.N;
  p.act = g.act;
  p.bias = (const float*)op.in[2];
  p.res = op.in[3];
  p.out = op.out[0];
  const int swz = p.kb * 2;

  // ---- M tiling ----------------------------------------------------------------------------------------
  long long m_tiles;
  if (rows) {
    p.M = g.N * g.Ho * g.Wo;
    p.bw = 128; p.bh = 1; p.bn = 1; p.tiles_x = ceil_div(p.M, 128); p.tiles_y = 1;
    m_tiles = p.tiles_x;
  } else {
    choose_box(g, 256 / g.stride, p.bw, p.bh, p.bn);
    p.tiles_x = ceil_div(g.Wo, p.bw);
    p.tiles_y = ceil_div(g.Ho, p.bh);
    m_tiles = (long long)p.tiles_x * p.tiles_y * ceil_div(g.N, p.bn);
    p.M = g.N * g.Ho * g.Wo;
  }

  // ---- N tiling: BN | Cout, multiple of 16, <= 256; fewest (waves x per-tile cost) ------------------------
  int best_bn = 0;
  double best_cost = 1e300;
  for (int bn = 16; bn <= 256 && bn <= g.Cout; bn += 16) {
    if (g.Cout % bn) continue;
    long long tiles = m_tiles * (g.Cout / bn);
    long long waves = (tiles + g_num_sms - 1) / g_num_sms;
    double cost = (double)waves * (bn + 24.0);
    if (cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && bn > best_bn)) { best_cost = cost; best_bn = bn; }
  }
  if (!best_bn) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "tc conv: no column tile"); }
  p.BN = best_bn;
  p.n_tiles_n = g.Cout / p.BN;
  long long nt = m_tiles * p.n_tiles_n;
  if (nt >= (1ll << 31)) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "tc conv: too many tiles"); }
  p.num_tiles = (int)nt;

  p.a_chunk_bytes = 128 * p.kb * 2;
  p.b_chunk_bytes = p.BN * p.kb * 2;
  p.stage_bytes = TC_A_STAGE_BYTES + p.BN * 128;
  const int box_rows = rows ? 128 : p.bw * p.bh * p.bn;
  p.tx_bytes_per_chunk = (box_rows + p.BN) * p.kb * 2;
  int stages = (TC_SMEM_LIMIT - TC_HEADER_BYTES - 1024) / p.stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  const int k_stages = ceil_div(p.num_chunks, p.cps);
  if (stages > 2 * k_stages && 2 * k_stages >= 2) stages = 2 * k_stages;   // no point in a ring deeper than two tiles
  if (stages < 2) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "tc conv: tile does not fit shared memory"); }
  p.num_stages = stages;
  s->smem_bytes = TC_HEADER_BYTES + 1024 + stages * p.stage_bytes;
  if (s->smem_bytes < 120 * 1024) s->smem_bytes = 120 * 1024;   // one CTA per SM: the CTA owns the SM's TMEM columns
  int cols = 32;
  while (cols < 2 * p.BN) cols <<= 1;
  p.tmem_cols = cols;
  p.idesc = tc_idesc(op.dtype_in == CAPF_BF16, p.BN);
  p.desc_hi = tc_desc_hi(swz, 8 * swz);
  s->grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
  s->dtype_out = op.dtype_out;

  // ---- tensor maps -------------------------------------------------------------------------------------
  const CUtensorMapDataType dt = op.dtype_in == CAPF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const int K = g.KH * g.KW * g.Cin;
  if (rows) {
    cuuint64_t dims[2] = {(cuuint64_t)g.Cin, (cuuint64_t)p.M};
    cuuint64_t strides[1] = {(cuuint64_t)g.Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kb, 128};
    cuuint32_t es[2] = {1, 1};
    e = tc_encode_map(&s->mapA, dt, 2, op.in[0], dims, strides, box, es, swz, "A rows");
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.W * g.Cin * 2, (cuuint64_t)g.H * g.W * g.Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.kb, (cuuint32_t)(p.bw * g.stride), (cuuint32_t)(p.bh * g.stride), (cuuint32_t)p.bn};
    cuuint32_t es[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
    e = tc_encode_map(&s->mapA, dt, 4, op.in[0], dims, strides, box, es, swz, "A conv");
  }
  if (!e) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)g.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kb, (cuuint32_t)p.BN};
    cuuint32_t es[2] = {1, 1};
    e = tc_encode_map(&s->mapB, dt, 2, op.in[1], dims, strides, box, es, swz, "B weights");
  }
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

template <typename TO>
static int tc_launch_typed(const TcConvState* s, cudaStream_t st) {
  static int max_smem = 0;   // opt-in once per instantiation (outside graph capture: Plan.capture warms up first)
  if (s->smem_bytes > max_smem) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_gemm_kernel smem opt-in: %s", cudaGetErrorString(e));
    max_smem = TC_SMEM_LIMIT;
  }
  launch_k(tc_gemm_kernel<TO>, dim3(s->grid), dim3(TC_THREADS), s->smem_bytes, st, s->mapA, s->mapB, s->p);
  return check_launch("tc_gemm_kernel");
}

int tc_conv_launch(const capf_op&, const TcConvState* s, cudaStream_t st) {
  if (!s) return set_error(CAPF_ERR_ARG, "tc conv: op was not prepared");
  if (s->halo) return tc_halo_launch(s->halo, st);
  switch (s->dtype_out) {
    case CAPF_F32: return tc_launch_typed<float>(s, st);
    case CAPF_F16: return tc_launch_typed<__half>(s, st);
    case CAPF_BF16: return tc_launch_typed<__nv_bfloat16>(s, st);
    default: return set_error(CAPF_ERR_UNSUPPORTED, "tc conv: dtype_out");
  }
}

void tc_conv_release(TcConvState* s) {
  if (s && s->halo) tc_halo_release(s->halo);
  delete s;
}

}  // namespace capf
