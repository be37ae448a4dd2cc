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
#include <cstdio>
#include <cstdlib>
#include <new>

#include "capf_tc.cuh"

namespace capf {

// =======================================================================================================
// kernel parameters
// =======================================================================================================
constexpr int TC_THREADS = 384;          // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_HEADER_BYTES = 1024;       // barriers + TMEM base pointer
constexpr int TC_A_SUB_BYTES = 128 * 128;   // one 128-row sub-tile of a stage: 128 rows x 64 elements x 2 B
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_STG_BYTES = 32 * 128;      // epilogue staging tile of one warp: 32 rows x 128 bytes

struct TcP {
  int mode;                 // 0 = rows ([M][K] matrix), 1 = conv (NHWC, 4-D boxes)
  int M;                    // rows mode: valid rows
  int msub;                 // 128-row sub-tiles per tile (1 | 2): both share every B stage, one accumulator each
  int a_stage_bytes;        // msub * TC_A_SUB_BYTES (B follows A inside a stage)
  int a_sub_bytes;          // offset of sub-tile 1 inside one A chunk = 128 * kb * 2
  int nacc;                 // TMEM accumulator stages (2, or 1 when 2 * msub * BN > 512 columns)
  int stg_bufs;             // staging buffers per epilogue warp (2 with a residual, else 1)
  uint32_t stg_off;         // staging region offset from the first pipeline stage
  uint32_t bias_off;        // per-warp bias slices (8 x 256 floats), offset from the first pipeline stage
  int Cout, BN, n_tiles_n;
  int num_chunks;           // K chunks = taps * (Cin / kb)
  int kb;                   // elements per chunk: 16 | 32 | 64  (32 | 64 | 128-byte swizzled rows)
  int cpt;                  // chunks per filter tap = Cin / kb (split operands: pass 0 walks the hi|lo planes, 2 * C / kb)
  int cpt1;                 // split operands only: chunks per tap of pass 1 (the hi plane alone, C / kb); else 0
  int KH;
  int c_a1;                 // rows mode with TWO A matrices (K-concatenation): chunks that come from the first one; else num_chunks
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
  int tmem_cols;            // allocated TMEM columns (power of two >= nacc * msub * BN)
  int acc_stride;           // TMEM columns per accumulator stage = msub * BN
  int act;
  const float* bias;
  const void* res;
  void* out;
  // output segments (CONV2D i[20] > 1, kernel MODE bit 2): columns [seg_beg[s], seg_beg[s] + seg_c[s]) of the GEMM go to the dense
  // tensor seg_out[s] [rows][seg_c[s]]; unused entries have seg_beg = INT_MAX.  Bit s of seg_noact: no ReLU on segment s.
  int nseg;
  int seg_beg[4], seg_c[4];
  void* seg_out[4];
  uint32_t seg_noact;
  long long* trace;         // optional (debug, op.in[4]): clock64 timeline of CTA 0, see tools/gemm_trace.py
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
// MODE: bit 0 = residual add, bit 1 = GELU, bit 2 = output segments (compile-time epilogue variants)
template <typename TO, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapA2, const TcP p) {
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
  const bool tr = p.trace != nullptr && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[0] = clock64();
    p.trace[1] = (long long)gt;
  }

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapA);
    ptx::prefetch_tmap(&mapB);
    if (p.c_a1 < p.num_chunks) ptx::prefetch_tmap(&mapA2);
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
  const uint32_t acc_stride = (uint32_t)p.acc_stride;
  // PDL: this one-wave persistent grid is fully resident -> let the successor be scheduled as SMs drain.  griddepcontrol.wait
  // is per thread: every role waits right before it first touches what the predecessor produced (activations, residual,
  // output).  The producer first does its tile arithmetic and requests the WEIGHT boxes of its first stages (constant
  // data), so their L2 / HBM latency and its own set-up overlap the predecessor's tail; the MMA issuer never touches
  // global memory and does not wait at all.
  if (tr && threadIdx.x == 0) p.trace[2] = clock64();
  pdl_trigger();
  int t0, t1;
  tile_range(p, t0, t1);
  const int tile_rows = 128 * p.msub;

  if (warp == 0) {
    // ===================================== TMA producer (one elected thread) ================
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      TileWalk w;
      w.init(p, t0);
      // weights of the first stages of the first tile, before the dependency wait
      int pre = 0;
      if (t0 < t1) {
        const int k_stages = (p.num_chunks + p.cps - 1) / p.cps;
        pre = min(p.num_stages, k_stages);
        const int nb0 = w.n_tile * p.BN;
        for (int s2 = 0; s2 < pre; ++s2) {
          const int c0 = s2 * p.cps, nc = min(p.cps, p.num_chunks - c0);
          const uint32_t full = bar_full + 8 * s2;
          ptx::mbar_arrive_expect_tx(full, (uint32_t)(nc * p.tx_bytes_per_chunk));
          uint32_t b_dst = stage0 + s2 * p.stage_bytes + p.a_stage_bytes;
          for (int j = 0; j < nc; ++j) {
            ptx::tma_load_2d(&mapB, full, b_dst, (c0 + j) * p.kb, nb0);
            b_dst += p.b_chunk_bytes;
          }
        }
      }
      pdl_wait();
      if (tr) p.trace[3] = clock64();
      for (int tile = t0; tile < t1; ++tile, w.next(p)) {
        const int m0 = (w.tx + p.tiles_x * (w.ty + p.tiles_y * w.tn)) * tile_rows;      // rows mode (tiles_y == 1)
        const int ix_base = w.tx * p.bw * p.stride - p.pad, iy_base = w.ty * p.bh * p.stride - p.pad, n0 = w.tn * p.bn;
        const int nb0 = w.n_tile * p.BN;
        int r = 0, sx = 0, cc = 0, kcol = 0;       // filter tap (r, sx), channel chunk inside the tap, B column
        int cpt = p.cpt;                           // split operands: a second pass over the taps with the hi plane only
        for (int c0 = 0; c0 < p.num_chunks; c0 += p.cps) {
          const bool primed = tile == t0 && c0 < pre * p.cps;      // B already requested, barrier already armed
          if (!primed) ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const int nc = min(p.cps, p.num_chunks - c0);
          const uint32_t full = bar_full + 8 * stage;
          if (!primed) ptx::mbar_arrive_expect_tx(full, (uint32_t)(nc * p.tx_bytes_per_chunk));
          uint32_t a_dst = stage0 + stage * p.stage_bytes;
          uint32_t b_dst = a_dst + p.a_stage_bytes;
          for (int j = 0; j < nc; ++j) {
            if (p.mode == 1) ptx::tma_load_4d(&mapA, full, a_dst, cc * p.kb, ix_base + sx, iy_base + r, n0);
            else if (cc < p.c_a1) ptx::tma_load_2d(&mapA, full, a_dst, cc * p.kb, m0);
            else ptx::tma_load_2d(&mapA2, full, a_dst, (cc - p.c_a1) * p.kb, m0);      // second operand of a K-concatenated Linear
            if (!primed) ptx::tma_load_2d(&mapB, full, b_dst, kcol, nb0);
            a_dst += p.a_chunk_bytes;
            b_dst += p.b_chunk_bytes;
            kcol += p.kb;
            if (++cc == cpt) {
              cc = 0;
              if (++sx == p.KW) {
                sx = 0;
                if (++r == p.KH && p.cpt1) { r = 0; cpt = p.cpt1; }
              }
            }
          }
          if (tr) { const int n = (tile - t0) * ((p.num_chunks + p.cps - 1) / p.cps) + c0 / p.cps; if (n < 64) p.trace[64 + n] = clock64(); }
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (one elected thread) ====================
    // One thread walks the pipeline: barrier waits, descriptor arithmetic and the tcgen05 ops are all single-thread work, and
    // a stage of an N <= 64 GEMM is only 4-8 MMAs of ~40 clk -- a per-stage warp re-convergence + election would be a
    // visible fraction of it.  A full stage is always 4 MMAs of K = 16 per 128-row sub-tile: (64 / kb) chunks x (kb / 16) steps;
    // the msub sub-tiles of a tile read the same B stage and accumulate into neighbouring TMEM column ranges.
    if (ptx::elect_one()) {
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
      const uint32_t a_sub16 = (uint32_t)p.a_sub_bytes >> 4;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = t0; tile < t1; ++tile) {
        ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        uint32_t accumulate = 0;
        for (int c0 = 0; c0 < p.num_chunks; c0 += p.cps) {
          ptx::mbar_wait(bar_full + 8 * stage, phase);
          ptx::tc_fence_after();
          const int nmma = min(p.cps, p.num_chunks - c0) * mma_per_chunk;
          const bool last = c0 + p.cps >= p.num_chunks;
          const uint32_t a_src = stage0 + stage * p.stage_bytes;
          const uint64_t a_desc = make_desc(a_src, p.desc_hi), b_desc = make_desc(a_src + p.a_stage_bytes, p.desc_hi);
          for (int sub = 0; sub < p.msub; ++sub) {
            const uint64_t a_sub = a_desc + (uint64_t)(sub * a_sub16);
            const uint32_t d_sub = d_tmem + (uint32_t)(sub * p.BN);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (i < nmma) ptx::umma_f16(d_sub, a_sub + a_off[i], b_desc + b_off[i], p.idesc, (accumulate | (uint32_t)i) ? 1u : 0u);
            }
          }
          ptx::umma_commit(bar_empty + 8 * stage);      // smem slot reusable once these MMAs have read it
          if (last) ptx::umma_commit(bar_tfull + 8 * acc);  // accumulator complete -> epilogue
          if (tr) { const int n = (tile - t0) * ((p.num_chunks + p.cps - 1) / p.cps) + c0 / p.cps; if (n < 64) p.trace[128 + n] = clock64(); }
          accumulate = 1;
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == (uint32_t)p.nacc) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    // 8 warps: quadrant q = warp & 3 owns TMEM lanes [32q, 32q+32) (= rows of a 128-row sub-tile).  With one sub-tile
    // per tile the two warps of a quadrant split the BN accumulator columns; with two sub-tiles each takes one.
    //
    // All global traffic is coalesced through a warp-private, XOR-swizzled staging tile of 32 rows x 128 bytes (one
    // "slab" = 64 16-bit / 32 fp32 columns): the residual slab is requested with cp.async one slab ahead (and pulled
    // into L2 a whole tile ahead), every thread adds its own row in place, and the finished slab leaves with 16 bytes
    // per lane, consecutive lanes -> consecutive addresses.  The epilogue is a single warp's dependent instruction
    // stream per 32 rows, so its instruction count is what bounds short GEMMs: activation / residual are compile-time
    // variants (MODE), the bias slice of the column tile sits in shared memory before the accumulator is waited for,
    // and the rows-mode store addresses are affine (no shuffles, no 64-bit multiplies in the loop).
    constexpr bool HAS_RES = (MODE & 1) != 0;
    constexpr bool SEG = (MODE & 4) != 0;
    constexpr int EMODE = MODE & 3;
    constexpr int SW = 128 / (int)sizeof(TO);          // columns per slab
    constexpr int CPG = 16 * (int)sizeof(TO) / 16;     // 16-byte chunks per 16 columns
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int sub = p.msub == 2 ? half : 0;
    const int row = sub * 128 + q * 32 + lane;
    const int split = ((p.BN / 16 + 1) / 2) * 16;
    const int cbeg = (p.msub == 1 && half) ? split : 0, cend = (p.msub == 1 && !half) ? split : p.BN;
    const int nslabs = (cend - cbeg + SW - 1) / SW;
    const TO* res = reinterpret_cast<const TO*>(p.res);
    TO* out = reinterpret_cast<TO*>(p.out);
    // tile-invariant position of this thread's row inside the output-pixel box
    const int bx = row % p.bw;
    const int by = (row / p.bw) % p.bh, bi = row / (p.bw * p.bh);
    const uint32_t tcol0 = (uint32_t)(sub * p.BN) + ((uint32_t)(q * 32) << 16);
    const float floor_v = p.act == CAPF_ACT_RELU ? 0.f : -__int_as_float(0x7f800000);
    // segment of GEMM column `col` (boundaries are multiples of 16 columns, so a 16-column group / a 16-byte chunk has one)
    auto seg_of = [&](int col) -> int { return (col >= p.seg_beg[1] ? 1 : 0) + (col >= p.seg_beg[2] ? 1 : 0) + (col >= p.seg_beg[3] ? 1 : 0); };
    auto seg_floor = [&](int col) -> float { return ((p.seg_noact >> seg_of(col)) & 1u) ? -__int_as_float(0x7f800000) : floor_v; };
    uint8_t* const stg_ptr = smem_raw + (stage0 - raw) + p.stg_off + (uint32_t)(warp - 4) * (uint32_t)(p.stg_bufs * TC_STG_BYTES);
    const uint32_t stg = stage0 + p.stg_off + (uint32_t)(warp - 4) * (uint32_t)(p.stg_bufs * TC_STG_BYTES);
    float* const sbias = reinterpret_cast<float*>(smem_raw + (stage0 - raw) + p.bias_off) + (warp - 4) * 256;
    auto slot_off = [&](uint32_t buf, int r, int c) { return buf * TC_STG_BYTES + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); };
    const uint64_t pol_out = ptx::policy_evict_last();
    const uint32_t row_bytes = (uint32_t)p.Cout * (uint32_t)sizeof(TO);
    // output row (pixel) index of accumulator row `row` of the tile the walker points at, -1 when it is padding
    auto row_index = [&](const TileWalk& t) -> int {
      if (p.mode == 1) {
        const int ox = t.tx * p.bw + bx, oy = t.ty * p.bh + by, n = t.tn * p.bn + bi;
        const bool ok = bi < p.bn && ox < p.Wo && oy < p.Ho && n < p.Nimg;
        return ok ? (n * p.Ho + oy) * p.Wo + ox : -1;
      }
      const int m = t.tx * tile_rows + row;
      return m < p.M ? m : -1;
    };
    uint32_t acc = 0, acc_phase = 0;
    TileWalk w;
    w.init(p, t0);
    pdl_wait();                                       // residual reads / output writes start below
    for (int tile = t0; tile < t1; ++tile) {
      const int myrow = row_index(w);
      const int ncol0 = w.n_tile * p.BN;
      const uint32_t taddr = tmem_base + acc * acc_stride + tcol0;
      // bias slice of this warp's columns -> shared memory (zeros without a bias)
      for (int c = 4 * lane; c < cend - cbeg; c += 128)
        *reinterpret_cast<float4*>(sbias + c) = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + ncol0 + cbeg + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      // item = i * 32 + lane -> (row r = item / chs, chunk c = item % chs) of a slab with chs 16-byte chunks per row
      auto prefetch_res = [&](int s, uint32_t buf) {
        const int c0 = cbeg + s * SW;
        const int chs = (min(SW, cend - c0) * (int)sizeof(TO)) >> 4;
        const uint32_t magic = (65536u + (uint32_t)chs - 1u) / (uint32_t)chs;
        for (int i = 0; i < chs; ++i) {
          const int item = i * 32 + lane, r = (int)(((uint32_t)item * magic) >> 16), c = item - r * chs;
          const int grow = __shfl_sync(0xffffffffu, myrow, r);
          if (grow >= 0) ptx::cp_async16(stg + slot_off(buf, r, c), reinterpret_cast<const uint8_t*>(res + (size_t)grow * p.Cout + ncol0 + c0) + 16 * c);
        }
        ptx::cp_async_commit();
      };
      TileWalk wn = w;
      wn.next(p);
      if (HAS_RES) {
        prefetch_res(0, 0);                          // independent of the MMAs: issue before waiting for them
        if (tile + 1 < t1) {                         // next tile's residual rows: HBM -> L2 while this tile is processed
          const int nrow = row_index(wn);
          if (nrow >= 0) {
            const uint8_t* a = reinterpret_cast<const uint8_t*>(res + (size_t)nrow * p.Cout + wn.n_tile * p.BN + cbeg);
            for (int b = 0; b < (cend - cbeg) * (int)sizeof(TO); b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + b));
          }
        }
      }
      __syncwarp();
      ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
      ptx::tc_fence_after();
      if (tr && warp == 4 && lane == 0 && tile - t0 < 16) p.trace[192 + 2 * (tile - t0)] = clock64();
      int fine = 224;
#define CAPF_STAMP() do { if (tr && warp == 4 && lane == 0 && tile == t0 && fine < 256) p.trace[fine++] = clock64(); } while (0)
      uint32_t buf = 0;
      for (int s = 0; s < nslabs; ++s) {
        const int c0 = cbeg + s * SW;
        const int ncol = min(SW, cend - c0);
        if (HAS_RES && s + 1 < nslabs) prefetch_res(s + 1, buf ^ 1u);
#pragma unroll
        for (int g = 0; g < SW; g += 32) {
          if (g < ncol) {
            const bool two = g + 16 < ncol;
            uint32_t a0[16], a1[16];
            ptx::tmem_ld16(taddr + (uint32_t)(c0 + g), a0);
            if (two) ptx::tmem_ld16(taddr + (uint32_t)(c0 + g + 16), a1);
            CAPF_STAMP();
            ptx::tmem_ld_wait();
            CAPF_STAMP();
            if (HAS_RES && g == 0) {                   // this slab's residual has landed (the next one may be in flight)
              if (s + 1 < nslabs) ptx::cp_async_wait_group1(); else ptx::cp_async_wait_all();
              __syncwarp();
            }
            CAPF_STAMP();
            float fl0 = floor_v, fl1 = floor_v;
            if constexpr (SEG) { fl0 = seg_floor(ncol0 + c0 + g); fl1 = seg_floor(ncol0 + c0 + g + 16); }
            epi16<TO, EMODE>(a0, sbias + (c0 - cbeg) + g, fl0, stg_ptr + buf * TC_STG_BYTES + lane * 128, (uint32_t)((g / 16) * CPG), (uint32_t)lane & 7u);
            if (two) epi16<TO, EMODE>(a1, sbias + (c0 - cbeg) + g + 16, fl1, stg_ptr + buf * TC_STG_BYTES + lane * 128, (uint32_t)((g / 16 + 1) * CPG), (uint32_t)lane & 7u);
            CAPF_STAMP();
          }
        }
        if (s + 1 == nslabs) {                        // the accumulator has been read completely
          ptx::tc_fence_before();
          ptx::mbar_arrive(bar_tempty + 8 * acc);     // 256 arrivals free this accumulator stage
        }
        __syncwarp();
        const int chs = (ncol * (int)sizeof(TO)) >> 4;
        if constexpr (SEG) {
          // every 16-byte chunk goes to the tensor of the segment its columns belong to (rows and conv mode alike)
          constexpr int EPC = 16 / (int)sizeof(TO);
          const uint32_t magic = (65536u + (uint32_t)chs - 1u) / (uint32_t)chs;
          for (int i = 0; i < chs; ++i) {
            const int item = i * 32 + lane, r = (int)(((uint32_t)item * magic) >> 16), c = item - r * chs;
            const int grow = __shfl_sync(0xffffffffu, myrow, r);
            const int col = ncol0 + c0 + c * EPC;
            const int sg = seg_of(col);
            TO* const so = reinterpret_cast<TO*>(sg == 0 ? p.seg_out[0] : sg == 1 ? p.seg_out[1] : sg == 2 ? p.seg_out[2] : p.seg_out[3]);
            const int sc = sg == 0 ? p.seg_c[0] : sg == 1 ? p.seg_c[1] : sg == 2 ? p.seg_c[2] : p.seg_c[3];
            const int sb = sg == 0 ? 0 : sg == 1 ? p.seg_beg[1] : sg == 2 ? p.seg_beg[2] : p.seg_beg[3];
            if (grow >= 0)
              st16_hint(reinterpret_cast<uint8_t*>(so + (size_t)grow * sc + (col - sb)), *reinterpret_cast<const uint4*>(stg_ptr + slot_off(buf, r, c)), pol_out);
          }
        } else if (p.mode == 0) {
          // rows mode: the warp's 32 rows are consecutive matrix rows -> affine addresses
          const int m_w0 = w.tx * tile_rows + sub * 128 + q * 32;
          const int rows_live = p.M - m_w0;
          uint8_t* gbase = reinterpret_cast<uint8_t*>(out + (size_t)m_w0 * p.Cout + ncol0 + c0);
          if (chs == 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3), c = lane & 7;
              if (r < rows_live) st16_hint(gbase + (uint32_t)r * row_bytes + 16 * c, *reinterpret_cast<const uint4*>(stg_ptr + slot_off(buf, r, c)), pol_out);
            }
          } else {
            const uint32_t magic = (65536u + (uint32_t)chs - 1u) / (uint32_t)chs;
            for (int i = 0; i < chs; ++i) {
              const int item = i * 32 + lane, r = (int)(((uint32_t)item * magic) >> 16), c = item - r * chs;
              if (r < rows_live) st16_hint(gbase + (uint32_t)r * row_bytes + 16 * c, *reinterpret_cast<const uint4*>(stg_ptr + slot_off(buf, r, c)), pol_out);
            }
          }
        } else {
          const uint32_t magic = (65536u + (uint32_t)chs - 1u) / (uint32_t)chs;
          for (int i = 0; i < chs; ++i) {
            const int item = i * 32 + lane, r = (int)(((uint32_t)item * magic) >> 16), c = item - r * chs;
            const int grow = __shfl_sync(0xffffffffu, myrow, r);
            if (grow >= 0)
              st16_hint(reinterpret_cast<uint8_t*>(out + (size_t)grow * p.Cout + ncol0 + c0) + 16 * c, *reinterpret_cast<const uint4*>(stg_ptr + slot_off(buf, r, c)),
                        pol_out);
          }
        }
        __syncwarp();
        CAPF_STAMP();
        if (HAS_RES) buf ^= 1u;
      }
#undef CAPF_STAMP
      if (nslabs == 0) {                              // BN = 16: the second warp of the quadrant has no columns
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_tempty + 8 * acc);
      }
      if (tr && warp == 4 && lane == 0 && tile - t0 < 16) p.trace[193 + 2 * (tile - t0)] = clock64();
      if (++acc == (uint32_t)p.nacc) { acc = 0; acc_phase ^= 1u; }
      w = wn;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (tr && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[4] = clock64();
    p.trace[5] = (long long)gt;
    p.trace[6] = p.msub; p.trace[7] = p.BN; p.trace[8] = p.num_stages; p.trace[9] = p.num_tiles; p.trace[10] = gridDim.x; p.trace[11] = p.nacc;
  }
}

// =======================================================================================================
// host side: tensor maps + launch geometry
// =======================================================================================================
struct TcConvState {
  TcHaloState* halo = nullptr;   // non-null: the op runs on the halo-band kernel instead of tc_gemm_kernel
  TcHalo128State* h128 = nullptr; // non-null: halo band + streamed weights (C = Cout = 128)
  TcHalo256State* h256 = nullptr; // non-null: halo tile in four planes + streamed weights (C = 256 -> 32)
  Tc2State* two = nullptr;       // non-null: the op runs on the 2-CTA GEMM kernel (capf_tc2.cu)
  TcBlockState* blk = nullptr;   // non-null: a fused BasicBlock op (capf_tc_block.cu)
  TcChainState* chain = nullptr; // non-null: a CAPF_OP_EXPAND_REDUCE op (capf_tc_chain.cu)
  TcMlpState* mlp = nullptr;     // non-null: a CAPF_OP_MLP op (capf_tc_mlp.cu)
  CUtensorMap mapA, mapB, mapA2;
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
  if (op.i[18] != 0 && (op.i[18] != 1 || op.dtype_in != CAPF_BF16)) return 0;     // split operands are bf16 hi|lo planes
  if (op.i[19] != 0) {        // second A matrix (K-concatenation): plain 1x1 / stride 1 rows only
    if (op.i[19] < 0 || op.i[19] % 16 || op.i[18] != 0 || g.KH != 1 || g.KW != 1 || g.stride != 1 || g.pad != 0) return 0;
    if (!op.in[5] || ((uintptr_t)op.in[5] & 15)) return 0;
  }
  if (op.i[20] != 0) {        // output segments: plain 16-bit/fp32 conv, no residual / GELU / split operands / second input
    const int S = op.i[20];
    if (S < 2 || S > 4 || op.in[3] || g.act == CAPF_ACT_GELU || op.i[18] != 0 || op.i[19] != 0 || op.i[14] < 0 || op.i[14] >= (1 << S)) return 0;
    int used = 0;
    for (int sgm = 0; sgm < S; ++sgm) {
      const int wd = sgm + 1 < S ? op.i[21 + sgm] : g.Cout - used;
      if (wd <= 0 || wd % 16) return 0;
      used += wd;
      if (!op.out[sgm] || ((uintptr_t)op.out[sgm] & 15)) return 0;
    }
    if (used != g.Cout) return 0;
  }
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

// Output-pixel box (bw x bh x bn <= max_rows accumulator rows) that wastes the fewest of them.
static void choose_box(const ConvGeo& g, int max_w, int max_rows, int& bw, int& bh, int& bn) {
  double best = -1.0;
  bw = bh = bn = 1;
  for (int w = 1; w <= g.Wo && w <= max_rows && w <= max_w; ++w) {
    for (int h = 1; h <= g.Ho && w * h <= max_rows && h * g.stride <= 256; ++h) {
      int n = max_rows / (w * h);
      if (n > g.N) n = g.N;
      if (n > 256) n = 256;
      if (n < 1) continue;
      double tiles = (double)ceil_div(g.Wo, w) * ceil_div(g.Ho, h) * ceil_div(g.N, n);
      double util = ((double)g.Wo * g.Ho * g.N) / (tiles * (double)max_rows);
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

int tc_blockop_prepare(const capf_op& op, TcConvState** out) {
  *out = nullptr;
  TcConvState* s = new (std::nothrow) TcConvState();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_blockop_prepare: out of host memory");
  int e = tc_block_prepare(op, &s->blk);
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

int tc_chainop_prepare(const capf_op& op, TcConvState** out) {
  *out = nullptr;
  TcConvState* s = new (std::nothrow) TcConvState();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_chainop_prepare: out of host memory");
  int e = tc_chain_prepare(op, &s->chain);
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

int tc_mlpop_prepare(const capf_op& op, TcConvState** out) {
  *out = nullptr;
  TcConvState* s = new (std::nothrow) TcConvState();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_mlpop_prepare: out of host memory");
  int e = tc_mlp_prepare(op, &s->mlp);
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

int tc_conv_prepare(const capf_op& op, TcConvState** out) {
  *out = nullptr;
  int e = tc_get_encoder();
  if (e) return e;
  const ConvGeo g = geo_of(op);
  TcConvState* s = new (std::nothrow) TcConvState();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_conv_prepare: out of host memory");
  const bool split = op.i[18] == 1;             // A = [.., 2 * Cin] bf16 (hi | lo planes), B = [Cout][taps * 3 * Cin], see capf_b200.h
  const int Cin2 = op.i[19];                    // second A matrix [M][Cin2] in in[5]: out = [x | x2] W^T, W = [Cout][Cin + Cin2]
  const int nseg = op.i[20] > 1 ? op.i[20] : 0;      // output segments: per-tap kernel only
  if (!nseg && !split && !Cin2 && op.i[13] == 0 && tc2_supported(op)) {     // wide Linears over many rows: CTA pairs (cta_group::2)
    e = tc2_prepare(op, &s->two);
    if (e) { delete s; return e; }
    *out = s;
    return CAPF_OK;
  }
  // i[13]: kernel variant hint (0 = automatic, 1 = per-tap TMA GEMM, 2 = halo band); tests use it for A/B parity
  if (!nseg && !split && !Cin2 && op.i[13] != 1 && tc_halo256_supported(op)) {
    e = tc_halo256_prepare(op, &s->h256);
    if (e) { delete s; return e; }
    *out = s;
    return CAPF_OK;
  }
  if (!nseg && !split && op.i[13] != 1 && tc_halo128_supported(op)) {
    e = tc_halo128_prepare(op, &s->h128);
    if (e) { delete s; return e; }
    *out = s;
    return CAPF_OK;
  }
  if (!nseg && !split && op.i[13] != 1 && tc_halo_supported(op)) {
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
  p.kb = (g.Cin | Cin2) % 64 == 0 ? 64 : (g.Cin | Cin2) % 32 == 0 ? 32 : 16;
  p.cpt = (split ? 2 : 1) * g.Cin / p.kb;
  p.cpt1 = split ? g.Cin / p.kb : 0;
  p.cps = 64 / p.kb;
  p.num_chunks = g.KH * g.KW * (p.cpt + p.cpt1) + Cin2 / p.kb;
  p.c_a1 = Cin2 ? p.cpt : p.num_chunks;
  if (Cin2) p.cpt = p.num_chunks;               // rows mode: one "tap" that spans both operands
  p.KW = g.KW; p.KH = g.KH; p.stride = g.stride; p.pad = g.pad;
  p.Cout = g.Cout; p.Ho = g.Ho; p.Wo = g.Wo; p.Nimg = g.N;
  p.act = g.act;
  p.bias = (const float*)op.in[2];
  p.res = op.in[3];
  p.out = op.out[0];
  p.trace = (long long*)op.in[4];     // debug only (NULL in every program the host layer builds)
  p.nseg = nseg;
  for (int sgm = 0, used = 0; sgm < 4; ++sgm) {
    const bool live = sgm < nseg;
    const int wd = !live ? 0 : sgm + 1 < nseg ? op.i[21 + sgm] : g.Cout - used;
    p.seg_beg[sgm] = live ? used : 0x7fffffff;
    p.seg_c[sgm] = wd;
    p.seg_out[sgm] = live ? op.out[sgm] : nullptr;
    used += wd;
  }
  p.seg_noact = nseg ? (uint32_t)op.i[14] : 0u;
  p.M = g.N * g.Ho * g.Wo;
  const int swz = p.kb * 2;
  const int K = g.KH * g.KW * g.Cin * (split ? 3 : 1) + Cin2;      // GEMM depth (3x with split operands: hi*Wh + lo*Wh + hi*Wl)
  const int Ca = g.Cin * (split ? 2 : 1);                   // channels of the A tensor in memory
  const int osz = op.dtype_out == CAPF_F32 ? 4 : 2;

  // ---- epilogue staging: two tiles per warp with a residual (one being filled by cp.async), else one ------------
  p.stg_bufs = p.res ? 2 : 1;
  const int stg_bytes = TC_EPI_WARPS * p.stg_bufs * TC_STG_BYTES + TC_EPI_WARPS * 256 * 4;   // staging tiles + bias slices

  // ---- tile shape: (msub x 128 rows) x BN columns ------------------------------------------------------------
  // Cycle model per tile (SM clocks): the tensor pipe (N / 2 cycles per 128 x N x 16 MMA, or the shared-memory operand
  // read of (128 + N) * 32 B at 128 B/clk when that is slower), the L2 -> SM operand feed (the chip-wide L2 cap is
  // ~6300 B/clk = 42 B/clk/SM: every stage moves (rows + BN) * 128 B), and the epilogue's global traffic.  Two
  // sub-tiles share every B stage, which is what makes the wide transformer GEMMs (QKV: 256 x 240 tiles, one per SM)
  // feed less per FLOP.  i[15] forces msub, i[16] forces BN (tests / experiments).
  int msub_lo = 1, msub_hi = 2;
  { const char* ev = getenv("CAPF_TC_MSUB"); if (ev && (ev[0] == '1' || ev[0] == '2')) msub_lo = msub_hi = ev[0] - '0'; }
  if (op.i[15] == 1 || op.i[15] == 2) msub_lo = msub_hi = op.i[15];
  double best_cost = 1e300;
  int best_msub = 0, best_bn = 0, best_bw = 0, best_bh = 0, best_bnn = 0;
  for (int msub = msub_lo; msub <= msub_hi; ++msub) {
    int bw = 128 * msub, bh = 1, bnn = 1, box_rows = 128 * msub;
    long long m_tiles;
    if (rows) {
      m_tiles = ceil_div(p.M, 128 * msub);
    } else {
      choose_box(g, 256 / g.stride, 128 * msub, bw, bh, bnn);
      box_rows = bw * bh * bnn;
      if (msub == 2 && box_rows <= 128 && msub_lo != 2) continue;
      m_tiles = (long long)ceil_div(g.Wo, bw) * ceil_div(g.Ho, bh) * ceil_div(g.N, bnn);
    }
    for (int bn = 16; bn <= 256 && bn <= g.Cout; bn += 16) {
      if (g.Cout % bn) continue;
      if (op.i[16] > 0 && bn != op.i[16]) continue;
      if (msub * bn > 512) continue;
      const int stage_bytes = msub * TC_A_SUB_BYTES + bn * 128;
      int stages = (TC_SMEM_LIMIT - TC_HEADER_BYTES - 1024 - stg_bytes) / stage_bytes;
      if (stages < 2) continue;
      const int nacc = 2 * msub * bn <= 512 ? 2 : 1;
      const long long tiles = m_tiles * (g.Cout / bn);
      const long long grid = tiles < num_sms() ? tiles : num_sms();
      const double per_cta = (double)((tiles + grid - 1) / grid);
      const double k16 = K / 16.0;
      const double t_mma = msub * k16 * ((bn / 2.0) > ((128 + bn) / 4.0) ? (bn / 2.0) : ((128 + bn) / 4.0));
      const double t_feed = (double)(box_rows + bn) * K * 2.0 / 42.0;
      const double t_epi = (double)msub * 128.0 * bn * osz * (p.res ? 2.0 : 1.0) / 24.0 + 300.0;
      double t_main = t_mma > t_feed ? t_mma : t_feed;
      double cost;
      if (nacc == 2) cost = per_cta * (t_main > t_epi ? t_main : t_epi) + (t_main < t_epi ? t_main : t_epi);
      else cost = per_cta * (t_main + t_epi);
      if (stages < 3) cost *= 1.15;
      cost += 600.0 * per_cta;    // per-tile pipeline bubbles (barrier round trips, accumulator hand-over)
      if (cost < best_cost * (1.0 - 1e-9)) {
        best_cost = cost; best_msub = msub; best_bn = bn; best_bw = bw; best_bh = bh; best_bnn = bnn;
      }
    }
  }
  if (!best_bn) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "tc conv: no tile shape fits"); }
  p.msub = best_msub;
  p.BN = best_bn;
  long long m_tiles;
  if (rows) {
    p.bw = 128 * p.msub; p.bh = 1; p.bn = 1; p.tiles_x = ceil_div(p.M, 128 * p.msub); p.tiles_y = 1;
    m_tiles = p.tiles_x;
  } else {
    p.bw = best_bw; p.bh = best_bh; p.bn = best_bnn;
    p.tiles_x = ceil_div(g.Wo, p.bw);
    p.tiles_y = ceil_div(g.Ho, p.bh);
    m_tiles = (long long)p.tiles_x * p.tiles_y * ceil_div(g.N, p.bn);
  }
  p.n_tiles_n = g.Cout / p.BN;
  long long nt = m_tiles * p.n_tiles_n;
  if (nt >= (1ll << 31)) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "tc conv: too many tiles"); }
  p.num_tiles = (int)nt;

  p.a_sub_bytes = 128 * p.kb * 2;
  p.a_chunk_bytes = p.msub * p.a_sub_bytes;
  p.a_stage_bytes = p.msub * TC_A_SUB_BYTES;
  p.b_chunk_bytes = p.BN * p.kb * 2;
  p.stage_bytes = p.a_stage_bytes + p.BN * 128;
  const int box_rows = rows ? 128 * p.msub : p.bw * p.bh * p.bn;
  p.tx_bytes_per_chunk = (box_rows + p.BN) * p.kb * 2;
  int stages = (TC_SMEM_LIMIT - TC_HEADER_BYTES - 1024 - stg_bytes) / p.stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  const int k_stages = ceil_div(p.num_chunks, p.cps);
  if (stages > 2 * k_stages && 2 * k_stages >= 2) stages = 2 * k_stages;   // no point in a ring deeper than two tiles
  if (stages < 2) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "tc conv: tile does not fit shared memory"); }
  p.num_stages = stages;
  p.stg_off = (uint32_t)(stages * p.stage_bytes);
  p.bias_off = p.stg_off + (uint32_t)(TC_EPI_WARPS * p.stg_bufs * TC_STG_BYTES);
  s->smem_bytes = TC_HEADER_BYTES + 1024 + stages * p.stage_bytes + stg_bytes;
  if (s->smem_bytes < 120 * 1024) s->smem_bytes = 120 * 1024;   // one CTA per SM: the CTA owns the SM's TMEM columns
  p.nacc = 2 * p.msub * p.BN <= 512 ? 2 : 1;
  p.acc_stride = p.msub * p.BN;
  int cols = 32;
  while (cols < p.nacc * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  p.idesc = tc_idesc(op.dtype_in == CAPF_BF16, p.BN);
  p.desc_hi = tc_desc_hi(swz, 8 * swz);
  s->grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  s->dtype_out = op.dtype_out;

  // ---- tensor maps -------------------------------------------------------------------------------------
  const CUtensorMapDataType dt = op.dtype_in == CAPF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  if (rows) {
    cuuint64_t dims[2] = {(cuuint64_t)Ca, (cuuint64_t)p.M};
    cuuint64_t strides[1] = {(cuuint64_t)Ca * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kb, (cuuint32_t)(128 * p.msub)};
    cuuint32_t es[2] = {1, 1};
    e = tc_encode_map(&s->mapA, dt, 2, op.in[0], dims, strides, box, es, swz, "A rows");
    s->mapA2 = s->mapA;
    if (!e && Cin2) {
      cuuint64_t dims2[2] = {(cuuint64_t)Cin2, (cuuint64_t)p.M};
      cuuint64_t strides2[1] = {(cuuint64_t)Cin2 * 2};
      e = tc_encode_map(&s->mapA2, dt, 2, op.in[5], dims2, strides2, box, es, swz, "A2 rows");
    }
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)Ca, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)Ca * 2, (cuuint64_t)g.W * Ca * 2, (cuuint64_t)g.H * g.W * Ca * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.kb, (cuuint32_t)(p.bw * g.stride), (cuuint32_t)(p.bh * g.stride), (cuuint32_t)p.bn};
    cuuint32_t es[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
    e = tc_encode_map(&s->mapA, dt, 4, op.in[0], dims, strides, box, es, swz, "A conv");
    s->mapA2 = s->mapA;
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

template <typename TO, int MODE>
static int tc_launch_mode(const TcConvState* s, cudaStream_t st) {
  static PerDevice<int> max_smem_;   // opt-in once per instantiation and device (outside graph capture: Plan.capture warms up first)
  std::atomic<int>& max_smem = max_smem_.get();
  if (s->smem_bytes > max_smem) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<TO, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_gemm_kernel smem opt-in: %s", cudaGetErrorString(e));
    max_smem = TC_SMEM_LIMIT;
  }
  launch_k(tc_gemm_kernel<TO, MODE>, dim3(s->grid), dim3(TC_THREADS), s->smem_bytes, st, s->mapA, s->mapB, s->mapA2, s->p);
  return check_launch("tc_gemm_kernel");
}

template <typename TO>
static int tc_launch_typed(const TcConvState* s, cudaStream_t st) {
  const int mode = (s->p.res ? 1 : 0) | (s->p.act == CAPF_ACT_GELU ? 2 : 0) | (s->p.nseg ? 4 : 0);
  switch (mode) {
    case 4: return tc_launch_mode<TO, 4>(s, st);
    case 0: return tc_launch_mode<TO, 0>(s, st);
    case 1: return tc_launch_mode<TO, 1>(s, st);
    case 2: return tc_launch_mode<TO, 2>(s, st);
    default: return tc_launch_mode<TO, 3>(s, st);
  }
}

int tc_conv_launch(const capf_op&, const TcConvState* s, cudaStream_t st) {
  if (!s) return set_error(CAPF_ERR_ARG, "tc conv: op was not prepared");
  if (s->blk) return tc_block_launch(s->blk, st);
  if (s->chain) return tc_chain_launch(s->chain, st);
  if (s->mlp) return tc_mlp_launch(s->mlp, st);
  if (s->two) return tc2_launch(s->two, st);
  if (s->halo) return tc_halo_launch(s->halo, st);
  if (s->h128) return tc_halo128_launch(s->h128, st);
  if (s->h256) return tc_halo256_launch(s->h256, st);
  switch (s->dtype_out) {
    case CAPF_F32: return tc_launch_typed<float>(s, st);
    case CAPF_F16: return tc_launch_typed<__half>(s, st);
    case CAPF_BF16: return tc_launch_typed<__nv_bfloat16>(s, st);
    default: return set_error(CAPF_ERR_UNSUPPORTED, "tc conv: dtype_out");
  }
}

void tc_conv_describe(const TcConvState* s, char* buf, int cap) {
  if (!s) { snprintf(buf, cap, "?"); return; }
  if (s->blk) { tc_block_describe(s->blk, buf, cap); return; }
  if (s->chain) { tc_chain_describe(s->chain, buf, cap); return; }
  if (s->mlp) { tc_mlp_describe(s->mlp, buf, cap); return; }
  if (s->two) { tc2_describe(s->two, buf, cap); return; }
  if (s->halo) { tc_halo_describe(s->halo, buf, cap); return; }
  if (s->h128) { tc_halo128_describe(s->h128, buf, cap); return; }
  if (s->h256) { tc_halo256_describe(s->h256, buf, cap); return; }
  snprintf(buf, cap, "tc_gemm_kernel[%dx%d tile, %d stages%s%s%s]", 128 * s->p.msub, s->p.BN, s->p.num_stages, s->p.cpt1 ? ", bf16x3 split operands" : "",
           s->p.c_a1 < s->p.num_chunks ? ", two A operands" : "", s->p.nseg ? ", output segments" : "");
}

void tc_conv_release(TcConvState* s) {
  if (s && s->blk) tc_block_release(s->blk);
  if (s && s->chain) tc_chain_release(s->chain);
  if (s && s->mlp) tc_mlp_release(s->mlp);
  if (s && s->two) tc2_release(s->two);
  if (s && s->halo) tc_halo_release(s->halo);
  if (s && s->h128) tc_halo128_release(s->h128);
  if (s && s->h256) tc_halo256_release(s->h256);
  delete s;
}

}  // namespace capf
