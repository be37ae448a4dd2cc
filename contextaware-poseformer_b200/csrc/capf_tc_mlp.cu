// CAPF_OP_MLP: the Mlp of a 128-wide transformer block (pose_dformer.py:25-31 inside DeformableBlock :138-141 and Block :78) as ONE
// kernel over tiles of 128 token rows:
//     h = gelu(t . W1^T + b1)            fc1, 128 -> 256, GELU (erf form)           -> 16-bit, stays in shared memory
//     X = X + h . W2^T + b2              fc2, 256 -> 128, residual add in fp32      -> the fp32 token stream, in place
// As two launches the hidden tensor (rows x 256) is written by one GEMM kernel and read back by the next, and each of the two pays
// the fixed cost of a tcgen05 kernel (barrier / TMEM / tensor-map set-up, pipeline fill, tail) for ~1 us of tensor work: 8 such
// pairs per forward (4 context blocks, 4 res blocks).  Here the first epilogue writes h straight into the K-major swizzled layout
// the tensor pipe reads (as capf_tc_chain.cu does for its y) and the second GEMM consumes it there.
//
// Shared memory is exactly the 227 KB of an SM: W (64 KB, TIME-SHARED: W1 for GEMM 1, then W2 streams in while epilogue 1 runs the
// GELUs, then W1 of the next tile while epilogue 2 runs) | T[2] (2 x 32 KB, the t tile: two 64-column chunks; next tile prefetched) |
// H (64 KB, four 64-column chunks of h) | staging (8 warps x 4 KB: residual in by cp.async, result out with 16-byte
// coalesced stores, the epilogue of capf_tc.cu).  TMEM: GEMM 1 = two 128-column halves (epilogue 1 of half 0 overlaps the MMAs of
// half 1), GEMM 2 = 128 columns.
// Results are bit-identical to the two CAPF_OP_CONV2D ops (same K order, same epilogue arithmetic): tests/test_mlp.py.
//
// Roles (640 threads): warp 0 TMA loads, warp 1 MMA issuer, warp 2 TMEM, warps 4-19 = SIXTEEN epilogue warps (lane quadrant x column
// quarter).  All sixteen run epilogue 1 -- 32 K erff evaluations per tile are what bounds the kernel, the first version (8 warps) spent
// ~5 us per tile there --; the warps of column quarters 0 and 1 then also run epilogue 2 (two 32-column fp32 slabs each).
#include <cstdio>
#include <cstdlib>
#include <new>

#include "capf_tc.cuh"

namespace capf {

constexpr int ML_THREADS = 640;
constexpr int ML_K1 = 128, ML_N1 = 256, ML_N2 = 128;
constexpr int ML_CHUNK = 128 * 128;                 // 128 rows x 64 elements x 2 B
constexpr int ML_T_BYTES = 2 * ML_CHUNK;            // t tile: K1 = two 64-column chunks
constexpr int ML_H_BYTES = 4 * ML_CHUNK;            // h tile: N1 = four 64-column chunks
constexpr int ML_W_BYTES = 65536;                   // W1 [256][128] as two K chunks of [256][64]; W2 [128][256] as four K chunks of [128][64]
constexpr int ML_STG_BYTES = 32 * 128;              // staging tile of one warp: 32 rows x 32 fp32 columns
constexpr int ML_HEADER = 2048;                     // barriers, TMEM slot, then b1[256] | b2[128] f32
constexpr int MB_TFULL = 0, MB_TFREE = 16, MB_W1FULL = 32, MB_W2FULL = 40, MB_G1DONE = 48, MB_G2DONE = 56, MB_A1FULL = 64, MB_A1EMPTY = 80,
              MB_HREADY = 96, MB_A2FULL = 104, MB_A2EMPTY = 112, MB_TMEM = 120;
constexpr int MB_BIAS = 512;                        // 384 floats = 1536 B
constexpr int ML_SMEM = 1024 + ML_HEADER + ML_W_BYTES + 2 * ML_T_BYTES + ML_H_BYTES + 8 * ML_STG_BYTES;
static_assert(ML_SMEM <= TC_SMEM_LIMIT, "fused MLP kernel: shared memory budget");

struct MlpP {
  int M, num_tiles;
  uint32_t idesc, desc_hi;
  const float* b1;
  const float* b2;
  const float* res;
  float* out;
};

__device__ __forceinline__ uint32_t ml_chunk(uint32_t row, uint32_t c) { return row * 128u + ((c ^ (row & 7u)) << 4); }

template <typename T>
__global__ void __launch_bounds__(ML_THREADS, 1)
tc_mlp128_kernel(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2, const MlpP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_tfull = base + MB_TFULL, bar_tfree = base + MB_TFREE, bar_w1full = base + MB_W1FULL, bar_w2full = base + MB_W2FULL;
  const uint32_t bar_g1done = base + MB_G1DONE, bar_g2done = base + MB_G2DONE, bar_a1full = base + MB_A1FULL, bar_a1empty = base + MB_A1EMPTY;
  const uint32_t bar_hready = base + MB_HREADY, bar_a2full = base + MB_A2FULL, bar_a2empty = base + MB_A2EMPTY;
  const uint32_t smem_w = base + ML_HEADER;
  const uint32_t smem_t = smem_w + ML_W_BYTES;                                  // T[0], T[1]
  const uint32_t smem_h = smem_t + 2 * ML_T_BYTES;
  const uint32_t smem_stg = smem_h + ML_H_BYTES;
  uint8_t* const gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + MB_TMEM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapT); ptx::prefetch_tmap(&mapW1); ptx::prefetch_tmap(&mapW2);
  }
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_tfull + 8 * b, 1);
      ptx::mbar_init(bar_tfree + 8 * b, 1);
      ptx::mbar_init(bar_a1full + 8 * b, 1);           // b = column half of GEMM 1
      ptx::mbar_init(bar_a1empty + 8 * b, 8);          // the eight epilogue-1 warps of the half
    }
    ptx::mbar_init(bar_w1full, 1);
    ptx::mbar_init(bar_w2full, 1);
    ptx::mbar_init(bar_g1done, 1);
    ptx::mbar_init(bar_g2done, 1);
    ptx::mbar_init(bar_hready, 16);                    // all sixteen epilogue-1 warps
    ptx::mbar_init(bar_a2full, 1);
    ptx::mbar_init(bar_a2empty, 8);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(base + MB_TMEM, 512u);
    ptx::tmem_relinquish();
  }
  if (warp >= 4) {                                      // biases (constant data): b1[256] | b2[128]
    const int i = threadIdx.x - 128;
    if (i < ML_N1 + ML_N2) {
      const float v = i < ML_N1 ? (p.b1 ? __ldg(p.b1 + i) : 0.f) : (p.b2 ? __ldg(p.b2 + i - ML_N1) : 0.f);
      reinterpret_cast<float*>(gen + MB_BIAS)[i] = v;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();
  if (warp != 0) pdl_wait();

  const int t0 = (int)(((long long)p.num_tiles * blockIdx.x) / gridDim.x);
  const int t1 = (int)(((long long)p.num_tiles * (blockIdx.x + 1)) / gridDim.x);
  const int ntiles = t1 - t0;
  const float* const sbias = reinterpret_cast<const float*>(gen + MB_BIAS);

  if (warp == 0) {
    // ===================================== TMA loads =========================================
    if (ptx::elect_one() && ntiles > 0) {
      auto load_w1 = [&]() {
        ptx::mbar_arrive_expect_tx(bar_w1full, (uint32_t)ML_W_BYTES);
        for (int c = 0; c < 2; ++c) ptx::tma_load_2d(&mapW1, bar_w1full, smem_w + c * (ML_N1 * 128), c * 64, 0);
      };
      auto load_t = [&](int k) {
        const uint32_t b = (uint32_t)k & 1u;
        ptx::mbar_arrive_expect_tx(bar_tfull + 8 * b, (uint32_t)ML_T_BYTES);
        for (int c = 0; c < 2; ++c) ptx::tma_load_2d(&mapT, bar_tfull + 8 * b, smem_t + b * ML_T_BYTES + c * ML_CHUNK, c * 64, (t0 + k) * 128);
      };
      load_w1();                                           // constant data: before the dependency wait
      pdl_wait();
      load_t(0);
      for (int k = 0; k < ntiles; ++k) {
        const uint32_t ph = (uint32_t)k & 1u;
        ptx::mbar_wait(bar_g1done, ph);                    // GEMM 1 of tile k has read W1 (and T[k & 1]): W2 takes its place
        ptx::mbar_arrive_expect_tx(bar_w2full, (uint32_t)ML_W_BYTES);
        for (int c = 0; c < 4; ++c) ptx::tma_load_2d(&mapW2, bar_w2full, smem_w + c * (ML_N2 * 128), c * 64, 0);
        if (k + 1 < ntiles) {
          if (k >= 1) ptx::mbar_wait(bar_tfree + 8 * (((uint32_t)k + 1u) & 1u), (((uint32_t)k - 1u) >> 1) & 1u);   // GEMM 1 of tile k - 1 done with that buffer
          load_t(k + 1);
          ptx::mbar_wait(bar_g2done, ph);                  // GEMM 2 of tile k has read W2: W1 of the next tile
          load_w1();
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (one elected thread) ====================
    if (ptx::elect_one()) {
      const uint32_t w_lo = tc_desc_lo(smem_w, 1u), h_lo = tc_desc_lo(smem_h, 1u);
      for (int k = 0; k < ntiles; ++k) {
        const uint32_t b = (uint32_t)k & 1u, ph = (uint32_t)k & 1u, tph = ((uint32_t)k >> 1) & 1u;
        // ---- GEMM 1: two column halves of 128, K = 128 (two 64-deep chunks)
        ptx::mbar_wait(bar_w1full, ph);
        ptx::mbar_wait(bar_tfull + 8 * b, tph);
        ptx::tc_fence_after();
        const uint32_t t_lo = tc_desc_lo(smem_t + b * ML_T_BYTES, 1u);
        for (int h = 0; h < 2; ++h) {
          ptx::mbar_wait(bar_a1empty + 8 * h, ph ^ 1u);
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16_lohi(tmem_base + (uint32_t)(128 * h), t_lo + (uint32_t)c * (ML_CHUNK >> 4) + 2u * kk, p.desc_hi,
                                 w_lo + (uint32_t)((c * (ML_N1 * 128) + h * ML_CHUNK) >> 4) + 2u * kk, p.desc_hi, p.idesc, (c | kk) ? 1u : 0u);
          }
          ptx::umma_commit(bar_a1full + 8 * h);
        }
        ptx::umma_commit(bar_tfree + 8 * b);
        ptx::umma_commit(bar_g1done);
        // ---- GEMM 2: K = 256 (four 64-deep chunks of h), N = 128
        ptx::mbar_wait(bar_w2full, ph);
        ptx::mbar_wait(bar_hready, ph);
        ptx::mbar_wait(bar_a2empty, ph ^ 1u);
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ptx::umma_f16_lohi(tmem_base + 256u, h_lo + (uint32_t)c * (ML_CHUNK >> 4) + 2u * kk, p.desc_hi,
                               w_lo + (uint32_t)c * ((ML_N2 * 128) >> 4) + 2u * kk, p.desc_hi, p.idesc, (c | kk) ? 1u : 0u);
        }
        ptx::umma_commit(bar_a2full);
        ptx::umma_commit(bar_g2done);
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue warps: (lane quadrant q, column quarter cq) ===================================
    const int q = warp & 3, cq = (warp - 4) >> 2;            // rows 32 q .. 32 q + 31
    const int h = cq >> 1;                                   // epilogue 1: columns 64 cq .. 64 cq + 63 of h = chunk cq, accumulator half h
    const uint32_t row = (uint32_t)(q * 32 + lane);
    uint8_t* const hb = gen + (smem_h - base);
    // epilogue 2 (cq < 2): fp32 slabs 2 cq and 2 cq + 1 (32 columns each) through one private staging tile
    const uint32_t stg = smem_stg + (uint32_t)(q + 4 * (cq & 1)) * (uint32_t)ML_STG_BYTES;
    uint8_t* const stg_ptr = gen + (stg - base);
    auto slot_off = [](int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); };
    const uint64_t pol_out = ptx::policy_evict_last();
    const float ninf = -__int_as_float(0x7f800000);
    for (int k = 0; k < ntiles; ++k) {
      const uint32_t ph = (uint32_t)k & 1u;
      const int m_w0 = (t0 + k) * 128 + q * 32;                // first token row of this warp
      const int rows_live = p.M - m_w0;
      auto prefetch_res = [&](int slab) {
        const uint8_t* gbase = reinterpret_cast<const uint8_t*>(p.res + (size_t)m_w0 * ML_N2 + 32 * slab);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3), c = lane & 7;
          if (r < rows_live) ptx::cp_async16(stg + slot_off(r, c), gbase + (size_t)r * (ML_N2 * 4) + 16 * c);
        }
        ptx::cp_async_commit();
      };
      if (cq < 2) prefetch_res(2 * cq);                        // residual of the first slab: independent of everything on chip
      // ---- epilogue 1: h = gelu(acc + b1) -> H (swizzled operand layout), 64 columns
      ptx::mbar_wait(bar_a1full + 8 * h, ph);                  // every MMA issued before it -- GEMM 2 of the previous tile too -- is complete
      ptx::tc_fence_after();
      const uint32_t taddr1 = tmem_base + (uint32_t)(64 * cq) + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int g = 0; g < 4; g += 2) {                         // 32 columns per step
        uint32_t a0[16], a1[16];
        ptx::tmem_ld16(taddr1 + (uint32_t)(16 * g), a0);
        ptx::tmem_ld16(taddr1 + (uint32_t)(16 * g + 16), a1);
        ptx::tmem_ld_wait();
        if (g == 2) {                                          // this warp's part of the accumulator half has been read
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_a1empty + 8 * h);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {                          // 8 columns = one 16-byte chunk
          const int col = 16 * g + 8 * c;                      // column inside the quarter = inside chunk cq
          const float4 bA = *reinterpret_cast<const float4*>(sbias + 64 * cq + col), bB = *reinterpret_cast<const float4*>(sbias + 64 * cq + col + 4);
          const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = gelu_erf(__uint_as_float(c < 2 ? a0[8 * c + e] : a1[8 * (c - 2) + e]) + bb[e]);
          *reinterpret_cast<uint4*>(hb + (uint32_t)cq * ML_CHUNK + ml_chunk(row, (uint32_t)(col >> 3))) = pack8<T>(f);
        }
      }
      ptx::fence_proxy_async();                                // generic-proxy writes of H -> tensor-pipe reads
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_hready);
      if (cq >= 2) continue;
      // ---- epilogue 2: X = X + acc2 + b2 (fp32), two slabs of 32 columns through the staging tile
      ptx::mbar_wait(bar_a2full, ph);
      ptx::tc_fence_after();
      const uint32_t taddr2 = tmem_base + 256u + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int sl = 0; sl < 2; ++sl) {
        const int slab = 2 * cq + sl;
        uint32_t a0[16], a1[16];
        ptx::tmem_ld16(taddr2 + (uint32_t)(32 * slab), a0);
        ptx::tmem_ld16(taddr2 + (uint32_t)(32 * slab + 16), a1);
        ptx::tmem_ld_wait();
        ptx::cp_async_wait_all();
        __syncwarp();
        epi16<float, 1>(a0, sbias + ML_N1 + 32 * slab, ninf, stg_ptr + lane * 128, 0u, (uint32_t)lane & 7u);
        epi16<float, 1>(a1, sbias + ML_N1 + 32 * slab + 16, ninf, stg_ptr + lane * 128, 4u, (uint32_t)lane & 7u);
        if (sl == 1) {                                         // this warp's part of the accumulator has been read
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_a2empty);
        }
        __syncwarp();
        uint8_t* gout = reinterpret_cast<uint8_t*>(p.out + (size_t)m_w0 * ML_N2 + 32 * slab);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3), c = lane & 7;
          if (r < rows_live) st16_hint(gout + (size_t)r * (ML_N2 * 4) + 16 * c, *reinterpret_cast<const uint4*>(stg_ptr + slot_off(r, c)), pol_out);
        }
        __syncwarp();
        if (sl == 0) prefetch_res(slab + 1);                   // the staging tile is free again: residual of the second slab
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512u);
}

// =======================================================================================================
// host side
// =======================================================================================================
struct TcMlpState {
  CUtensorMap mapT, mapW1, mapW2;
  MlpP p;
  int grid, dtype;
};

int tc_mlp_supported(const capf_op& op) {
  if (op.kind != CAPF_OP_MLP) return 0;
  if (op.i[0] <= 0 || op.i[1] != ML_K1 || op.i[2] != ML_N1 || op.i[3] != ML_N2) return 0;
  if ((op.dtype_in != CAPF_F16 && op.dtype_in != CAPF_BF16) || op.dtype_out != CAPF_F32) return 0;
  if ((long long)op.i[0] * ML_N1 >= (1ll << 31)) return 0;
  if (!op.in[0] || !op.in[1] || !op.in[3] || !op.in[4] || !op.out[0]) return 0;
  if (((uintptr_t)op.in[0] | (uintptr_t)op.in[1] | (uintptr_t)op.in[3] | (uintptr_t)op.in[4] | (uintptr_t)op.out[0]) & 15) return 0;
  return 1;
}

int tc_mlp_prepare(const capf_op& op, TcMlpState** out) {
  *out = nullptr;
  int e = tc_get_encoder();
  if (e) return e;
  if (!tc_mlp_supported(op)) return set_error(CAPF_ERR_UNSUPPORTED, "fused MLP: shape / dtype not supported (128 -> 256 -> 128, 16-bit operands, fp32 stream)");
  TcMlpState* s = new (std::nothrow) TcMlpState();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_mlp_prepare: out of host memory");
  MlpP& p = s->p;
  memset(&p, 0, sizeof(p));
  p.M = op.i[0];
  p.num_tiles = (p.M + 127) / 128;
  const bool bf16 = op.dtype_in == CAPF_BF16;
  p.idesc = tc_idesc(bf16, 128);
  p.desc_hi = tc_desc_hi(128, 1024);
  p.b1 = (const float*)op.in[2];
  p.b2 = (const float*)op.in[5];
  p.res = (const float*)op.in[3];
  p.out = (float*)op.out[0];
  s->grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  s->dtype = op.dtype_in;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  auto rows_map = [&](CUtensorMap* m, const void* ptr, int cols, int rows, int box_rows, const char* what) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return tc_encode_map(m, dt, 2, ptr, dims, strides, box, es, 128, what);
  };
  e = rows_map(&s->mapT, op.in[0], ML_K1, p.M, 128, "t tile (fused MLP)");
  if (!e) e = rows_map(&s->mapW1, op.in[1], ML_K1, ML_N1, ML_N1, "W1 (fused MLP)");
  if (!e) e = rows_map(&s->mapW2, op.in[4], ML_N1, ML_N2, ML_N2, "W2 (fused MLP)");
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

template <typename T>
static int mlp_launch_typed(const TcMlpState* s, cudaStream_t st) {
  static PerDevice<bool> opted_;
  std::atomic<bool>& opted = opted_.get();
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(tc_mlp128_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_mlp128_kernel smem opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  launch_k(tc_mlp128_kernel<T>, dim3(s->grid), dim3(ML_THREADS), ML_SMEM, st, s->mapT, s->mapW1, s->mapW2, s->p);
  return check_launch("tc_mlp128_kernel");
}

int tc_mlp_launch(const TcMlpState* s, cudaStream_t st) {
  return s->dtype == CAPF_F16 ? mlp_launch_typed<__half>(s, st) : mlp_launch_typed<__nv_bfloat16>(s, st);
}

void tc_mlp_release(TcMlpState* s) { delete s; }

void tc_mlp_describe(const TcMlpState* s, char* buf, int cap) {
  snprintf(buf, cap, "tc_mlp128_kernel[fc1 + GELU + fc2 + residual, 128-row tiles, %d tiles]", s->p.num_tiles);
}

}  // namespace capf
