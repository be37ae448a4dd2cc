// CAPF_OP_EXPAND_REDUCE: the tail of one Bottleneck and the head of the next as ONE kernel (pose_hrnet.py:116-136, layer1 :421-427;
// networks/resnet.py Bottleneck):
//     y = relu(t . W3^T + b3 + x)        conv3 (1x1, 64 -> 256) + bn3 + residual + ReLU of block i      -> written to HBM (block i + 1's residual)
//     u = relu(y . W1^T + b1)            conv1 (1x1, 256 -> 64) + bn1 + ReLU of block i + 1              -> written to HBM
// Both convolutions are per-pixel, so a tile of 128 pixels never needs a neighbour.  As two launches the 256-channel tensor y
// (537 MB at bs = 256) is written by the first kernel and read back by the second, which then is nothing but that read (5.8 TB/s);
// here y goes from the first epilogue into shared memory in the K-major swizzled layout the tensor pipe reads, the second
// GEMM consumes it there, and the same bytes leave for HBM with TMA tensor stores.  HBM per pixel: 128 + 512 in, 512 + 128 out.
//
// Per tile of 128 pixels (two tiles in flight, buffers b = tile & 1):
//   T[b]  16 KB  t tile (TMA), later the staging tile of u          Y[b]  64 KB  four 64-channel chunks: the residual x (TMA), overwritten
//                                                                                in place by y (epilogue 1) = A operand of GEMM 2 + source of the y store
//   GEMM 1: N = 256 as two halves of 128 columns with their own accumulators, so epilogue 1 of the first half overlaps the MMAs of
//           the second and the next tile's GEMM 1 starts as soon as a half has been read;  GEMM 2: K = 256, N = 64, accumulator per buffer.
// Roles (512 threads): warp 0 TMA loads, warp 1 MMA issuer (GEMM 1 of tile k, then GEMM 2 of tile k - 1), warp 2 TMEM, warp 3 TMA
// stores, warps 4-11 epilogue 1 (two groups = the two column halves), warps 12-15 epilogue 2.
#include <cstdio>
#include <cstdlib>
#include <new>

#include "capf_tc.cuh"

namespace capf {

constexpr int CH_THREADS = 512;
constexpr int CH_K1 = 64, CH_N1 = 256, CH_N2 = 64;
constexpr int CH_T_BYTES = 128 * 128;               // 128 pixels x 64 channels x 2 B
constexpr int CH_Y_BYTES = 4 * CH_T_BYTES;          // 128 pixels x 256 channels
constexpr int CH_W3_BYTES = CH_N1 * 128;            // [256][64]
constexpr int CH_W1_BYTES = 4 * CH_N2 * 128;        // four K chunks of [64][64]
constexpr int CH_HEADER = 2048;                     // barriers, TMEM slot, then the two bias vectors
// header: barriers (8 B each) then the two bias vectors
constexpr int CB_W = 0, CB_LDFULL = 8, CB_BUFFREE = 24, CB_A1FULL = 40, CB_A1EMPTY = 56, CB_YREADY = 72, CB_A2FULL = 88, CB_A2EMPTY = 104,
              CB_UREADY = 120, CB_TMEM = 136;
constexpr int CB_BIAS = 256;                        // b3[256] | b1[64] f32 = 1280 B

struct ChainP {
  int P, num_tiles;
  uint32_t idesc1, idesc2, desc_hi;
  const float* b3;
  const float* b1;
};

__device__ __forceinline__ uint32_t ch_chunk(uint32_t row, uint32_t c) { return row * 128u + ((c ^ (row & 7u)) << 4); }

template <typename T>
__global__ void __launch_bounds__(CH_THREADS, 1)
tc_expand_reduce_kernel(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapW3,
                        const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapU,
                        const ChainP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_w = base + CB_W, bar_ldfull = base + CB_LDFULL, bar_buffree = base + CB_BUFFREE, bar_a1full = base + CB_A1FULL;
  const uint32_t bar_a1empty = base + CB_A1EMPTY, bar_yready = base + CB_YREADY, bar_a2full = base + CB_A2FULL, bar_a2empty = base + CB_A2EMPTY;
  const uint32_t bar_uready = base + CB_UREADY, tmem_slot = base + CB_TMEM;
  const uint32_t smem_w3 = base + CH_HEADER, smem_w1 = smem_w3 + CH_W3_BYTES;
  const uint32_t smem_t = smem_w1 + CH_W1_BYTES;                               // T[0], T[1]
  const uint32_t smem_y = smem_t + 2 * CH_T_BYTES;                             // Y[0], Y[1]
  uint8_t* const gen = smem_raw + (base - raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + CB_TMEM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapT); ptx::prefetch_tmap(&mapR); ptx::prefetch_tmap(&mapW3);
    ptx::prefetch_tmap(&mapW1); ptx::prefetch_tmap(&mapY); ptx::prefetch_tmap(&mapU);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(bar_w, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_ldfull + 8 * b, 1);
      ptx::mbar_init(bar_buffree + 8 * b, 1);          // the store warp: both stores of the tile have read T[b] / Y[b]
      ptx::mbar_init(bar_a1full + 8 * b, 1);           // b = column half
      ptx::mbar_init(bar_a1empty + 8 * b, 4);          // the four epilogue-1 warps of the half
      ptx::mbar_init(bar_yready + 8 * b, 8);           // all eight epilogue-1 warps
      ptx::mbar_init(bar_a2full + 8 * b, 1);
      ptx::mbar_init(bar_a2empty + 8 * b, 4);
      ptx::mbar_init(bar_uready + 8 * b, 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512u);
    ptx::tmem_relinquish();
  }
  if (warp >= 4) {                                      // folded-BN shifts (constant data): b3[256] | b1[64]
    const int i = threadIdx.x - 128;
    if (i < CH_N1 + CH_N2) {
      const float v = i < CH_N1 ? (p.b3 ? __ldg(p.b3 + i) : 0.f) : (p.b1 ? __ldg(p.b1 + i - CH_N1) : 0.f);
      reinterpret_cast<float*>(gen + CB_BIAS)[i] = v;
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
  const float* const sbias = reinterpret_cast<const float*>(gen + CB_BIAS);

  if (warp == 0) {
    // ===================================== TMA loads =========================================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(bar_w, (uint32_t)(CH_W3_BYTES + CH_W1_BYTES));
      ptx::tma_load_2d(&mapW3, bar_w, smem_w3, 0, 0);
      for (int c = 0; c < 4; ++c) ptx::tma_load_2d(&mapW1, bar_w, smem_w1 + c * (CH_N2 * 128), c * 64, 0);
      pdl_wait();
      const uint64_t pol_in = ptx::policy_evict_first();
      for (int k = 0; k < ntiles; ++k) {
        const uint32_t b = (uint32_t)k & 1u, ph = ((uint32_t)k >> 1) & 1u;
        const int row0 = (t0 + k) * 128;
        ptx::mbar_wait(bar_buffree + 8 * b, ph ^ 1u);
        ptx::mbar_arrive_expect_tx(bar_ldfull + 8 * b, (uint32_t)(CH_T_BYTES + CH_Y_BYTES));
        ptx::tma_load_2d_hint(&mapT, bar_ldfull + 8 * b, smem_t + b * CH_T_BYTES, 0, row0, pol_in);
        for (int c = 0; c < 4; ++c) ptx::tma_load_2d_hint(&mapR, bar_ldfull + 8 * b, smem_y + b * CH_Y_BYTES + c * CH_T_BYTES, c * 64, row0, pol_in);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    ptx::mbar_wait(bar_w, 0);
    ptx::tc_fence_after();
    const uint32_t w3_lo = tc_desc_lo(smem_w3, 1u), w1_lo = tc_desc_lo(smem_w1, 1u);
    for (int k = 0; k <= ntiles; ++k) {
      if (k < ntiles) {                                   // GEMM 1 of tile k: two column halves
        const uint32_t b = (uint32_t)k & 1u, ph = ((uint32_t)k >> 1) & 1u;
        ptx::mbar_wait(bar_ldfull + 8 * b, ph);
        ptx::tc_fence_after();
        const uint32_t t_lo = tc_desc_lo(smem_t + b * CH_T_BYTES, 1u);
        for (int h = 0; h < 2; ++h) {
          ptx::mbar_wait(bar_a1empty + 8 * h, ((uint32_t)k & 1u) ^ 1u);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16_lohi(tmem_base + (uint32_t)(128 * h), t_lo + 2u * kk, p.desc_hi, w3_lo + (uint32_t)h * ((128 * 128) >> 4) + 2u * kk, p.desc_hi, p.idesc1,
                                 kk ? 1u : 0u);
            ptx::umma_commit(bar_a1full + 8 * h);
          }
          __syncwarp();
        }
      }
      if (k > 0) {                                        // GEMM 2 of tile k - 1
        const int j = k - 1;
        const uint32_t b = (uint32_t)j & 1u, ph = ((uint32_t)j >> 1) & 1u;
        ptx::mbar_wait(bar_yready + 8 * b, ph);
        ptx::mbar_wait(bar_a2empty + 8 * b, ph ^ 1u);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t y_lo = tc_desc_lo(smem_y + b * CH_Y_BYTES, 1u);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16_lohi(tmem_base + 256u + 64u * b, y_lo + (uint32_t)c * (CH_T_BYTES >> 4) + 2u * kk, p.desc_hi,
                                 w1_lo + (uint32_t)c * ((CH_N2 * 128) >> 4) + 2u * kk, p.desc_hi, p.idesc2, (c | kk) ? 1u : 0u);
          }
          ptx::umma_commit(bar_a2full + 8 * b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ===================================== TMA stores ========================================
    if (ptx::elect_one()) {
      const uint64_t pol_out = ptx::policy_evict_last();
      for (int k = 0; k < ntiles; ++k) {
        const uint32_t b = (uint32_t)k & 1u, ph = ((uint32_t)k >> 1) & 1u;
        const int row0 = (t0 + k) * 128;
        ptx::mbar_wait(bar_yready + 8 * b, ph);            // y complete in Y[b] (the writers fenced towards the async proxy)
        for (int c = 0; c < 4; ++c) ptx::tma_store_2d_hint(&mapY, smem_y + b * CH_Y_BYTES + c * CH_T_BYTES, c * 64, row0, pol_out);
        ptx::bulk_commit();
        ptx::mbar_wait(bar_uready + 8 * b, ph);            // u staged in T[b]; GEMM 2 (which read Y[b]) completed before epilogue 2 ran
        ptx::tma_store_2d_hint(&mapU, smem_t + b * CH_T_BYTES, 0, row0, pol_out);
        ptx::bulk_commit();
        ptx::bulk_wait_read_all();                         // both stores have read their source: the buffers may be refilled
        ptx::mbar_arrive(bar_buffree + 8 * b);
      }
      ptx::bulk_wait_all();
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================== epilogue 1: y = relu(acc + b3 + x), in place over x ====
    const int q = warp & 3, h = (warp - 4) >> 2;             // rows 32 q .. 32 q + 31, columns 128 h .. 128 h + 127 (chunks 2 h, 2 h + 1)
    const uint32_t row = (uint32_t)(q * 32 + lane);
    for (int k = 0; k < ntiles; ++k) {
      const uint32_t b = (uint32_t)k & 1u, ph = ((uint32_t)k >> 1) & 1u;
      uint8_t* const yb = gen + (smem_y - base) + b * CH_Y_BYTES;
      ptx::mbar_wait(bar_ldfull + 8 * b, ph);                // the residual tile (same transaction barrier as t)
      ptx::mbar_wait(bar_a1full + 8 * h, (uint32_t)k & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(128 * h) + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int g = 0; g < 8; g += 2) {                        // 32 columns per step
        uint32_t a0[16], a1[16];
        ptx::tmem_ld16(taddr + (uint32_t)(16 * g), a0);
        ptx::tmem_ld16(taddr + (uint32_t)(16 * g + 16), a1);
        ptx::tmem_ld_wait();
        if (g == 6) {                                         // accumulator half read completely
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_a1empty + 8 * h);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {                         // 8 channels = one 16-byte chunk
          const int col = 16 * g + 8 * c;                     // column inside the half
          uint4* const slot_p = reinterpret_cast<uint4*>(yb + (uint32_t)(2 * h + (col >> 6)) * CH_T_BYTES + ch_chunk(row, (uint32_t)((col & 63) >> 3)));
          const float4 bA = *reinterpret_cast<const float4*>(sbias + 128 * h + col), bB = *reinterpret_cast<const float4*>(sbias + 128 * h + col + 4);
          const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
          float f[8], r[8];
          unpack8<T>(*slot_p, r);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float acc = __uint_as_float(c < 2 ? a0[8 * c + e] : a1[8 * (c - 2) + e]);
            f[e] = fmaxf(acc + bb[e] + r[e], 0.f);
          }
          *slot_p = pack8<T>(f);
        }
      }
      ptx::fence_proxy_async();                               // generic-proxy writes of Y[b] -> tensor-pipe reads and the TMA store
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_yready + 8 * b);
    }
  } else if (warp >= 12) {
    // ===================================== epilogue 2: u = relu(acc2 + b1) -> T[b] (staging of the u store) ====
    const int q = warp & 3;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    for (int k = 0; k < ntiles; ++k) {
      const uint32_t b = (uint32_t)k & 1u, ph = ((uint32_t)k >> 1) & 1u;
      uint8_t* const tb = gen + (smem_t - base) + b * CH_T_BYTES;
      ptx::mbar_wait(bar_a2full + 8 * b, ph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + 256u + 64u * b + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int g = 0; g < 4; g += 2) {
        uint32_t a0[16], a1[16];
        ptx::tmem_ld16(taddr + (uint32_t)(16 * g), a0);
        ptx::tmem_ld16(taddr + (uint32_t)(16 * g + 16), a1);
        ptx::tmem_ld_wait();
        if (g == 2) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_a2empty + 8 * b);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = 16 * g + 8 * c;
          const float4 bA = *reinterpret_cast<const float4*>(sbias + CH_N1 + col), bB = *reinterpret_cast<const float4*>(sbias + CH_N1 + col + 4);
          const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float acc = __uint_as_float(c < 2 ? a0[8 * c + e] : a1[8 * (c - 2) + e]);
            f[e] = fmaxf(acc + bb[e], 0.f);
          }
          *reinterpret_cast<uint4*>(tb + ch_chunk(row, (uint32_t)(col >> 3))) = pack8<T>(f);
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_uready + 8 * b);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512u);
}

// =======================================================================================================
// host side
// =======================================================================================================
struct TcChainState {
  CUtensorMap mapT, mapR, mapW3, mapW1, mapY, mapU;
  ChainP p;
  int grid, smem_bytes, dtype;
};

constexpr int CH_SMEM = 1024 + CH_HEADER + CH_W3_BYTES + CH_W1_BYTES + 2 * (CH_T_BYTES + CH_Y_BYTES);
static_assert(CH_SMEM <= TC_SMEM_LIMIT, "expand-reduce kernel: shared memory budget");

int tc_chain_supported(const capf_op& op) {
  const char* ev = getenv("CAPF_FUSE_CHAIN");
  if (ev && ev[0] == '0') return 0;
  if (op.kind != CAPF_OP_EXPAND_REDUCE) return 0;
  if (op.i[0] <= 0 || op.i[1] != CH_K1 || op.i[2] != CH_N1 || op.i[3] != CH_N2) return 0;
  if (op.dtype_in != op.dtype_out || (op.dtype_in != CAPF_F16 && op.dtype_in != CAPF_BF16)) return 0;
  if ((long long)op.i[0] * CH_N1 >= (1ll << 31)) return 0;
  if (!op.in[0] || !op.in[1] || !op.in[3] || !op.in[4] || !op.out[0] || !op.out[1]) return 0;
  if (((uintptr_t)op.in[0] | (uintptr_t)op.in[1] | (uintptr_t)op.in[3] | (uintptr_t)op.in[4] | (uintptr_t)op.out[0] | (uintptr_t)op.out[1]) & 15) return 0;
  return 1;
}

int tc_chain_prepare(const capf_op& op, TcChainState** out) {
  *out = nullptr;
  int e = tc_get_encoder();
  if (e) return e;
  if (!tc_chain_supported(op)) return set_error(CAPF_ERR_UNSUPPORTED, "expand-reduce: shape / dtype not supported (K1 = 64, N1 = 256, N2 = 64, 16-bit)");
  TcChainState* s = new (std::nothrow) TcChainState();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_chain_prepare: out of host memory");
  ChainP& p = s->p;
  memset(&p, 0, sizeof(p));
  p.P = op.i[0];
  p.num_tiles = (p.P + 127) / 128;
  const bool bf16 = op.dtype_in == CAPF_BF16;
  p.idesc1 = tc_idesc(bf16, 128);
  p.idesc2 = tc_idesc(bf16, CH_N2);
  p.desc_hi = tc_desc_hi(128, 1024);
  p.b3 = (const float*)op.in[2];
  p.b1 = (const float*)op.in[5];
  s->grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  s->smem_bytes = CH_SMEM;
  s->dtype = op.dtype_in;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  auto rows_map = [&](CUtensorMap* m, const void* ptr, int cols, int rows, int box_rows, const char* what) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return tc_encode_map(m, dt, 2, ptr, dims, strides, box, es, 128, what);
  };
  e = rows_map(&s->mapT, op.in[0], CH_K1, p.P, 128, "t tile (expand-reduce)");
  if (!e) e = rows_map(&s->mapR, op.in[3], CH_N1, p.P, 128, "residual tile (expand-reduce)");
  if (!e) e = rows_map(&s->mapW3, op.in[1], CH_K1, CH_N1, CH_N1, "W3 (expand-reduce)");
  if (!e) e = rows_map(&s->mapW1, op.in[4], CH_N1, CH_N2, CH_N2, "W1 (expand-reduce)");
  if (!e) e = rows_map(&s->mapY, op.out[0], CH_N1, p.P, 128, "y store (expand-reduce)");
  if (!e) e = rows_map(&s->mapU, op.out[1], CH_N2, p.P, 128, "u store (expand-reduce)");
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

template <typename T>
static int chain_launch_typed(const TcChainState* s, cudaStream_t st) {
  static PerDevice<bool> opted_;
  std::atomic<bool>& opted = opted_.get();
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(tc_expand_reduce_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_expand_reduce_kernel smem opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  launch_k(tc_expand_reduce_kernel<T>, dim3(s->grid), dim3(CH_THREADS), s->smem_bytes, st, s->mapT, s->mapR, s->mapW3, s->mapW1, s->mapY, s->mapU, s->p);
  return check_launch("tc_expand_reduce_kernel");
}

int tc_chain_launch(const TcChainState* s, cudaStream_t st) {
  return s->dtype == CAPF_F16 ? chain_launch_typed<__half>(s, st) : chain_launch_typed<__nv_bfloat16>(s, st);
}

void tc_chain_release(TcChainState* s) { delete s; }

void tc_chain_describe(const TcChainState* s, char* buf, int cap) {
  snprintf(buf, cap, "tc_expand_reduce_kernel[conv3 + residual -> conv1 of the next block, 128-pixel tiles, %d tiles]", s->p.num_tiles);
}

}  // namespace capf
