// Backbone stem conv1 on the tensor pipe, straight from the caller's fp32 NHWC image, folded BatchNorm + ReLU, 16-bit NHWC
// output: 3x3 / stride 2 / pad 1, 3 -> 64 channels for HRNet (pose_hrnet.py:321-322, :465-467) and 7x7 / stride 2 / pad 3
// for the ResNet-50 of CPN (networks/resnet.py:100-104, :137-139).  The text below describes the 3x3 case; the 7x7 case
// differs in K = 147 -> 160 and single (un-split) 16-bit operands, see StemGeo.
//
// The op is bound by its output (2.6x the input bytes); the CUDA-core version (capf_simt.cu) needs 1728 FMAs and 432
// shared-memory weight loads per output pixel and ran at ~1/6 of the HBM roofline.  Here a tile is 128 consecutive
// output pixels of one output row:
//   builder warps (4)   cp.async the three input rows of the tile (fp32, zero-filled outside the image) into a 4-deep
//                       shared-memory ring, then every thread gathers the 27 taps of its pixel and writes one im2col
//                       row, split into a 16-bit "hi" part and a 16-bit "lo" remainder (x = hi + lo to ~2^-22), into
//                       the un-swizzled K-major core-matrix layout tcgen05 reads (K = 27 padded to 32 per part);
//   MMA warp (1 thread) D[128 x 64] = Ahi*Whi + Alo*Whi + Ahi*Wlo  (6 tcgen05.mma of K = 16, fp32 accumulate in TMEM):
//                       the hi/lo split keeps the fp32 image and the fp32 folded weights at fp32-class accuracy;
//   epilogue warps (4)  tcgen05.ld -> bias + ReLU -> 16-bit -> XOR-swizzled staging tile -> coalesced 16-byte stores
//                       (the 32 pixels of a warp are 4 KB of contiguous NHWC output).
// A tiles and TMEM accumulators are double buffered; two CTAs per SM keep enough loads in flight for HBM.
#include <cstdlib>

#include "capf_tc.cuh"

namespace capf {

constexpr int STEM_THREADS = 288;              // warps 0-3 builders, warp 4 MMA issuer / TMEM owner, warps 5-8 epilogue
constexpr int STEM_RING = 4;

// Compile-time geometry of one stem variant: KS x KS taps, stride 2, pad KS / 2, 3 input channels, 64 output channels.
//   KS = 3 (HRNet conv1, pose_hrnet.py:321):   K = 27 -> 32,  hi/lo split operands (fp32-class accuracy), 2 CTAs per SM
//   KS = 7 (CPN ResNet conv1, resnet.py:100):  K = 147 -> 160, single 16-bit operands, 1 CTA per SM
template <int KS, bool SPLIT>
struct StemGeo {
  static constexpr int PAD = KS / 2;
  static constexpr int TAPW = 3 * KS;                         // floats of one filter row of one pixel (contiguous)
  static constexpr int K = KS * TAPW;
  static constexpr int KPAD = (K + 15) / 16 * 16;
  static constexpr int NP = KPAD / 8;                         // 8-value K planes
  static constexpr int LEAD = (4 - (3 * PAD) % 4) % 4;        // floats in front of the first tap (16-byte aligned row start)
  static constexpr int ROW_FLOATS = (LEAD + 3 * (254 + KS) + 3) / 4 * 4;
  static constexpr int ROW_CHUNKS = ROW_FLOATS / 4;
  static constexpr int PATCH_BYTES = KS * ROW_FLOATS * 4;
  static constexpr int PARTS = SPLIT ? 2 : 1;
  static constexpr int A_PART = NP * 128 * 16;                // one operand part of a 128-row A tile
  static constexpr int B_PART = NP * 64 * 16;
  static constexpr int OFF_B = 1024;
  static constexpr int OFF_A = OFF_B + PARTS * B_PART;
  static constexpr int OFF_PATCH = OFF_A + 2 * PARTS * A_PART;
  static constexpr int OFF_STG = OFF_PATCH + STEM_RING * PATCH_BYTES;
  static constexpr int SMEM = OFF_STG + 4 * 4096 + 1024;
  static constexpr int CTAS_PER_SM = SMEM <= 110 * 1024 ? 2 : 1;
};

struct StemP {
  int N, H, W, Ho, Wo, tiles_x, num_tiles, relu;
  uint32_t idesc, a_desc_hi, b_desc_hi;
  const float* x;
  const float* w;      // [KS*KS*3][64] fp32, k = (r*KS+s)*3+c (folded BN scale applied)
  const float* bias;   // [64] or NULL
  void* y;
};

template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

template <typename TO> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <typename TO, int KS, bool SPLIT>
__global__ void __launch_bounds__(STEM_THREADS, (StemGeo<KS, SPLIT>::CTAS_PER_SM)) stem_tc_kernel(const StemP p) {
  using G = StemGeo<KS, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_afull = base, bar_aempty = base + 16, bar_tfull = base + 32, bar_tempty = base + 48, tmem_slot = base + 64;
  const uint32_t smem_b = base + G::OFF_B, smem_a = base + G::OFF_A, smem_patch = base + G::OFF_PATCH, smem_stg = base + G::OFF_STG;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar_afull + 8 * b, 128);
      ptx::mbar_init(bar_aempty + 8 * b, 1);
      ptx::mbar_init(bar_tfull + 8 * b, 1);
      ptx::mbar_init(bar_tempty + 8 * b, 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(tmem_slot, 128u);
    ptx::tmem_relinquish();
  }
  // weights: fp32 [K][64] -> 16-bit (hi / lo parts when SPLIT) in the K-major core-matrix layout (plane = 8 K values:
  // [64 rows][16 B])
  for (int idx = tid; idx < 64 * G::KPAD; idx += STEM_THREADS) {
    const int n = idx / G::KPAD, k = idx - n * G::KPAD;
    const float v = k < G::K ? __ldg(p.w + k * 64 + n) : 0.f;
    const TO hi = from_f<TO>(v);
    const uint32_t off = (uint32_t)((k >> 3) * 1024 + n * 16 + (k & 7) * 2);
    *reinterpret_cast<TO*>(smem_raw + (smem_b - raw) + off) = hi;
    if (SPLIT) *reinterpret_cast<TO*>(smem_raw + (smem_b - raw) + G::B_PART + off) = from_f<TO>(v - to_f<TO>(hi));
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_trigger();
  pdl_wait();      // the output buffer may still be read by the predecessor (activation memory is reused)

  const int t0 = (int)(((long long)p.num_tiles * blockIdx.x) / gridDim.x);
  const int t1 = (int)(((long long)p.num_tiles * (blockIdx.x + 1)) / gridDim.x);

  if (warp < 4) {
    // ===================================== builders ==========================================
    auto issue_loads = [&](int tile, int slot) {
      if (tile < t1) {
        const int txi = tile % p.tiles_x, rowid = tile / p.tiles_x;
        const int oy = rowid % p.Ho, n = rowid / p.Ho;
        const int col_base = 6 * (txi * 128) - 3 * G::PAD - G::LEAD;
        const float* img = p.x + (size_t)n * p.H * p.W * 3;
        const uint32_t dst0 = smem_patch + (uint32_t)slot * G::PATCH_BYTES;
        for (int q = tid; q < KS * G::ROW_CHUNKS; q += 128) {
          const int r = q / G::ROW_CHUNKS, j = q - r * G::ROW_CHUNKS;
          const int iy = 2 * oy - G::PAD + r, col0 = col_base + 4 * j;
          const bool ok = iy >= 0 && iy < p.H && col0 >= 0 && col0 + 4 <= 3 * p.W;
          const float* src = ok ? img + (size_t)iy * p.W * 3 + col0 : p.x;
          cp_async16_zfill(dst0 + (uint32_t)(r * G::ROW_FLOATS + 4 * j) * 4u, src, ok ? 16u : 0u);
        }
      }
      ptx::cp_async_commit();
    };
    for (int d = 0; d < STEM_RING - 1; ++d) issue_loads(t0 + d, d);
    uint32_t it = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      cp_async_wait_group<STEM_RING - 2>();            // this tile's rows have landed (two younger groups may be in flight)
      named_bar_sync(1, 128);                          // ... for every builder thread; everyone is also done with tile - 1
      issue_loads(tile + STEM_RING - 1, (int)((it + STEM_RING - 1) & (STEM_RING - 1)));
      const uint32_t buf = it & 1u, aph = (it >> 1) & 1u;
      const uint32_t patch = smem_patch + (it & (STEM_RING - 1)) * G::PATCH_BYTES + (uint32_t)(6 * tid + G::LEAD) * 4u;
      ptx::mbar_wait(bar_aempty + 8 * buf, aph ^ 1u);  // the MMAs that read this A buffer two tiles ago have completed
      const uint32_t a_hi = smem_a + buf * (G::PARTS * G::A_PART) + (uint32_t)tid * 16u, a_lo = a_hi + G::A_PART;
#pragma unroll
      for (int kc = 0; kc < G::NP; ++kc) {
        // the 8 im2col values of K plane kc of this pixel: k = (filter row r) * TAPW + j, contiguous in j inside a row
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = kc * 8 + e;
          x[e] = k < G::K ? ld_shared_f32(patch + (uint32_t)((k / G::TAPW) * G::ROW_FLOATS + (k % G::TAPW)) * 4u) : 0.f;
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float h0 = to_f<TO>(from_f<TO>(x[2 * e])), h1 = to_f<TO>(from_f<TO>(x[2 * e + 1]));
          h[e] = pack2<TO>(h0, h1);
          l[e] = pack2<TO>(x[2 * e] - h0, x[2 * e + 1] - h1);
        }
        ptx::st_shared_v4(a_hi + kc * 2048, make_uint4(h[0], h[1], h[2], h[3]));
        if (SPLIT) ptx::st_shared_v4(a_lo + kc * 2048, make_uint4(l[0], l[1], l[2], l[3]));
      }
      ptx::fence_proxy_async();                        // generic-proxy writes -> visible to the tensor pipe's reads
      ptx::mbar_arrive(bar_afull + 8 * buf);
    }
    ptx::cp_async_wait_all();
  } else if (warp == 4) {
    // ===================================== MMA issuer ========================================
    uint32_t it = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
      ptx::mbar_wait(bar_tempty + 8 * buf, ph ^ 1u);
      ptx::mbar_wait(bar_afull + 8 * buf, ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t d_tmem = tmem_base + buf * 64u;
        const uint32_t a0 = smem_a + buf * (G::PARTS * G::A_PART);
        // (A part, B part): hi*hi, lo*hi, hi*lo when the operands are split, else the single product
        constexpr int NC = SPLIT ? 3 : 1;
        const uint32_t a_part[3] = {0u, (uint32_t)G::A_PART, 0u}, b_part[3] = {0u, 0u, (uint32_t)G::B_PART};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
#pragma unroll
          for (int kk = 0; kk < G::KPAD / 16; ++kk) {
            const uint32_t a_lo = tc_desc_lo(a0 + a_part[c] + (uint32_t)kk * 4096u, 128u);     // LBO: 2048 B between K planes
            const uint32_t b_lo = tc_desc_lo(smem_b + b_part[c] + (uint32_t)kk * 2048u, 64u);   // LBO: 1024 B
            ptx::umma_f16_lohi(d_tmem, a_lo, p.a_desc_hi, b_lo, p.b_desc_hi, p.idesc, (c | kk) ? 1u : 0u);
          }
        }
        ptx::umma_commit(bar_aempty + 8 * buf);
        ptx::umma_commit(bar_tfull + 8 * buf);
      }
      __syncwarp();
    }
  } else {
    // ===================================== epilogue ==========================================
    const int q = warp & 3;                            // TMEM lane quadrant this warp may read
    TO* out = reinterpret_cast<TO*>(p.y);
    const uint32_t stg = smem_stg + (uint32_t)(warp - 5) * 4096u;
    auto slot = [&](int r, int c) { return stg + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); };
    const uint64_t pol_out = ptx::policy_evict_last();
    const int act = p.relu ? CAPF_ACT_RELU : CAPF_ACT_NONE;
    uint32_t it = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
      const int txi = tile % p.tiles_x, rowid = tile / p.tiles_x;     // rowid = n * Ho + oy
      const int ox_w = txi * 128 + q * 32;                            // first output pixel of this warp
      const uint32_t taddr = tmem_base + buf * 64u + ((uint32_t)(q * 32) << 16);
      ptx::mbar_wait(bar_tfull + 8 * buf, ph);
      ptx::tc_fence_after();
#pragma unroll
      for (int v = 0; v < 4; v += 2) {
        uint32_t a0[16], a1[16];
        ptx::tmem_ld16(taddr + (uint32_t)(16 * v), a0);
        ptx::tmem_ld16(taddr + (uint32_t)(16 * v + 16), a1);
        Bias16 b0, b1;
        b0.load(p.bias, 16 * v);
        b1.load(p.bias, 16 * v + 16);
        ptx::tmem_ld_wait();
        finish16_smem<TO>(b0, act, a0, false, slot(lane, 2 * v), slot(lane, 2 * v + 1));
        finish16_smem<TO>(b1, act, a1, false, slot(lane, 2 * v + 2), slot(lane, 2 * v + 3));
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_tempty + 8 * buf);
      __syncwarp();
      TO* dst = out + ((size_t)rowid * p.Wo + ox_w) * 64;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int item = i * 32 + lane, r = item >> 3, c = item & 7;
        if (ox_w + r < p.Wo) ptx::st_global_v4_hint(dst + r * 64 + c * 8, ptx::ld_shared_v4(slot(r, c)), pol_out);
      }
      __syncwarp();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem_base, 128u);
}

int stem_tc_supported(const capf_op& op) {
  const char* ev = getenv("CAPF_STEM_TC");
  if (ev && ev[0] == '0') return 0;
  const int ks = op.i[5];
  if (op.i[3] != 3 || op.i[4] != 64 || (ks != 3 && ks != 7) || op.i[6] != ks || op.i[7] != 2 || op.i[8] != ks / 2) return 0;
  if (op.dtype_in != CAPF_F32 || (op.dtype_out != CAPF_F16 && op.dtype_out != CAPF_BF16)) return 0;
  if (op.in[3] || op.i[11] == CAPF_ACT_GELU) return 0;
  if (op.i[2] % 4 || op.i[2] < 4) return 0;                                   // 16-byte chunks never straddle the row end
  if (((uintptr_t)op.in[0] | (uintptr_t)op.out[0]) & 15) return 0;
  if ((long long)op.i[0] * op.i[9] * ((op.i[10] + 127) / 128) >= (1ll << 31)) return 0;
  return 1;
}

template <typename TO, int KS, bool SPLIT>
static int stem_launch_typed(const StemP& p, cudaStream_t st) {
  using G = StemGeo<KS, SPLIT>;
  static PerDevice<bool> opted_;
  std::atomic<bool>& opted = opted_.get();
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel<TO, KS, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "stem_tc_kernel smem opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  int grid = G::CTAS_PER_SM * num_sms();
  if (grid > p.num_tiles) grid = p.num_tiles;
  launch_k(stem_tc_kernel<TO, KS, SPLIT>, dim3(grid), dim3(STEM_THREADS), G::SMEM, st, p);
  return check_launch("stem_tc_kernel");
}

int launch_stem_tc(const capf_op& op, cudaStream_t st) {
  StemP p;
  p.N = op.i[0]; p.H = op.i[1]; p.W = op.i[2]; p.Ho = op.i[9]; p.Wo = op.i[10];
  p.tiles_x = (p.Wo + 127) / 128;
  p.num_tiles = p.N * p.Ho * p.tiles_x;
  p.relu = op.i[11] == CAPF_ACT_RELU;
  p.idesc = tc_idesc(op.dtype_out == CAPF_BF16, 64);
  p.a_desc_hi = tc_desc_hi(0, 128);
  p.b_desc_hi = tc_desc_hi(0, 128);
  p.x = (const float*)op.in[0];
  p.w = (const float*)op.in[1];
  p.bias = (const float*)op.in[2];
  p.y = op.out[0];
  const bool f16 = op.dtype_out == CAPF_F16;
  if (op.i[5] == 3) return f16 ? stem_launch_typed<__half, 3, true>(p, st) : stem_launch_typed<__nv_bfloat16, 3, true>(p, st);
  return f16 ? stem_launch_typed<__half, 7, false>(p, st) : stem_launch_typed<__nv_bfloat16, 7, false>(p, st);
}

}  // namespace capf
