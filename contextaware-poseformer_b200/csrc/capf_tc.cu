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
// Roles (256 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator,
// warps 4..7 = epilogue (tcgen05.ld -> bias/GELU/residual/ReLU -> global).  Two TMEM accumulator stages let the
// epilogue of tile i overlap the MMAs of tile i+1; a ring of smem stages decouples TMA from the tensor pipe.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <new>

#include "capf_common.cuh"
#include "capf_internal.h"

namespace capf {

// =======================================================================================================
// device-side PTX wrappers
// =======================================================================================================
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0x3ff) == 0 && clock64() - t0 > 4000000000ll) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every tcgen05 op previously issued by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp receives row (lane base + t).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace ptx

// =======================================================================================================
// kernel parameters
// =======================================================================================================
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_HEADER_BYTES = 1024;       // barriers + TMEM base pointer
constexpr int TC_A_STAGE_BYTES = 128 * 128; // 128 rows x 64 elements x 2 B
constexpr int TC_SMEM_LIMIT = 232448;       // 227 KB opt-in maximum per CTA

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

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t hi) {
  // bits [0,14) start address >> 4, bits [16,30) leading byte offset >> 4 (ignored for swizzled K-major; canonical 1)
  return ((uint64_t)hi << 32) | (uint64_t)(((smem_addr >> 4) & 0x3fffu) | (1u << 16));
}

struct TileCoord {
  int n_tile;      // column tile
  int m0;          // rows mode: first row
  int ox0, oy0, n0;  // conv mode: first output pixel / image of the box
};

__device__ __forceinline__ TileCoord decode_tile(const TcP& p, int tile) {
  TileCoord t;
  t.n_tile = tile % p.n_tiles_n;
  int mt = tile / p.n_tiles_n;
  t.m0 = mt * 128;
  int tx = mt % p.tiles_x;
  int r = mt / p.tiles_x;
  int ty = r % p.tiles_y;
  t.ox0 = tx * p.bw;
  t.oy0 = ty * p.bh;
  t.n0 = (r / p.tiles_y) * p.bn;
  return t;
}

template <typename TO> struct Pack16;
template <> struct Pack16<float> {
  static __device__ __forceinline__ void load(const float* p, float (&r)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
      r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
    }
  }
  static __device__ __forceinline__ void store(float* p, const float (&r)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
  }
};
template <> struct Pack16<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&r)[16]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint4 v = *reinterpret_cast<const uint4*>(p + 8 * q);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __half22float2(h[e]);
        r[8 * q + 2 * e] = f.x; r[8 * q + 2 * e + 1] = f.y;
      }
    }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&r)[16]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint4 v;
      __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(r[8 * q + 2 * e], r[8 * q + 2 * e + 1]);
      *reinterpret_cast<uint4*>(p + 8 * q) = v;
    }
  }
};
template <> struct Pack16<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&r)[16]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint4 v = *reinterpret_cast<const uint4*>(p + 8 * q);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __bfloat1622float2(h[e]);
        r[8 * q + 2 * e] = f.x; r[8 * q + 2 * e + 1] = f.y;
      }
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&r)[16]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint4 v;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(r[8 * q + 2 * e], r[8 * q + 2 * e + 1]);
      *reinterpret_cast<uint4*>(p + 8 * q) = v;
    }
  }
};

// =======================================================================================================
// the kernel
// =======================================================================================================
template <typename TO>
__global__ void __launch_bounds__(256, 1)
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
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (base - raw) + (tmem_slot - base));

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
  const uint32_t acc_stride = (uint32_t)p.tmem_cols >> 1;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const int ix_base = t.ox0 * p.stride - p.pad, iy_base = t.oy0 * p.stride - p.pad;
        for (int c0 = 0; c0 < p.num_chunks; c0 += p.cps) {
          ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const int nc = min(p.cps, p.num_chunks - c0);
          const uint32_t full = bar_full + 8 * stage;
          ptx::mbar_arrive_expect_tx(full, (uint32_t)(nc * p.tx_bytes_per_chunk));
          const uint32_t a_dst = stage0 + stage * p.stage_bytes;
          const uint32_t b_dst = a_dst + TC_A_STAGE_BYTES;
          for (int j = 0; j < nc; ++j) {
            const int chunk = c0 + j;
            if (p.mode == 1) {
              const int tap = chunk / p.cpt;
              const int ci0 = (chunk - tap * p.cpt) * p.kb;
              const int r = tap / p.KW, s = tap - r * p.KW;
              ptx::tma_load_4d(&mapA, full, a_dst + j * p.a_chunk_bytes, ci0, ix_base + s, iy_base + r, t.n0);
            } else {
              ptx::tma_load_2d(&mapA, full, a_dst + j * p.a_chunk_bytes, chunk * p.kb, t.m0);
            }
            ptx::tma_load_2d(&mapB, full, b_dst + j * p.b_chunk_bytes, chunk * p.kb, t.n_tile * p.BN);
          }
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t it = 0;
      const int ksteps = p.kb >> 4;  // UMMA_K = 16 elements = 32 bytes inside the swizzled row
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
        ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        uint32_t accumulate = 0;
        for (int c0 = 0; c0 < p.num_chunks; c0 += p.cps) {
          ptx::mbar_wait(bar_full + 8 * stage, phase);
          ptx::tc_fence_after();
          const int nc = min(p.cps, p.num_chunks - c0);
          const uint32_t a_src = stage0 + stage * p.stage_bytes;
          const uint32_t b_src = a_src + TC_A_STAGE_BYTES;
          for (int j = 0; j < nc; ++j) {
            const uint32_t a_addr = a_src + j * p.a_chunk_bytes, b_addr = b_src + j * p.b_chunk_bytes;
            for (int k = 0; k < ksteps; ++k) {
              ptx::umma_f16(d_tmem, make_desc(a_addr + 32 * k, p.desc_hi), make_desc(b_addr + 32 * k, p.desc_hi), p.idesc, accumulate);
              accumulate = 1;
            }
          }
          ptx::umma_commit(bar_empty + 8 * stage);  // smem slot reusable once these MMAs have read it
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit(bar_tfull + 8 * acc);      // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;          // accumulator row == tile row
    const TO* res = reinterpret_cast<const TO*>(p.res);
    TO* out = reinterpret_cast<TO*>(p.out);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
      const TileCoord t = decode_tile(p, tile);
      long long orow;  // output row (pixel) index, -1 when this tile row is padding
      if (p.mode == 1) {
        const int ix = row % p.bw;
        const int r2 = row / p.bw;
        const int iy = r2 % p.bh, in_ = r2 / p.bh;
        const int ox = t.ox0 + ix, oy = t.oy0 + iy, n = t.n0 + in_;
        const bool ok = in_ < p.bn && ox < p.Wo && oy < p.Ho && n < p.Nimg;
        orow = ok ? ((long long)n * p.Ho + oy) * p.Wo + ox : -1;
      } else {
        orow = (t.m0 + row < p.M) ? (long long)(t.m0 + row) : -1;
      }
      ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);
      const int ncol0 = t.n_tile * p.BN;
      for (int c = 0; c < p.BN; c += 16) {
        uint32_t raw16[16];
        ptx::tmem_ld16(taddr + (uint32_t)c, raw16);
        ptx::tmem_ld_wait();
        if (orow >= 0) {
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(raw16[e]);
          const int n = ncol0 + c;
          if (p.bias) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + q4);
              v[4 * q4] += b4.x; v[4 * q4 + 1] += b4.y; v[4 * q4 + 2] += b4.z; v[4 * q4 + 3] += b4.w;
            }
          }
          if (p.act == CAPF_ACT_GELU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = gelu_erf(v[e]);
          }
          const size_t off = (size_t)orow * p.Cout + n;
          if (res) {
            float r16[16];
            Pack16<TO>::load(res + off, r16);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += r16[e];
          }
          if (p.act == CAPF_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          Pack16<TO>::store(out + off, v);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_tempty + 8 * acc);   // 128 arrivals free this accumulator stage
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
  CUtensorMap mapA, mapB;
  TcP p;
  int grid;
  int smem_bytes;
  int dtype_out;
};

static PFN_cuTensorMapEncodeTiled g_encode = nullptr;

static int get_encoder() {
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

static int encode_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* ptr, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, int swz_bytes, const char* what) {
  CUtensorMapSwizzle sw = swz_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = g_encode(m, dt, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_errorf(CAPF_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return CAPF_OK;
}

int tc_conv_prepare(const capf_op& op, TcConvState** out) {
  *out = nullptr;
  int e = get_encoder();
  if (e) return e;
  const ConvGeo g = geo_of(op);
  TcConvState* s = new (std::nothrow) TcConvState();
  if (!s) return set_error(CAPF_ERR_ARG, "tc_conv_prepare: out of host memory");
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
  const uint32_t fmt = op.dtype_in == CAPF_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t layout = swz == 128 ? 2u : swz == 64 ? 4u : 6u;   // UMMA LayoutType: SWIZZLE_128B / 64B / 32B
  const uint32_t sbo = (uint32_t)(8 * swz) >> 4;                    // byte distance between 8-row groups, >> 4
  p.desc_hi = sbo | (1u << 14) | (layout << 29);                    // version = 1 at bit 46, layout at bits 61..63
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
    e = encode_map(&s->mapA, dt, 2, op.in[0], dims, strides, box, es, swz, "A rows");
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.W * g.Cin * 2, (cuuint64_t)g.H * g.W * g.Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.kb, (cuuint32_t)(p.bw * g.stride), (cuuint32_t)(p.bh * g.stride), (cuuint32_t)p.bn};
    cuuint32_t es[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
    e = encode_map(&s->mapA, dt, 4, op.in[0], dims, strides, box, es, swz, "A conv");
  }
  if (!e) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)g.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kb, (cuuint32_t)p.BN};
    cuuint32_t es[2] = {1, 1};
    e = encode_map(&s->mapB, dt, 2, op.in[1], dims, strides, box, es, swz, "B weights");
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
  tc_gemm_kernel<TO><<<s->grid, 256, s->smem_bytes, st>>>(s->mapA, s->mapB, s->p);
  return check_launch("tc_gemm_kernel");
}

int tc_conv_launch(const capf_op&, const TcConvState* s, cudaStream_t st) {
  if (!s) return set_error(CAPF_ERR_ARG, "tc conv: op was not prepared");
  switch (s->dtype_out) {
    case CAPF_F32: return tc_launch_typed<float>(s, st);
    case CAPF_F16: return tc_launch_typed<__half>(s, st);
    case CAPF_BF16: return tc_launch_typed<__nv_bfloat16>(s, st);
    default: return set_error(CAPF_ERR_UNSUPPORTED, "tc conv: dtype_out");
  }
}

void tc_conv_release(TcConvState* s) { delete s; }

}  // namespace capf
