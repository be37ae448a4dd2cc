// CUDA-core (SIMT) kernels of libcapf_b200: the fp32-parity implementation of every operator on the
// CA_PF.forward path, plus the memory-bound operators (samplers, LayerNorm, tiny attention, fuse-sum)
// that the fp16/bf16 tensor-core path shares.  Reference citations are relative to
// /root/reference/ContextPose/mvn/models/.
#include <cstdlib>
#include <type_traits>

#include "capf_common.cuh"
#include "capf_internal.h"

namespace capf {

// =======================================================================================================
// conv2d / linear: implicit GEMM  y[m][co] = sum_k A[m][k] * Wt[k][co],  m = (n,oy,ox), k = (r,s,ci)
// replaces nn.Conv2d+BatchNorm2d(+ReLU)(+residual) (pose_hrnet.py:79-136) and nn.Linear (pose_dformer.py:25-31)
// =======================================================================================================
struct ConvP {
  int N, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo, act;
  int M, K;
};

template <typename TI, typename TW, typename TO, int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(256)
conv_nhwc_simt(ConvP p, const TI* __restrict__ x, const TW* __restrict__ w, const float* __restrict__ bias,
               const TO* res, TO* y) {
  pdl_wait();
  constexpr int BK = 16;
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const bool pointwise = (p.KH == 1 && p.KW == 1 && p.stride == 1 && p.pad == 0);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // ---- per-thread A-load coordinates --------------------------------------------------------------
  constexpr int A_CHUNKS = VEC ? (BM * BK / 4) / 256 : 1;  // VEC: float4 chunks; scalar: one pixel/thread
  int a_pix[A_CHUNKS], a_n[A_CHUNKS], a_iy0[A_CHUNKS], a_ix0[A_CHUNKS];
  bool a_ok[A_CHUNKS];
#pragma unroll
  for (int c = 0; c < A_CHUNKS; ++c) {
    int pix = VEC ? (tid + c * 256) / 4 : tid / 2;
    a_pix[c] = pix;
    int m = m0 + pix;
    a_ok[c] = (pix < BM) && (m < p.M);
    int mm = a_ok[c] ? m : 0;
    int ox = mm % p.Wo, t = mm / p.Wo;
    int oy = t % p.Ho;
    a_n[c] = t / p.Ho;
    a_iy0[c] = oy * p.stride - p.pad;
    a_ix0[c] = ox * p.stride - p.pad;
  }

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    if (VEC) {
      // Cin % 16 == 0: the 16 k's of this slab share one filter tap.
      int tap = k0 / p.Cin, ci0 = k0 - tap * p.Cin;
      int r = tap / p.KW, s = tap - r * p.KW;
      int kq = tid & 3;
#pragma unroll
      for (int c = 0; c < A_CHUNKS; ++c) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int iy = a_iy0[c] + r, ix = a_ix0[c] + s;
        if (a_ok[c] && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
          size_t off = pointwise ? (size_t)(m0 + a_pix[c]) * p.Cin
                                 : ((size_t)(a_n[c] * p.H + iy) * p.W + ix) * p.Cin;
          v = ld4<TI>(x + off + ci0 + kq * 4);
        }
        As[kq * 4 + 0][a_pix[c]] = v.x;
        As[kq * 4 + 1][a_pix[c]] = v.y;
        As[kq * 4 + 2][a_pix[c]] = v.z;
        As[kq * 4 + 3][a_pix[c]] = v.w;
      }
      constexpr int B_CHUNKS = (BK * BN / 4 + 255) / 256;
#pragma unroll
      for (int c = 0; c < B_CHUNKS; ++c) {
        int id = tid + c * 256;
        if (id < BK * BN / 4) {
          int kk = id / (BN / 4), nq = id - kk * (BN / 4);
          int n = n0 + nq * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n < p.Cout) v = ld4<TW>(w + (size_t)(k0 + kk) * p.Cout + n);
          *reinterpret_cast<float4*>(&Bs[kk][nq * 4]) = v;
        }
      }
    } else {
      // generic path (stem Cin=3, head Cout=3, ragged K): scalar loads with full decode
      int kb = (tid & 1) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int kk = kb + e, k = k0 + kk;
        float v = 0.f;
        if (a_ok[0] && k < p.K) {
          int tap = k / p.Cin, ci = k - tap * p.Cin;
          int r = tap / p.KW, s = tap - r * p.KW;
          int iy = a_iy0[0] + r, ix = a_ix0[0] + s;
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
            v = to_f<TI>(x[((size_t)(a_n[0] * p.H + iy) * p.W + ix) * p.Cin + ci]);
        }
        As[kk][a_pix[0]] = v;
      }
      for (int id = tid; id < BK * BN; id += 256) {
        int kk = id / BN, nn = id - kk * BN;
        int k = k0 + kk, n = n0 + nn;
        Bs[kk][nn] = (k < p.K && n < p.Cout) ? to_f<TW>(w[(size_t)k * p.Cout + n]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: y = relu?( gelu?(acc + bias) + residual ) ---------------------------------------------
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      int n = n0 + tx * TN + j;
      if (n >= p.Cout) continue;
      float v[4] = {acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]};
      size_t off = (size_t)m * p.Cout + n;
      if (VEC) {
        if (bias) {
          float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
          v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
        }
        if (p.act == CAPF_ACT_GELU) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = gelu_erf(v[e]);
        }
        if (res) {
          float4 r4 = ld4<TO>(res + off);
          v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
        }
        if (p.act == CAPF_ACT_RELU) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
        }
        st4<TO>(y + off, make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (n + e >= p.Cout) break;
          float t = v[e] + (bias ? bias[n + e] : 0.f);
          if (p.act == CAPF_ACT_GELU) t = gelu_erf(t);
          if (res) t += to_f<TO>(res[off + e]);
          if (p.act == CAPF_ACT_RELU) t = fmaxf(t, 0.f);
          y[off + e] = from_f<TO>(t);
        }
      }
    }
  }
}

template <typename TI, typename TW, typename TO>
static int conv_dispatch(const ConvP& p, const capf_op& op, cudaStream_t st) {
  const TI* x = (const TI*)op.in[0];
  const TW* w = (const TW*)op.in[1];
  const float* bias = (const float*)op.in[2];
  const TO* res = (const TO*)op.in[3];
  TO* y = (TO*)op.out[0];
  bool vec = (p.Cin % 16 == 0) && (p.Cout % 4 == 0);
  if (!vec) {
    dim3 g((p.M + 127) / 128, (p.Cout + 31) / 32);
    launch_k(conv_nhwc_simt<TI, TW, TO, 128, 32, 4, 4, false>, dim3(g), dim3(256), 0, st, p, x, w, bias, res, y);
  } else if (p.Cout <= 32 || (p.Cout % 64 != 0 && p.Cout % 32 == 0 && p.Cout < 128)) {
    dim3 g((p.M + 127) / 128, (p.Cout + 31) / 32);
    launch_k(conv_nhwc_simt<TI, TW, TO, 128, 32, 4, 4, true>, dim3(g), dim3(256), 0, st, p, x, w, bias, res, y);
  } else {
    dim3 g((p.M + 127) / 128, (p.Cout + 63) / 64);
    launch_k(conv_nhwc_simt<TI, TW, TO, 128, 64, 8, 4, true>, dim3(g), dim3(256), 0, st, p, x, w, bias, res, y);
  }
  return check_launch("conv_nhwc_simt");
}

// =======================================================================================================
// Stem conv1 of HRNet: 3x3 / stride 2 / pad 1, 3 -> 64 channels on the caller's fp32 NHWC image, folded BN + ReLU,
// 16-bit output (pose_hrnet.py:321-322, :465-467).  K = 27 is too thin for the tensor path and the op is bound by
// its 2.6x larger output, so: one thread per output pixel, the 33x33x3 input patch of a 16x16 output tile and the
// 27x64 weights in shared memory, weights read as broadcast 128-bit loads, fp32 accumulate in tap order.
// =======================================================================================================
template <typename TO>
__global__ void __launch_bounds__(256) stem_conv3x3s2_c3_kernel(int H, int W, int Ho, int Wo, int relu, const float* __restrict__ x,
                                                                const float* __restrict__ w, const float* __restrict__ bias,
                                                                TO* __restrict__ y) {
  pdl_wait();
  __shared__ float patch[33 * 33 * 3 + 1];
  __shared__ __align__(16) float ws[27 * 64];
  __shared__ __align__(16) float bs[64];
  const int tid = threadIdx.x;
  const int ox0 = blockIdx.x * 16, oy0 = blockIdx.y * 16, n = blockIdx.z;
  const int iy0 = oy0 * 2 - 1, ix0 = ox0 * 2 - 1;
  const float* img = x + (size_t)n * H * W * 3;
  for (int i = tid; i < 33 * 99; i += 256) {
    const int r = i / 99, c = i - r * 99;          // c = (column, channel) flattened: contiguous in the image row
    const int iy = iy0 + r, ixc = ix0 * 3 + c;
    float v = 0.f;
    if (iy >= 0 && iy < H && ixc >= 0 && ixc < W * 3) v = __ldg(img + (size_t)iy * W * 3 + ixc);
    patch[i] = v;
  }
  for (int i = tid; i < 27 * 64; i += 256) ws[i] = __ldg(w + i);
  if (tid < 64) bs[tid] = bias ? __ldg(bias + tid) : 0.f;
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  const int ox = ox0 + tx, oy = oy0 + ty;
  float in[27];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int c = 0; c < 3; ++c) in[(r * 3 + s) * 3 + c] = patch[((2 * ty + r) * 33 + 2 * tx + s) * 3 + c];
  if (ox >= Wo || oy >= Ho) return;
  TO* dst = y + (((size_t)n * Ho + oy) * Wo + ox) * 64;
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float4* wk = reinterpret_cast<const float4*>(ws + k * 64 + g * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w4 = wk[q];
        acc[4 * q] = fmaf(in[k], w4.x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(in[k], w4.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(in[k], w4.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(in[k], w4.w, acc[4 * q + 3]);
      }
    }
    float o[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      o[c] = acc[c] + bs[g * 16 + c];
      if (relu) o[c] = fmaxf(o[c], 0.f);
    }
    // 16 channels = 32 bytes: two 128-bit stores
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 v;
      uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if constexpr (sizeof(TO) == 2 && std::is_same<TO, __half>::value) {
          __half2 t = __floats2half2_rn(o[8 * h + 2 * e], o[8 * h + 2 * e + 1]);
          pv[e] = *reinterpret_cast<uint32_t*>(&t);
        } else {
          __nv_bfloat162 t = __floats2bfloat162_rn(o[8 * h + 2 * e], o[8 * h + 2 * e + 1]);
          pv[e] = *reinterpret_cast<uint32_t*>(&t);
        }
      }
      *reinterpret_cast<uint4*>(dst + g * 16 + 8 * h) = v;
    }
  }
}

static bool is_stem(const ConvP& p, const capf_op& op) {
  return p.Cin == 3 && p.Cout == 64 && p.KH == 3 && p.KW == 3 && p.stride == 2 && p.pad == 1 && op.dtype_in == CAPF_F32 &&
         (op.dtype_out == CAPF_F16 || op.dtype_out == CAPF_BF16) && !op.in[3] && p.act != CAPF_ACT_GELU && p.N <= 65535;
}

static int launch_stem(const ConvP& p, const capf_op& op, cudaStream_t st) {
  dim3 g((p.Wo + 15) / 16, (p.Ho + 15) / 16, p.N);
  const float *x = (const float*)op.in[0], *w = (const float*)op.in[1], *b = (const float*)op.in[2];
  const int relu = p.act == CAPF_ACT_RELU;
  if (op.dtype_out == CAPF_F16)
    launch_k(stem_conv3x3s2_c3_kernel<__half>, dim3(g), dim3(256), 0, st, p.H, p.W, p.Ho, p.Wo, relu, x, w, b, (__half*)op.out[0]);
  else
    launch_k(stem_conv3x3s2_c3_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, st, p.H, p.W, p.Ho, p.Wo, relu, x, w, b, (__nv_bfloat16*)op.out[0]);
  return check_launch("stem_conv3x3s2_c3");
}

int launch_conv_simt(const capf_op& op, cudaStream_t st) {
  ConvP p;
  p.N = op.i[0]; p.H = op.i[1]; p.W = op.i[2]; p.Cin = op.i[3]; p.Cout = op.i[4];
  p.KH = op.i[5]; p.KW = op.i[6]; p.stride = op.i[7]; p.pad = op.i[8]; p.Ho = op.i[9]; p.Wo = op.i[10];
  p.act = op.i[11];
  long long M = (long long)p.N * p.Ho * p.Wo;
  if (M <= 0 || M >= (1ll << 31) || p.Cin <= 0 || p.Cout <= 0) return set_error(CAPF_ERR_ARG, "conv2d: bad shape");
  if (p.Ho != (p.H + 2 * p.pad - p.KH) / p.stride + 1 || p.Wo != (p.W + 2 * p.pad - p.KW) / p.stride + 1)
    return set_error(CAPF_ERR_ARG, "conv2d: Ho/Wo inconsistent with H,W,k,stride,pad");
  p.M = (int)M;
  p.K = p.KH * p.KW * p.Cin;
  if (!op.in[0] || !op.in[1] || !op.out[0]) return set_error(CAPF_ERR_ARG, "conv2d: null pointer");
  int di = op.dtype_in, dd = op.dtype_out;
  if (stem_tc_supported(op)) return launch_stem_tc(op, st);   // tensor-pipe stem (capf_stem.cu); env CAPF_STEM_TC=0 -> below
  if (is_stem(p, op)) return launch_stem(p, op, st);
  if (di == CAPF_F32 && dd == CAPF_F32) return conv_dispatch<float, float, float>(p, op, st);
  if (di == CAPF_F32 && dd == CAPF_F16) return conv_dispatch<float, float, __half>(p, op, st);
  if (di == CAPF_F32 && dd == CAPF_BF16) return conv_dispatch<float, float, __nv_bfloat16>(p, op, st);
  if (di == CAPF_F16 && dd == CAPF_F16) return conv_dispatch<__half, __half, __half>(p, op, st);
  if (di == CAPF_F16 && dd == CAPF_F32) return conv_dispatch<__half, __half, float>(p, op, st);
  if (di == CAPF_BF16 && dd == CAPF_BF16) return conv_dispatch<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(p, op, st);
  if (di == CAPF_BF16 && dd == CAPF_F32) return conv_dispatch<__nv_bfloat16, __nv_bfloat16, float>(p, op, st);
  return set_error(CAPF_ERR_UNSUPPORTED, "conv2d: dtype combination");
}

// =======================================================================================================
// HRNet fuse: y = relu(sum_t up_{2^s_t}(term_t))   (pose_hrnet.py:294-301, nearest upsample :244)
// =======================================================================================================
struct FuseP {
  int N, H, W, C, nt, relu;
  int sh[4];
  const void* t[4];
};

// 16 bytes (8 x 16-bit or 4 x fp32 elements) per thread, 32-bit index arithmetic.
template <typename T> struct Vec16B;
template <> struct Vec16B<float> {
  static constexpr int E = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec16B<__half> {
  static constexpr int E = 8;
  float v[8];
  __device__ __forceinline__ void load(const __half* p) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int e = 0; e < 4; ++e) { float2 f = __half22float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  }
  __device__ __forceinline__ void store(__half* p) const {
    uint4 t;
    __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};
template <> struct Vec16B<__nv_bfloat16> {
  static constexpr int E = 8;
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// Every block owns a contiguous range of output rows (n, y) and every thread a fixed 16-byte slot of the row (for the HRNet
// shapes a row is exactly 256 vectors: one per thread), so the per-vector work is one 32-bit multiply-add per term for the
// address, the loads, the fp32 sum and one store -- the first version spent three integer divisions and four 64-bit address
// chains per vector and was bound by instruction issue (ncu: SM 58 %, DRAM 34 %, profiles/r2u_ncu_full_summary.md).
template <typename T>
__global__ void __launch_bounds__(256) fuse_sum_kernel(FuseP p, T* __restrict__ y) {
  pdl_wait();
  constexpr int E = Vec16B<T>::E;
  const int CV = p.C / E;                                     // 16-byte vectors per pixel
  const int rowv = p.W * CV;                                  // ... per output row
  const int rows = p.N * p.H;
  const int r0 = (int)(((long long)rows * blockIdx.x) / gridDim.x);
  const int r1 = (int)(((long long)rows * (blockIdx.x + 1)) / gridDim.x);
  if (r0 >= r1) return;
  const int n0 = r0 / p.H, y0 = r0 - n0 * p.H;
  uint32_t hs[4], rs[4];                                      // rows per image and vectors per row of term t
#pragma unroll
  for (int t = 0; t < 4; ++t) { hs[t] = (uint32_t)(p.H >> p.sh[t]); rs[t] = (uint32_t)((p.W >> p.sh[t]) * CV); }
  for (int i = threadIdx.x; i < rowv; i += blockDim.x) {
    const int xx = i / CV, cv = i - xx * CV;
    uint32_t col[4];                                          // this thread's vector inside a source row of term t
#pragma unroll
    for (int t = 0; t < 4; ++t) col[t] = (uint32_t)((xx >> p.sh[t]) * CV + cv);
    int n = n0, yy = y0;
    uint4* dst = reinterpret_cast<uint4*>(y) + (size_t)r0 * rowv + i;
    for (int r = r0; r < r1; ++r, dst += rowv) {
      Vec16B<T> a;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t < p.nt) {
          const uint32_t off = ((uint32_t)n * hs[t] + (uint32_t)(yy >> p.sh[t])) * rs[t] + col[t];
          Vec16B<T> v;
          v.load(reinterpret_cast<const T*>(reinterpret_cast<const uint4*>(p.t[t]) + off));
          // first term initialises (reference: y = x[0] ... then y = y + term)
#pragma unroll
          for (int e = 0; e < E; ++e) a.v[e] = t == 0 ? v.v[e] : a.v[e] + v.v[e];
        }
      }
      if (p.relu) {
#pragma unroll
        for (int e = 0; e < E; ++e) a.v[e] = fmaxf(a.v[e], 0.f);
      }
      a.store(reinterpret_cast<T*>(dst));
      if (++yy == p.H) { yy = 0; ++n; }
    }
  }
}

int launch_fuse_sum(const capf_op& op, cudaStream_t st) {
  FuseP p;
  p.N = op.i[0]; p.H = op.i[1]; p.W = op.i[2]; p.C = op.i[3]; p.nt = op.i[4]; p.relu = op.i[9];
  if (p.nt < 1 || p.nt > 4 || (p.C & 3)) return set_error(CAPF_ERR_ARG, "fuse_sum: 1..4 terms, C%4==0");
  for (int t = 0; t < 4; ++t) {
    p.sh[t] = op.i[5 + t];
    p.t[t] = t < p.nt ? op.in[t] : nullptr;
    if (t < p.nt && (!p.t[t] || p.sh[t] < 0 || (p.H & ((1 << p.sh[t]) - 1)) || (p.W & ((1 << p.sh[t]) - 1))))
      return set_error(CAPF_ERR_ARG, "fuse_sum: bad term");
  }
  if (op.dtype_in != op.dtype_out) return set_error(CAPF_ERR_UNSUPPORTED, "fuse_sum: dtype_in != dtype_out");
  const int E = op.dtype_out == CAPF_F32 ? 4 : 8;              // elements per 16-byte vector
  if (p.C % E) return set_error(CAPF_ERR_ARG, "fuse_sum: C must be a multiple of 16 bytes of elements");
  size_t total = (size_t)p.N * p.H * p.W * (p.C / E);
  if (total >= (1ull << 31)) return set_error(CAPF_ERR_UNSUPPORTED, "fuse_sum: tensor too large for 32-bit indexing");
  for (int t = 0; t < p.nt; ++t)
    if ((uintptr_t)p.t[t] & 15) return set_error(CAPF_ERR_ARG, "fuse_sum: terms must be 16-byte aligned");
  const long long rows = (long long)p.N * p.H;                 // a block owns whole output rows
  int blocks = (int)(rows < (long long)num_sms() * 16 ? rows : (long long)num_sms() * 16);
  if (blocks < 1) blocks = 1;
  switch (op.dtype_out) {
    case CAPF_F32: launch_k(fuse_sum_kernel<float>, dim3(blocks), dim3(256), 0, st, p, (float*)op.out[0]); break;
    case CAPF_F16: launch_k(fuse_sum_kernel<__half>, dim3(blocks), dim3(256), 0, st, p, (__half*)op.out[0]); break;
    case CAPF_BF16: launch_k(fuse_sum_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, p, (__nv_bfloat16*)op.out[0]); break;
    default: return set_error(CAPF_ERR_UNSUPPORTED, "fuse_sum: dtype");
  }
  return check_launch("fuse_sum");
}

// =======================================================================================================
// CPN helpers: MaxPool2d(3,2,1) (networks/resnet.py:105), bilinear align_corners=True resize
// (networks/globalNet.py:40, networks/refineNet.py:61; ATen UpSample.h area_pixel_compute_source_index)
// =======================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(int N, int H, int W, int C, int Ho, int Wo,
                                                           const T* __restrict__ x, T* __restrict__ y) {
  pdl_wait();
  const int C4 = C >> 2;
  size_t total = (size_t)N * Ho * Wo * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    size_t pix = i / C4;
    int ox = (int)(pix % Wo);
    size_t t2 = pix / Wo;
    int oy = (int)(t2 % Ho), n = (int)(t2 / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int iy = oy * 2 - 1 + r;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int ix = ox * 2 - 1 + s;
        if (ix < 0 || ix >= W) continue;
        float4 v = ld4<T>(x + (((size_t)n * H + iy) * W + ix) * C + c4 * 4);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    st4<T>(y + i * 4, m);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bilinear_ac_kernel(int N, int H, int W, int C, int Ho, int Wo, float sy, float sx,
                                                          const T* __restrict__ x, T* __restrict__ y) {
  pdl_wait();
  const int C4 = C >> 2;
  size_t total = (size_t)N * Ho * Wo * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    size_t pix = i / C4;
    int ox = (int)(pix % Wo);
    size_t t2 = pix / Wo;
    int oy = (int)(t2 % Ho), n = (int)(t2 / Ho);
    float fy = __fmul_rn(sy, (float)oy), fx = __fmul_rn(sx, (float)ox);
    int y0 = (int)fy, x0 = (int)fx;
    int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    float ly = fy - (float)y0, lx = fx - (float)x0;
    float hy = 1.f - ly, hx = 1.f - lx;
    const T* b = x + (size_t)n * H * W * C + c4 * 4;
    float4 v00 = ld4<T>(b + ((size_t)y0 * W + x0) * C), v01 = ld4<T>(b + ((size_t)y0 * W + x1) * C);
    float4 v10 = ld4<T>(b + ((size_t)y1 * W + x0) * C), v11 = ld4<T>(b + ((size_t)y1 * W + x1) * C);
    float4 o;
    o.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    o.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    o.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
    o.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    st4<T>(y + i * 4, o);
  }
}

// 16-bit variant: 8 channels (16 bytes) per thread, 32-bit index arithmetic, one thread per output element group.

template <typename T>
__global__ void __launch_bounds__(256) bilinear_ac16_kernel(int N, int H, int W, int C8, int Ho, int Wo, float sy, float sx,
                                                            const uint4* __restrict__ x, uint4* __restrict__ y, unsigned total) {
  pdl_wait();
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= total) return;
  const unsigned c8 = i % (unsigned)C8, pix = i / (unsigned)C8;
  const int ox = (int)(pix % (unsigned)Wo);
  const unsigned t2 = pix / (unsigned)Wo;
  const int oy = (int)(t2 % (unsigned)Ho), n = (int)(t2 / (unsigned)Ho);
  const float fy = __fmul_rn(sy, (float)oy), fx = __fmul_rn(sx, (float)ox);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const uint4* b = x + (size_t)n * H * W * C8 + c8;
  float v00[8], v01[8], v10[8], v11[8], o[8];
  unpack8<T>(__ldg(b + (size_t)(y0 * W + x0) * C8), v00);
  unpack8<T>(__ldg(b + (size_t)(y0 * W + x1) * C8), v01);
  unpack8<T>(__ldg(b + (size_t)(y1 * W + x0) * C8), v10);
  unpack8<T>(__ldg(b + (size_t)(y1 * W + x1) * C8), v11);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = hy * (hx * v00[e] + lx * v01[e]) + ly * (hx * v10[e] + lx * v11[e]);
  y[i] = pack8<T>(o);
}

static int ew_blocks(size_t total) {
  size_t b = (total + 255) / 256, cap = (size_t)num_sms() * 16;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

int launch_maxpool(const capf_op& op, cudaStream_t st) {
  int N = op.i[0], H = op.i[1], W = op.i[2], C = op.i[3], Ho = op.i[4], Wo = op.i[5];
  if ((C & 3) || Ho != (H + 2 - 3) / 2 + 1 || Wo != (W + 2 - 3) / 2 + 1 || op.dtype_in != op.dtype_out)
    return set_error(CAPF_ERR_ARG, "maxpool: bad shape/dtype");
  int blocks = ew_blocks((size_t)N * Ho * Wo * (C / 4));
  switch (op.dtype_in) {
    case CAPF_F32: launch_k(maxpool3x3s2_kernel<float>, dim3(blocks), dim3(256), 0, st, N, H, W, C, Ho, Wo, (const float*)op.in[0], (float*)op.out[0]); break;
    case CAPF_F16: launch_k(maxpool3x3s2_kernel<__half>, dim3(blocks), dim3(256), 0, st, N, H, W, C, Ho, Wo, (const __half*)op.in[0], (__half*)op.out[0]); break;
    case CAPF_BF16: launch_k(maxpool3x3s2_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, N, H, W, C, Ho, Wo, (const __nv_bfloat16*)op.in[0], (__nv_bfloat16*)op.out[0]); break;
    default: return set_error(CAPF_ERR_UNSUPPORTED, "maxpool: dtype");
  }
  return check_launch("maxpool3x3s2");
}

int launch_bilinear(const capf_op& op, cudaStream_t st) {
  int N = op.i[0], H = op.i[1], W = op.i[2], C = op.i[3], Ho = op.i[4], Wo = op.i[5];
  if ((C & 3) || op.dtype_in != op.dtype_out) return set_error(CAPF_ERR_ARG, "bilinear: bad shape/dtype");
  // ATen area_pixel_compute_scale(align_corners=True): (in-1)/(out-1), 0 when out == 1
  float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  float sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const size_t tot8 = (size_t)N * Ho * Wo * (C / 8);
  if (op.dtype_in != CAPF_F32 && C % 8 == 0 && tot8 < (1ull << 31) && (size_t)N * H * W * C < (1ull << 31) &&
      !(((uintptr_t)op.in[0] | (uintptr_t)op.out[0]) & 15)) {
    const unsigned nb = (unsigned)((tot8 + 255) / 256);
    if (op.dtype_in == CAPF_F16)
      launch_k(bilinear_ac16_kernel<__half>, dim3(nb), dim3(256), 0, st, N, H, W, C / 8, Ho, Wo, sy, sx, (const uint4*)op.in[0], (uint4*)op.out[0], (unsigned)tot8);
    else
      launch_k(bilinear_ac16_kernel<__nv_bfloat16>, dim3(nb), dim3(256), 0, st, N, H, W, C / 8, Ho, Wo, sy, sx, (const uint4*)op.in[0], (uint4*)op.out[0], (unsigned)tot8);
    return check_launch("bilinear_ac16");
  }
  int blocks = ew_blocks((size_t)N * Ho * Wo * (C / 4));
  switch (op.dtype_in) {
    case CAPF_F32: launch_k(bilinear_ac_kernel<float>, dim3(blocks), dim3(256), 0, st, N, H, W, C, Ho, Wo, sy, sx, (const float*)op.in[0], (float*)op.out[0]); break;
    case CAPF_F16: launch_k(bilinear_ac_kernel<__half>, dim3(blocks), dim3(256), 0, st, N, H, W, C, Ho, Wo, sy, sx, (const __half*)op.in[0], (__half*)op.out[0]); break;
    case CAPF_BF16: launch_k(bilinear_ac_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, N, H, W, C, Ho, Wo, sy, sx, (const __nv_bfloat16*)op.in[0], (__nv_bfloat16*)op.out[0]); break;
    default: return set_error(CAPF_ERR_UNSUPPORTED, "bilinear: dtype");
  }
  return check_launch("bilinear_ac");
}

// =======================================================================================================
// LayerNorm over the last dim, one warp per row (pose_dformer.py:65,72,120,138,206)
// =======================================================================================================
template <typename TO, int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(int rows, int D, int period, float eps,
                                                        const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ x0,
                                                        TO* __restrict__ y) {
  pdl_wait();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (size_t)warp * D;
  const float* ar = period > 0 ? x0 + (size_t)(warp % period) * D : nullptr;
  const int nv = (D + 127) >> 7;  // float4 slots per lane; lanes past the row end (D % 128 != 0) contribute nothing
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nv && (i * 32 + lane) * 4 < D) {
      int c = (i * 32 + lane) * 4;
      float4 t = __ldg(reinterpret_cast<const float4*>(xr + c));
      if (ar) {
        float4 a = __ldg(reinterpret_cast<const float4*>(ar + c));
        t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
      }
      v[i] = t;
      s += (t.x + t.y) + (t.z + t.w);
    }
  }
  float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv && (i * 32 + lane) * 4 < D) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv && (i * 32 + lane) * 4 < D) {
      int c = (i * 32 + lane) * 4;
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      st4<TO>(y + (size_t)warp * D + c, o);
    }
  }
}

// Narrow rows (D <= 128, the 128-wide LayerNorms of the context / res blocks: 16 of the 24 launches of a step): one
// float4 per lane covers a row, so a warp takes RPW rows at once -- RPW independent loads and RPW interleaved shuffle
// reductions in flight instead of one dependent chain per warp (the one-row version ran at ~1.3 TB/s, latency-bound).
template <typename TO, int RPW>
__global__ void __launch_bounds__(256) layernorm_narrow_kernel(int rows, int D, int period, float eps, const float* __restrict__ x,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               const float* __restrict__ x0, TO* __restrict__ y) {
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int row0 = warp * RPW;
  if (row0 >= rows) return;
  const int c = lane * 4;
  const bool on = c < D;
  float4 v[RPW];
  float s[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int row = row0 + r;
    if (on && row < rows) {
      v[r] = __ldg(reinterpret_cast<const float4*>(x + (size_t)row * D + c));
      if (period > 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x0 + (size_t)(row % period) * D + c));
        v[r].x += a.x; v[r].y += a.y; v[r].z += a.z; v[r].w += a.w;
      }
    }
    s[r] = (v[r].x + v[r].y) + (v[r].z + v[r].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
  }
  float mean[RPW], q[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    mean[r] = s[r] / (float)D;
    const float a = v[r].x - mean[r], b = v[r].y - mean[r], cc = v[r].z - mean[r], d = v[r].w - mean[r];
    q[r] = on ? (a * a + b * b) + (cc * cc + d * d) : 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int r = 0; r < RPW; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
  }
  if (!on) return;
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c));
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = row0 + r;
    if (row < rows) {
      const float rstd = rsqrtf(q[r] / (float)D + eps);
      float4 o;
      o.x = (v[r].x - mean[r]) * rstd * g.x + bb.x;
      o.y = (v[r].y - mean[r]) * rstd * g.y + bb.y;
      o.z = (v[r].z - mean[r]) * rstd * g.z + bb.z;
      o.w = (v[r].w - mean[r]) * rstd * g.w + bb.w;
      st4<TO>(y + (size_t)row * D + c, o);
    }
  }
}

// LayerNorm + narrow Linear (the head, pose_dformer.py:205-208,240): one warp per row, everything in fp32.
template <int MAXV, int MAXP>
__global__ void __launch_bounds__(256) layernorm_proj_kernel(int rows, int D, int nproj, float eps, const float* __restrict__ x,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ Wp, const float* __restrict__ bp,
                                                             float* __restrict__ y) {
  pdl_wait();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (size_t)warp * D;
  const int nv = (D + 127) >> 7;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nv && (i * 32 + lane) * 4 < D) {
      v[i] = __ldg(reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4));
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv && (i * 32 + lane) * 4 < D) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  float acc[MAXP];
#pragma unroll
  for (int o = 0; o < MAXP; ++o) acc[o] = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv && (i * 32 + lane) * 4 < D) {
      const int c = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 n;
      n.x = (v[i].x - mean) * rstd * g.x + b.x;
      n.y = (v[i].y - mean) * rstd * g.y + b.y;
      n.z = (v[i].z - mean) * rstd * g.z + b.z;
      n.w = (v[i].w - mean) * rstd * g.w + b.w;
#pragma unroll
      for (int o = 0; o < MAXP; ++o) {
        if (o < nproj) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(Wp + (size_t)o * D + c));
          acc[o] = fmaf(n.x, w.x, fmaf(n.y, w.y, fmaf(n.z, w.z, fmaf(n.w, w.w, acc[o]))));
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < MAXP; ++o) {
    if (o < nproj) {
      const float t = warp_sum(acc[o]);
      if (lane == 0) y[(size_t)warp * nproj + o] = t + (bp ? __ldg(bp + o) : 0.f);
    }
  }
}

int launch_layernorm(const capf_op& op, cudaStream_t st) {
  int rows = op.i[0], D = op.i[1], period = op.i[2];
  if (op.i[3] > 0) {
    const int np = op.i[3];
    if (np > 8 || period != 0 || op.dtype_in != CAPF_F32 || op.dtype_out != CAPF_F32 || !op.in[4] || (D & 3) || D > 1024 || rows <= 0)
      return set_error(CAPF_ERR_ARG, "layernorm+proj: n_proj <= 8, no x0, f32 in/out, D % 4 == 0, D <= 1024");
    launch_k(layernorm_proj_kernel<8, 8>, dim3((rows + 7) / 8), dim3(256), 0, st, rows, D, np, op.f[0], (const float*)op.in[0],
             (const float*)op.in[1], (const float*)op.in[2], (const float*)op.in[4], (const float*)op.in[5], (float*)op.out[0]);
    return check_launch("layernorm_proj");
  }
  if (rows <= 0 || D <= 0 || (D & 3) || D > 128 * 8) return set_error(CAPF_ERR_ARG, "layernorm: D must be a multiple of 4, <= 1024");
  if (op.dtype_in != CAPF_F32) return set_error(CAPF_ERR_UNSUPPORTED, "layernorm: input stream is f32");
  if (period > 0 && !op.in[3]) return set_error(CAPF_ERR_ARG, "layernorm: period without x0");
  int blocks = (rows + 7) / 8;
  const float *x = (const float*)op.in[0], *g = (const float*)op.in[1], *b = (const float*)op.in[2], *x0 = (const float*)op.in[3];
  if (D <= 128) {            // same arithmetic (sum / variance association per lane, shuffle order) as the one-row kernel
    constexpr int RPW = 4;
    const int nb = (rows + 8 * RPW - 1) / (8 * RPW);
    switch (op.dtype_out) {
      case CAPF_F32: launch_k(layernorm_narrow_kernel<float, RPW>, dim3(nb), dim3(256), 0, st, rows, D, period, op.f[0], x, g, b, x0, (float*)op.out[0]); break;
      case CAPF_F16: launch_k(layernorm_narrow_kernel<__half, RPW>, dim3(nb), dim3(256), 0, st, rows, D, period, op.f[0], x, g, b, x0, (__half*)op.out[0]); break;
      case CAPF_BF16: launch_k(layernorm_narrow_kernel<__nv_bfloat16, RPW>, dim3(nb), dim3(256), 0, st, rows, D, period, op.f[0], x, g, b, x0, (__nv_bfloat16*)op.out[0]); break;
      default: return set_error(CAPF_ERR_UNSUPPORTED, "layernorm: dtype_out");
    }
    return check_launch("layernorm_narrow");
  }
  switch (op.dtype_out) {
    case CAPF_F32: launch_k(layernorm_kernel<float, 8>, dim3(blocks), dim3(256), 0, st, rows, D, period, op.f[0], x, g, b, x0, (float*)op.out[0]); break;
    case CAPF_F16: launch_k(layernorm_kernel<__half, 8>, dim3(blocks), dim3(256), 0, st, rows, D, period, op.f[0], x, g, b, x0, (__half*)op.out[0]); break;
    case CAPF_BF16: launch_k(layernorm_kernel<__nv_bfloat16, 8>, dim3(blocks), dim3(256), 0, st, rows, D, period, op.f[0], x, g, b, x0, (__nv_bfloat16*)op.out[0]); break;
    default: return set_error(CAPF_ERR_UNSUPPORTED, "layernorm: dtype_out");
  }
  return check_launch("layernorm");
}

// =======================================================================================================
// Tiny-sequence attention (pose_dformer.py:47-55): seq 5 (levels of one joint) or 17 (joints of a frame).
// One lane per query token; a warp packs floor(32/SEQ) (group, head) items; K/V/Q staged in shared memory.
// =======================================================================================================
template <typename TI, typename TO, int SEQ>
__global__ void __launch_bounds__(128) attention_small_kernel(int groups, int heads, int hd, int tok_stride, int grp_stride,
                                                              float scale, const TI* __restrict__ qkv, TO* __restrict__ out) {
  pdl_wait();
  constexpr int IPW = 32 / SEQ;  // items per warp
  extern __shared__ float sm[];
  const int warp_in_blk = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp = hd + 1;  // padded row pitch (bank-conflict free row-strided access)
  float* base = sm + (size_t)warp_in_blk * (IPW * SEQ * hp * 3);
  float* sq = base;
  float* sk = sq + IPW * SEQ * hp;
  float* sv = sk + IPW * SEQ * hp;
  const int D = heads * hd;
  const long long n_items = (long long)groups * heads;
  const long long item0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp_in_blk) * IPW;
  if (item0 >= n_items) return;

  // cooperative, coalesced staging of q/k/v rows
  for (int it = 0; it < IPW; ++it) {
    long long item = item0 + it;
    if (item >= n_items) break;
    int g = (int)(item / heads), h = (int)(item % heads);
    for (int t = 0; t < SEQ; ++t) {
      size_t row = (size_t)g * grp_stride + (size_t)t * tok_stride;
      const TI* src = qkv + row * (size_t)(3 * D) + h * hd;
      float* dq = sq + (it * SEQ + t) * hp;
      float* dk = sk + (it * SEQ + t) * hp;
      float* dv = sv + (it * SEQ + t) * hp;
      for (int d = lane; d < hd; d += 32) {
        dq[d] = to_f<TI>(src[d]);
        dk[d] = to_f<TI>(src[D + d]);
        dv[d] = to_f<TI>(src[2 * D + d]);
      }
    }
  }
  __syncwarp();

  const int it = lane / SEQ, i = lane - it * SEQ;
  const bool active = (it < IPW) && (item0 + it < n_items);
  if (active) {
    const float* q = sq + (it * SEQ + i) * hp;
    const float* kb = sk + it * SEQ * hp;
    const float* vb = sv + it * SEQ * hp;
    float sc[SEQ];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SEQ; ++j) {
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a = fmaf(q[d], kb[j * hp + d], a);
      a *= scale;
      sc[j] = a;
      mx = fmaxf(mx, a);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < SEQ; ++j) {
      sc[j] = __expf(sc[j] - mx);
      den += sc[j];
    }
    float inv = 1.0f / den;
    // write the output row over this lane's own q row (no other lane reads it any more)
    float* o = sq + (it * SEQ + i) * hp;
    for (int d = 0; d < hd; ++d) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < SEQ; ++j) a = fmaf(sc[j], vb[j * hp + d], a);
      o[d] = a * inv;
    }
  }
  __syncwarp();
  for (int it2 = 0; it2 < IPW; ++it2) {
    long long item = item0 + it2;
    if (item >= n_items) break;
    int g = (int)(item / heads), h = (int)(item % heads);
    for (int t = 0; t < SEQ; ++t) {
      size_t row = (size_t)g * grp_stride + (size_t)t * tok_stride;
      TO* dst = out + row * (size_t)D + h * hd;
      const float* o = sq + (it2 * SEQ + t) * hp;
      for (int d = lane; d < hd; d += 32) dst[d] = from_f<TO>(o[d]);
    }
  }
}

// ---- 16-bit fast paths -------------------------------------------------------------------------------------
// Attention over the 17 joints of a frame (pose_dformer.py:235-238; 8 heads x 80): one warp per (frame, head).  K and V
// (17 x 80, 16-bit) are staged in shared memory with 16-byte loads; lane i < 17 keeps query row i and its 80 fp32 output
// accumulators in registers and reads K / V rows as warp-wide broadcasts, so the inner loops are 8 FMAs per LDS.128.
template <typename T, int HDC>   // HDC = head_dim / 8
__global__ void __launch_bounds__(128) attention17_kernel(int groups, int heads, int grp_stride, float scale, const T* __restrict__ qkv,
                                                          T* __restrict__ out) {
  pdl_wait();
  constexpr int SEQ = 17, HD = HDC * 8;
  __shared__ uint4 skv[4][2][SEQ * HDC];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * 4 + wib;
  if (item >= (long long)groups * heads) return;
  const int g = (int)(item / heads), h = (int)(item % heads);
  const int D = heads * HD;
  const T* base = qkv + (size_t)g * grp_stride * (3 * D) + h * HD;       // token t: + t * 3D (tok_stride == 1)
  for (int c = lane; c < SEQ * HDC; c += 32) {
    const int t = c / HDC, k = c - t * HDC;
    const uint4* src = reinterpret_cast<const uint4*>(base + (size_t)t * (3 * D) + D) + k;
    skv[wib][0][c] = __ldg(src);
    skv[wib][1][c] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(src) + D));
  }
  const int i = lane < SEQ ? lane : SEQ - 1;                               // idle lanes shadow the last query (no divergence)
  uint4 q[HDC];
#pragma unroll
  for (int k = 0; k < HDC; ++k) q[k] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)i * (3 * D)) + k);
  __syncwarp();
  float sc[SEQ], mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < SEQ; ++j) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < HDC; ++k) {
      float qf[8], kf[8];
      unpack8<T>(q[k], qf);
      unpack8<T>(skv[wib][0][j * HDC + k], kf);
#pragma unroll
      for (int e = 0; e < 8; ++e) a = fmaf(qf[e], kf[e], a);
    }
    sc[j] = a * scale;
    mx = fmaxf(mx, sc[j]);
  }
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < SEQ; ++j) { sc[j] = __expf(sc[j] - mx); den += sc[j]; }
  const float inv = 1.0f / den;
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < SEQ; ++j) {
#pragma unroll
    for (int k = 0; k < HDC; ++k) {
      float vf[8];
      unpack8<T>(skv[wib][1][j * HDC + k], vf);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[8 * k + e] = fmaf(sc[j], vf[e], o[8 * k + e]);
    }
  }
  if (lane < SEQ) {
    uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)g * grp_stride + lane) * D + h * HD);
#pragma unroll
    for (int k = 0; k < HDC; ++k) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = o[8 * k + e] * inv;
      dst[k] = pack8<T>(f);
    }
  }
}

// ---- attention over the 17 joints of a frame on the warp-level tensor-core path (mma.sync m16n8k16) ----------------------------
// The kernel above spends ~9 k instructions per lane on 17 x 17 x 80 scalar FMAs and half->float conversions (30 us per joint
// block at bs = 256, instruction-bound with 15 of 32 lanes idle).  Here one warp still owns a (frame, head) pair, but
//   S = Q K^T   is 2 x 3 x 5 MMAs (queries padded to 32 rows, keys to 24; padded rows re-read row 16, so no value is ever NaN),
//   softmax     runs on the accumulator fragments (a row lives in the 4 lanes of a quad: two shuffles per reduction),
//   O = P V     re-uses the S fragments as the A operand (two 16 x 8 C tiles = one 16 x 16 A tile) with P SPLIT into
//               hi = T(p) and lo = T(p - hi): two MMAs per tile keep P at fp32-class precision (2^-22 / 2^-16 for fp16 / bf16),
//               so the result matches the fp32-math kernel to the output rounding; V fragments come from ldmatrix.trans.
// tcgen05 is the wrong tool for 17 x 17 problems (M = 128 minimum); this is the one place the legacy warp MMA is the right size.
template <typename T> struct Mma16816;
template <> struct Mma16816<__half> {
  static __device__ __forceinline__ void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
  static __device__ __forceinline__ uint32_t pack2(float x, float y) { const __half2 h = __floats2half2_rn(x, y); return *reinterpret_cast<const uint32_t*>(&h); }
};
template <> struct Mma16816<__nv_bfloat16> {
  static __device__ __forceinline__ void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
  static __device__ __forceinline__ uint32_t pack2(float x, float y) { const __nv_bfloat162 h = __floats2bfloat162_rn(x, y); return *reinterpret_cast<const uint32_t*>(&h); }
};

template <typename T>
__global__ void __launch_bounds__(128) attention17_mma_kernel(int groups, int heads, int grp_stride, float scale, const T* __restrict__ qkv,
                                                              T* __restrict__ out) {
  pdl_wait();
  constexpr int SEQ = 17, HD = 80, LD = 88;      // rows padded to 176 bytes: the 32-bit fragment loads of a quad-row pattern hit 32 distinct banks
  __shared__ __align__(16) T sq[4][SEQ * LD];
  __shared__ __align__(16) T sk[4][SEQ * LD];
  __shared__ __align__(16) T sv[4][SEQ * LD];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * 4 + wib;
  if (item >= (long long)groups * heads) return;
  const int g = (int)(item / heads), h = (int)(item % heads);
  const int D = heads * HD;
  const T* base = qkv + (size_t)g * grp_stride * (3 * D) + h * HD;       // token t: + t * 3D (tok_stride == 1)
  T* const q_s = sq[wib];
  T* const k_s = sk[wib];
  T* const v_s = sv[wib];
  for (int c = lane; c < SEQ * (HD / 8); c += 32) {
    const int t = c / (HD / 8), k = c - t * (HD / 8);
    const T* src = base + (size_t)t * (3 * D) + 8 * k;
    *reinterpret_cast<uint4*>(q_s + t * LD + 8 * k) = __ldg(reinterpret_cast<const uint4*>(src));
    *reinterpret_cast<uint4*>(k_s + t * LD + 8 * k) = __ldg(reinterpret_cast<const uint4*>(src + D));
    *reinterpret_cast<uint4*>(v_s + t * LD + 8 * k) = __ldg(reinterpret_cast<const uint4*>(src + 2 * D));
  }
  __syncwarp();
  const int gid = lane >> 2, t4 = lane & 3;
  auto ld32 = [](const T* p) { return *reinterpret_cast<const uint32_t*>(p); };
  // ---- S = Q K^T: s[mt][nt] = rows 16 mt + {gid, gid + 8}, keys 8 nt + 2 t4 + {0, 1}
  float s[2][3][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t a[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r0 = min(16 * mt + gid, SEQ - 1), r1 = min(16 * mt + gid + 8, SEQ - 1);
      a[mt][0] = ld32(q_s + r0 * LD + 16 * kk + 2 * t4);
      a[mt][1] = ld32(q_s + r1 * LD + 16 * kk + 2 * t4);
      a[mt][2] = ld32(q_s + r0 * LD + 16 * kk + 8 + 2 * t4);
      a[mt][3] = ld32(q_s + r1 * LD + 16 * kk + 8 + 2 * t4);
    }
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const int n = min(8 * nt + gid, SEQ - 1);
      const uint32_t b0 = ld32(k_s + n * LD + 16 * kk + 2 * t4), b1 = ld32(k_s + n * LD + 16 * kk + 8 + 2 * t4);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) Mma16816<T>::run(s[mt][nt], a[mt], b0, b1);
    }
  }
  // ---- softmax over the 17 keys of every row; P fragments (hi | lo) for the second GEMM
  uint32_t phi[2][4][2], plo[2][4][2];            // [mt][key tile 0..3][row half]: keys 8 nt + 2 t4 + {0, 1}
  float inv[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      float v[3][2], mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool ok = 8 * nt + 2 * t4 + e < SEQ;
          v[nt][e] = ok ? s[mt][nt][2 * hr + e] * scale : -INFINITY;
          mx = fmaxf(mx, v[nt][e]);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float den = 0.f;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) { v[nt][e] = __expf(v[nt][e] - mx); den += v[nt][e]; }      // exp(-inf) = 0 for the padded keys
      den += __shfl_xor_sync(0xffffffffu, den, 1);
      den += __shfl_xor_sync(0xffffffffu, den, 2);
      inv[mt][hr] = 1.0f / den;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) Mma16816<T>::split2(v[nt][0], v[nt][1], phi[mt][nt][hr], plo[mt][nt][hr]);
      phi[mt][3][hr] = 0u;
      plo[mt][3][hr] = 0u;
    }
  }
  // ---- O = P V: o[mt][ct] = rows 16 mt + {gid, gid + 8}, channels 8 ct + 2 t4 + {0, 1}
  float o[2][HD / 8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ct = 0; ct < HD / 8; ++ct)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][ct][e] = 0.f;
  const uint32_t v_addr = (uint32_t)__cvta_generic_to_shared(v_s);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      ah[mt][0] = phi[mt][2 * ks][0]; ah[mt][1] = phi[mt][2 * ks][1]; ah[mt][2] = phi[mt][2 * ks + 1][0]; ah[mt][3] = phi[mt][2 * ks + 1][1];
      al[mt][0] = plo[mt][2 * ks][0]; al[mt][1] = plo[mt][2 * ks][1]; al[mt][2] = plo[mt][2 * ks + 1][0]; al[mt][3] = plo[mt][2 * ks + 1][1];
    }
    // ldmatrix row address of this lane: lanes 0-7 = keys 16 ks + lane, lanes 8-15 = keys 16 ks + 8 + (lane - 8); padded keys re-read row 16 (P = 0 there)
    const int key = min(16 * ks + (lane & 15), SEQ - 1);
    const uint32_t row_addr = v_addr + (uint32_t)(key * LD) * (uint32_t)sizeof(T);
#pragma unroll
    for (int ct = 0; ct < HD / 8; ++ct) {
      uint32_t b0, b1;
      asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(row_addr + 16u * ct));
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        Mma16816<T>::run(o[mt][ct], ah[mt], b0, b1);
        Mma16816<T>::run(o[mt][ct], al[mt], b0, b1);
      }
    }
  }
  // ---- out[row][h * 80 + channel] = o / den
  T* const obase = out + (size_t)g * grp_stride * D + h * HD;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int row = 16 * mt + gid + 8 * hr;
      if (row < SEQ) {
        const float sc = inv[mt][hr];
#pragma unroll
        for (int ct = 0; ct < HD / 8; ++ct)
          *reinterpret_cast<uint32_t*>(obase + (size_t)row * D + 8 * ct + 2 * t4) = Mma16816<T>::pack2(o[mt][ct][2 * hr] * sc, o[mt][ct][2 * hr + 1] * sc);
      }
    }
  }
}

// Attention over the 5 level tokens of one joint (pose_dformer.py:231-234; 8 heads x 16): one LANE per (joint, head), a
// warp covers 4 joints.  q/k/v of a lane are 15 x 32 bytes, loaded straight into registers (the 8 heads of a token are
// 256 contiguous bytes, so a warp's loads are full sectors); no shared memory, no synchronisation.
template <typename T>
__global__ void __launch_bounds__(128) attention5_kernel(int groups, int tok_stride, int grp_stride, float scale, const T* __restrict__ qkv,
                                                         T* __restrict__ out) {
  pdl_wait();
  constexpr int SEQ = 5, D = 128;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long g = warp * 4 + (lane >> 3);
  if (g >= groups) return;
  const int h = lane & 7;
  uint4 qv[SEQ][2], kv[SEQ][2], vv[SEQ][2];
#pragma unroll
  for (int t = 0; t < SEQ; ++t) {
    const uint4* src = reinterpret_cast<const uint4*>(qkv + ((size_t)g * grp_stride + (size_t)t * tok_stride) * (3 * D) + h * 16);
    qv[t][0] = __ldg(src); qv[t][1] = __ldg(src + 1);
    kv[t][0] = __ldg(src + D / 8); kv[t][1] = __ldg(src + D / 8 + 1);
    vv[t][0] = __ldg(src + 2 * D / 8); vv[t][1] = __ldg(src + 2 * D / 8 + 1);
  }
#pragma unroll
  for (int i = 0; i < SEQ; ++i) {
    float qf[16];
    { float a[8], b[8]; unpack8<T>(qv[i][0], a); unpack8<T>(qv[i][1], b);
#pragma unroll
      for (int e = 0; e < 8; ++e) { qf[e] = a[e]; qf[8 + e] = b[e]; } }
    float sc[SEQ], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SEQ; ++j) {
      float a[8], b[8], acc = 0.f;
      unpack8<T>(kv[j][0], a); unpack8<T>(kv[j][1], b);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(qf[e], a[e], acc);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(qf[8 + e], b[e], acc);
      sc[j] = acc * scale;
      mx = fmaxf(mx, sc[j]);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < SEQ; ++j) { sc[j] = __expf(sc[j] - mx); den += sc[j]; }
    const float inv = 1.0f / den;
    float o[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < SEQ; ++j) {
      float a[8], b[8];
      unpack8<T>(vv[j][0], a); unpack8<T>(vv[j][1], b);
#pragma unroll
      for (int e = 0; e < 8; ++e) { o[e] = fmaf(sc[j], a[e], o[e]); o[8 + e] = fmaf(sc[j], b[e], o[8 + e]); }
    }
    float f0[8], f1[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { f0[e] = o[e] * inv; f1[e] = o[8 + e] * inv; }
    uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)g * grp_stride + (size_t)i * tok_stride) * D + h * 16);
    dst[0] = pack8<T>(f0);
    dst[1] = pack8<T>(f1);
  }
}

template <typename T>
static int attention_fast(const capf_op& op, cudaStream_t st, bool& taken) {
  const int groups = op.i[0], seq = op.i[1], heads = op.i[2], hd = op.i[3], ts = op.i[4], gs = op.i[5];
  taken = false;
  { const char* ev = getenv("CAPF_ATTN_FAST"); if (ev && ev[0] == '0') return CAPF_OK; }
  if ((((uintptr_t)op.in[0]) | ((uintptr_t)op.out[0])) & 15) return CAPF_OK;
  const T* qkv = (const T*)op.in[0];
  T* out = (T*)op.out[0];
  if (seq == 17 && hd == 80 && ts == 1) {
    const long long items = (long long)groups * heads;
    static const bool use_mma = []() { const char* ev = getenv("CAPF_ATTN_MMA"); return !(ev && ev[0] == '0'); }();
    if (use_mma) {
      launch_k(attention17_mma_kernel<T>, dim3((unsigned)((items + 3) / 4)), dim3(128), 0, st, groups, heads, gs, op.f[0], qkv, out);
      taken = true;
      return check_launch("attention17_mma");
    }
    launch_k(attention17_kernel<T, 10>, dim3((unsigned)((items + 3) / 4)), dim3(128), 0, st, groups, heads, gs, op.f[0], qkv, out);
    taken = true;
    return check_launch("attention17");
  }
  if (seq == 5 && hd == 16 && heads == 8) {
    const long long warps = ((long long)groups + 3) / 4;
    launch_k(attention5_kernel<T>, dim3((unsigned)((warps + 3) / 4)), dim3(128), 0, st, groups, ts, gs, op.f[0], qkv, out);
    taken = true;
    return check_launch("attention5");
  }
  return CAPF_OK;
}

template <typename TI, typename TO>
static int attention_dispatch(const capf_op& op, cudaStream_t st) {
  int groups = op.i[0], seq = op.i[1], heads = op.i[2], hd = op.i[3], ts = op.i[4], gs = op.i[5];
  int ipw = 32 / seq;
  long long items = (long long)groups * heads;
  long long warps = (items + ipw - 1) / ipw;
  int blocks = (int)((warps + 3) / 4);
  size_t smem = (size_t)4 * ipw * seq * (hd + 1) * 3 * sizeof(float);
  const TI* qkv = (const TI*)op.in[0];
  TO* out = (TO*)op.out[0];
  // opt in to > 48 KB dynamic smem once per instantiation (not a stream operation; done outside graph capture
  // because Plan.capture() always runs one eager warm-up pass first)
  static PerDevice<size_t> max5_, max17_;
  std::atomic<size_t>&max5 = max5_.get(), &max17 = max17_.get();
  if (seq == 5) {
    if (smem > max5) {
      cudaFuncSetAttribute(attention_small_kernel<TI, TO, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      max5 = smem;
    }
    launch_k(attention_small_kernel<TI, TO, 5>, dim3(blocks), dim3(128), smem, st, groups, heads, hd, ts, gs, op.f[0], qkv, out);
  } else {
    if (smem > max17) {
      cudaFuncSetAttribute(attention_small_kernel<TI, TO, 17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      max17 = smem;
    }
    launch_k(attention_small_kernel<TI, TO, 17>, dim3(blocks), dim3(128), smem, st, groups, heads, hd, ts, gs, op.f[0], qkv, out);
  }
  return check_launch("attention_small");
}

int launch_attention(const capf_op& op, cudaStream_t st) {
  int seq = op.i[1], hd = op.i[3];
  if (op.i[0] <= 0 || (seq != 5 && seq != 17) || op.i[2] <= 0 || hd <= 0 || hd > 256)
    return set_error(CAPF_ERR_ARG, "attention: seq must be 5 or 17, head_dim <= 256");
  if (op.dtype_in == CAPF_F32 && op.dtype_out == CAPF_F32) return attention_dispatch<float, float>(op, st);
  if (op.dtype_in == CAPF_F16 && op.dtype_out == CAPF_F16) {
    bool taken;
    int e = attention_fast<__half>(op, st, taken);
    if (e || taken) return e;
    return attention_dispatch<__half, __half>(op, st);
  }
  if (op.dtype_in == CAPF_BF16 && op.dtype_out == CAPF_BF16) {
    bool taken;
    int e = attention_fast<__nv_bfloat16>(op, st, taken);
    if (e || taken) return e;
    return attention_dispatch<__nv_bfloat16, __nv_bfloat16>(op, st);
  }
  return set_error(CAPF_ERR_UNSUPPORTED, "attention: dtype combination");
}

// =======================================================================================================
// Joint-context samplers
// =======================================================================================================
struct SampP {
  int B, J, nl;
  int H[4], W[4], C[4];
  long long off[4];
  const void* map[4];
};

// (a7) reference-point gather, padding_mode='zeros' (pose_dformer.py:216-218).  One warp per (b, j).
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) ref_sample_kernel(SampP p, const float* __restrict__ ref, TO* __restrict__ out,
                                                         int* __restrict__ rec) {
  pdl_wait();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int R = p.B * p.J;
  if (warp >= R) return;
  int b = warp / p.J;
  float gx = __ldg(ref + 2 * warp), gy = __ldg(ref + 2 * warp + 1);
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    if (l >= p.nl) break;
    const int H = p.H[l], W = p.W[l], C = p.C[l];
    Corners c = make_corners<false>(gx, gy, W, H);
    if (rec && lane == 0) {
      int* r = rec + ((size_t)l * R + warp) * 8;
      r[0] = c.x0; r[1] = c.y0; r[2] = (int)c.mask; r[3] = 0;
      r[4] = __float_as_int(gx); r[5] = __float_as_int(gy); r[6] = 0; r[7] = 0;
    }
    const TI* m = (const TI*)p.map[l] + (size_t)b * H * W * C;
    TO* o = out + p.off[l] + (size_t)warp * C;
    for (int ch = lane * 4; ch < C; ch += 128) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c.mask & (1u << k)) {
          int xx = c.x0 + (k & 1), yy = c.y0 + (k >> 1);
          float4 v = ld4<TI>(m + ((size_t)yy * W + xx) * C + ch);
          float wk = c.w[k];
          // ATen accumulates out += val * weight corner by corner (no FMA on the CPU path)
          a.x = __fadd_rn(a.x, __fmul_rn(v.x, wk));
          a.y = __fadd_rn(a.y, __fmul_rn(v.y, wk));
          a.z = __fadd_rn(a.z, __fmul_rn(v.z, wk));
          a.w = __fadd_rn(a.w, __fmul_rn(v.w, wk));
        }
      }
      st4<TO>(o + ch, a);
    }
  }
}

// (a8) deformable gather, padding_mode='border' (pose_dformer.py:122-135). One warp per (level, b, j, head):
// softmax over the head's 4 logits, tanh offsets, 4 samples x 4 corners, sample-weighted sum.  Narrow levels do not fill
// a warp with channels (C = 32 is 8 lanes of 4), so the 32 lanes are split into G = 32 / L groups (L = lanes that
// cover the channels) and each group takes 4 / G of the samples; the partial sums meet through shuffles.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) deform_sample_kernel(SampP p, const float* __restrict__ ref,
                                                            const float* __restrict__ ow, TO* __restrict__ out,
                                                            int* __restrict__ rec) {
  pdl_wait();
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int R = p.B * p.J;
  long long total = (long long)p.nl * R * 4;
  if (warp >= total) return;
  int h = (int)(warp & 3);
  long long t = warp >> 2;
  int rj = (int)(t % R), l = (int)(t / R);
  int b = rj / p.J;
  const int H = p.H[l], W = p.W[l], C = p.C[l];
  const int L = C <= 32 ? 8 : C <= 64 ? 16 : 32;      // lanes per sample group
  const int G = 32 / L;                               // sample groups: 4, 2 or 1
  const int grp = lane / L, li = lane - grp * L;
  const float* row = ow + ((size_t)l * R + rj) * 48;
  float lg[4], mx = -INFINITY;
#pragma unroll
  for (int s = 0; s < 4; ++s) { lg[s] = __ldg(row + h * 4 + s); mx = fmaxf(mx, lg[s]); }
  float den = 0.f;
#pragma unroll
  for (int s = 0; s < 4; ++s) { lg[s] = expf(lg[s] - mx); den += lg[s]; }
  float gx = __ldg(ref + 2 * rj), gy = __ldg(ref + 2 * rj + 1);
  const TI* m = (const TI*)p.map[l] + (size_t)b * H * W * C;
  TO* o = out + p.off[l] + ((size_t)rj * 4 + h) * C;
  float4 acc[2];                                        // up to 2 channel steps per lane (C <= 256); more loop below
  for (int ch0 = 0; ch0 < C; ch0 += 256) {
    acc[0] = acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      // every lane evaluates every sample's position (cheap, keeps the debug record complete); it only gathers its own
      const float aw = lg[s] / den;
      float ox = tanhf(__ldg(row + 16 + (h * 4 + s) * 2)), oy = tanhf(__ldg(row + 16 + (h * 4 + s) * 2 + 1));
      const float px = __fadd_rn(ox, gx), py = __fadd_rn(oy, gy);
      const Corners c = make_corners<true>(px, py, W, H);
      if (rec && lane == 0 && ch0 == 0) {
        int* r = rec + ((((size_t)l * R + rj) * 16) + h * 4 + s) * 8;
        r[0] = c.x0; r[1] = c.y0; r[2] = (int)c.mask; r[3] = 0;
        r[4] = __float_as_int(px); r[5] = __float_as_int(py); r[6] = 0; r[7] = 0;
      }
      if ((s & (G - 1)) != grp) continue;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int ch = ch0 + (it * L + li) * 4;
        if (ch < C && (it == 0 || L == 32)) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (c.mask & (1u << k)) {
              int xx = c.x0 + (k & 1), yy = c.y0 + (k >> 1);
              float4 v = ld4<TI>(m + ((size_t)yy * W + xx) * C + ch);
              float wk = c.w[k];
              a.x = __fadd_rn(a.x, __fmul_rn(v.x, wk));
              a.y = __fadd_rn(a.y, __fmul_rn(v.y, wk));
              a.z = __fadd_rn(a.z, __fmul_rn(v.z, wk));
              a.w = __fadd_rn(a.w, __fmul_rn(v.w, wk));
            }
          }
          acc[it].x = fmaf(aw, a.x, acc[it].x);
          acc[it].y = fmaf(aw, a.y, acc[it].y);
          acc[it].z = fmaf(aw, a.z, acc[it].z);
          acc[it].w = fmaf(aw, a.w, acc[it].w);
        }
      }
    }
    // combine the sample groups (lanes li, li + L, ...)
    for (int off = L; off < 32; off <<= 1) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        acc[it].x += __shfl_xor_sync(0xffffffffu, acc[it].x, off);
        acc[it].y += __shfl_xor_sync(0xffffffffu, acc[it].y, off);
        acc[it].z += __shfl_xor_sync(0xffffffffu, acc[it].z, off);
        acc[it].w += __shfl_xor_sync(0xffffffffu, acc[it].w, off);
      }
    }
    if (grp == 0) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int ch = ch0 + (it * L + li) * 4;
        if (ch < C && (it == 0 || L == 32)) st4<TO>(o + ch, acc[it]);
      }
    }
  }
}

// 16-bit fast path of the deformable gather: one warp per (frame, joint, level).  Lane s < 16 evaluates sample s
// (head s / 4, sample s % 4) ONCE -- softmax weight through two shuffles inside the head's 4 lanes, tanh offsets, the
// ATen corner record -- and publishes element offsets / blend weights through shared memory; then the lanes regroup as
// 32 / LPG sample groups of LPG lanes that each fetch 8 channels (16 bytes) per corner, so all four corners of up to
// eight samples are in flight per instruction.  Corner order and the unfused multiply-add of ATen are kept inside a
// sample; the per-head sum over the head's samples meets through shuffles.
struct DSample {
  int off[4];       // element offset of the corner pixel in the map (-1: corner outside the map)
  float w[4];       // bilinear weights nw, ne, sw, se
};

// Gather + blend + per-head sum for one (frame, joint, level) with LPG lanes per sample group.
template <typename T, int LPG>     // LPG = pow2 >= min(C / 8, 32)
__device__ __forceinline__ void deform_gather(const DSample* __restrict__ samp, const float* __restrict__ aws, const T* __restrict__ m, T* __restrict__ o,
                                              int C, int lane) {
  constexpr int NG = 32 / LPG;                     // sample groups per instruction
  constexpr int ITS = 16 / NG;                     // instructions per pass
  constexpr int IPH = NG >= 4 ? 1 : 4 / NG;        // iterations that make up one head
  constexpr int GH = NG >= 4 ? 4 : NG;             // groups that hold samples of the same head
  const int grp = lane / LPG, li = lane - grp * LPG;
  const int CP = C >> 3;                           // 16-byte chunks per pixel
  const int n_pass = (CP + LPG - 1) / LPG;         // 1 unless C > 256 (LPG = 32)
  for (int pass = 0; pass < n_pass; ++pass) {
    const int ch = pass * LPG + li;
    const bool active = ch < CP;                   // C / 8 not a power of two (HRNet-48): some lanes only shuffle
    float acc[8];
#pragma unroll
    for (int it = 0; it < ITS; ++it) {
      const int sidx = it * NG + grp;
      if (it % IPH == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      }
      const DSample d = samp[sidx];
      const float aw = aws[sidx];
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = (active && d.off[k] >= 0) ? __ldg(reinterpret_cast<const uint4*>(m + d.off[k]) + ch) : make_uint4(0u, 0u, 0u, 0u);
      float a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (d.off[k] >= 0) {                       // ATen skips out-of-map corners; inside a sample: out += val * weight
          float f[8];
          unpack8<T>(v[k], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] = __fadd_rn(a[e], __fmul_rn(f[e], d.w[k]));
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(aw, a[e], acc[e]);
      if (it % IPH == IPH - 1) {                   // the head's four samples are in: meet across its groups, store
#pragma unroll
        for (int off = LPG; off < LPG * GH; off <<= 1) {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
        }
        const int h = sidx >> 2;
        if (active && (grp % GH) == 0) *reinterpret_cast<uint4*>(o + (size_t)h * C + 8 * ch) = pack8<T>(acc);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 3) deform_sample16_kernel(SampP p, const float* __restrict__ ref, const float* __restrict__ ow,
                                                              T* __restrict__ out, int* __restrict__ rec) {
  pdl_wait();
  __shared__ DSample ssamp[8][16];
  __shared__ float saw[8][16];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * 8 + wib;
  const int R = p.B * p.J;
  if (warp >= (long long)R * p.nl) return;
  // level-major, widest level first: the warps of a block do equal work (a block lives as long as its slowest warp; C = 256
  // gathers 8x the bytes of C = 32) and the long blocks are scheduled before the short ones
  const int l = p.nl - 1 - (int)(warp / R), rj = (int)(warp % R);
  const int b = rj / p.J;
  const int H = p.H[l], W = p.W[l], C = p.C[l];
  const float* row = ow + ((size_t)l * R + rj) * 48;
  if (lane < 16) {
    const float lg = __ldg(row + lane);
    float mx = fmaxf(lg, __shfl_xor_sync(0x0000ffffu, lg, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0x0000ffffu, mx, 2));
    const float e = expf(lg - mx);
    float den = e + __shfl_xor_sync(0x0000ffffu, e, 1);
    den += __shfl_xor_sync(0x0000ffffu, den, 2);
    const float gx = __ldg(ref + 2 * rj), gy = __ldg(ref + 2 * rj + 1);
    const float ox = tanhf(__ldg(row + 16 + 2 * lane)), oy = tanhf(__ldg(row + 16 + 2 * lane + 1));
    const float px = __fadd_rn(ox, gx), py = __fadd_rn(oy, gy);
    const Corners c = make_corners<true>(px, py, W, H);
    if (rec) {
      int* r = rec + ((((size_t)l * R + rj) * 16) + lane) * 8;
      r[0] = c.x0; r[1] = c.y0; r[2] = (int)c.mask; r[3] = 0;
      r[4] = __float_as_int(px); r[5] = __float_as_int(py); r[6] = 0; r[7] = 0;
    }
    DSample d;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = c.x0 + (k & 1), yy = c.y0 + (k >> 1);
      d.off[k] = (c.mask & (1u << k)) ? (yy * W + xx) * C : -1;
      d.w[k] = c.w[k];
    }
    ssamp[wib][lane] = d;
    saw[wib][lane] = e / den;
  }
  __syncwarp();
  const T* m = (const T*)p.map[l] + (size_t)b * H * W * C;
  T* o = out + p.off[l] + (size_t)rj * 4 * C;
  const int CP = C >> 3;
  if (CP <= 4) deform_gather<T, 4>(ssamp[wib], saw[wib], m, o, C, lane);
  else if (CP <= 8) deform_gather<T, 8>(ssamp[wib], saw[wib], m, o, C, lane);
  else if (CP <= 16) deform_gather<T, 16>(ssamp[wib], saw[wib], m, o, C, lane);
  else deform_gather<T, 32>(ssamp[wib], saw[wib], m, o, C, lane);
}

template <typename T>
static bool deform_fast_ok(const capf_op& op, const SampP& p) {
  const char* ev = getenv("CAPF_DEFORM_FAST");
  if (ev && ev[0] == '0') return false;
  for (int l = 0; l < p.nl; ++l) {
    if (p.C[l] % 8 || ((uintptr_t)p.map[l] & 15) || (p.off[l] & 7)) return false;
    if ((long long)p.H[l] * p.W[l] * p.C[l] >= (1ll << 31)) return false;
  }
  return ((uintptr_t)op.out[0] & 15) == 0;
}

static int fill_samp(const capf_op& op, SampP& p, const char* who) {
  p.B = op.i[0]; p.J = op.i[1]; p.nl = op.i[2];
  if (p.B <= 0 || p.J <= 0 || p.nl < 1 || p.nl > 4) return set_error(CAPF_ERR_ARG, who);
  for (int l = 0; l < 4; ++l) {
    p.H[l] = p.W[l] = p.C[l] = 0; p.off[l] = 0; p.map[l] = nullptr;
    if (l < p.nl) {
      p.H[l] = op.i[3 + 3 * l]; p.W[l] = op.i[4 + 3 * l]; p.C[l] = op.i[5 + 3 * l];
      p.off[l] = op.i[15 + l];
      p.map[l] = op.in[1 + l];
      if (p.H[l] <= 0 || p.W[l] <= 0 || p.C[l] <= 0 || (p.C[l] & 3) || !p.map[l] || (p.off[l] & 3))
        return set_error(CAPF_ERR_ARG, who);
    }
  }
  if (!op.in[0] || !op.out[0]) return set_error(CAPF_ERR_ARG, who);
  return 0;
}

template <typename TI, typename TO>
static int sample_dispatch(const capf_op& op, const SampP& p, cudaStream_t st) {
  if (op.kind == CAPF_OP_REF_SAMPLE) {
    int R = p.B * p.J;
    launch_k(ref_sample_kernel<TI, TO>, dim3((R + 7) / 8), dim3(256), 0, st, p, (const float*)op.in[0], (TO*)op.out[0], (int*)op.out[1]);
    return check_launch("ref_sample");
  }
  if constexpr (std::is_same<TI, TO>::value && sizeof(TI) == 2) {
    if (deform_fast_ok<TI>(op, p)) {
      const long long w16 = (long long)p.nl * p.B * p.J;
      launch_k(deform_sample16_kernel<TI>, dim3((unsigned)((w16 + 7) / 8)), dim3(256), 0, st, p, (const float*)op.in[0], (const float*)op.in[5],
               (TI*)op.out[0], (int*)op.out[1]);
      return check_launch("deform_sample16");
    }
  }
  long long warps = (long long)p.nl * p.B * p.J * 4;
  launch_k(deform_sample_kernel<TI, TO>, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, st, p, (const float*)op.in[0], (const float*)op.in[5],
                                                                           (TO*)op.out[0], (int*)op.out[1]);
  return check_launch("deform_sample");
}

int launch_sample(const capf_op& op, cudaStream_t st) {
  SampP p;
  int e = fill_samp(op, p, op.kind == CAPF_OP_REF_SAMPLE ? "ref_sample: bad arguments" : "deform_sample: bad arguments");
  if (e) return e;
  if (op.kind == CAPF_OP_DEFORM_SAMPLE && !op.in[5]) return set_error(CAPF_ERR_ARG, "deform_sample: ow is null");
  int di = op.dtype_in, dd = op.dtype_out;
  if (di == CAPF_F32 && dd == CAPF_F32) return sample_dispatch<float, float>(op, p, st);
  if (di == CAPF_F16 && dd == CAPF_F16) return sample_dispatch<__half, __half>(op, p, st);
  if (di == CAPF_F16 && dd == CAPF_F32) return sample_dispatch<__half, float>(op, p, st);
  if (di == CAPF_BF16 && dd == CAPF_BF16) return sample_dispatch<__nv_bfloat16, __nv_bfloat16>(op, p, st);
  if (di == CAPF_BF16 && dd == CAPF_F32) return sample_dispatch<__nv_bfloat16, float>(op, p, st);
  return set_error(CAPF_ERR_UNSUPPORTED, "sample: dtype combination");
}

// =======================================================================================================
// Token-stream glue
// =======================================================================================================
// coord_embed + Spatial_pos_embed (pose_dformer.py:214,223-225) into the level-major stream X[slab][b*J+j][D]
__global__ void __launch_bounds__(256) embed_coord_kernel(int R, int J, int D, int slabs, const float* __restrict__ kp,
                                                          const float* __restrict__ Wc, const float* __restrict__ bc,
                                                          const float* __restrict__ pos, float* __restrict__ X) {
  pdl_wait();
  size_t total = (size_t)slabs * R * D;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int d = (int)(i % D);
    size_t t = i / D;
    int r = (int)(t % R), s = (int)(t / R);
    int j = r % J;
    float v = __ldg(pos + ((size_t)s * J + j) * D + d);
    if (s == 0) {
      float kx = __ldg(kp + 2 * r), ky = __ldg(kp + 2 * r + 1);
      // nn.Linear(2, D): x @ W^T + b, accumulated in k order like a GEMM
      float e = fmaf(ky, __ldg(Wc + 2 * d + 1), __fmul_rn(kx, __ldg(Wc + 2 * d)));
      v = __fadd_rn(__fadd_rn(e, __ldg(bc + d)), v);
    }
    X[i] = v;
  }
}

// '(b p) l c -> b p (l c)' (pose_dformer.py:235)
__global__ void __launch_bounds__(256) levels_to_joint_kernel(int R, int slabs, int D, const float* __restrict__ X,
                                                              float* __restrict__ Y) {
  pdl_wait();
  const int D4 = D >> 2;
  size_t total = (size_t)R * slabs * D4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int d4 = (int)(i % D4);
    size_t t = i / D4;
    int s = (int)(t % slabs), r = (int)(t / slabs);
    float4 v = __ldg(reinterpret_cast<const float4*>(X + ((size_t)s * R + r) * D + d4 * 4));
    *reinterpret_cast<float4*>(Y + i * 4) = v;
  }
}

// keypoints_2d_cpn_crop[..., :2] /= (96, 128);  -= (1, 1)   (conpose.py:34-35), in place
__global__ void crop_normalize_kernel(int n, float* __restrict__ c) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float2 v = reinterpret_cast<float2*>(c)[i];
    v.x = __fsub_rn(__fdiv_rn(v.x, 96.0f), 1.0f);
    v.y = __fsub_rn(__fdiv_rn(v.y, 128.0f), 1.0f);
    reinterpret_cast<float2*>(c)[i] = v;
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(size_t n, const TI* __restrict__ x, TO* __restrict__ y) {
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = from_f<TO>(to_f<TI>(x[i]));
}

// fp32 -> bf16 hi | lo planes: row r of x ([rows][C]) becomes [hi(x) (C) | bf16(x - hi) (C)] -- the A operand of the
// split-operand (bf16x3) tcgen05 GEMMs: hi * Wh + lo * Wh + hi * Wl reproduces the fp32 product to ~2^-16.
__global__ void __launch_bounds__(256) cast_split_kernel(size_t groups, int C8, const float* __restrict__ x, __nv_bfloat16* __restrict__ y) {
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < groups; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / C8;
    const int c = (int)(i - row * C8) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + i * 8)), b = __ldg(reinterpret_cast<const float4*>(x + i * 8) + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      hi[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
      lo[e] = __fsub_rn(v[e], hi[e]);
    }
    __nv_bfloat16* dst = y + row * (size_t)(16 * C8) + c;
    *reinterpret_cast<uint4*>(dst) = pack8<__nv_bfloat16>(hi);
    *reinterpret_cast<uint4*>(dst + 8 * C8) = pack8<__nv_bfloat16>(lo);
  }
}

// fp32 weight matrix [rows][C] (optionally read transposed from [C][rows]) -> the B operand of a split-operand GEMM:
// [rows][hi (C) | hi (C) | lo (C)] bf16 (program.split_weight_packer for one tap), for weights that change every step (training).
__global__ void __launch_bounds__(256) pack_split_weight_kernel(int rows, int C, int transposed, const float* __restrict__ w,
                                                                __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  const size_t total = (size_t)rows * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i - (size_t)r * C);
    const float v = __ldg(transposed ? w + (size_t)c * rows + r : w + i);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(hi)));
    __nv_bfloat16* o = out + (size_t)r * 3 * C + c;
    o[0] = hi;
    o[C] = hi;
    o[2 * C] = lo;
  }
}

int launch_embed_coord(const capf_op& op, cudaStream_t st) {
  int B = op.i[0], J = op.i[1], D = op.i[2], slabs = op.i[3];
  if (B <= 0 || J <= 0 || D <= 0 || slabs <= 0 || !op.in[0] || !op.in[1] || !op.in[2] || !op.in[3] || !op.out[0])
    return set_error(CAPF_ERR_ARG, "embed_coord: bad arguments");
  size_t total = (size_t)slabs * B * J * D;
  launch_k(embed_coord_kernel, dim3(ew_blocks(total)), dim3(256), 0, st, B * J, J, D, slabs, (const float*)op.in[0], (const float*)op.in[1],
                                                       (const float*)op.in[2], (const float*)op.in[3], (float*)op.out[0]);
  return check_launch("embed_coord");
}

int launch_levels_to_joint(const capf_op& op, cudaStream_t st) {
  int R = op.i[0], slabs = op.i[1], D = op.i[2];
  if (R <= 0 || slabs <= 0 || D <= 0 || (D & 3)) return set_error(CAPF_ERR_ARG, "levels_to_joint: bad arguments");
  launch_k(levels_to_joint_kernel, dim3(ew_blocks((size_t)R * slabs * (D / 4))), dim3(256), 0, st, R, slabs, D, (const float*)op.in[0], (float*)op.out[0]);
  return check_launch("levels_to_joint");
}

int launch_crop_normalize(const capf_op& op, cudaStream_t st) {
  int n = op.i[0];
  if (n <= 0 || !op.out[0]) return set_error(CAPF_ERR_ARG, "crop_normalize: bad arguments");
  launch_k(crop_normalize_kernel, dim3((n + 255) / 256), dim3(256), 0, st, n, (float*)op.out[0]);
  return check_launch("crop_normalize");
}

int launch_cast(const capf_op& op, cudaStream_t st) {
  size_t n = (size_t)(uint32_t)op.i[0] | ((size_t)(uint32_t)op.i[1] << 31);
  if (!n || !op.in[0] || !op.out[0]) return set_error(CAPF_ERR_ARG, "cast: bad arguments");
  int blocks = ew_blocks(n);
  if (op.i[2] > 0 && op.i[3] > 0) {        // split WEIGHT layout: i[2] = C (input features), i[3] = 1 plain / 2 source stored transposed
    if (op.dtype_in != CAPF_F32 || op.dtype_out != CAPF_BF16 || n % (size_t)op.i[2])
      return set_error(CAPF_ERR_UNSUPPORTED, "cast (split weights): needs f32 -> bf16");
    launch_k(pack_split_weight_kernel, dim3(blocks), dim3(256), 0, st, (int)(n / op.i[2]), op.i[2], op.i[3] == 2 ? 1 : 0, (const float*)op.in[0],
             (__nv_bfloat16*)op.out[0]);
    return check_launch("pack_split_weight");
  }
  if (op.i[2] > 0) {                       // split planes: i[2] = channels per row
    if (op.dtype_in != CAPF_F32 || op.dtype_out != CAPF_BF16 || (op.i[2] & 7) || n % (size_t)op.i[2])
      return set_error(CAPF_ERR_UNSUPPORTED, "cast (split planes): needs f32 -> bf16 and a channel count that is a multiple of 8");
    launch_k(cast_split_kernel, dim3(ew_blocks(n / 8)), dim3(256), 0, st, n / 8, op.i[2] / 8, (const float*)op.in[0], (__nv_bfloat16*)op.out[0]);
    return check_launch("cast_split");
  }
  if (op.dtype_in == CAPF_F32 && op.dtype_out == CAPF_F16)
    launch_k(cast_kernel<float, __half>, dim3(blocks), dim3(256), 0, st, n, (const float*)op.in[0], (__half*)op.out[0]);
  else if (op.dtype_in == CAPF_F32 && op.dtype_out == CAPF_BF16)
    launch_k(cast_kernel<float, __nv_bfloat16>, dim3(blocks), dim3(256), 0, st, n, (const float*)op.in[0], (__nv_bfloat16*)op.out[0]);
  else if (op.dtype_in == CAPF_F16 && op.dtype_out == CAPF_F32)
    launch_k(cast_kernel<__half, float>, dim3(blocks), dim3(256), 0, st, n, (const __half*)op.in[0], (float*)op.out[0]);
  else if (op.dtype_in == CAPF_BF16 && op.dtype_out == CAPF_F32)
    launch_k(cast_kernel<__nv_bfloat16, float>, dim3(blocks), dim3(256), 0, st, n, (const __nv_bfloat16*)op.in[0], (float*)op.out[0]);
  else if (op.dtype_in == CAPF_F32 && op.dtype_out == CAPF_F32)
    launch_k(cast_kernel<float, float>, dim3(blocks), dim3(256), 0, st, n, (const float*)op.in[0], (float*)op.out[0]);
  else
    return set_error(CAPF_ERR_UNSUPPORTED, "cast: dtype combination");
  return check_launch("cast");
}

// =======================================================================================================
// Pre-processing front end: data_prefetcher.preload's image path, mvn/datasets/utils.py:45-50
//   images = torch.flip(images_u8, [-1])                   BGR -> RGB
//   images = (images / 255.0 - mean) / std                 HRNet      |   images / 255.0 - mean      CPN
// and the flip-test copy torch.flip(images, [2]) (:67), fused into one pass: 3 bytes in, 12 bytes out per pixel.
// Four pixels per thread (12 bytes -> three float4) when W % 4 == 0; IEEE divisions keep it bit-exact with torch.
// =======================================================================================================
__device__ __forceinline__ float prep_px(unsigned v, float m, float s, int apply_std) {
  float t = __fsub_rn(__fdiv_rn((float)v, 255.0f), m);
  return apply_std ? __fdiv_rn(t, s) : t;
}

__global__ void __launch_bounds__(256) preprocess_u8_kernel(int H, int W, int mirror, int apply_std, unsigned total_groups, int gpr,
                                                            const uint8_t* __restrict__ x, const float* __restrict__ ms, float* __restrict__ y) {
  pdl_wait();
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= total_groups) return;
  const float mr = __ldg(ms), mg = __ldg(ms + 1), mb = __ldg(ms + 2), sr = __ldg(ms + 3), sg = __ldg(ms + 4), sb = __ldg(ms + 5);
  const unsigned row = i / (unsigned)gpr, g = i - row * (unsigned)gpr;       // row = b * H + h, g = group of 4 pixels in the row
  const int w0 = 4 * (int)g;
  const int n = min(4, W - w0);
  const uint8_t* src_row = x + (size_t)row * W * 3;
  float* dst = y + ((size_t)row * W + w0) * 3;
  float o[12];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < n) {
      const int ws = mirror ? (W - 1 - (w0 + k)) : (w0 + k);
      const uint8_t* px = src_row + 3 * ws;
      o[3 * k] = prep_px(px[2], mr, sr, apply_std);        // R <- byte 2
      o[3 * k + 1] = prep_px(px[1], mg, sg, apply_std);
      o[3 * k + 2] = prep_px(px[0], mb, sb, apply_std);    // B <- byte 0
    }
  }
  if (n == 4 && ((W & 3) == 0)) {
#pragma unroll
    for (int q = 0; q < 3; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  } else {
    for (int k = 0; k < 3 * n; ++k) dst[k] = o[k];
  }
}

int launch_preprocess_u8(const capf_op& op, cudaStream_t st) {
  const int B = op.i[0], H = op.i[1], W = op.i[2];
  if (B <= 0 || H <= 0 || W <= 0 || !op.in[0] || !op.in[1] || !op.out[0]) return set_error(CAPF_ERR_ARG, "preprocess_u8: bad arguments");
  if (op.dtype_out != CAPF_F32) return set_error(CAPF_ERR_UNSUPPORTED, "preprocess_u8: output must be f32 (the dtype CA_PF.forward takes)");
  if (((uintptr_t)op.out[0]) & 15) return set_error(CAPF_ERR_ARG, "preprocess_u8: output must be 16-byte aligned");
  const int gpr = (W + 3) / 4;
  const long long groups = (long long)B * H * gpr;
  if (groups >= (1ll << 32)) return set_error(CAPF_ERR_UNSUPPORTED, "preprocess_u8: too many pixels");
  launch_k(preprocess_u8_kernel, dim3((unsigned)((groups + 255) / 256)), dim3(256), 0, st, H, W, op.i[3] ? 1 : 0, op.i[4] ? 1 : 0, (unsigned)groups, gpr,
           (const uint8_t*)op.in[0], (const float*)op.in[1], (float*)op.out[0]);
  return check_launch("preprocess_u8");
}

// =======================================================================================================
// crop_image (mvn/utils/img.py:51-69): cv2.warpAffine(image, trans, (Wo, Ho), flags=INTER_LINEAR) on uint8 BGR frames,
// constant (0) border -- the per-frame CPU work of Human36MSingleViewDataset.__getitem__ (human36m.py:554-584).
// OpenCV's algorithm (imgproc/imgwarp.cpp, WarpAffineInvoker + remapBilinear, restated from its published source; the
// restatement is pinned against cv2 itself in oracle/capf_oracle.py::warp_affine_u8) is integer past the first step:
//   X = (round((M1*y + M2) * 1024) + 16 + round(M0*x*1024)) >> 5      (source x in 1/32 pixel; Y alike with M3..M5)
//   corner (X >> 5, Y >> 5), weights (32-fx, fx) x (32-fy, fy) with fx = X & 31, fy = Y & 31 (sum 1024)
//   value = (sum of corner * weight + 512) >> 10, corners outside the frame count as 0
// where M is the INVERSE (crop -> frame) map in fp64 and round() is round-half-even.  The fp64 products are formed
// with explicit _rn intrinsics so that no FMA contraction changes a rounding.  Result: the same bytes as cv2.
// mode 1 writes the normalised fp32 RGB pixel of data_prefetcher.preload instead (optionally mirrored along W), i.e.
// crop + CAPF_OP_PREPROCESS_U8 in one pass without the uint8 round trip.
// =======================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256) warp_affine_u8_kernel(int Hs, int Ws, int Ho, int Wo, unsigned total, int mirror, int apply_std,
                                                             const uint8_t* __restrict__ frames, const double* __restrict__ minv,
                                                             const int* __restrict__ sizes, const float* __restrict__ ms, void* __restrict__ outp) {
  pdl_wait();
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= total) return;
  const unsigned per = (unsigned)(Ho * Wo);
  const unsigned b = i / per, r = i - b * per;
  const int y = (int)(r / (unsigned)Wo), xo = (int)(r - (unsigned)y * (unsigned)Wo);
  const int x = (MODE == 1 && mirror) ? (Wo - 1 - xo) : xo;          // crop column this output pixel shows
  const double* M = minv + 6 * (size_t)b;
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(M[0], (double)x), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(M[3], (double)x), 1024.0));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[1], (double)y), M[2]), 1024.0)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[4], (double)y), M[5]), 1024.0)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));     // saturate_cast<short>
  const int fx = X & 31, fy = Y & 31;
  const int h = sizes ? sizes[2 * b] : Hs, w = sizes ? sizes[2 * b + 1] : Ws;               // live part of the padded frame
  const uint8_t* f = frames + (size_t)b * Hs * Ws * 3;
  int v[3] = {0, 0, 0};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int yy = sy + (c >> 1), xx = sx + (c & 1);
    const int wgt = ((c & 1) ? fx : 32 - fx) * ((c >> 1) ? fy : 32 - fy);
    if (wgt != 0 && yy >= 0 && yy < h && xx >= 0 && xx < w) {
      const uint8_t* px = f + ((size_t)yy * Ws + xx) * 3;
      v[0] += wgt * (int)__ldg(px);
      v[1] += wgt * (int)__ldg(px + 1);
      v[2] += wgt * (int)__ldg(px + 2);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) v[c] = (v[c] + 512) >> 10;
  if (MODE == 0) {
    uint8_t* o = reinterpret_cast<uint8_t*>(outp) + (size_t)i * 3;
    o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
  } else {
    float* o = reinterpret_cast<float*>(outp) + (size_t)i * 3;
    o[0] = prep_px((uint8_t)v[2], __ldg(ms), __ldg(ms + 3), apply_std);        // R <- byte 2
    o[1] = prep_px((uint8_t)v[1], __ldg(ms + 1), __ldg(ms + 4), apply_std);
    o[2] = prep_px((uint8_t)v[0], __ldg(ms + 2), __ldg(ms + 5), apply_std);    // B <- byte 0
  }
}

int launch_warp_affine_u8(const capf_op& op, cudaStream_t st) {
  const int B = op.i[0], Hs = op.i[1], Ws = op.i[2], Ho = op.i[3], Wo = op.i[4], mode = op.i[5];
  if (B <= 0 || Hs <= 0 || Ws <= 0 || Ho <= 0 || Wo <= 0 || !op.in[0] || !op.in[1] || !op.out[0]) return set_error(CAPF_ERR_ARG, "warp_affine_u8: bad arguments");
  if (mode != 0 && mode != 1) return set_error(CAPF_ERR_ARG, "warp_affine_u8: mode must be 0 (uint8 crop) or 1 (normalised fp32)");
  if (mode == 1 && !op.in[3]) return set_error(CAPF_ERR_ARG, "warp_affine_u8: mode 1 needs the mean/std vector");
  if (Hs > 32767 || Ws > 32767) return set_error(CAPF_ERR_UNSUPPORTED, "warp_affine_u8: frames larger than 32767 pixels a side (OpenCV's own limit)");
  if (((uintptr_t)op.in[1]) & 7) return set_error(CAPF_ERR_ARG, "warp_affine_u8: matrices must be 8-byte aligned fp64");
  const long long total = (long long)B * Ho * Wo;
  if (total >= (1ll << 32)) return set_error(CAPF_ERR_UNSUPPORTED, "warp_affine_u8: too many output pixels");
  const dim3 grid((unsigned)((total + 255) / 256));
  if (mode == 0)
    launch_k(warp_affine_u8_kernel<0>, grid, dim3(256), 0, st, Hs, Ws, Ho, Wo, (unsigned)total, 0, 0, (const uint8_t*)op.in[0], (const double*)op.in[1],
             (const int*)op.in[2], (const float*)nullptr, (void*)op.out[0]);
  else
    launch_k(warp_affine_u8_kernel<1>, grid, dim3(256), 0, st, Hs, Ws, Ho, Wo, (unsigned)total, op.i[6] ? 1 : 0, op.i[7] ? 1 : 0, (const uint8_t*)op.in[0],
             (const double*)op.in[1], (const int*)op.in[2], (const float*)op.in[3], (void*)op.out[0]);
  return check_launch("warp_affine_u8");
}

// =======================================================================================================
// Evaluation errors per frame (mvn/models/loss.py:16-22 MPJPE, :25-68 P_MPJPE, :87-101 MPJVE; reduced per action by
// evaluate_using_pred, mvn/datasets/human36m.py:358-422).  One thread per frame, fp64 inside:
//   out[n][0] = mean_j |pred - gt|
//   out[n][1] = the same after the similarity (scale, rotation, translation) that best maps pred onto gt.  The reference
//               takes numpy's SVD of H = X0^T Y0; here V and the singular values come from a Jacobi eigen-decomposition
//               of H^T H and U from H V, with u3 = u1 x u2 and v3 = v1 x v2, which makes R = V U^T a proper rotation and
//               gives the third singular value the sign the reference's det(R) fix-up produces.
//   out[n][2] = mean_j |(pred_n - pred_p) - (gt_n - gt_p)| with p = prev[n] (the frame before n inside its action, the
//               np.diff of the action-masked sequence), 0 when prev[n] < 0.
// =======================================================================================================
__device__ __forceinline__ void jacobi_rot(double (&a)[3][3], double (&v)[3][3], int p, int q) {
  if (fabs(a[p][q]) < 1e-300) return;
  const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
  const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
  for (int k = 0; k < 3; ++k) {            // A <- A J
    const double akp = a[k][p], akq = a[k][q];
    a[k][p] = c * akp - sn * akq;
    a[k][q] = sn * akp + c * akq;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {            // A <- J^T A
    const double apk = a[p][k], aqk = a[q][k];
    a[p][k] = c * apk - sn * aqk;
    a[q][k] = sn * apk + c * aqk;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {            // V <- V J
    const double vkp = v[k][p], vkq = v[k][q];
    v[k][p] = c * vkp - sn * vkq;
    v[k][q] = sn * vkp + c * vkq;
  }
}

__global__ void __launch_bounds__(128) pose_errors_kernel(int n_frames, int J, const float* __restrict__ pred, const float* __restrict__ gt,
                                                          const int* __restrict__ prev, double* __restrict__ out) {
  pdl_wait();
  const int n = blockIdx.x * 128 + threadIdx.x;
  if (n >= n_frames) return;
  const float* Y = pred + (size_t)n * J * 3;
  const float* X = gt + (size_t)n * J * 3;
  double mx[3] = {0, 0, 0}, my[3] = {0, 0, 0}, e1 = 0;
  for (int j = 0; j < J; ++j) {
    double d2 = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double x = X[3 * j + c], y = Y[3 * j + c];
      mx[c] += x; my[c] += y;
      d2 += (y - x) * (y - x);
    }
    e1 += sqrt(d2);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) { mx[c] /= J; my[c] /= J; }
  double nx = 0, ny = 0, h[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int j = 0; j < J; ++j) {
    double x[3], y[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { x[c] = X[3 * j + c] - mx[c]; y[c] = Y[3 * j + c] - my[c]; nx += x[c] * x[c]; ny += y[c] * y[c]; }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) h[a][b] += x[a] * y[b];
  }
  nx = sqrt(nx); ny = sqrt(ny);
  const double inv = 1.0 / (nx * ny);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) h[a][b] *= inv;          // H = X0^T Y0 of the normalised point sets
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) A[a][b] = h[0][a] * h[0][b] + h[1][a] * h[1][b] + h[2][a] * h[2][b];
  for (int sweep = 0; sweep < 10; ++sweep) { jacobi_rot(A, V, 0, 1); jacobi_rot(A, V, 0, 2); jacobi_rot(A, V, 1, 2); }
  // order the two largest eigenvalues first
  int i0 = 0, i1 = 1, i2 = 2;
  if (A[i0][i0] < A[i1][i1]) { const int t = i0; i0 = i1; i1 = t; }
  if (A[i0][i0] < A[i2][i2]) { const int t = i0; i0 = i2; i2 = t; }
  if (A[i1][i1] < A[i2][i2]) { const int t = i1; i1 = i2; i2 = t; }
  double v1[3], v2[3], v3[3], u1[3], u2[3], u3[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { v1[c] = V[c][i0]; v2[c] = V[c][i1]; }
  v3[0] = v1[1] * v2[2] - v1[2] * v2[1]; v3[1] = v1[2] * v2[0] - v1[0] * v2[2]; v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
  auto mulH = [&](const double (&v)[3], double (&o)[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) o[a] = h[a][0] * v[0] + h[a][1] * v[1] + h[a][2] * v[2];
  };
  double hv1[3], hv2[3], hv3[3];
  mulH(v1, hv1); mulH(v2, hv2); mulH(v3, hv3);
  const double s1 = sqrt(hv1[0] * hv1[0] + hv1[1] * hv1[1] + hv1[2] * hv1[2]);
#pragma unroll
  for (int c = 0; c < 3; ++c) u1[c] = hv1[c] / s1;
  const double d12 = u1[0] * hv2[0] + u1[1] * hv2[1] + u1[2] * hv2[2];
#pragma unroll
  for (int c = 0; c < 3; ++c) u2[c] = hv2[c] - d12 * u1[c];
  double s2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  if (s2 > 1e-150) {
#pragma unroll
    for (int c = 0; c < 3; ++c) u2[c] /= s2;
  } else {                                  // rank-1 H (collinear points): any unit vector orthogonal to u1
    s2 = 0;
    const int k = fabs(u1[0]) < fabs(u1[1]) ? (fabs(u1[0]) < fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) < fabs(u1[2]) ? 1 : 2);
    double e[3] = {0, 0, 0};
    e[k] = 1.0;
    const double d = u1[k];
    double nn = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) { u2[c] = e[c] - d * u1[c]; nn += u2[c] * u2[c]; }
    nn = sqrt(nn);
#pragma unroll
    for (int c = 0; c < 3; ++c) u2[c] /= nn;
  }
  u3[0] = u1[1] * u2[2] - u1[2] * u2[1]; u3[1] = u1[2] * u2[0] - u1[0] * u2[2]; u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
  const double s3 = u3[0] * hv3[0] + u3[1] * hv3[1] + u3[2] * hv3[2];        // signed: negative = the reflection case
  double R[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) R[a][b] = v1[a] * u1[b] + v2[a] * u2[b] + v3[a] * u3[b];
  const double scale = (s1 + s2 + s3) * nx / ny;
  double t[3];
#pragma unroll
  for (int b = 0; b < 3; ++b) t[b] = mx[b] - scale * (my[0] * R[0][b] + my[1] * R[1][b] + my[2] * R[2][b]);
  double e2 = 0, e3 = 0;
  const int p = prev ? prev[n] : -1;
  const float* Yp = pred + (size_t)(p < 0 ? 0 : p) * J * 3;
  const float* Xp = gt + (size_t)(p < 0 ? 0 : p) * J * 3;
  for (int j = 0; j < J; ++j) {
    const double y0 = Y[3 * j], y1 = Y[3 * j + 1], y2 = Y[3 * j + 2];
    double d2 = 0, dv = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double al = scale * (y0 * R[0][b] + y1 * R[1][b] + y2 * R[2][b]) + t[b] - (double)X[3 * j + b];
      d2 += al * al;
      if (p >= 0) {
        // the reference differences fp32 arrays (np.diff) before subtracting: keep those two roundings
        const float vy = Y[3 * j + b] - Yp[3 * j + b], vx = X[3 * j + b] - Xp[3 * j + b];
        const double w = (double)(vy - vx);
        dv += w * w;
      }
    }
    e2 += sqrt(d2);
    e3 += sqrt(dv);
  }
  out[3 * (size_t)n] = e1 / J;
  out[3 * (size_t)n + 1] = e2 / J;
  out[3 * (size_t)n + 2] = p >= 0 ? e3 / J : 0.0;
}

int launch_pose_errors(const capf_op& op, cudaStream_t st) {
  const int N = op.i[0], J = op.i[1];
  if (N <= 0 || J < 3 || !op.in[0] || !op.in[1] || !op.out[0]) return set_error(CAPF_ERR_ARG, "pose_errors: bad arguments");
  if (((uintptr_t)op.out[0]) & 7) return set_error(CAPF_ERR_ARG, "pose_errors: output must be 8-byte aligned fp64");
  launch_k(pose_errors_kernel, dim3((unsigned)((N + 127) / 128)), dim3(128), 0, st, N, J, (const float*)op.in[0], (const float*)op.in[1], (const int*)op.in[2],
           (double*)op.out[0]);
  return check_launch("pose_errors");
}

}  // namespace capf
