// Shared device helpers for libcapf_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "../../include/capf_b200.h"

namespace capf {

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel may
// start while its predecessor in the stream is still draining.  Contract: a kernel calls pdl_wait() before it reads
// or writes anything a previous kernel may have produced (it returns once the predecessor grid has completed and
// its memory is visible), and pdl_trigger() once it no longer minds the successor being scheduled (persistent
// one-wave kernels: at entry; multi-wave kernels: implicitly at block exit).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_use_pdl;   // capf_api.cu; env CAPF_PDL=0 turns the launch attribute off (kernels stay correct)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- storage <-> fp32 ---------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements as float4 (16-byte load for f32, 8-byte for 16-bit storage).
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 ld4<__half>(const __half* p) {
  uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  __half2 a = *reinterpret_cast<__half2*>(&r.x), b = *reinterpret_cast<__half2*>(&r.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x), b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<__half>(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// 8 consecutive 16-bit elements (one 16-byte vector) <-> fp32
template <typename T> __device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]);
template <typename T> __device__ __forceinline__ uint4 pack8(const float (&f)[8]);
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) { float2 t = __half22float2(h[e]); f[2 * e] = t.x; f[2 * e + 1] = t.y; }
}
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) { float2 t = __bfloat1622float2(h[e]); f[2 * e] = t.x; f[2 * e + 1] = t.y; }
}
template <> __device__ __forceinline__ uint4 pack8<__half>(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
  return u;
}
template <> __device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return u;
}

// nn.GELU() default (erf form), pose_dformer.py:15-22
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- bilinear corner record shared by both samplers (ATen/native/GridSampler.h:27-36, cuda/GridSampler.cu) --
struct Corners {
  int x0, y0;        // north-west integer corner
  float w[4];        // nw, ne, sw, se blend weights
  unsigned mask;     // bit k set: corner k inside the map
};

// align_corners=True un-normalisation + optional border clip, fp32 op-for-op as ATen (no FMA contraction).
template <bool BORDER>
__device__ __forceinline__ Corners make_corners(float gx, float gy, int W, int H) {
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(W - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(H - 1));
  if (BORDER) {
    ix = fminf((float)(W - 1), fmaxf(ix, 0.0f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.0f));
  }
  float fx = floorf(ix), fy = floorf(iy);
  Corners c;
  // keep far-out-of-range points representable as int (their mask is 0 anyway)
  c.x0 = (int)fminf(fmaxf(fx, -1.0e6f), 1.0e6f);
  c.y0 = (int)fminf(fmaxf(fy, -1.0e6f), 1.0e6f);
  float x1 = __fadd_rn(fx, 1.0f), y1 = __fadd_rn(fy, 1.0f);
  c.w[0] = __fmul_rn(__fsub_rn(x1, ix), __fsub_rn(y1, iy));
  c.w[1] = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(y1, iy));
  c.w[2] = __fmul_rn(__fsub_rn(x1, ix), __fsub_rn(iy, fy));
  c.w[3] = __fmul_rn(__fsub_rn(ix, fx), __fsub_rn(iy, fy));
  bool xl = c.x0 >= 0 && c.x0 < W, xr = c.x0 + 1 >= 0 && c.x0 + 1 < W;
  bool yt = c.y0 >= 0 && c.y0 < H, yb = c.y0 + 1 >= 0 && c.y0 + 1 < H;
  c.mask = (xl && yt ? 1u : 0u) | (xr && yt ? 2u : 0u) | (xl && yb ? 4u : 0u) | (xr && yb ? 8u : 0u);
  return c;
}

}  // namespace capf
