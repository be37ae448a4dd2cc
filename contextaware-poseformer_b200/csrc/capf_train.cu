// Training-step kernels of libcapf_b200: the backward pass of the lifter (`volume_net`,
// reference train.py:186-201 drives it through autograd) and the AdamW update train.py:337-345 configures.
// The backbone is frozen (conpose.py:22-25): nothing here touches it, and the feature maps are constants.
//
// Everything runs in fp32 on CUDA cores with fixed summation orders (no atomics: every reduction is a tree over
// fixed partials), so a training step is bit-reproducible.  Kernels:
//   gemm_f32_kernel        C = op(A) op(B) [+ bias] [+ C]      forward Linear, dgrad (dy W), wgrad (dy^T x; split over rows)
//   colsum_kernel          bias gradients, Spatial_pos_embed gradient, reduction of partials
//   layernorm_bwd_kernel   dx of nn.LayerNorm + per-block partial dgamma / dbeta
//   gelu / gelu_bwd        nn.GELU() (erf form) kept apart from the GEMM so the pre-activation is saved
//   attention_bwd_kernel   softmax(q k^T scale) v backward for the two tiny attentions (5 levels / 17 joints)
//   deform_bwd_kernel      DeformableBlock sampling (pose_dformer.py:124-135) backward w.r.t. the attention logits and the
//                          sampling offsets: F.grid_sample(padding_mode='border', align_corners=True) w.r.t. its grid
//                          (ATen GridSampler.cuh backward, bilinear), tanh and softmax
//   rows_axpy_kernel       y (+)= scale[row group] * t   -- DropPath (stochastic depth) residual adds and their backward
//   joint_to_levels_kernel inverse of the 'b p (l c)' regrouping
//   adamw_kernel           torch.optim.AdamW update over one flat parameter buffer
#include <cstdio>

#include "capf_common.cuh"
#include "capf_internal.h"

namespace capf {

static int ew_grid(size_t total, int threads = 256) {
  size_t b = (total + threads - 1) / threads, cap = (size_t)num_sms() * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

// =======================================================================================================
// fp32 GEMM, 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread.  A(m,k) / B(k,n) are addressed through
// (row stride, column stride) pairs, so all four transpose combinations share the kernel; grid.z splits K.
// =======================================================================================================
struct GemmP {
  int M, N, K;
  long long a_sm, a_sk, b_sk, b_sn;     // element strides: A(m,k) = A[m * a_sm + k * a_sk], B(k,n) = B[k * b_sk + n * b_sn]
  int ldc;
  int ksplit, kchunk;                   // grid.z slices of kchunk (multiple of 16) reduction steps
  int accumulate;                       // C += result (only with ksplit == 1)
  const float* A;
  const float* B;
  const float* bias;
  float* C;                             // ksplit > 1: partials [ksplit][M][N] (dense, ld = N)
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmP p) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  pdl_wait();
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kbeg = blockIdx.z * p.kchunk, kend = min(p.K, kbeg + p.kchunk);
  const int tx = tid & 15, ty = tid >> 4;             // 16 x 16 threads -> 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: the fastest-varying thread index walks the unit-stride dimension of each operand
  const bool a_k_fast = p.a_sk == 1, b_n_fast = p.b_sn == 1;
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;                    // 1024 elements of a 64 x 16 tile
      int am, ak, bk, bn;
      if (a_k_fast) { ak = e & 15; am = e >> 4; } else { am = e & 63; ak = e >> 6; }
      if (b_n_fast) { bn = e & 63; bk = e >> 6; } else { bk = e & 15; bn = e >> 4; }
      const int gm = m0 + am, gk = k0 + ak;
      sA[ak][am] = (gm < p.M && gk < kend) ? __ldg(p.A + gm * p.a_sm + gk * p.a_sk) : 0.f;
      const int gn = n0 + bn, gk2 = k0 + bk;
      sB[bk][bn] = (gn < p.N && gk2 < kend) ? __ldg(p.B + gk2 * p.b_sk + gn * p.b_sn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* C = p.C + (p.ksplit > 1 ? (size_t)blockIdx.z * p.M * p.N : 0);
  const int ldc = p.ksplit > 1 ? p.N : p.ldc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.ksplit == 1) {
        if (p.bias) v += __ldg(p.bias + n);
        if (p.accumulate) v += C[(size_t)m * ldc + n];
      }
      C[(size_t)m * ldc + n] = v;
    }
  }
}

// out[m][n] (ld) = [out +] bias[n] + sum_z part[z][m][n]   (fixed order over z)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(int M, int N, int ld, int ksplit, int accumulate, const float* __restrict__ part,
                                                            const float* __restrict__ bias, float* __restrict__ out) {
  pdl_wait();
  const size_t total = (size_t)M * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), n = (int)(i - (size_t)m * N);
    float v = 0.f;
    for (int z = 0; z < ksplit; ++z) v += part[(size_t)z * total + i];
    if (bias) v += __ldg(bias + n);
    float* o = out + (size_t)m * ld + n;
    *o = accumulate ? *o + v : v;
  }
}

int launch_gemm_f32(const capf_op& op, cudaStream_t st) {
  GemmP p;
  p.M = op.i[0]; p.N = op.i[1]; p.K = op.i[2];
  const int ta = op.i[3], tb = op.i[4], lda = op.i[5], ldb = op.i[6];
  p.ldc = op.i[7];
  int ksplit = op.i[8] > 1 ? op.i[8] : 1;
  p.accumulate = op.i[9] ? 1 : 0;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || lda <= 0 || ldb <= 0 || p.ldc < p.N || !op.in[0] || !op.in[1] || !op.out[0])
    return set_error(CAPF_ERR_ARG, "gemm_f32: bad arguments");
  // A: [M][K] (ta = 0) or stored transposed [K][M] (ta = 1); B: [K][N] (tb = 0) or stored [N][K] (tb = 1, C = A B^T)
  p.a_sm = ta ? 1 : lda; p.a_sk = ta ? lda : 1;
  p.b_sk = tb ? 1 : ldb; p.b_sn = tb ? ldb : 1;
  p.A = (const float*)op.in[0]; p.B = (const float*)op.in[1]; p.bias = (const float*)op.in[2];
  int kchunk = ((p.K + ksplit - 1) / ksplit + 15) / 16 * 16;
  ksplit = (p.K + kchunk - 1) / kchunk;
  if (ksplit > 1 && !op.out[1]) return set_error(CAPF_ERR_ARG, "gemm_f32: split-K needs a partials workspace in out[1]");
  p.ksplit = ksplit; p.kchunk = kchunk;
  p.C = ksplit > 1 ? (float*)op.out[1] : (float*)op.out[0];
  dim3 grid((p.N + 63) / 64, (p.M + 63) / 64, ksplit);
  launch_k(gemm_f32_kernel, grid, dim3(256), 0, st, p);
  int e = check_launch("gemm_f32");
  if (e || ksplit == 1) return e;
  launch_k(splitk_reduce_kernel, dim3(ew_grid((size_t)p.M * p.N)), dim3(256), 0, st, p.M, p.N, p.ldc, ksplit, p.accumulate, (const float*)op.out[1],
           p.bias, (float*)op.out[0]);
  return check_launch("splitk_reduce");
}

// =======================================================================================================
// column sums: out[n] = [out[n] +] sum_m x[m * ld + n].  Two deterministic stages: row chunks -> partials -> sum.
// =======================================================================================================
__global__ void __launch_bounds__(256) colsum_kernel(int M, int N, long long ld, int rows_per_chunk, int accumulate, const float* __restrict__ x,
                                                     float* __restrict__ out /* [gridDim.y][N] */) {
  __shared__ float red[8][33];
  pdl_wait();
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float s = 0.f;
  if (c < N)
    for (int r = r0 + rl; r < r1; r += 8) s += __ldg(x + (size_t)r * ld + c);
  red[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    float* o = out + (size_t)blockIdx.y * N + c;
    *o = (accumulate && gridDim.y == 1) ? *o + t : t;
  }
}

int launch_colsum(const capf_op& op, cudaStream_t st) {
  const int M = op.i[0], N = op.i[1], ld = op.i[2], accumulate = op.i[3] ? 1 : 0;
  if (M <= 0 || N <= 0 || ld < N || !op.in[0] || !op.out[0]) return set_error(CAPF_ERR_ARG, "colsum: bad arguments");
  int chunks = (M + 511) / 512;
  if (chunks > 256) chunks = 256;
  if (chunks <= 1 || !op.out[1]) {
    launch_k(colsum_kernel, dim3((N + 31) / 32, 1), dim3(256), 0, st, M, N, (long long)ld, M, accumulate, (const float*)op.in[0], (float*)op.out[0]);
    return check_launch("colsum");
  }
  const int rpc = (M + chunks - 1) / chunks;
  chunks = (M + rpc - 1) / rpc;
  launch_k(colsum_kernel, dim3((N + 31) / 32, chunks), dim3(256), 0, st, M, N, (long long)ld, rpc, 0, (const float*)op.in[0], (float*)op.out[1]);
  int e = check_launch("colsum");
  if (e) return e;
  launch_k(colsum_kernel, dim3((N + 31) / 32, 1), dim3(256), 0, st, chunks, N, (long long)N, chunks, accumulate, (const float*)op.out[1], (float*)op.out[0]);
  return check_launch("colsum(partials)");
}

// =======================================================================================================
// LayerNorm backward.  One warp per row (grid-stride); lanes own columns lane, lane + 32, ...
//   xhat = (x - mean) * rstd,  g = dy * gamma,  dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
// `x0` (period rows, broadcast over rows / period) is added to x first -- DeformableBlock.norm1(x + x_0), :120.
// Each block leaves its partial sums of dy * xhat and dy in part[block][2][D]; colsum finishes dgamma / dbeta.
// =======================================================================================================
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(int rows, int D, int period, float eps, const float* __restrict__ x,
                                                            const float* __restrict__ x0, const float* __restrict__ gamma,
                                                            const float* __restrict__ dy, float* __restrict__ dx, int dx_accumulate,
                                                            float* __restrict__ part) {
  extern __shared__ float sred[];        // [8 warps][2][D]
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = D / 32;                 // D % 32 == 0
  float gsum[MAXV], bsum[MAXV], gam[MAXV];
#pragma unroll
  for (int v = 0; v < MAXV; ++v) { gsum[v] = bsum[v] = 0.f; gam[v] = v < nv ? __ldg(gamma + lane + 32 * v) : 0.f; }
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    float xv[MAXV], dv[MAXV];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
      if (v < nv) {
        const int c = lane + 32 * v;
        float t = __ldg(x + (size_t)r * D + c);
        if (period) t += __ldg(x0 + (size_t)(r % period) * D + c);
        xv[v] = t;
        dv[v] = __ldg(dy + (size_t)r * D + c);
        s += t;
      } else { xv[v] = dv[v] = 0.f; }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < MAXV; ++v)
      if (v < nv) { const float d = xv[v] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
      if (v < nv) {
        xv[v] = (xv[v] - mean) * rstd;             // xhat
        const float g = dv[v] * gam[v];
        m1 += g;
        m2 += g * xv[v];
        gsum[v] += dv[v] * xv[v];
        bsum[v] += dv[v];
      }
    }
    m1 = warp_sum(m1) / (float)D;
    m2 = warp_sum(m2) / (float)D;
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
      if (v < nv) {
        const float o = rstd * (dv[v] * gam[v] - m1 - xv[v] * m2);
        float* d = dx + (size_t)r * D + lane + 32 * v;
        *d = dx_accumulate ? *d + o : o;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < MAXV; ++v) {
    if (v < nv) {
      sred[(warp * 2 + 0) * D + lane + 32 * v] = gsum[v];
      sred[(warp * 2 + 1) * D + lane + 32 * v] = bsum[v];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sred[w * 2 * D + i];
    part[(size_t)blockIdx.x * 2 * D + i] = t;
  }
}

int launch_layernorm_bwd(const capf_op& op, cudaStream_t st) {
  const int rows = op.i[0], D = op.i[1], period = op.i[2], dx_acc = op.i[3] ? 1 : 0, nblocks = op.i[4];
  if (rows <= 0 || D <= 0 || (D & 31) || D > 32 * 20 || nblocks <= 0 || !op.in[0] || !op.in[1] || !op.in[2] || !op.out[0] || !op.out[1] ||
      (period && (!op.in[3] || rows % period)))
    return set_error(CAPF_ERR_ARG, "layernorm_bwd: bad arguments (D % 32 == 0, D <= 640)");
  const size_t smem = (size_t)8 * 2 * D * sizeof(float);
  const float* x = (const float*)op.in[0];
  const float* g = (const float*)op.in[1];
  const float* dy = (const float*)op.in[2];
  const float* x0 = (const float*)op.in[3];
  if (D <= 128) launch_k(layernorm_bwd_kernel<4>, dim3(nblocks), dim3(256), smem, st, rows, D, period, op.f[0], x, x0, g, dy, (float*)op.out[0], dx_acc, (float*)op.out[1]);
  else launch_k(layernorm_bwd_kernel<20>, dim3(nblocks), dim3(256), smem, st, rows, D, period, op.f[0], x, x0, g, dy, (float*)op.out[0], dx_acc, (float*)op.out[1]);
  return check_launch("layernorm_bwd");
}

// =======================================================================================================
// GELU (erf form) forward / backward on dense arrays
// =======================================================================================================
__global__ void __launch_bounds__(256) gelu_kernel(size_t n4, const float4* __restrict__ h, float4* __restrict__ y) {
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(h + i);
    y[i] = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
  }
}
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(size_t n4, const float4* __restrict__ h, const float4* __restrict__ dy, float4* __restrict__ dh) {
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(h + i), d = __ldg(dy + i);
    dh[i] = make_float4(d.x * gelu_grad(v.x), d.y * gelu_grad(v.y), d.z * gelu_grad(v.z), d.w * gelu_grad(v.w));
  }
}

int launch_gelu(const capf_op& op, cudaStream_t st) {
  const size_t n = (size_t)(uint32_t)op.i[0] | ((size_t)(uint32_t)op.i[1] << 31);
  const bool bwd = op.kind == CAPF_OP_GELU_BWD;
  if (!n || (n & 3) || !op.in[0] || !op.out[0] || (bwd && !op.in[1])) return set_error(CAPF_ERR_ARG, "gelu: bad arguments (count % 4 == 0)");
  if (bwd) launch_k(gelu_bwd_kernel, dim3(ew_grid(n / 4)), dim3(256), 0, st, n / 4, (const float4*)op.in[0], (const float4*)op.in[1], (float4*)op.out[0]);
  else launch_k(gelu_kernel, dim3(ew_grid(n / 4)), dim3(256), 0, st, n / 4, (const float4*)op.in[0], (float4*)op.out[0]);
  return check_launch(bwd ? "gelu_bwd" : "gelu");
}

// =======================================================================================================
// Attention backward (Attention.forward, pose_dformer.py:47-59): one warp per (group, head), everything in shared
// memory.  qkv rows hold (3, heads, hd); token t of group g is row g * grp_stride + t * tok_stride.
// =======================================================================================================
template <int SEQ>
__global__ void __launch_bounds__(128) attention_bwd_kernel(int groups, int heads, int hd, int tok_stride, int grp_stride, float scale,
                                                            const float* __restrict__ qkv, const float* __restrict__ dout, float* __restrict__ dqkv) {
  extern __shared__ float sm[];
  pdl_wait();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long item = (long long)blockIdx.x * 4 + wib;
  if (item >= (long long)groups * heads) return;
  const int g = (int)(item / heads), h = (int)(item % heads);
  const int D = heads * hd, hp = hd + 1;
  float* q = sm + (size_t)wib * (4 * SEQ * hp + 2 * SEQ * (SEQ + 1));
  float* k = q + SEQ * hp;
  float* v = k + SEQ * hp;
  float* go = v + SEQ * hp;
  float* P = go + SEQ * hp;                 // [SEQ][SEQ + 1]
  float* dS = P + SEQ * (SEQ + 1);
  for (int e = lane; e < SEQ * hd; e += 32) {
    const int t = e / hd, d = e - t * hd;
    const size_t row = (size_t)g * grp_stride + (size_t)t * tok_stride;
    const float* r = qkv + row * 3 * D + h * hd + d;
    q[t * hp + d] = __ldg(r);
    k[t * hp + d] = __ldg(r + D);
    v[t * hp + d] = __ldg(r + 2 * D);
    go[t * hp + d] = __ldg(dout + row * D + h * hd + d);
  }
  __syncwarp();
  for (int e = lane; e < SEQ * SEQ; e += 32) {
    const int i = e / SEQ, j = e - i * SEQ;
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < hd; ++d) { s = fmaf(q[i * hp + d], k[j * hp + d], s); dp = fmaf(go[i * hp + d], v[j * hp + d], dp); }
    P[i * (SEQ + 1) + j] = s * scale;
    dS[i * (SEQ + 1) + j] = dp;              // dP for now
  }
  __syncwarp();
  if (lane < SEQ) {
    float* pr = P + lane * (SEQ + 1);
    float* dr = dS + lane * (SEQ + 1);
    float mx = -INFINITY;
    for (int j = 0; j < SEQ; ++j) mx = fmaxf(mx, pr[j]);
    float den = 0.f;
    for (int j = 0; j < SEQ; ++j) { pr[j] = expf(pr[j] - mx); den += pr[j]; }
    float dot = 0.f;
    for (int j = 0; j < SEQ; ++j) { pr[j] /= den; dot = fmaf(pr[j], dr[j], dot); }
    for (int j = 0; j < SEQ; ++j) dr[j] = pr[j] * (dr[j] - dot);      // dS
  }
  __syncwarp();
  for (int e = lane; e < SEQ * hd; e += 32) {
    const int t = e / hd, d = e - t * hd;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < SEQ; ++j) {
      dq = fmaf(dS[t * (SEQ + 1) + j], k[j * hp + d], dq);
      dk = fmaf(dS[j * (SEQ + 1) + t], q[j * hp + d], dk);
      dv = fmaf(P[j * (SEQ + 1) + t], go[j * hp + d], dv);
    }
    const size_t row = (size_t)g * grp_stride + (size_t)t * tok_stride;
    float* r = dqkv + row * 3 * D + h * hd + d;
    r[0] = dq * scale;
    r[D] = dk * scale;
    r[2 * D] = dv;
  }
}

int launch_attention_bwd(const capf_op& op, cudaStream_t st) {
  const int groups = op.i[0], seq = op.i[1], heads = op.i[2], hd = op.i[3], ts = op.i[4], gs = op.i[5];
  if (groups <= 0 || (seq != 5 && seq != 17) || heads <= 0 || hd <= 0 || hd > 128 || !op.in[0] || !op.in[1] || !op.out[0])
    return set_error(CAPF_ERR_ARG, "attention_bwd: bad arguments (seq 5 | 17, head_dim <= 128)");
  const long long items = (long long)groups * heads;
  const int blocks = (int)((items + 3) / 4);
  const size_t smem = (size_t)4 * (4 * seq * (hd + 1) + 2 * seq * (seq + 1)) * sizeof(float);
  static PerDevice<size_t> max5_, max17_;
  if (seq == 5) {
    std::atomic<size_t>& mx = max5_.get();
    if (smem > mx) { cudaFuncSetAttribute(attention_bwd_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); mx = smem; }
    launch_k(attention_bwd_kernel<5>, dim3(blocks), dim3(128), smem, st, groups, heads, hd, ts, gs, op.f[0], (const float*)op.in[0], (const float*)op.in[1], (float*)op.out[0]);
  } else {
    std::atomic<size_t>& mx = max17_.get();
    if (smem > mx) { cudaFuncSetAttribute(attention_bwd_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); mx = smem; }
    launch_k(attention_bwd_kernel<17>, dim3(blocks), dim3(128), smem, st, groups, heads, hd, ts, gs, op.f[0], (const float*)op.in[0], (const float*)op.in[1], (float*)op.out[0]);
  }
  return check_launch("attention_bwd");
}

// =======================================================================================================
// DeformableBlock sampling backward.  Forward (deform_sample_kernel): for row (b, j), level l, head h
//   g[h][:] = sum_s softmax_s(ow[h*4+s]) * bilinear_border(map_l[b], ref[b,j] + tanh(ow[16 + (h*4+s)*2 + {0,1}]))
// Given dg = d loss / d g this kernel returns d loss / d ow[48] per (level, row).  One warp per (level, row, head); lanes
// stride over the channels, three dot products per sample (value, d/dx, d/dy) meet through warp sums.
// grid_sample w.r.t. its grid follows ATen's grid_sampler_2d_backward (bilinear, align_corners=True -> multiplier
// (size - 1) / 2, padding 'border' -> zero gradient where the coordinate was clipped, out-of-map corners read as 0).
// =======================================================================================================
struct DeformBwdP {
  int B, J, nl;
  int H[4], W[4], C[4];
  int off[4];
  const void* map[4];
};

template <typename TI>
__global__ void __launch_bounds__(256) deform_bwd_kernel(DeformBwdP p, const float* __restrict__ ref, const float* __restrict__ ow,
                                                         const float* __restrict__ dg, float* __restrict__ dow) {
  pdl_wait();
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int R = p.B * p.J;
  if (warp >= (long long)p.nl * R * 4) return;
  const int h = (int)(warp & 3);
  const long long t = warp >> 2;
  const int rj = (int)(t % R), l = (int)(t / R), b = rj / p.J;
  const int H = p.H[l], W = p.W[l], C = p.C[l];
  const float* row = ow + ((size_t)l * R + rj) * 48;
  float wgt[4], mx = -INFINITY;
#pragma unroll
  for (int s = 0; s < 4; ++s) { wgt[s] = __ldg(row + h * 4 + s); mx = fmaxf(mx, wgt[s]); }
  float den = 0.f;
#pragma unroll
  for (int s = 0; s < 4; ++s) { wgt[s] = expf(wgt[s] - mx); den += wgt[s]; }
  const float gx = __ldg(ref + 2 * rj), gy = __ldg(ref + 2 * rj + 1);
  const TI* m = (const TI*)p.map[l] + (size_t)b * H * W * C;
  const float* d = dg + p.off[l] + ((size_t)rj * 4 + h) * C;
  float dval[4], dpx[4], dpy[4], th[4][2];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    wgt[s] /= den;
    const float ox = tanhf(__ldg(row + 16 + (h * 4 + s) * 2)), oy = tanhf(__ldg(row + 16 + (h * 4 + s) * 2 + 1));
    th[s][0] = ox; th[s][1] = oy;
    const float px = __fadd_rn(ox, gx), py = __fadd_rn(oy, gy);
    const Corners c = make_corners<true>(px, py, W, H);
    // un-normalised, clipped coordinates again for the interpolation fractions and the clip mask
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(px, 1.0f), 0.5f), (float)(W - 1));
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(py, 1.0f), 0.5f), (float)(H - 1));
    const float mxs = (ix <= 0.0f || ix >= (float)(W - 1)) ? 0.0f : 0.5f * (float)(W - 1);     // clip_coordinates_set_grad
    const float mys = (iy <= 0.0f || iy >= (float)(H - 1)) ? 0.0f : 0.5f * (float)(H - 1);
    ix = fminf((float)(W - 1), fmaxf(ix, 0.0f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.0f));
    const float tx = ix - floorf(ix), ty = iy - floorf(iy);
    float a = 0.f, ax = 0.f, ay = 0.f;
    for (int ch = lane * 4; ch < C; ch += 128) {
      const float4 gv = __ldg(reinterpret_cast<const float4*>(d + ch));
      float4 cv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        cv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c.mask & (1u << k)) cv[k] = ld4<TI>(m + ((size_t)(c.y0 + (k >> 1)) * W + c.x0 + (k & 1)) * C + ch);
      }
      const float gq[4] = {gv.x, gv.y, gv.z, gv.w};
      const float nw[4] = {cv[0].x, cv[0].y, cv[0].z, cv[0].w}, ne[4] = {cv[1].x, cv[1].y, cv[1].z, cv[1].w};
      const float sw[4] = {cv[2].x, cv[2].y, cv[2].z, cv[2].w}, se[4] = {cv[3].x, cv[3].y, cv[3].z, cv[3].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float val = nw[e] * c.w[0] + ne[e] * c.w[1] + sw[e] * c.w[2] + se[e] * c.w[3];
        a = fmaf(gq[e], val, a);
        ax = fmaf(gq[e], (ne[e] - nw[e]) * (1.0f - ty) + (se[e] - sw[e]) * ty, ax);
        ay = fmaf(gq[e], (sw[e] - nw[e]) * (1.0f - tx) + (se[e] - ne[e]) * tx, ay);
      }
    }
    dval[s] = warp_sum(a);
    dpx[s] = warp_sum(ax) * mxs;
    dpy[s] = warp_sum(ay) * mys;
  }
  if (lane == 0) {
    float dot = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) dot = fmaf(wgt[s], dval[s], dot);
    float* o = dow + ((size_t)l * R + rj) * 48;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      o[h * 4 + s] = wgt[s] * (dval[s] - dot);                                        // softmax backward
      o[16 + (h * 4 + s) * 2] = wgt[s] * dpx[s] * (1.0f - th[s][0] * th[s][0]);       // d sample / d pos, tanh backward
      o[16 + (h * 4 + s) * 2 + 1] = wgt[s] * dpy[s] * (1.0f - th[s][1] * th[s][1]);
    }
  }
}

int launch_deform_bwd(const capf_op& op, cudaStream_t st) {
  DeformBwdP p;
  p.B = op.i[0]; p.J = op.i[1]; p.nl = op.i[2];
  if (p.B <= 0 || p.J <= 0 || p.nl < 1 || p.nl > 4 || !op.in[0] || !op.in[5] || !op.out[0] || !op.out[1])
    return set_error(CAPF_ERR_ARG, "deform_bwd: bad arguments");
  for (int l = 0; l < 4; ++l) {
    p.H[l] = p.W[l] = p.C[l] = 0; p.off[l] = 0; p.map[l] = nullptr;
    if (l < p.nl) {
      p.H[l] = op.i[3 + 3 * l]; p.W[l] = op.i[4 + 3 * l]; p.C[l] = op.i[5 + 3 * l];
      p.off[l] = op.i[15 + l];
      p.map[l] = op.in[1 + l];
      if (p.H[l] <= 0 || p.W[l] <= 0 || p.C[l] <= 0 || (p.C[l] & 3) || !p.map[l] || (p.off[l] & 3)) return set_error(CAPF_ERR_ARG, "deform_bwd: bad level geometry");
    }
  }
  // in[0] = ref, in[1..4] = maps (dtype_in), in[5] = ow; out[0] = d ow, out[1] = dg (an INPUT: d loss / d sampled sums)
  const long long warps = (long long)p.nl * p.B * p.J * 4;
  const dim3 grid((unsigned)((warps + 7) / 8));
  const float* ref = (const float*)op.in[0];
  const float* ow = (const float*)op.in[5];
  const float* dg = (const float*)op.out[1];
  float* dow = (float*)op.out[0];
  switch (op.dtype_in) {
    case CAPF_F32: launch_k(deform_bwd_kernel<float>, grid, dim3(256), 0, st, p, ref, ow, dg, dow); break;
    case CAPF_F16: launch_k(deform_bwd_kernel<__half>, grid, dim3(256), 0, st, p, ref, ow, dg, dow); break;
    case CAPF_BF16: launch_k(deform_bwd_kernel<__nv_bfloat16>, grid, dim3(256), 0, st, p, ref, ow, dg, dow); break;
    default: return set_error(CAPF_ERR_UNSUPPORTED, "deform_bwd: map dtype");
  }
  return check_launch("deform_bwd");
}

// =======================================================================================================
// y[r][:] = [y[r][:] +] scale[(r % mod) / div] * t[r][:]      (scale == NULL: 1)
// DropPath: timm's drop_path multiplies a residual branch by bernoulli(keep) / keep per sample of its leading dimension
// (pose_dformer.py:77-78, :137-138); the mask is drawn by the host, the same kernel serves the backward (dt = scale * dy).
// =======================================================================================================
__global__ void __launch_bounds__(256) rows_axpy_kernel(size_t rows, int D4, int mod, int div, int accumulate, const float* __restrict__ scale,
                                                        const float4* __restrict__ t, float4* __restrict__ y) {
  pdl_wait();
  const size_t total = rows * D4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / D4;
    const float s = scale ? __ldg(scale + (r % mod) / div) : 1.0f;
    const float4 v = __ldg(t + i);
    float4 o = make_float4(s * v.x, s * v.y, s * v.z, s * v.w);
    if (accumulate) { const float4 c = y[i]; o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
    y[i] = o;
  }
}

int launch_rows_axpy(const capf_op& op, cudaStream_t st) {
  const int rows = op.i[0], D = op.i[1], mod = op.i[2], div = op.i[3], acc = op.i[4] ? 1 : 0;
  if (rows <= 0 || D <= 0 || (D & 3) || mod <= 0 || div <= 0 || !op.in[0] || !op.out[0]) return set_error(CAPF_ERR_ARG, "rows_axpy: bad arguments");
  launch_k(rows_axpy_kernel, dim3(ew_grid((size_t)rows * (D / 4))), dim3(256), 0, st, (size_t)rows, D / 4, mod, div, acc, (const float*)op.in[1],
           (const float4*)op.in[0], (float4*)op.out[0]);
  return check_launch("rows_axpy");
}

// dX[s][r][:] = dY[r][s * D : (s + 1) * D]   (backward of levels_to_joint)
__global__ void __launch_bounds__(256) joint_to_levels_kernel(int R, int slabs, int D4, const float4* __restrict__ dY, float4* __restrict__ dX) {
  pdl_wait();
  const size_t total = (size_t)R * slabs * D4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D4);
    const size_t q = i / D4;
    const int r = (int)(q % R), s = (int)(q / R);
    dX[i] = __ldg(dY + ((size_t)r * slabs + s) * D4 + c);
  }
}

int launch_joint_to_levels(const capf_op& op, cudaStream_t st) {
  const int R = op.i[0], slabs = op.i[1], D = op.i[2];
  if (R <= 0 || slabs <= 0 || D <= 0 || (D & 3) || !op.in[0] || !op.out[0]) return set_error(CAPF_ERR_ARG, "joint_to_levels: bad arguments");
  launch_k(joint_to_levels_kernel, dim3(ew_grid((size_t)R * slabs * (D / 4))), dim3(256), 0, st, R, slabs, D / 4, (const float4*)op.in[0], (float4*)op.out[0]);
  return check_launch("joint_to_levels");
}

// =======================================================================================================
// AdamW (torch.optim.AdamW semantics, decoupled weight decay; train.py:337-345) over one flat buffer.
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// hp = {lr, beta1, beta2, eps, weight_decay, 1 - beta1^t, sqrt(1 - beta2^t)} computed by the host in double.
// =======================================================================================================
__global__ void __launch_bounds__(256) adamw_kernel(size_t n, const float* __restrict__ hp, float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v) {
  pdl_wait();
  const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], wd = hp[4], bc1 = hp[5], bc2s = hp[6];
  const float step = lr / bc1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2s + eps;
    pi -= step * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

int launch_adamw(const capf_op& op, cudaStream_t st) {
  const size_t n = (size_t)(uint32_t)op.i[0] | ((size_t)(uint32_t)op.i[1] << 31);
  if (!n || !op.in[0] || !op.in[1] || !op.out[0] || !op.out[1] || !op.in[2]) return set_error(CAPF_ERR_ARG, "adamw: bad arguments");
  launch_k(adamw_kernel, dim3(ew_grid(n)), dim3(256), 0, st, n, (const float*)op.in[0], (float*)op.out[0], (const float*)op.in[1], (float*)op.out[1],
           (float*)op.in[2]);
  return check_launch("adamw");
}

}  // namespace capf
