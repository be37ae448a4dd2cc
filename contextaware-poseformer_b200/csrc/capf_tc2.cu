// 2-CTA (cta_group::2) tcgen05 GEMM of libcapf_b200 for the wide transformer Linears of the lifter (pose_dformer.py:49,
// 56, 25-31: joint-block qkv / proj / fc1 / fc2, M = 17 * B rows).
//
// A single-CTA SS-mode MMA reads all of B (BN x 16 x 2 B) from its own shared memory for every 128-row MMA, and TMA
// writes share the same 128 B/clk port: measured on the QKV GEMM the MMA phase tops out at 56 % (128 x 240 tiles) to
// 81 % (256 x 240) of the tensor peak (tools/gemm_trace.py).  Here a CTA PAIR (cluster of two SMs of one TPC) computes a
// 256 x BN tile: each CTA loads its own 128 rows of A and HALF of the B tile, the leader's single thread issues
// tcgen05.mma.cta_group::2 (M = 256), each tensor core reads its local A and both B halves through the pair datapath,
// and each CTA's TMEM receives its 128 accumulator rows.  Shared-memory traffic per SM per 64-deep stage drops to
// write 31 KB + read 31 KB for 480 tensor clocks -- MMA-bound.
//
// Protocol (per pipeline stage s, accumulator stage a; barriers live at the same offset in both CTAs):
//   full[s]    leader only: 1 arrival (leader's producer, expect_tx = bytes of BOTH CTAs); the TMA loads of both CTAs
//              complete_tx on the leader's barrier (cp.async.bulk.tensor...cta_group::2 with a mapa'd mbarrier address)
//   empty[s]   both CTAs: tcgen05.commit.cta_group::2...multicast::cluster from the leader's MMA thread
//   tfull[a]   both CTAs: same multicast commit after the last stage of a tile
//   tempty[a]  leader only: 2 x 8 arrivals -- the epilogue warps of both CTAs (the peer's arrive remotely)
#include <cstdio>
#include <cstdlib>
#include <new>

#include "capf_tc.cuh"

namespace capf {

constexpr int T2_THREADS = 384;             // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue
constexpr int T2_MAX_STAGES = 8;
constexpr int T2_HEADER = 1024;
constexpr int T2_STG_BYTES = 32 * 128;      // per-warp staging tile: 32 rows x 128 B
constexpr int T2_A_BYTES = 128 * 128;       // A stage of one CTA: 128 rows x 64 elements x 2 B

struct Tc2P {
  int M, K, Cout, BN, n_tiles_n, num_tiles, num_k, num_stages, nacc, tmem_cols;
  int stage_bytes, stg_bufs, act;
  // conv mode (3x3 / stride 1 / pad 1 over WHOLE small images: H * W divides 128, e.g. the 8 x 8 maps of HRNet's 256-channel branch):
  // a CTA's 128 rows are img_per_cta complete images = contiguous NHWC rows, so only the A loads differ from the Linear case --
  // K step ks = (tap, 64-channel chunk) is one 4-D box {64, W, H, img_per_cta} at (chunk, s - 1, r - 1, image), zero-filled outside
  int conv, cpt, img_hw;
  uint32_t idesc, desc_hi, stg_off, bias_off;
  const float* bias;
  const void* res;
  void* out;
  long long* trace;          // optional (debug, op.in[4]): clock64 timeline of the leader CTA of pair 0, tools/gemm_trace.py
};

// MODE: bit 0 = residual add, bit 1 = GELU
template <typename TO, int MODE>
__global__ void __launch_bounds__(T2_THREADS, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Tc2P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar_full = base;                           // [T2_MAX_STAGES]
  const uint32_t bar_empty = base + 8 * T2_MAX_STAGES;      // [T2_MAX_STAGES]
  const uint32_t bar_tfull = base + 16 * T2_MAX_STAGES;     // [2]
  const uint32_t bar_tempty = bar_tfull + 16;               // [2]
  const uint32_t tmem_slot = bar_tempty + 16;
  const uint32_t stage0 = base + T2_HEADER;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx2::cluster_ctarank();
  const bool leader = rank == 0;
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
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);
      ptx::mbar_init(bar_tempty + 8 * a, 2 * 8);                 // one arrival per epilogue warp, both CTAs
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {                    // one warp of EACH CTA of the pair executes the cta_group::2 allocation
    ptx2::tmem_alloc2(tmem_slot, (uint32_t)p.tmem_cols);
    ptx2::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx2::cluster_sync();               // barriers of both CTAs initialised, TMEM of both allocated
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (tr && threadIdx.x == 0) p.trace[2] = clock64();
  pdl_trigger();
  // griddepcontrol.wait is per thread: every role waits before it touches anything the predecessor produced (A, the
  // residual, the output).  The producer first requests the WEIGHT halves of its first stages -- they do not depend
  // on the predecessor -- so their HBM/L2 latency overlaps the predecessor's tail.
  if (warp != 0) pdl_wait();

  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int t0 = (int)(((long long)p.num_tiles * pair) / n_pairs);
  const int t1 = (int)(((long long)p.num_tiles * (pair + 1)) / n_pairs);
  const int half_bn = p.BN >> 1;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =========================
    if (ptx::elect_one()) {
      int pre = 0;                                  // stages of the first tile whose B half is already in flight
      if (t0 < t1) {
        const int nb0 = (t0 % p.n_tiles_n) * p.BN + (int)rank * half_bn;
        pre = min(p.num_stages, p.num_k);
        for (int ks = 0; ks < pre; ++ks) {
          if (leader) ptx::mbar_arrive_expect_tx(bar_full + 8 * ks, (uint32_t)(2 * p.stage_bytes));
          ptx2::tma_load_2d_2sm(&mapB, ptx2::mapa(bar_full + 8 * ks, 0u), stage0 + ks * p.stage_bytes + T2_A_BYTES, ks * 64, nb0);
        }
      }
      pdl_wait();
      if (tr) p.trace[3] = clock64();
      uint32_t stage = 0, phase = 0;
      const uint64_t pol_a = ptx::policy_evict_last();          // conv mode: every input pixel is fetched once per tap
      for (int tile = t0; tile < t1; ++tile) {
        const int n_tile = tile % p.n_tiles_n, m_tile = tile / p.n_tiles_n;
        const int m0 = m_tile * 256 + (int)rank * 128;
        const int nb0 = n_tile * p.BN + (int)rank * half_bn;
        const int img0 = p.conv ? m0 / p.img_hw : 0;
        int cc = 0, tap_s = 0, tap_r = 0;                        // conv mode: channel chunk inside the tap, filter column / row
        for (int ks = 0; ks < p.num_k; ++ks) {
          const bool primed = tile == t0 && ks < pre;
          if (!primed) ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
          const uint32_t full_leader = ptx2::mapa(bar_full + 8 * stage, 0u);
          if (leader && !primed) ptx::mbar_arrive_expect_tx(bar_full + 8 * stage, (uint32_t)(2 * p.stage_bytes));
          const uint32_t a_dst = stage0 + stage * p.stage_bytes;
          if (p.conv) {
            ptx2::tma_load_4d_2sm_hint(&mapA, full_leader, a_dst, cc * 64, tap_s - 1, tap_r - 1, img0, pol_a);
            if (++cc == p.cpt) {
              cc = 0;
              if (++tap_s == 3) { tap_s = 0; ++tap_r; }
            }
          } else {
            ptx2::tma_load_2d_2sm(&mapA, full_leader, a_dst, ks * 64, m0);
          }
          if (!primed) ptx2::tma_load_2d_2sm(&mapB, full_leader, a_dst + T2_A_BYTES, ks * 64, nb0);
          if (tr) { const int n = (tile - t0) * p.num_k + ks; if (n < 64) p.trace[64 + n] = clock64(); }
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) ======================
    if (leader) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = t0; tile < t1; ++tile) {
        ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.BN;
        uint32_t accumulate = 0;
        for (int ks = 0; ks < p.num_k; ++ks) {
          ptx::mbar_wait(bar_full + 8 * stage, phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t a_src = stage0 + stage * p.stage_bytes;
            const uint64_t a_desc = tc_make_desc(a_src, 1u, p.desc_hi), b_desc = tc_make_desc(a_src + T2_A_BYTES, 1u, p.desc_hi);
#pragma unroll
            for (int i = 0; i < 4; ++i) ptx2::umma2_f16(d_tmem, a_desc + 2u * i, b_desc + 2u * i, p.idesc, (accumulate | (uint32_t)i) ? 1u : 0u);
            ptx2::umma2_commit_mc(bar_empty + 8 * stage);
            if (ks + 1 == p.num_k) ptx2::umma2_commit_mc(bar_tfull + 8 * acc);
            if (tr) { const int n = (tile - t0) * p.num_k + ks; if (n < 64) p.trace[128 + n] = clock64(); }
          }
          __syncwarp();
          accumulate = 1;
          if (++stage == (uint32_t)p.num_stages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == (uint32_t)p.nacc) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue (both CTAs: own 128 rows) ================
    constexpr bool HAS_RES = (MODE & 1) != 0;
    constexpr int SW = 128 / (int)sizeof(TO);
    constexpr int CPG = 16 * (int)sizeof(TO) / 16;
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int split = ((p.BN / 16 + 1) / 2) * 16;
    const int cbeg = half ? split : 0, cend = half ? p.BN : split;
    const int nslabs = (cend - cbeg + SW - 1) / SW;
    const TO* res = reinterpret_cast<const TO*>(p.res);
    TO* out = reinterpret_cast<TO*>(p.out);
    const float floor_v = p.act == CAPF_ACT_RELU ? 0.f : -__int_as_float(0x7f800000);
    uint8_t* const stg_ptr = smem_raw + (stage0 - raw) + p.stg_off + (uint32_t)(warp - 4) * (uint32_t)(p.stg_bufs * T2_STG_BYTES);
    const uint32_t stg = stage0 + p.stg_off + (uint32_t)(warp - 4) * (uint32_t)(p.stg_bufs * T2_STG_BYTES);
    float* const sbias = reinterpret_cast<float*>(smem_raw + (stage0 - raw) + p.bias_off) + (warp - 4) * 128;
    auto slot_off = [&](uint32_t buf, int r, int c) { return buf * T2_STG_BYTES + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); };
    const uint64_t pol_out = ptx::policy_evict_last();
    const uint32_t row_bytes = (uint32_t)p.Cout * (uint32_t)sizeof(TO);
    const uint32_t tempty_leader0 = ptx2::mapa(bar_tempty, 0u);
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = t0; tile < t1; ++tile) {
      const int n_tile = tile % p.n_tiles_n, m_tile = tile / p.n_tiles_n;
      const int ncol0 = n_tile * p.BN;
      const int m_w0 = m_tile * 256 + (int)rank * 128 + q * 32;        // first matrix row of this warp
      const int rows_live = p.M - m_w0;
      const uint32_t taddr = tmem_base + acc * (uint32_t)p.BN + ((uint32_t)(q * 32) << 16);
      if (4 * lane < cend - cbeg)
        *reinterpret_cast<float4*>(sbias + 4 * lane) = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + ncol0 + cbeg + 4 * lane)) : make_float4(0.f, 0.f, 0.f, 0.f);
      auto prefetch_res = [&](int s, uint32_t buf) {
        const int c0 = cbeg + s * SW;
        const int chs = (min(SW, cend - c0) * (int)sizeof(TO)) >> 4;
        const uint32_t magic = (65536u + (uint32_t)chs - 1u) / (uint32_t)chs;
        const uint8_t* gbase = reinterpret_cast<const uint8_t*>(res + (size_t)m_w0 * p.Cout + ncol0 + c0);
        for (int i = 0; i < chs; ++i) {
          const int item = i * 32 + lane, r = (int)(((uint32_t)item * magic) >> 16), c = item - r * chs;
          if (r < rows_live) ptx::cp_async16(stg + slot_off(buf, r, c), gbase + (uint32_t)r * row_bytes + 16 * c);
        }
        ptx::cp_async_commit();
      };
      if (HAS_RES) prefetch_res(0, 0);
      __syncwarp();
      ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
      ptx::tc_fence_after();
      if (tr && warp == 4 && lane == 0 && tile - t0 < 16) p.trace[192 + 2 * (tile - t0)] = clock64();
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
            ptx::tmem_ld_wait();
            if (HAS_RES && g == 0) {
              if (s + 1 < nslabs) ptx::cp_async_wait_group1(); else ptx::cp_async_wait_all();
              __syncwarp();
            }
            epi16<TO, MODE>(a0, sbias + (c0 - cbeg) + g, floor_v, stg_ptr + buf * T2_STG_BYTES + lane * 128, (uint32_t)((g / 16) * CPG), (uint32_t)lane & 7u);
            if (two) epi16<TO, MODE>(a1, sbias + (c0 - cbeg) + g + 16, floor_v, stg_ptr + buf * T2_STG_BYTES + lane * 128, (uint32_t)((g / 16 + 1) * CPG), (uint32_t)lane & 7u);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        // accumulator read completely (tcgen05.wait::ld above): free it at the leader, one relaxed arrival per warp -- a
        // release.cluster arrive per thread cost ~800 clk in front of the slab's stores (measured in capf_tc_block64.cu)
        if (s + 1 == nslabs && lane == 0) ptx2::mbar_arrive_cluster_relaxed(tempty_leader0 + 8 * acc);
        const int chs = (ncol * (int)sizeof(TO)) >> 4;
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
        __syncwarp();
        if (HAS_RES) buf ^= 1u;
      }
      if (nslabs == 0) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx2::mbar_arrive_cluster_relaxed(tempty_leader0 + 8 * acc);
      }
      if (tr && warp == 4 && lane == 0 && tile - t0 < 16) p.trace[193 + 2 * (tile - t0)] = clock64();
      if (++acc == (uint32_t)p.nacc) { acc = 0; acc_phase ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx2::cluster_sync();               // the peer may still be reading operands of / arriving at this CTA
  if (warp == 2) ptx2::tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
  if (tr && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.trace[4] = clock64();
    p.trace[5] = (long long)gt;
    p.trace[6] = 2; p.trace[7] = p.BN; p.trace[8] = p.num_stages; p.trace[9] = p.num_tiles; p.trace[10] = gridDim.x; p.trace[11] = p.nacc;
  }
}

// =======================================================================================================
// host side
// =======================================================================================================
struct Tc2State {
  CUtensorMap mapA, mapB;
  Tc2P p;
  int grid, smem_bytes, dtype_out;
};

static int ceil_div2(int a, int b) { return (a + b - 1) / b; }

// 1 if the 2-CTA kernel should take this op: a plain Linear over rows with enough work for CTA pairs to pay off.
int tc2_supported(const capf_op& op) {
  const char* ev = getenv("CAPF_TC2");
  if (ev && ev[0] == '0') return 0;
  if (op.i[17] == 1) return 0;                                  // i[17]: 1 = never, 2 = force (tests)
  if (op.dtype_in != CAPF_F16 && op.dtype_in != CAPF_BF16) return 0;
  if (op.kind == CAPF_OP_CONV2D && op.i[5] == 3 && op.i[6] == 3 && op.i[7] == 1 && op.i[8] == 1) {
    // conv mode: whole small images per CTA (see Tc2P); taken by default for the 256-channel branch, whose per-tap launches are bound by
    // every CTA streaming the whole 1.2 MB weight matrix for 128 rows -- a pair streams half of it per CTA
    const int hw = op.i[1] * op.i[2], Cin = op.i[3], N = op.i[4];
    if (hw < 1 || hw > 128 || 128 % hw || Cin % 64 || N % 16 || op.i[0] < 1) return 0;
    if (op.i[17] == 2) return 1;
    return (Cin >= 256 && N >= 256 && (long long)op.i[0] * hw >= 2048) ? 1 : 0;
  }
  if (op.kind != CAPF_OP_CONV2D || op.i[5] != 1 || op.i[6] != 1 || op.i[7] != 1 || op.i[8] != 0) return 0;
  const int M = op.i[0] * op.i[1] * op.i[2], K = op.i[3], N = op.i[4];
  if (K % 64 || N % 16 || M < 1) return 0;
  if (op.i[17] == 2) return 1;
  { const char* t = getenv("CAPF_TC2_MIN_N"); const int min_n = t ? atoi(t) : 512; return (M >= 2048 && K >= 512 && N >= min_n) ? 1 : 0; }   // the joint-block Linears (qkv 1920, fc1 1280, proj / fc2 640 wide: 311 -> 290 us per forward with the 640-wide ones on pairs too)
}

int tc2_prepare(const capf_op& op, Tc2State** out) {
  *out = nullptr;
  int e = tc_get_encoder();
  if (e) return e;
  Tc2State* s = new (std::nothrow) Tc2State();
  if (!s) return set_error(CAPF_ERR_ARG, "tc2_prepare: out of host memory");
  Tc2P& p = s->p;
  memset(&p, 0, sizeof(p));
  p.M = op.i[0] * op.i[1] * op.i[2]; p.K = op.i[3]; p.Cout = op.i[4];
  p.conv = op.i[5] == 3 ? 1 : 0;
  if (p.conv) { p.cpt = op.i[3] / 64; p.img_hw = op.i[1] * op.i[2]; p.K = 9 * op.i[3]; }
  p.act = op.i[11];
  p.bias = (const float*)op.in[2];
  p.res = op.in[3];
  p.out = op.out[0];
  p.trace = (long long*)op.in[4];     // debug only (NULL in every program the host layer builds)
  p.num_k = p.K / 64;
  const int osz = op.dtype_out == CAPF_F32 ? 4 : 2;
  p.stg_bufs = p.res ? 2 : 1;
  const int stg_bytes = 8 * p.stg_bufs * T2_STG_BYTES + 8 * 128 * 4;
  const int pairs = num_sms() / 2;
  const int m_tiles = ceil_div2(p.M, 256);
  // column tile: multiple of 16 (each CTA holds BN / 2 rows of B, whole 8-row swizzle groups) dividing Cout, <= 256; cost = rounds of
  // pair tiles x per-tile tensor time (+ an epilogue term), with the i[16] override used by the tests
  double best = 1e300;
  int best_bn = 0;
  int env_bn = 0;
  { const char* ev = getenv("CAPF_TC2_BN"); if (ev) env_bn = atoi(ev); }     // experiments: force the column tile where it divides Cout
  for (int bn = 16; bn <= 256 && bn <= p.Cout; bn += 16) {
    if (p.Cout % bn) continue;
    if (op.i[16] > 0 && bn != op.i[16]) continue;
    if (env_bn > 0 && op.i[16] == 0 && p.Cout % env_bn == 0 && bn != env_bn) continue;
    const int stage_bytes = T2_A_BYTES + (bn / 2) * 128;
    if ((TC_SMEM_LIMIT - T2_HEADER - 1024 - stg_bytes) / stage_bytes < 3) continue;
    const long long tiles = (long long)m_tiles * (p.Cout / bn);
    const long long g = tiles < pairs ? tiles : pairs;
    const double per_pair = (double)((tiles + g - 1) / g);
    // per pair tile and SM: tensor clocks, and the L2 -> SM operand feed of (128 + bn / 2) rows x K at the chip-wide L2
    // cap of ~42 B/clk/SM (what actually bounds these GEMMs: every operand byte is fetched ~10x through L2)
    const double t_mma = p.num_k * 4.0 * (bn / 2.0 > 40.0 ? bn / 2.0 : 40.0);
    const double t_feed = (128.0 + bn / 2.0) * p.K * 2.0 / 42.0;
    const double t_main = t_mma > t_feed ? t_mma : t_feed;
    const double t_epi = 128.0 * bn * osz * (p.res ? 2.0 : 1.0) / 16.0 + 400.0;
    const bool dbl = 2 * bn <= 512;
    double cost = dbl ? per_pair * (t_main > t_epi ? t_main : t_epi) + (t_main < t_epi ? t_main : t_epi) : per_pair * (t_main + t_epi);
    cost += 600.0 * per_pair;
    if (cost < best) { best = cost; best_bn = bn; }
  }
  if (!best_bn) { delete s; return set_error(CAPF_ERR_UNSUPPORTED, "tc2: no column tile"); }
  p.BN = best_bn;
  p.n_tiles_n = p.Cout / p.BN;
  p.num_tiles = m_tiles * p.n_tiles_n;
  p.stage_bytes = T2_A_BYTES + (p.BN / 2) * 128;
  int stages = (TC_SMEM_LIMIT - T2_HEADER - 1024 - stg_bytes) / p.stage_bytes;
  if (stages > T2_MAX_STAGES) stages = T2_MAX_STAGES;
  p.num_stages = stages;
  p.stg_off = (uint32_t)(stages * p.stage_bytes);
  p.bias_off = p.stg_off + (uint32_t)(8 * p.stg_bufs * T2_STG_BYTES);
  s->smem_bytes = T2_HEADER + 1024 + stages * p.stage_bytes + stg_bytes;
  if (s->smem_bytes < 120 * 1024) s->smem_bytes = 120 * 1024;
  p.nacc = 2 * p.BN <= 512 ? 2 : 1;
  int cols = 32;
  while (cols < p.nacc * p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const uint32_t fmt = op.dtype_in == CAPF_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // M = 256
  p.desc_hi = tc_desc_hi(128, 1024);
  const int g = p.num_tiles < pairs ? p.num_tiles : pairs;
  s->grid = 2 * g;
  s->dtype_out = op.dtype_out;
  const CUtensorMapDataType dt = op.dtype_in == CAPF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  if (p.conv) {
    const int Cin = op.i[3], H = op.i[1], W = op.i[2];
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)op.i[0]};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)H, (cuuint32_t)(128 / p.img_hw)};
    cuuint32_t es[4] = {1, 1, 1, 1};
    e = tc_encode_map(&s->mapA, dt, 4, op.in[0], dims, strides, box, es, 128, "A whole-image boxes (2-CTA)");
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
    cuuint64_t strides[1] = {(cuuint64_t)p.K * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    e = tc_encode_map(&s->mapA, dt, 2, op.in[0], dims, strides, box, es, 128, "A rows (2-CTA)");
  }
  if (!e) {
    cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.Cout};
    cuuint64_t strides[1] = {(cuuint64_t)p.K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(p.BN / 2)};
    cuuint32_t es[2] = {1, 1};
    e = tc_encode_map(&s->mapB, dt, 2, op.in[1], dims, strides, box, es, 128, "B weights (2-CTA)");
  }
  if (e) { delete s; return e; }
  *out = s;
  return CAPF_OK;
}

template <typename TO, int MODE>
static int tc2_launch_mode(const Tc2State* s, cudaStream_t st) {
  static PerDevice<bool> opted_;
  std::atomic<bool>& opted = opted_.get();
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm2_kernel<TO, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT);
    if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_gemm2_kernel smem opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(s->grid);
  cfg.blockDim = dim3(T2_THREADS);
  cfg.dynamicSmemBytes = s->smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_gemm2_kernel<TO, MODE>, s->mapA, s->mapB, s->p);
  if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "tc_gemm2_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("tc_gemm2_kernel");
}

template <typename TO>
static int tc2_launch_typed(const Tc2State* s, cudaStream_t st) {
  const int mode = (s->p.res ? 1 : 0) | (s->p.act == CAPF_ACT_GELU ? 2 : 0);
  switch (mode) {
    case 0: return tc2_launch_mode<TO, 0>(s, st);
    case 1: return tc2_launch_mode<TO, 1>(s, st);
    case 2: return tc2_launch_mode<TO, 2>(s, st);
    default: return tc2_launch_mode<TO, 3>(s, st);
  }
}

int tc2_launch(const Tc2State* s, cudaStream_t st) {
  switch (s->dtype_out) {
    case CAPF_F32: return tc2_launch_typed<float>(s, st);
    case CAPF_F16: return tc2_launch_typed<__half>(s, st);
    case CAPF_BF16: return tc2_launch_typed<__nv_bfloat16>(s, st);
    default: return set_error(CAPF_ERR_UNSUPPORTED, "tc2: dtype_out");
  }
}

void tc2_release(Tc2State* s) { delete s; }

void tc2_describe(const Tc2State* s, char* buf, int cap) {
  snprintf(buf, cap, "tc_gemm2_kernel[2-CTA 256x%d tile, %d stages%s]", s->p.BN, s->p.num_stages, s->p.conv ? ", 3x3 over whole-image boxes" : "");
}

}  // namespace capf
