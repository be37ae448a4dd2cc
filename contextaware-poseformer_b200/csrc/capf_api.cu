// C ABI of libcapf_b200: error handling, device info, plans (validated op programs) and dispatch.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <new>
#include <vector>

#include "capf_internal.h"

namespace capf {

int g_use_pdl = []() { const char* e = getenv("CAPF_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
static thread_local char g_err[512] = "";
static thread_local int g_sel_device = 0;                 // innermost DeviceGuard of this thread
static std::atomic<int> g_sms[CAPF_MAX_DEVICES];          // per-device SM count (0 = not probed yet)

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int set_errorf(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* name) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_errorf(CAPF_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e));
  return CAPF_OK;
}

int current_device() { return g_sel_device; }

int num_sms() {
  int n = g_sms[g_sel_device].load(std::memory_order_relaxed);
  return n > 0 ? n : 148;
}

// Validates `device` (an sm_100 GPU), makes it current for the lifetime of the guard and restores the caller's device.
DeviceGuard::DeviceGuard(int device) : status(CAPF_OK), prev_cuda_(-1), prev_sel_(g_sel_device), switched_(false) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    status = set_errorf(CAPF_ERR_CUDA, "no CUDA device (%s): libcapf_b200 has no CPU path", cudaGetErrorString(e));
    return;
  }
  if (device < 0 || device >= n || device >= CAPF_MAX_DEVICES) {
    status = set_errorf(CAPF_ERR_ARG, "device %d out of range (%d devices)", device, n);
    return;
  }
  if (g_sms[device].load(std::memory_order_relaxed) <= 0) {
    int sms = 0, major = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10) {
      status = set_errorf(CAPF_ERR_UNSUPPORTED, "device %d is sm_%dx; this library is built for sm_100a only", device, major);
      return;
    }
    g_sms[device].store(sms > 0 ? sms : 148, std::memory_order_relaxed);
  }
  cudaGetDevice(&prev_cuda_);
  if (prev_cuda_ != device) {
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
      status = set_errorf(CAPF_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
      return;
    }
    switched_ = true;
  }
  g_sel_device = device;
}

DeviceGuard::~DeviceGuard() {
  g_sel_device = prev_sel_;
  if (switched_) cudaSetDevice(prev_cuda_);
}

static int dispatch(const capf_op& op, const TcConvState* tc, cudaStream_t st) {
  switch (op.kind) {
    case CAPF_OP_CONV2D:
      if (op.i[12] == CAPF_IMPL_TCGEN05) return tc_conv_launch(op, tc, st);
      return launch_conv_simt(op, st);
    case CAPF_OP_BASICBLOCK: return tc_conv_launch(op, tc, st);
    case CAPF_OP_EXPAND_REDUCE: return tc_conv_launch(op, tc, st);
    case CAPF_OP_MLP: return tc_conv_launch(op, tc, st);
    case CAPF_OP_FUSE_SUM: return launch_fuse_sum(op, st);
    case CAPF_OP_MAXPOOL3X3S2: return launch_maxpool(op, st);
    case CAPF_OP_BILINEAR: return launch_bilinear(op, st);
    case CAPF_OP_LAYERNORM: return launch_layernorm(op, st);
    case CAPF_OP_ATTENTION: return launch_attention(op, st);
    case CAPF_OP_REF_SAMPLE:
    case CAPF_OP_DEFORM_SAMPLE: return launch_sample(op, st);
    case CAPF_OP_EMBED_COORD: return launch_embed_coord(op, st);
    case CAPF_OP_LEVELS_TO_JOINT: return launch_levels_to_joint(op, st);
    case CAPF_OP_CROP_NORMALIZE: return launch_crop_normalize(op, st);
    case CAPF_OP_CAST: return launch_cast(op, st);
    case CAPF_OP_PREPROCESS_U8: return launch_preprocess_u8(op, st);
    case CAPF_OP_WARP_AFFINE_U8: return launch_warp_affine_u8(op, st);
    case CAPF_OP_POSE_ERRORS: return launch_pose_errors(op, st);
    case CAPF_OP_GEMM_F32: return launch_gemm_f32(op, st);
    case CAPF_OP_COLSUM: return launch_colsum(op, st);
    case CAPF_OP_LAYERNORM_BWD: return launch_layernorm_bwd(op, st);
    case CAPF_OP_GELU:
    case CAPF_OP_GELU_BWD: return launch_gelu(op, st);
    case CAPF_OP_ATTENTION_BWD: return launch_attention_bwd(op, st);
    case CAPF_OP_DEFORM_BWD: return launch_deform_bwd(op, st);
    case CAPF_OP_ROWS_AXPY: return launch_rows_axpy(op, st);
    case CAPF_OP_JOINT_TO_LEVELS: return launch_joint_to_levels(op, st);
    case CAPF_OP_ADAMW: return launch_adamw(op, st);
    default: return set_errorf(CAPF_ERR_ARG, "unknown op kind %d", op.kind);
  }
}

}  // namespace capf

using namespace capf;

struct capf_plan {
  int device;
  std::vector<capf_op> ops;
  std::vector<TcConvState*> tc;  // parallel to ops; null unless impl == TCGEN05
};

extern "C" {

int capf_abi_version(void) { return CAPF_ABI_VERSION; }

const char* capf_last_error(void) { return g_err; }

int capf_device_info(int device, int64_t* out4) {
  if (!out4) return set_error(CAPF_ERR_ARG, "capf_device_info: null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || device < 0 || device >= n)
    return set_errorf(CAPF_ERR_CUDA, "capf_device_info: no such device (%s)", cudaGetErrorString(e));
  cudaDeviceProp p;
  if ((e = cudaGetDeviceProperties(&p, device)) != cudaSuccess)
    return set_errorf(CAPF_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  out4[0] = p.multiProcessorCount;
  out4[1] = p.major;
  out4[2] = p.minor;
  out4[3] = (int64_t)p.totalGlobalMem;
  return CAPF_OK;
}

int capf_plan_create(const capf_op* ops, int n_ops, int device, capf_plan** out_plan) {
  if (!ops || n_ops <= 0 || !out_plan) return set_error(CAPF_ERR_ARG, "capf_plan_create: bad arguments");
  DeviceGuard guard(device);
  int e = guard.status;
  if (e) return e;
  capf_plan* pl = new (std::nothrow) capf_plan();
  if (!pl) return set_error(CAPF_ERR_ARG, "capf_plan_create: out of host memory");
  pl->device = device;
  pl->ops.assign(ops, ops + n_ops);
  pl->tc.assign(n_ops, nullptr);
  for (int k = 0; k < n_ops; ++k) {
    capf_op& op = pl->ops[k];
    if (op.kind == CAPF_OP_BASICBLOCK) {
      if (!tc_block_supported(op)) {
        capf_plan_destroy(pl);
        return set_errorf(CAPF_ERR_UNSUPPORTED, "op %d: fused BasicBlock shape/dtype not supported", k);
      }
      e = tc_blockop_prepare(op, &pl->tc[k]);
      if (e) {
        capf_plan_destroy(pl);
        return e;
      }
    }
    if (op.kind == CAPF_OP_EXPAND_REDUCE) {
      e = tc_chainop_prepare(op, &pl->tc[k]);
      if (e) {
        capf_plan_destroy(pl);
        return e;
      }
    }
    if (op.kind == CAPF_OP_MLP) {
      e = tc_mlpop_prepare(op, &pl->tc[k]);
      if (e) {
        capf_plan_destroy(pl);
        return e;
      }
    }
    if (op.kind == CAPF_OP_CONV2D && op.i[12] != CAPF_IMPL_TCGEN05 && op.i[20] != 0) {
      capf_plan_destroy(pl);
      return set_errorf(CAPF_ERR_UNSUPPORTED, "op %d: output segments (i[20]) need the tcgen05 kernel", k);
    }
    if (op.kind == CAPF_OP_CONV2D && op.i[12] == CAPF_IMPL_TCGEN05) {
      if (!tc_conv_supported(op)) {
        capf_plan_destroy(pl);
        return set_errorf(CAPF_ERR_UNSUPPORTED, "op %d: conv2d shape/dtype not supported by the tcgen05 kernel", k);
      }
      e = tc_conv_prepare(op, &pl->tc[k]);
      if (e) {
        capf_plan_destroy(pl);
        return e;
      }
    }
  }
  *out_plan = pl;
  return CAPF_OK;
}

int capf_plan_run(const capf_plan* plan, int first, int count, void* stream) {
  if (!plan) return set_error(CAPF_ERR_ARG, "capf_plan_run: null plan");
  int n = (int)plan->ops.size();
  if (count < 0) count = n - first;
  if (first < 0 || first + count > n) return set_error(CAPF_ERR_ARG, "capf_plan_run: range out of bounds");
  DeviceGuard guard(plan->device);     // launches go to the plan's device whatever the caller's current device is
  if (guard.status) return guard.status;
  cudaStream_t st = (cudaStream_t)stream;
  for (int k = first; k < first + count; ++k) {
    int e = dispatch(plan->ops[k], plan->tc[k], st);
    if (e) {
      char prev[400];
      snprintf(prev, sizeof(prev), "%s", capf_last_error());
      return set_errorf(e, "op %d (kind %d): %s", k, plan->ops[k].kind, prev);
    }
  }
  return CAPF_OK;
}

int capf_plan_num_launches(const capf_plan* plan) { return plan ? (int)plan->ops.size() : 0; }

int capf_plan_op_kernel(const capf_plan* plan, int k, char* buf, int cap) {
  if (!plan || !buf || cap <= 0 || k < 0 || k >= (int)plan->ops.size()) return set_error(CAPF_ERR_ARG, "capf_plan_op_kernel: bad arguments");
  const capf_op& op = plan->ops[k];
  switch (op.kind) {
    case CAPF_OP_CONV2D:
      if (op.i[12] == CAPF_IMPL_TCGEN05) tc_conv_describe(plan->tc[k], buf, cap);
      else snprintf(buf, cap, "%s", stem_tc_supported(op) ? "stem_tc_kernel" : "conv_nhwc_simt");
      break;
    case CAPF_OP_BASICBLOCK: tc_conv_describe(plan->tc[k], buf, cap); break;
    case CAPF_OP_EXPAND_REDUCE: tc_conv_describe(plan->tc[k], buf, cap); break;
    case CAPF_OP_MLP: tc_conv_describe(plan->tc[k], buf, cap); break;
    case CAPF_OP_FUSE_SUM: snprintf(buf, cap, "fuse_sum_kernel"); break;
    case CAPF_OP_MAXPOOL3X3S2: snprintf(buf, cap, "maxpool3x3s2_kernel"); break;
    case CAPF_OP_BILINEAR: snprintf(buf, cap, "bilinear_ac_kernel"); break;
    case CAPF_OP_LAYERNORM: snprintf(buf, cap, op.i[3] > 0 ? "layernorm_proj_kernel" : "layernorm_kernel"); break;
    case CAPF_OP_ATTENTION: snprintf(buf, cap, "attention_kernel[seq %d]", op.i[1]); break;
    case CAPF_OP_REF_SAMPLE: snprintf(buf, cap, "ref_sample_kernel"); break;
    case CAPF_OP_DEFORM_SAMPLE: snprintf(buf, cap, "deform_sample_kernel"); break;
    case CAPF_OP_EMBED_COORD: snprintf(buf, cap, "embed_coord_kernel"); break;
    case CAPF_OP_LEVELS_TO_JOINT: snprintf(buf, cap, "levels_to_joint_kernel"); break;
    case CAPF_OP_CROP_NORMALIZE: snprintf(buf, cap, "crop_normalize_kernel"); break;
    case CAPF_OP_CAST: snprintf(buf, cap, "cast_kernel"); break;
    case CAPF_OP_PREPROCESS_U8: snprintf(buf, cap, "preprocess_u8_kernel"); break;
    case CAPF_OP_WARP_AFFINE_U8: snprintf(buf, cap, "warp_affine_u8_kernel"); break;
    case CAPF_OP_POSE_ERRORS: snprintf(buf, cap, "pose_errors_kernel"); break;
    case CAPF_OP_GEMM_F32: snprintf(buf, cap, "gemm_f32_kernel"); break;
    case CAPF_OP_COLSUM: snprintf(buf, cap, "colsum_kernel"); break;
    case CAPF_OP_LAYERNORM_BWD: snprintf(buf, cap, "layernorm_bwd_kernel"); break;
    case CAPF_OP_GELU: snprintf(buf, cap, "gelu_kernel"); break;
    case CAPF_OP_GELU_BWD: snprintf(buf, cap, "gelu_bwd_kernel"); break;
    case CAPF_OP_ATTENTION_BWD: snprintf(buf, cap, "attention_bwd_kernel"); break;
    case CAPF_OP_DEFORM_BWD: snprintf(buf, cap, "deform_bwd_kernel"); break;
    case CAPF_OP_ROWS_AXPY: snprintf(buf, cap, "rows_axpy_kernel"); break;
    case CAPF_OP_JOINT_TO_LEVELS: snprintf(buf, cap, "joint_to_levels_kernel"); break;
    case CAPF_OP_ADAMW: snprintf(buf, cap, "adamw_kernel"); break;
    default: snprintf(buf, cap, "?");
  }
  return CAPF_OK;
}

int capf_plan_destroy(capf_plan* plan) {
  if (!plan) return CAPF_OK;
  for (TcConvState* s : plan->tc)
    if (s) tc_conv_release(s);
  delete plan;
  return CAPF_OK;
}

int capf_op_run(const capf_op* op, int device, void* stream) {
  DeviceGuard guard(device);
  if (guard.status) return guard.status;
  capf_plan* pl = nullptr;
  int e = capf_plan_create(op, 1, device, &pl);
  if (e) return e;
  e = capf_plan_run(pl, 0, 1, stream);
  if (!e && pl->tc[0]) {
    // tensor maps live in the throw-away plan: finish before freeing it
    cudaError_t ce = cudaStreamSynchronize((cudaStream_t)stream);
    if (ce != cudaSuccess) e = set_errorf(CAPF_ERR_CUDA, "capf_op_run: %s", cudaGetErrorString(ce));
  }
  capf_plan_destroy(pl);
  return e;
}

int capf_crop_normalize(float* crop_xy, int n_points, void* stream) {
  capf_op op;
  memset(&op, 0, sizeof(op));
  op.kind = CAPF_OP_CROP_NORMALIZE;
  op.i[0] = n_points;
  op.out[0] = crop_xy;
  // the device is the one that owns the caller's tensor, not whatever happens to be current
  if (!crop_xy || n_points < 0) return set_error(CAPF_ERR_ARG, "capf_crop_normalize: bad arguments");
  if (((uintptr_t)crop_xy & 7) != 0) return set_error(CAPF_ERR_ARG, "capf_crop_normalize: pointer must be 8-byte aligned (float2 access)");
  cudaPointerAttributes at;
  cudaError_t ce = cudaPointerGetAttributes(&at, crop_xy);
  if (ce != cudaSuccess || at.type != cudaMemoryTypeDevice) {
    cudaGetLastError();
    return set_error(CAPF_ERR_ARG, "capf_crop_normalize: crop_xy is not a device pointer (there is no CPU path)");
  }
  return capf_op_run(&op, at.device, stream);
}

}  // extern "C"
