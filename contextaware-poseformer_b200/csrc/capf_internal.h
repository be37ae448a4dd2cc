// Internal (non-ABI) declarations shared by the translation units of libcapf_b200.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "../../include/capf_b200.h"

namespace capf {

// Device selection.  Every extern "C" entry point that touches the GPU opens a DeviceGuard for the device it targets
// (the plan's device, or the device that owns the caller's pointer) and restores the caller's current device on
// exit, so the library never changes the current device behind torch's back and a process may drive several GPUs.
constexpr int CAPF_MAX_DEVICES = 64;
struct DeviceGuard {
  explicit DeviceGuard(int device);   // status != CAPF_OK: device unusable (message recorded)
  ~DeviceGuard();
  int status;
 private:
  int prev_cuda_, prev_sel_;
  bool switched_;
};
int current_device();   // ordinal selected by the innermost DeviceGuard of this thread
int num_sms();          // SM count of current_device() (148 on B200)

// Per-device state of a kernel instantiation: function attributes (the > 48 KB dynamic shared-memory opt-in) belong to
// a (function, device) pair, not to the process.  Races are benign (the guarded action is idempotent).
template <typename T>
struct PerDevice {
  std::atomic<T> v[CAPF_MAX_DEVICES];
  std::atomic<T>& get() { return v[current_device()]; }
};

int set_error(int code, const char* msg);       // records thread-local message, returns code
int set_errorf(int code, const char* fmt, ...);
int check_launch(const char* kernel_name);      // cudaGetLastError() -> status

// SIMT launchers (capf_simt.cu)
int launch_conv_simt(const capf_op& op, cudaStream_t st);
int launch_fuse_sum(const capf_op& op, cudaStream_t st);
int launch_maxpool(const capf_op& op, cudaStream_t st);
int launch_bilinear(const capf_op& op, cudaStream_t st);
int launch_layernorm(const capf_op& op, cudaStream_t st);
int launch_attention(const capf_op& op, cudaStream_t st);
int launch_sample(const capf_op& op, cudaStream_t st);
int launch_embed_coord(const capf_op& op, cudaStream_t st);
int launch_levels_to_joint(const capf_op& op, cudaStream_t st);
int launch_crop_normalize(const capf_op& op, cudaStream_t st);
int launch_cast(const capf_op& op, cudaStream_t st);
int launch_preprocess_u8(const capf_op& op, cudaStream_t st);
int launch_warp_affine_u8(const capf_op& op, cudaStream_t st);
int launch_pose_errors(const capf_op& op, cudaStream_t st);

// training step (capf_train.cu)
int launch_gemm_f32(const capf_op& op, cudaStream_t st);
int launch_colsum(const capf_op& op, cudaStream_t st);
int launch_layernorm_bwd(const capf_op& op, cudaStream_t st);
int launch_gelu(const capf_op& op, cudaStream_t st);
int launch_attention_bwd(const capf_op& op, cudaStream_t st);
int launch_deform_bwd(const capf_op& op, cudaStream_t st);
int launch_rows_axpy(const capf_op& op, cudaStream_t st);
int launch_joint_to_levels(const capf_op& op, cudaStream_t st);
int launch_adamw(const capf_op& op, cudaStream_t st);

// tensor-pipe HRNet stem conv1 (capf_stem.cu): fp32 NHWC image -> 64 channels, 3x3 / stride 2
int stem_tc_supported(const capf_op& op);
int launch_stem_tc(const capf_op& op, cudaStream_t st);

// tcgen05 path (capf_tc.cu): per-op prepared state lives in the plan
struct TcConvState;                                   // tensor maps + launch geometry
int tc_conv_supported(const capf_op& op);             // 1 if the tcgen05 kernel handles this op
int tc_conv_prepare(const capf_op& op, TcConvState** out);
int tc_block_supported(const capf_op& op);            // fused BasicBlock op (capf_tc_block.cu)
int tc_blockop_prepare(const capf_op& op, TcConvState** out);
int tc_chainop_prepare(const capf_op& op, TcConvState** out);   // CAPF_OP_EXPAND_REDUCE (capf_tc_chain.cu)
int tc_mlpop_prepare(const capf_op& op, TcConvState** out);     // CAPF_OP_MLP (capf_tc_mlp.cu)
int tc_conv_launch(const capf_op& op, const TcConvState* s, cudaStream_t st);
void tc_conv_release(TcConvState* s);
void tc_conv_describe(const TcConvState* s, char* buf, int cap);   // kernel name + tile shape

}  // namespace capf
