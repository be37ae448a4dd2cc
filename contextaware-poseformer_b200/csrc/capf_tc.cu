// tcgen05 / TMA path -- placeholder until the tensor-core kernels land.
#include "capf_internal.h"
namespace capf {
struct TcConvState {};
int tc_conv_supported(const capf_op&) { return 0; }
int tc_conv_prepare(const capf_op&, TcConvState** out) { *out = nullptr; return set_error(CAPF_ERR_UNSUPPORTED, "tcgen05 conv not built"); }
int tc_conv_launch(const capf_op&, const TcConvState*, cudaStream_t) { return set_error(CAPF_ERR_UNSUPPORTED, "tcgen05 conv not built"); }
void tc_conv_release(TcConvState* s) { delete s; }
}  // namespace capf
