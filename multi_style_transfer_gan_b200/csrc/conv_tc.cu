// placeholder -- replaced by the tcgen05 implicit-GEMM kernel
#include "common.cuh"
namespace msg {
bool conv2d_tc_supported(const msg_conv_desc*, const void*, const void*, const void*) { return false; }
int conv2d_tc(const msg_conv_desc*, const void*, const void*, const float*, void*, double*, const double*, cudaStream_t) {
  set_error("conv_tc: not built");
  return MSG_ERR_UNSUPPORTED;
}
}  // namespace msg
