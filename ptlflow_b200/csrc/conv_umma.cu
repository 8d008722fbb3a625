// tcgen05 implicit-GEMM convolution (f16/bf16).  Placeholder until the UMMA kernel lands:
// reports "unsupported" so pfb_conv2d(impl=auto) routes to the SIMT kernel.
#include "common.cuh"

namespace pfb {
bool conv2d_umma_supported(const pfb_conv_params*) { return false; }
int conv2d_umma(const pfb_conv_params*, cudaStream_t) {
  set_error("conv2d: tcgen05 path not built");
  return PFB_ERR_UNSUPPORTED;
}
}  // namespace pfb
