// tcgen05 all-pairs correlation GEMM with fused pyramid epilogue (f16/bf16).  Placeholder.
#include "common.cuh"

namespace pfb {
bool corr_volume_umma_supported(int, int, int, int, int, pfb_dtype) { return false; }
int corr_volume_umma(const void*, const void*, void* const*, int, int, int, int, int, pfb_dtype, cudaStream_t) {
  set_error("corr_volume_build: tcgen05 path not built");
  return PFB_ERR_UNSUPPORTED;
}
}  // namespace pfb
