// Host-side TMA descriptor creation.  cuTensorMapEncodeTiled is a driver API symbol; it is fetched
// through the runtime (cudaGetDriverEntryPoint) so the library does not link against libcuda.
#include "umma.cuh"

namespace pfb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_map_impl(CUtensorMap* out, const void* base, pfb_dtype dt, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);

int make_tensor_map(CUtensorMap* out, const void* base, pfb_dtype dt, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box) {
  return make_map_impl(out, base, dt, rank, dims, strides_bytes, box, true);
}
int make_tensor_map_linear(CUtensorMap* out, const void* base, pfb_dtype dt, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box) {
  return make_map_impl(out, base, dt, rank, dims, strides_bytes, box, false);
}

static int make_map_impl(CUtensorMap* out, const void* base, pfb_dtype dt, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return PFB_ERR_CUDA;
  }
  PFB_CHECK_ARG(rank >= 2 && rank <= 5, "tensor map rank %d", rank);
  PFB_CHECK_ARG(dt == PFB_F16 || dt == PFB_BF16, "tensor map: only 16-bit element types");
  PFB_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base pointer must be 16-byte aligned");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      PFB_CHECK_ARG(strides_bytes[i - 1] % 16 == 0, "tensor map: stride %d = %llu bytes not a multiple of 16", i,
                    (unsigned long long)strides_bytes[i - 1]);
      gstr[i - 1] = strides_bytes[i - 1];
    }
  }
  if (swizzle128) PFB_CHECK_ARG(box[0] * 2 == 128, "tensor map: inner box must span 128 bytes for SWIZZLE_128B");
  else PFB_CHECK_ARG((box[0] * 2) % 16 == 0, "tensor map: inner box must be a multiple of 16 bytes");
  CUresult r = enc(out, dt == PFB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                   const_cast<void*>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu ...)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1]);
    return PFB_ERR_CUDA;
  }
  return PFB_OK;
}

}  // namespace pfb
