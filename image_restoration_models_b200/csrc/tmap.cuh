// Host-side tensor-map construction for the TMA-fed kernels.
#pragma once
#include <cuda.h>
#include <string>

#include "common.cuh"

namespace irb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
static inline EncodeTiledFn tmap_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// rank-`rank` tiled map: dims[] / box[] in elements (fastest first), strides[] in BYTES for dims 1..rank-1.
// Out-of-range elements read as zero and are not written.
static inline int make_tmap(CUtensorMap* tm, const void* ptr, bool half, int rank, const cuuint64_t* dims,
                            const cuuint64_t* strides, const cuuint32_t* box, bool swizzle128) {
  EncodeTiledFn fn = tmap_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return IR_ERR_CUDA; }
  cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  const CUresult r = fn(tm, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                        const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return IR_ERR_CUDA;
  }
  return IR_OK;
}

}  // namespace irb
