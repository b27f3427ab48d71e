// Host-side construction of 2-D bf16 tensor maps (cuTensorMapEncodeTiled resolved through the runtime, no -lcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "host_util.h"

namespace rtts {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// `inner` contiguous elements per row, `outer` rows of `ld` elements; box = box_inner x box_outer; SWIZZLE_128B.
// A box of {64, 1} is the form tile::gather4 loads take (four such rows per instruction).
inline int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, int64_t ld, uint32_t box_inner,
                          uint32_t box_outer, CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(kErrCuda, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(kErrCuda, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return kOk;
}

}  // namespace rtts
