// Host-side CUtensorMap construction (TMA descriptors) shared by the tensor-core translation units.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace bci {

// driver entry point for tensor-map encoding, fetched through the runtime (no libcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D bf16 row-major [rows][cols] tensor, box = box_rows x box_cols (64 columns = 128 B), 128-byte swizzle
static inline int make_tmap_bf16(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  BCI_REQUIRE(enc, BCI_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BCI_REQUIRE(r == CUDA_SUCCESS, BCI_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return BCI_OK;
}

// 3D bf16 tensor [d2][d1][cols] (row-major), box = 1 x box_rows x box_cols, 128-byte swizzle; rows beyond d1 are
// clipped on store, so a partial last window tile cannot spill into the next time step.
static inline int make_tmap_bf16_3d(CUtensorMap* tm, const void* ptr, uint64_t d2, uint64_t d1, uint64_t cols, uint32_t box_cols,
                             uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  BCI_REQUIRE(enc, BCI_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {cols, d1, d2};
  cuuint64_t strides[2] = {cols * 2, d1 * cols * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BCI_REQUIRE(r == CUDA_SUCCESS, BCI_ECUDA, "cuTensorMapEncodeTiled(3D) failed with CUresult %d", (int)r);
  return BCI_OK;
}


}  // namespace bci
