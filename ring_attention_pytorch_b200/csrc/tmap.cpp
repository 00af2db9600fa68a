#include "tmap.h"

#include <cudaTypedefs.h>

#include <mutex>

namespace rab {

namespace {
PFN_cuTensorMapEncodeTiled_v12000 resolve_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      throw std::runtime_error("[ring_attention_b200] cannot resolve cuTensorMapEncodeTiled (no CUDA driver?)");
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}
}  // namespace

namespace {
CUtensorMap make_tmap_typed(CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, TmapSwizzle swizzle) {
  auto encode = resolve_encode();
  CUtensorMap map;
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstrides[i] = strides_bytes[i];
  }
  CUresult r = encode(&map, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdims,
                      gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      swizzle == TmapSwizzle::B128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    throw std::runtime_error("[ring_attention_b200] cuTensorMapEncodeTiled failed with code " +
                             std::to_string((int)r));
  }
  return map;
}
}  // namespace

CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, TmapSwizzle swizzle) {
  return make_tmap_typed(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle);
}

CUtensorMap make_tmap_u8(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, TmapSwizzle swizzle) {
  return make_tmap_typed(CU_TENSOR_MAP_DATA_TYPE_UINT8, base, rank, dims, strides_bytes, box, swizzle);
}

CUtensorMap make_tmap_f16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box, TmapSwizzle swizzle) {
  return make_tmap_typed(CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rank, dims, strides_bytes, box, swizzle);
}

CUtensorMap make_tmap_f32(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box, TmapSwizzle swizzle) {
  return make_tmap_typed(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle);
}

}  // namespace rab
