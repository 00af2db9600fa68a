// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled resolved at run time through the
// CUDA runtime so the extension links without libcuda and builds on a machine with no GPU).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace rab {

enum class TmapSwizzle { None = 0, B128 = 3 };

// dims/strides are innermost-first.  strides_bytes[i] is the byte stride of dimension i+1
// (dimension 0 is contiguous), i.e. rank-1 entries.
CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, TmapSwizzle swizzle);

// same for fp32 elements (accumulators that are the target of cp.reduce.async.bulk.tensor .add)
CUtensorMap make_tmap_f32(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box, TmapSwizzle swizzle);

// same for 8-bit elements (fp8 KV cache tiles consumed by kind::f8f6f4 MMAs)
CUtensorMap make_tmap_u8(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, TmapSwizzle swizzle);
CUtensorMap make_tmap_f16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box, TmapSwizzle swizzle);

inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    throw std::runtime_error(std::string("[ring_attention_b200] ") + what + ": " + cudaGetErrorString(e));
  }
}

}  // namespace rab
