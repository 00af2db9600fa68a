// Shared device-side helpers for the attention kernels: position maps, tile classification and the
// deterministic per-work-item KV tile schedule that every warp role replays independently.
#pragma once
#include "kernels.h"
#include "ptx.cuh"

namespace rab {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// position of local index i on ring rank r
__device__ __forceinline__ int pos_of(const PosMap& pm, int r, int i) {
  return i < pm.seg_len ? pm.base0[r] + pm.stride * i : pm.base1[r] + pm.stride * (i - pm.seg_len);
}

// [lo, hi] position range covered by local indices a..b (inclusive) on ring rank r
__device__ __forceinline__ void pos_range(const PosMap& pm, int r, int a, int b, int& lo, int& hi) {
  if (b < pm.seg_len) {
    lo = pm.base0[r] + pm.stride * a;
    hi = pm.base0[r] + pm.stride * b;
  } else if (a >= pm.seg_len) {
    lo = pm.base1[r] + pm.stride * (a - pm.seg_len);
    hi = pm.base1[r] + pm.stride * (b - pm.seg_len);
  } else {
    const int lo0 = pm.base0[r] + pm.stride * a, hi0 = pm.base0[r] + pm.stride * (pm.seg_len - 1);
    const int lo1 = pm.base1[r], hi1 = pm.base1[r] + pm.stride * (b - pm.seg_len);
    lo = min(lo0, lo1);
    hi = max(hi0, hi1);
  }
}

// Mask parameters shared by forward and backward kernels.
struct MaskCfg {
  int causal;
  int window;
  int has_kmask;
};

// need: at least one (q, k) pair of the tile pair may be visible.  partial: per-element masking required.
__device__ __forceinline__ void classify_tile(const MaskCfg& mc, int qlo, int qhi, int klo, int khi, bool k_tail,
                                              bool& need, bool& partial) {
  need = true;
  partial = k_tail || mc.has_kmask;
  if (mc.causal) {
    if (klo > qhi) {
      need = false;
    } else if (khi > qlo) {
      partial = true;
    }
    if (mc.window > 0) {
      if (qlo - khi > mc.window) {
        need = false;
      } else if (qhi - klo > mc.window) {
        partial = true;
      }
    }
  }
}

}  // namespace rab
