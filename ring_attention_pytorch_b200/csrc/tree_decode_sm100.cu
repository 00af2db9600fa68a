// Tree-attention decode for sm_100a: one query token per (batch, head) against a KV cache that is sharded
// along the sequence across the ranks of one NVSwitch box.
//
//   tree_decode_partial_kernel : split-KV flash-decoding over this rank's shard.  Bandwidth-bound, so it
//       is a CUDA-core kernel tuned for coalesced 16-byte loads: one CTA per (split, batch*kv_head); the g
//       query heads that share a KV head are processed together so K and V are read from HBM exactly once.
//       KV may be bf16 / fp16 or fp8-e4m3 with per-(batch*kv_head) dequantisation scales (serve path).
//   tree_decode_combine_kernel : merges the splits into this rank's (lse, out) partial and publishes it in
//       the rank's symmetric (peer-mapped) slot.
//   tree_decode_reduce_kernel  : after the device barrier, every rank reads all peers' partials straight
//       over NVLink (P2P loads) and applies the max-rescale identity once -- replacing the reference's
//       three latency-bound NCCL all-reduces (MAX lse, SUM den, SUM num; tree_attn_decoding.py:89-100).
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "kernels.h"
#include "ptx.cuh"

namespace rab {
namespace {

constexpr int TD_THREADS = 128;
constexpr int TD_TILE = 64;     // keys per inner tile
constexpr int TD_MAX_G = 4;     // query heads per kv head handled by one CTA (larger groups use blockIdx.z)

template <int KV_KIND>  // 0 bf16, 1 fp16, 2 fp8 e4m3
struct KvTraits;
template <>
struct KvTraits<0> {
  static constexpr int kElemsPer16B = 8;
  __device__ static void load8(const void* p, float* out) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      out[2 * i] = __uint_as_float(w[i] << 16);
      out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <>
struct KvTraits<1> {
  static constexpr int kElemsPer16B = 8;
  __device__ static void load8(const void* p, float* out) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      out[2 * i] = __low2float(h);
      out[2 * i + 1] = __high2float(h);
    }
  }
};
template <>
struct KvTraits<2> {
  __device__ static void load8(const void* p, float* out) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);  // 8 fp8 values
    const uint32_t w[2] = {v.x, v.y};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const __nv_fp8x2_storage_t pair = (__nv_fp8x2_storage_t)((w[i] >> (16 * j)) & 0xffffu);
        const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2(pair, __NV_E4M3);
        const __half2 h = *reinterpret_cast<const __half2*>(&hr);
        out[4 * i + 2 * j] = __low2float(h);
        out[4 * i + 2 * j + 1] = __high2float(h);
      }
    }
  }
};

template <int KV_KIND>
__device__ __forceinline__ size_t kv_elem_bytes() {
  return KV_KIND == 2 ? 1 : 2;
}

// q: [b, h, d] (fp32 staged by the host wrapper), k: [b*hk, n, d], v: [b*hk, n, dv]
// scratch: [b*hk][splits][g][dv + 2]  (acc..., m, l) in the log2 domain
//
// Per 64-key tile:   scores : 8 lanes per key, 4 keys per warp step, all 16-byte loads of the tile issued up
//                             front (memory-level parallelism), q held in registers, 3-step shuffle reduce
//                    softmax: warp gi owns head gi (online, log2 domain)
//                    P V    : thread = (key group of 8, 8-column chunk); the 8 V loads of a tile are issued
//                             before the FMAs; probabilities are broadcast from smem as one float4 per key
template <int D, int KV_KIND>
__global__ void __launch_bounds__(TD_THREADS)
tree_decode_partial_kernel(const float* __restrict__ q, const void* __restrict__ k, const void* __restrict__ v,
                           const float* __restrict__ k_scale, const float* __restrict__ v_scale,
                           float* __restrict__ scratch, int heads, int kv_heads, int n, int splits,
                           float scale_log2, int scale_block, int n_scale_blocks) {
  const int g_total = heads / kv_heads;
  const int g0 = blockIdx.z * TD_MAX_G;                 // first group member handled by this CTA
  const int g = min(TD_MAX_G, g_total - g0);
  const int bhk = blockIdx.y;
  const int b = bhk / kv_heads, kvh = bhk % kv_heads;
  const int split = blockIdx.x;
  const int per = ((n + splits - 1) / splits + TD_TILE - 1) / TD_TILE * TD_TILE;  // tile-aligned splits
  const int k0 = split * per, k1 = min(n, k0 + per);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

  __shared__ __align__(16) float s_s[TD_TILE][TD_MAX_G];  // scores, then probabilities: one float4 per key
  __shared__ float corr_s[TD_MAX_G];
  __shared__ float red_s[8][TD_MAX_G][D + 1];

  // Block-scaled KV (fp8 serve path): one fp32 scale per `scale_block` keys of every (batch, kv head), applied to
  // the scores (K) and folded into the probabilities (V).  scale_block is a multiple of the 64-key tile.
  const float* ksb = k_scale ? k_scale + (size_t)bhk * n_scale_blocks : nullptr;
  const float* vsb = v_scale ? v_scale + (size_t)bhk * n_scale_blocks : nullptr;

  // QK ownership: 8 lanes per key (each lane D/8 elements), 4 keys per warp step, 16 keys per warp per tile
  constexpr int EPL = D / 8;  // elements per lane: 16 (D=128) or 8 (D=64)
  const int sub = lane / 8, l8 = lane % 8;
  float2 qr[TD_MAX_G][EPL / 2];  // packed pairs: the dot products run on FFMA2
#pragma unroll
  for (int gi = 0; gi < TD_MAX_G; ++gi) {
#pragma unroll
    for (int e = 0; e < EPL / 2; ++e) {
      // query head j uses kv head j % kv_heads  ->  heads {kvh, kvh + hk, ...}
      const float* qp = q + ((size_t)b * heads + (g0 + gi) * kv_heads + kvh) * D + l8 * EPL + 2 * e;
      qr[gi][e] = gi < g ? make_float2(qp[0] * scale_log2, qp[1] * scale_log2) : make_float2(0.f, 0.f);
    }
  }

  // PV ownership: thread -> (key group kgrp, column chunk of 8 elements)
  constexpr int CHUNKS = D / 8;                 // 16 for D=128, 8 for D=64
  constexpr int KGROUPS = TD_THREADS / CHUNKS;  // 8 or 16
  constexpr int KPT = TD_TILE / KGROUPS;        // keys per thread per tile: 8 or 4
  const int chunk = tid % CHUNKS, kgrp = tid / CHUNKS;
  float2 acc[TD_MAX_G][4];
#pragma unroll
  for (int gi = 0; gi < TD_MAX_G; ++gi)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[gi][e] = make_float2(0.f, 0.f);
  float m_run = -INFINITY, l_run = 0.f;  // warp gi keeps the running stats of head gi (identical in all lanes)

  const size_t eb = kv_elem_bytes<KV_KIND>();
  const uint8_t* kbase = reinterpret_cast<const uint8_t*>(k) + (size_t)bhk * n * D * eb;
  const uint8_t* vbase = reinterpret_cast<const uint8_t*>(v) + (size_t)bhk * n * D * eb;

  for (int t0 = k0; t0 < k1; t0 += TD_TILE) {
    const float ks = ksb ? ksb[t0 / scale_block] : 1.f;
    const float vs = vsb ? vsb[t0 / scale_block] : 1.f;
    // ---- scores -------------------------------------------------------------------------------
    float kf[4][EPL];
#pragma unroll
    for (int step = 0; step < 4; ++step) {
      const int key = min(t0 + warp * 16 + step * 4 + sub, k1 - 1);  // clamp: loads stay in bounds
      const uint8_t* row = kbase + ((size_t)key * D + l8 * EPL) * eb;
#pragma unroll
      for (int c = 0; c < EPL; c += 8) KvTraits<KV_KIND>::load8(row + c * eb, &kf[step][c]);
    }
#pragma unroll
    for (int step = 0; step < 4; ++step) {
      const int kl = warp * 16 + step * 4 + sub;
      const bool live = (t0 + kl) < k1;
      float part[TD_MAX_G];
#pragma unroll
      for (int gi = 0; gi < TD_MAX_G; ++gi) {
        float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < EPL / 2; ++e)
          a2 = ffma2(make_float2(kf[step][2 * e], kf[step][2 * e + 1]), qr[gi][e], a2);
        float a = a2.x + a2.y;
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        part[gi] = live ? a * ks : -INFINITY;
      }
      if (l8 == 0) *reinterpret_cast<float4*>(&s_s[kl][0]) = make_float4(part[0], part[1], part[2], part[3]);
    }
    __syncthreads();
    // ---- online softmax: warp gi owns head gi -------------------------------------------------------
    if (warp < g) {
      const int gi = warp;
      const float a = s_s[lane][gi], bb = s_s[lane + 32][gi];
      float mx = fmaxf(a, bb);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_prev = m_run, l_prev = l_run;
      const float m_new = fmaxf(m_prev, mx);
      const float m_eff = m_new == -INFINITY ? 0.f : m_new;
      const float pa = fast_exp2(a - m_eff), pb = fast_exp2(bb - m_eff);
      float sum = pa + pb;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float corr = m_prev == -INFINITY ? 0.f : fast_exp2(m_prev - m_eff);
      s_s[lane][gi] = pa * vs;  // V block scale rides on the probabilities (the denominator uses the unscaled p)
      s_s[lane + 32][gi] = pb * vs;
      m_run = m_new;
      l_run = l_prev * corr + sum;
      if (lane == 0) corr_s[gi] = corr;
    } else if (warp < TD_MAX_G) {
      // unused head slots must read as zero probability in the float4 broadcast below
      s_s[lane][warp] = 0.f;
      s_s[lane + 32][warp] = 0.f;
      if (lane == 0) corr_s[warp] = 0.f;
    }
    __syncthreads();
    // ---- P V --------------------------------------------------------------------------------------
    float vf[KPT][8];
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      const int key = min(t0 + kgrp + i * KGROUPS, k1 - 1);
      KvTraits<KV_KIND>::load8(vbase + ((size_t)key * D + chunk * 8) * eb, vf[i]);
    }
#pragma unroll
    for (int gi = 0; gi < TD_MAX_G; ++gi) {
      const float c = corr_s[gi];
      const float2 c2 = make_float2(c, c);
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[gi][e] = fmul2(acc[gi][e], c2);
    }
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      const int kl = kgrp + i * KGROUPS;
      const float4 p4 = *reinterpret_cast<const float4*>(&s_s[kl][0]);  // 0 for keys beyond the shard
      const float pk[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int gi = 0; gi < TD_MAX_G; ++gi) {
        const float2 p2 = make_float2(pk[gi], pk[gi]);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          acc[gi][e] = ffma2(p2, make_float2(vf[i][2 * e], vf[i][2 * e + 1]), acc[gi][e]);
      }
    }
    __syncthreads();
  }

  // ---- reduce the key groups and write the split partial ------------------------------------------
  float* out = scratch + (((size_t)bhk * splits + split) * g_total + g0) * (D + 2);
  for (int gi = 0; gi < g; ++gi) {
    if (kgrp < 8) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        red_s[kgrp][gi][chunk * 8 + 2 * e] = acc[gi][e].x;
        red_s[kgrp][gi][chunk * 8 + 2 * e + 1] = acc[gi][e].y;
      }
    }
  }
  __syncthreads();
  if (KGROUPS > 8) {  // D = 64: 16 key groups, fold the upper 8 onto the lower 8
    for (int gi = 0; gi < g; ++gi) {
      if (kgrp >= 8) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          atomicAdd(&red_s[kgrp - 8][gi][chunk * 8 + 2 * e], acc[gi][e].x);
          atomicAdd(&red_s[kgrp - 8][gi][chunk * 8 + 2 * e + 1], acc[gi][e].y);
        }
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < g * D; i += TD_THREADS) {
    const int gi = i / D, c = i % D;
    float sacc = 0.f;
#pragma unroll
    for (int kg = 0; kg < 8; ++kg) sacc += red_s[kg][gi][c];
    out[gi * (D + 2) + c] = sacc;
  }
  if (warp < g && lane == 0) {
    out[warp * (D + 2) + D] = m_run;
    out[warp * (D + 2) + D + 1] = l_run;
  }
}

// scratch [b*hk][splits][g][D+2] -> partial [b*h][D+2] = (normalised out[D], lse2, valid)
template <int D>
__global__ void tree_decode_combine_kernel(const float* __restrict__ scratch, float* __restrict__ partial, int heads,
                                           int kv_heads, int splits, int n) {
  const int g = heads / kv_heads;
  const int bh = blockIdx.x;  // b * heads + head
  const int b = bh / heads, head = bh % heads;
  const int kvh = head % kv_heads, gi = head / kv_heads;
  const int bhk = b * kv_heads + kvh;
  const float* base = scratch + (size_t)bhk * splits * g * (D + 2) + gi * (D + 2);
  const size_t stride = (size_t)g * (D + 2);
  float m = -INFINITY;
  if (n > 0)
    for (int s = 0; s < splits; ++s) m = fmaxf(m, base[s * stride + D]);
  const float m_eff = m == -INFINITY ? 0.f : m;
  float l = 0.f;
  if (n > 0)
    for (int s = 0; s < splits; ++s) {
      const float ms = base[s * stride + D];
      l += ms == -INFINITY ? 0.f : base[s * stride + D + 1] * fast_exp2(ms - m_eff);
    }
  const int c = threadIdx.x;
  if (c < D) {
    float acc = 0.f;
    if (n > 0)
      for (int s = 0; s < splits; ++s) {
        const float ms = base[s * stride + D];
        if (ms != -INFINITY) acc += base[s * stride + c] * fast_exp2(ms - m_eff);
      }
    partial[(size_t)bh * (D + 2) + c] = l > 0.f ? acc / l : 0.f;
  }
  if (c == 0) {
    partial[(size_t)bh * (D + 2) + D] = l > 0.f ? m_eff + log2f(l) : -INFINITY;  // lse in log2 units
    partial[(size_t)bh * (D + 2) + D + 1] = l > 0.f ? 1.f : 0.f;
  }
}

// Every rank reads every peer's [b*h][D+2] partial over NVLink and merges them:
//   out = sum_r out_r * 2^(lse_r - M) / sum_r 2^(lse_r - M),  M = max_r lse_r
template <int D>
__global__ void tree_decode_reduce_kernel(const __grid_constant__ TreeReduceParams p) {
  const int bh = blockIdx.x;
  const int c = threadIdx.x;
  float mx = -INFINITY;
  for (int r = 0; r < p.world; ++r) mx = fmaxf(mx, p.partials[r][(size_t)bh * (D + 2) + D]);
  const float m_eff = mx == -INFINITY ? 0.f : mx;
  float den = 0.f, num = 0.f;
  for (int r = 0; r < p.world; ++r) {
    const float* row = p.partials[r] + (size_t)bh * (D + 2);
    const float lse = row[D];
    if (lse == -INFINITY) continue;
    const float wgt = fast_exp2(lse - m_eff);
    den += wgt;
    if (c < D) num += wgt * row[c];
  }
  if (c < D) {
    const float o = num / fmaxf(den, p.eps);
    if (p.out_is_bf16 == 1) {
      reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)bh * D + c] = __float2bfloat16(o);
    } else if (p.out_is_bf16 == 0) {
      reinterpret_cast<__half*>(p.out)[(size_t)bh * D + c] = __float2half(o);
    } else {
      reinterpret_cast<float*>(p.out)[(size_t)bh * D + c] = o;
    }
  }
}

}  // namespace

void launch_tree_decode_partial(const float* q, const void* k, const void* v, const float* k_scale,
                                const float* v_scale, float* scratch, float* partial, int batch, int heads,
                                int kv_heads, int n, int d, int splits, int kv_kind, float scale, int scale_block,
                                int n_scale_blocks, cudaStream_t stream) {
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid(splits, batch * kv_heads, (heads / kv_heads + TD_MAX_G - 1) / TD_MAX_G);
  if (n > 0) {
#define RAB_TD_LAUNCH(DD, KK)                                                                                   \
  tree_decode_partial_kernel<DD, KK><<<grid, TD_THREADS, 0, stream>>>(                                          \
      q, k, v, k_scale, v_scale, scratch, heads, kv_heads, n, splits, scale_log2, scale_block, n_scale_blocks)
    if (d == 128) {
      if (kv_kind == 0) RAB_TD_LAUNCH(128, 0);
      else if (kv_kind == 1) RAB_TD_LAUNCH(128, 1);
      else RAB_TD_LAUNCH(128, 2);
    } else {
      if (kv_kind == 0) RAB_TD_LAUNCH(64, 0);
      else if (kv_kind == 1) RAB_TD_LAUNCH(64, 1);
      else RAB_TD_LAUNCH(64, 2);
    }
#undef RAB_TD_LAUNCH
    cuda_check(cudaGetLastError(), "tree_decode_partial launch");
  }
  if (d == 128) {
    tree_decode_combine_kernel<128><<<batch * heads, 128, 0, stream>>>(scratch, partial, heads, kv_heads, splits, n);
  } else {
    tree_decode_combine_kernel<64><<<batch * heads, 64, 0, stream>>>(scratch, partial, heads, kv_heads, splits, n);
  }
  cuda_check(cudaGetLastError(), "tree_decode_combine launch");
}

void launch_tree_decode_reduce(const TreeReduceParams& p, int batch_heads, int d, cudaStream_t stream) {
  if (d == 128) {
    tree_decode_reduce_kernel<128><<<batch_heads, 128, 0, stream>>>(p);
  } else {
    tree_decode_reduce_kernel<64><<<batch_heads, 64, 0, stream>>>(p);
  }
  cuda_check(cudaGetLastError(), "tree_decode_reduce launch");
}

}  // namespace rab
