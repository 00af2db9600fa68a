// Tree-attention decode for sm_100a: one query token per (batch, head) against a KV cache that is sharded along the
// sequence across the ranks of one NVSwitch box — ONE persistent cooperative kernel per rank and step:
//
//   phase 1  split-KV flash decoding over this rank's shard.  Work units (batch*kv_head, group chunk, split) are handed
//            out by an atomic queue to a grid of co-resident CTAs.  Bandwidth bound, so the inner loop is built around
//            16-byte loads that are issued one tile AHEAD of their use (K of tile t+1 is in flight during the softmax and
//            P V of tile t, V of tile t+1 during the scores of tile t+1) and the g query heads that share a KV head are
//            processed together so K and V are read from HBM exactly once.  KV may be bf16 / fp16 or fp8-e4m3 with
//            per-head or per-block dequantisation scales (serve path).  The CTA that finishes the LAST split of a group
//            merges the splits and publishes (out, lse) in this rank's symmetric (peer-mapped) partial buffer.
//   phase 2  grid barrier (all partials of this rank are published) -> one st.release.sys per peer on its signal pad,
//            then every CTA waits until all peers have signalled this rank's pad (the pad is local memory).
//   phase 3  cross-rank merge with the max-rescale identity, one warp per (batch, head) row:
//              * P2P: every rank reads every peer's row straight over NVLink (W rows of d+2 floats), or
//              * NVLS: multimem.ld_reduce through the multicast mapping of the partial buffers — the NVSwitch returns
//                max_r(lse_r) and then sum_r(w_r out_r), sum_r(w_r) in two in-switch reductions.
//
// Replaces the reference's Triton launch padded to a 128-row tile plus three latency-bound NCCL all-reduces (MAX lse,
// SUM den, SUM num; tree_attn_decoding.py:60-102).  All counters are self-resetting and the cross-rank epoch lives in
// device memory, so the launch is CUDA-graph capturable; the host wrapper allocates nothing per call.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "tree_decode_common.cuh"

namespace rab {
namespace {

constexpr int TD_THREADS = 128;
constexpr int TD_TILE = 64;     // keys per inner tile
constexpr int TD_MAX_G = 4;     // query heads per kv head handled by one CTA (larger groups use more units)

// ---- raw 16-byte (bf16 / fp16) or 8-byte (fp8) vectors of 8 elements -> fp32 ---------------------------------------
template <int KV_KIND>  // 0 bf16, 1 fp16, 2 fp8 e4m3
struct KvVec;
template <>
struct KvVec<0> {
  using Raw = uint4;
  static constexpr int kBytes = 16;
  __device__ static void cvt(const Raw& v, float* out) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      out[2 * i] = __uint_as_float(w[i] << 16);
      out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <>
struct KvVec<1> {
  using Raw = uint4;
  static constexpr int kBytes = 16;
  __device__ static void cvt(const Raw& v, float* out) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      out[2 * i] = f.x;
      out[2 * i + 1] = f.y;
    }
  }
};
template <>
struct KvVec<2> {
  using Raw = uint2;
  static constexpr int kBytes = 8;
  __device__ static void cvt(const Raw& v, float* out) {
    const uint32_t w[2] = {v.x, v.y};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        // cvt.rn.f16x2.e4m3x2: two fp8 values per instruction, then one unpack per pair
        const __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)((w[i] >> (16 * j)) & 0xffffu), __NV_E4M3);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hr));
        out[4 * i + 2 * j] = f.x;
        out[4 * i + 2 * j + 1] = f.y;
      }
    }
  }
};


template <int D, int KV_KIND>
__global__ void __launch_bounds__(TD_THREADS)
tree_decode_kernel(const __grid_constant__ TreeDecodeParams p) {
  using Vec = KvVec<KV_KIND>;
  using Raw = typename Vec::Raw;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int g_total = p.heads / p.kv_heads;
  const int zchunks = (g_total + TD_MAX_G - 1) / TD_MAX_G;
  const int groups = p.batch * p.kv_heads * zchunks;       // units = groups x splits
  const int total_units = p.n > 0 ? groups * p.splits : 0;
  constexpr int row_stride = TdCall<D>::row_stride;
  uint32_t* const ctr = p.counters;                         // [0] queue head, [1] barrier count, [2] barrier gen, [3] epoch
  TdCall<D> cs;
  cs.init(p);
  float* const my_partial = cs.my_partial;

  __shared__ __align__(16) float s_s[2][TD_TILE][TD_MAX_G];  // scores, then probabilities (double buffered per tile)
  __shared__ float corr_s[2][TD_MAX_G];
  __shared__ float red_s[8][TD_MAX_G][D + 1];
  __shared__ int unit_s;
  __shared__ uint32_t last_s;

  constexpr int EPL = D / 8;                    // QK: 8 lanes per key, EPL elements per lane
  constexpr int KVEC = EPL / 8;                 // 16-byte (8-byte for fp8) vectors per lane and key: 2 (D=128) or 1
  constexpr int CHUNKS = D / 8;                 // PV: thread -> (key group, 8-column chunk)
  constexpr int KGROUPS = TD_THREADS / CHUNKS;  // 8 or 16
  constexpr int KPT = TD_TILE / KGROUPS;        // keys per thread per tile: 8 or 4
  const int sub = lane / 8, l8 = lane % 8;
  const int chunk = tid % CHUNKS, kgrp = tid / CHUNKS;
  const size_t eb = KV_KIND == 2 ? 1 : 2;

  // =========================================== phase 1: split-KV partials ============================================
  while (true) {
    if (tid == 0) unit_s = (int)atomicAdd(&ctr[0], 1u);
    __syncthreads();
    const int unit = unit_s;
    __syncthreads();
    if (unit >= total_units) break;
    const int split = unit % p.splits;
    const int grp = unit / p.splits;
    const int zc = grp % zchunks;
    const int bhk = grp / zchunks;
    const int b = bhk / p.kv_heads, kvh = bhk % p.kv_heads;
    const int g0 = zc * TD_MAX_G;
    const int g = min(TD_MAX_G, g_total - g0);
    const int per = ((p.n + p.splits - 1) / p.splits + TD_TILE - 1) / TD_TILE * TD_TILE;  // tile-aligned splits
    const int k0 = split * per, k1 = min(p.n, k0 + per);

    const float* ksb = p.k_scale ? p.k_scale + (size_t)bhk * p.n_scale_blocks : nullptr;
    const float* vsb = p.v_scale ? p.v_scale + (size_t)bhk * p.n_scale_blocks : nullptr;
    const uint8_t* kbase = reinterpret_cast<const uint8_t*>(p.k) + (size_t)bhk * p.n * D * eb;
    const uint8_t* vbase = reinterpret_cast<const uint8_t*>(p.v) + (size_t)bhk * p.n * D * eb;

    float2 qr[TD_MAX_G][EPL / 2];  // packed pairs: the dot products run on FFMA2
#pragma unroll
    for (int gi = 0; gi < TD_MAX_G; ++gi) {
#pragma unroll
      for (int e = 0; e < EPL / 2; ++e) {
        // query head j uses kv head j % kv_heads  ->  heads {kvh, kvh + hk, ...}
        const size_t qi = ((size_t)b * p.heads + (size_t)(g0 + gi) * p.kv_heads + kvh) * D + l8 * EPL + 2 * e;
        qr[gi][e] = gi < g ? make_float2(load_q(p.q, p.q_kind, qi) * p.scale_log2, load_q(p.q, p.q_kind, qi + 1) * p.scale_log2)
                           : make_float2(0.f, 0.f);
      }
    }
    float2 acc[TD_MAX_G][4];
#pragma unroll
    for (int gi = 0; gi < TD_MAX_G; ++gi)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[gi][e] = make_float2(0.f, 0.f);
    float m_run = -INFINITY, l_run = 0.f;  // warp gi keeps the running stats of head gi (identical in all lanes)

    Raw kraw[4][KVEC], vraw[KPT];
    auto load_k = [&](int t0) {
#pragma unroll
      for (int step = 0; step < 4; ++step) {
        const int key = min(t0 + warp * 16 + step * 4 + sub, k1 - 1);  // clamp: loads stay in bounds
        const uint8_t* row = kbase + ((size_t)key * D + l8 * EPL) * eb;
#pragma unroll
        for (int c = 0; c < KVEC; ++c) kraw[step][c] = *reinterpret_cast<const Raw*>(row + c * Vec::kBytes);
      }
    };
    auto load_v = [&](int t0) {
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        const int key = min(t0 + kgrp + i * KGROUPS, k1 - 1);
        vraw[i] = *reinterpret_cast<const Raw*>(vbase + ((size_t)key * D + chunk * 8) * eb);
      }
    };
    if (k0 < k1) {
      load_k(k0);
      load_v(k0);
    }
    uint32_t par = 0;
    for (int t0 = k0; t0 < k1; t0 += TD_TILE, par ^= 1u) {
      const float ks = ksb ? ksb[t0 / p.scale_block] : 1.f;
      const float vs = vsb ? vsb[t0 / p.scale_block] : 1.f;
      const bool more = t0 + TD_TILE < k1;
      // ---- scores: consumes kraw ---------------------------------------------------------------------------------
#pragma unroll
      for (int step = 0; step < 4; ++step) {
        float kf[EPL];
#pragma unroll
        for (int c = 0; c < KVEC; ++c) Vec::cvt(kraw[step][c], kf + 8 * c);
        const int kl = warp * 16 + step * 4 + sub;
        const bool live = (t0 + kl) < k1;
        float part[TD_MAX_G];
#pragma unroll
        for (int gi = 0; gi < TD_MAX_G; ++gi) {
          float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int e = 0; e < EPL / 2; ++e) a2 = ffma2(make_float2(kf[2 * e], kf[2 * e + 1]), qr[gi][e], a2);
          float a = a2.x + a2.y;
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          a += __shfl_xor_sync(0xffffffffu, a, 4);
          part[gi] = live ? a * ks : -INFINITY;
        }
        if (l8 == 0) *reinterpret_cast<float4*>(&s_s[par][kl][0]) = make_float4(part[0], part[1], part[2], part[3]);
      }
      if (more) load_k(t0 + TD_TILE);  // in flight during the softmax and P V of this tile
      __syncthreads();
      // ---- online softmax: warp gi owns head gi ----------------------------------------------------------------
      if (warp < g) {
        const int gi = warp;
        const float a = s_s[par][lane][gi], bb = s_s[par][lane + 32][gi];
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
        s_s[par][lane][gi] = pa * vs;  // V block scale rides on the probabilities (the denominator uses the unscaled p)
        s_s[par][lane + 32][gi] = pb * vs;
        m_run = m_new;
        l_run = l_prev * corr + sum;
        if (lane == 0) corr_s[par][gi] = corr;
      } else if (warp < TD_MAX_G) {
        // unused head slots must read as zero probability in the float4 broadcast below
        s_s[par][lane][warp] = 0.f;
        s_s[par][lane + 32][warp] = 0.f;
        if (lane == 0) corr_s[par][warp] = 0.f;
      }
      __syncthreads();
      // ---- P V: consumes vraw ---------------------------------------------------------------------------------------
#pragma unroll
      for (int gi = 0; gi < TD_MAX_G; ++gi) {
        const float c = corr_s[par][gi];
        const float2 c2 = make_float2(c, c);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[gi][e] = fmul2(acc[gi][e], c2);
      }
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        float vf[8];
        Vec::cvt(vraw[i], vf);
        const int kl = kgrp + i * KGROUPS;
        const float4 p4 = *reinterpret_cast<const float4*>(&s_s[par][kl][0]);  // 0 for keys beyond the shard
        const float pk[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
        for (int gi = 0; gi < TD_MAX_G; ++gi) {
          const float2 p2 = make_float2(pk[gi], pk[gi]);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[gi][e] = ffma2(p2, make_float2(vf[2 * e], vf[2 * e + 1]), acc[gi][e]);
        }
      }
      if (more) load_v(t0 + TD_TILE);  // in flight during the scores and the softmax of the next tile
      // no barrier here: the next tile writes the OTHER s_s / corr_s buffer; this one is rewritten two tiles later,
      // behind the two barriers of the next tile
    }

    // ---- reduce the key groups, write the split result ----------------------------------------------------------------
    __syncthreads();
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
    if (p.splits == 1) {
      // the unit IS the group: normalise and publish (out, lse2, valid) for its g heads directly
      if (warp < g && lane == 0) {
        corr_s[0][warp] = l_run > 0.f ? 1.f / l_run : 0.f;
        const int head = (g0 + warp) * p.kv_heads + kvh;
        float* row = my_partial + ((size_t)b * p.heads + head) * row_stride;
        row[D] = l_run > 0.f ? (m_run == -INFINITY ? 0.f : m_run) + log2f(l_run) : -INFINITY;
        row[D + 1] = l_run > 0.f ? 1.f : 0.f;
      }
      __syncthreads();
      for (int i = tid; i < g * D; i += TD_THREADS) {
        const int gi = i / D, c = i % D;
        float sacc = 0.f;
#pragma unroll
        for (int kg = 0; kg < 8; ++kg) sacc += red_s[kg][gi][c];
        const int head = (g0 + gi) * p.kv_heads + kvh;
        my_partial[((size_t)b * p.heads + head) * row_stride + c] = sacc * corr_s[0][gi];
      }
    } else {
      float* out = p.scratch + (((size_t)bhk * p.splits + split) * g_total + g0) * row_stride;
      for (int i = tid; i < g * D; i += TD_THREADS) {
        const int gi = i / D, c = i % D;
        float sacc = 0.f;
#pragma unroll
        for (int kg = 0; kg < 8; ++kg) sacc += red_s[kg][gi][c];
        out[gi * row_stride + c] = sacc;
      }
      if (warp < g && lane == 0) {
        out[warp * row_stride + D] = m_run;
        out[warp * row_stride + D + 1] = l_run;
      }
      // the CTA that completes the last split of the group merges the splits
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const uint32_t done = atomicAdd(&p.group_done[grp], 1u);
        last_s = (done == (uint32_t)p.splits - 1) ? 1u : 0u;
        if (last_s) p.group_done[grp] = 0;  // self-resetting
      }
      __syncthreads();
      if (last_s) {
        __threadfence();
        for (int gi = 0; gi < g; ++gi) {
          const float* base = p.scratch + ((size_t)bhk * p.splits * g_total + g0 + gi) * row_stride;
          const size_t stride = (size_t)g_total * row_stride;
          float m = -INFINITY;
          for (int s = 0; s < p.splits; ++s) m = fmaxf(m, __ldcg(&base[s * stride + D]));
          const float m_eff = m == -INFINITY ? 0.f : m;
          float l = 0.f;
          for (int s = 0; s < p.splits; ++s) {
            const float ms = __ldcg(&base[s * stride + D]);
            l += ms == -INFINITY ? 0.f : __ldcg(&base[s * stride + D + 1]) * fast_exp2(ms - m_eff);
          }
          const int head = (g0 + gi) * p.kv_heads + kvh;
          float* row = my_partial + ((size_t)b * p.heads + head) * row_stride;
          for (int c = tid; c < D; c += TD_THREADS) {
            float a = 0.f;
            for (int s = 0; s < p.splits; ++s) {
              const float ms = __ldcg(&base[s * stride + D]);
              if (ms != -INFINITY) a += __ldcg(&base[s * stride + c]) * fast_exp2(ms - m_eff);
            }
            row[c] = l > 0.f ? a / l : 0.f;
          }
          if (tid == 0) {
            row[D] = l > 0.f ? m_eff + log2f(l) : -INFINITY;
            row[D + 1] = l > 0.f ? 1.f : 0.f;
          }
        }
      }
    }
    __syncthreads();
  }
  td_cross_rank_merge<D>(p, cs, total_units);
}


}  // namespace

int tree_decode_max_ctas(int d, int kv_kind, int num_sms) {
  int per_sm = 0;
  const void* fn = nullptr;
#define RAB_TD_PICK(DD, KK) fn = (const void*)tree_decode_kernel<DD, KK>
  if (d == 128) {
    if (kv_kind == 0) RAB_TD_PICK(128, 0); else if (kv_kind == 1) RAB_TD_PICK(128, 1); else RAB_TD_PICK(128, 2);
  } else {
    if (kv_kind == 0) RAB_TD_PICK(64, 0); else if (kv_kind == 1) RAB_TD_PICK(64, 1); else RAB_TD_PICK(64, 2);
  }
#undef RAB_TD_PICK
  cuda_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, TD_THREADS, 0), "tree_decode occupancy");
  return per_sm * num_sms;
}

void launch_tree_decode(const TreeDecodeParams& p, int d, int grid, cudaStream_t stream) {
  const void* fn = nullptr;
#define RAB_TD_PICK(DD, KK) fn = (const void*)tree_decode_kernel<DD, KK>
  if (d == 128) {
    if (p.kv_kind == 0) RAB_TD_PICK(128, 0); else if (p.kv_kind == 1) RAB_TD_PICK(128, 1); else RAB_TD_PICK(128, 2);
  } else {
    if (p.kv_kind == 0) RAB_TD_PICK(64, 0); else if (p.kv_kind == 1) RAB_TD_PICK(64, 1); else RAB_TD_PICK(64, 2);
  }
#undef RAB_TD_PICK
  void* args[] = {(void*)&p};
  // cooperative: the grid barrier and the cross-rank waits need every CTA of the grid to be resident
  cuda_check(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(TD_THREADS), args, 0, stream), "tree_decode launch");
}

}  // namespace rab
