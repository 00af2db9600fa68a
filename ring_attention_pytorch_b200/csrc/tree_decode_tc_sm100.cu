// Tree-attention decode on the 5th-generation tensor cores (head dim 128, up to 16 query heads per KV head).
//
// A decode step is bandwidth bound, but with grouped-query heads the CUDA-core split-KV kernel (tree_decode_sm100.cu)
// spends g FMAs per loaded element and stalls on its own dependency chains long before HBM is saturated (0.68 of the
// measured copy rate for bf16, no gain at all from an fp8 cache).  Here the two products run as tcgen05.mma in the
// TRANSPOSED form, so that the 128 TMEM lanes are keys / head-dim entries and the (few) query heads are MMA columns:
//
//      S^T [128 keys x 16 heads] = K_tile [128 x d] . Q^T [d x 16]        A = K tile (TMA, K-major), B = Q^T (smem)
//      O^T [128 d    x 16 heads] += V_tile^T [d x 128] . P [128 x 16]     A = V tile read MN-major, B = P (smem, K-major)
//
// K / V tiles are streamed by TMA through a multi-stage shared-memory ring and consumed by the tensor core directly:
// bf16 / fp16 caches with kind::f16, fp8-e4m3 caches with kind::f8f6f4 (Q and P are quantised to e4m3 on the fly, P
// scaled by 2^4 so that small probabilities stay above the e4m3 subnormals) — no dequantisation pass, half the bytes, half
// the time.  The only per-element CUDA-core work left is the online softmax: one thread per key, g values each.
//
//   warp 0      TMA producer (K and V tiles; a tile is skipped entirely when it lies beyond the shard)
//   warp 1      MMA issuer + TMEM allocator: S(t+1) is issued before P V(t), so the softmax of tile t overlaps S(t+1)
//   warps 2-5   softmax warpgroup (thread = key row = TMEM lane): lazy running maximum per head (raised only when a
//               tile exceeds it by more than 2^8, voted with one barrier.red), P written to shared memory as the B
//               operand, per-thread partial row sums; epilogue: O^T / l -> this rank's partial rows (or split scratch)
//
// Phases 2 and 3 of the step (publish -> cross-rank signal -> merge over NVLink loads or NVLS multimem) are shared with
// the CUDA-core kernel (tree_decode_common.cuh); the whole step is still ONE cooperative launch.
// Reference: tree_attn_decoding.py:60-102.
#include <cuda_fp8.h>

#include "tree_decode_common.cuh"

namespace rab {
namespace {

constexpr int TC_THREADS = 192;
constexpr int TC_TILE = 128;   // keys per tile
constexpr int TC_NH = 16;      // MMA N: query heads per unit (zero padded)
constexpr int TC_D = 128;

template <bool KV8>
struct TcSmem {
  static constexpr int NST = KV8 ? 6 : 3;
  static constexpr int TILE_BYTES = TC_TILE * TC_D * (KV8 ? 1 : 2);
  static constexpr int QP_BYTES = TC_NH * TC_D * (KV8 ? 1 : 2);  // Q^T tile; P tile has the same size (128 keys)
  alignas(1024) uint8_t k[NST][TILE_BYTES];
  alignas(1024) uint8_t v[NST][TILE_BYTES];
  alignas(1024) uint8_t q[QP_BYTES];
  alignas(1024) uint8_t p[2][QP_BYTES];
  uint64_t k_full[NST], k_empty[NST], v_full[NST], v_empty[NST];
  uint64_t s_full[2], s_free[2], p_ready[2], p_free[2];
  uint64_t o_done;
  float red[4][TC_NH];   // cross-warp reductions (tile maxima, row sums)
  float qscale[TC_NH];   // fp8: per-head dequantisation scale of the quantised query
  int unit, unit_next;
  uint32_t last;
  uint32_t tmem_base;
};

// kind::f16 instruction descriptor lives in ptx.cuh (umma_idesc_bf16); kind::f8f6f4 with e4m3 operands, fp32 accumulate:
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (0 = e4m3)  [10,13) B fmt (0 = e4m3)  [15] A major  [16] B major  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_e4m3(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_ss_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// barrier.red.or over the 128 threads of the softmax warpgroup (named barrier 1)
__device__ __forceinline__ bool wg_vote_any(bool pred) {
  uint32_t r;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %1, 0;\n"
      "barrier.red.or.pred q, 1, 128, p;\n"
      "selp.u32 %0, 1, 0, q;\n"
      "}\n"
      : "=r"(r)
      : "r"((uint32_t)pred)
      : "memory");
  return r != 0;
}
__device__ __forceinline__ void wg_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// GM: compile-time bound on the query heads of a unit (4 or 16).  The MMA always has N = 16 columns, but the softmax
// warpgroup runs ONE warp per sub-partition, so every instruction of its per-tile loop is exposed latency: with the
// usual group sizes (<= 4) the loops, the TMEM loads and the P stores are a quarter of the 16-wide version.
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* r) {
  if constexpr (N == 4) tmem_ld4(taddr, r);
  else tmem_ld16(taddr, r);
}
template <int N>
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t* r) {
  if constexpr (N == 4) tmem_st4(taddr, r);
  else tmem_st16(taddr, r);
}

template <bool KV8, int Q16 /* 0: bf16 cache, 1: fp16 cache (ignored for fp8) */, int GM>
__global__ void __launch_bounds__(TC_THREADS, 1)
tree_decode_tc_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                      const __grid_constant__ TreeDecodeParams p) {
  constexpr int D = TC_D;
  using Smem = TcSmem<KV8>;
  constexpr int NST = Smem::NST;
  constexpr int EB = KV8 ? 1 : 2;
  constexpr int NSUB = KV8 ? 1 : 2;                    // 128-byte wide sub-tiles per K/V tile row
  constexpr int SUBK = TC_TILE * 128;                  // bytes of one K/V sub-tile (128 rows x 128 B)
  constexpr int SUBQ = TC_NH * 128;                    // bytes of one Q^T / P sub-tile (16 rows x 128 B)
  constexpr int KSTEP = KV8 ? 32 : 16;                 // MMA K per instruction
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int g_total = p.heads / p.kv_heads;
  const int zchunks = (g_total + GM - 1) / GM;
  const int groups = p.batch * p.kv_heads * zchunks;
  const int total_units = p.n > 0 ? groups * p.splits : 0;
  constexpr int row_stride = TdCall<D>::row_stride;
  uint32_t* const ctr = p.counters;
  TdCall<D> cs;
  cs.init(p);
  float* const my_partial = cs.my_partial;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&sm.k_full[i], 1);
      mbar_init(&sm.k_empty[i], 1);
      mbar_init(&sm.v_full[i], 1);
      mbar_init(&sm.v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.s_full[i], 1);
      mbar_init(&sm.s_free[i], 128);
      mbar_init(&sm.p_ready[i], 128);
      mbar_init(&sm.p_free[i], 1);
    }
    mbar_init(&sm.o_done, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
  }
  if (warp == 1) {
    tmem_alloc(&sm.tmem_base, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t s_tm[2] = {tmem + 0, tmem + 16};
  const uint32_t o_tm = tmem + 32;

  // cumulative pipeline counters (continue across units)
  uint32_t n_tile = 0;   // tiles of this CTA so far: stage = n_tile % NST, S / P buffer = n_tile & 1
  uint32_t n_unit = 0;

  // The work queue is read one unit AHEAD: at the end of a unit the TMA producer already streams the first tiles of the
  // next one, so the shared-memory ring does not drain while the warpgroup runs the epilogue and builds the next Q^T.
  uint32_t n_prod = 0;    // tiles issued by the producer so far (runs ahead of n_tile by the prefetched tiles)
  int pre_issued = 0;     // tiles of the CURRENT unit that were issued during the previous unit's tail
  bool first = true;
  while (true) {
    __syncthreads();
    if (tid == 0) {
      sm.unit = first ? (int)atomicAdd(&ctr[0], 1u) : sm.unit_next;
      sm.unit_next = (int)atomicAdd(&ctr[0], 1u);
    }
    first = false;
    __syncthreads();
    const int unit = sm.unit;
    const int unit_next = sm.unit_next;
    if (unit >= total_units) break;
    const int split = unit % p.splits;
    const int grp = unit / p.splits;
    const int zc = grp % zchunks;
    const int bhk = grp / zchunks;
    const int b = bhk / p.kv_heads, kvh = bhk % p.kv_heads;
    const int g0 = zc * GM;
    const int g = min(GM, g_total - g0);
    const int per = ((p.n + p.splits - 1) / p.splits + TC_TILE - 1) / TC_TILE * TC_TILE;
    const int k0 = split * per, k1 = min(p.n, k0 + per);
    const int ntiles = k1 > k0 ? (k1 - k0 + TC_TILE - 1) / TC_TILE : 0;

    if (warp >= 2) {
      // ---- Q^T tile: 16 rows (heads, zero padded) x 128 d, K-major with the 128B swizzle applied by hand -------------
      const int wt = tid - 64;  // 0..127
      {
        const int row = wt / 8, seg = wt % 8;  // 16 elements per thread
        float qv[16];
        float amax = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const size_t qi = ((size_t)b * p.heads + (size_t)(g0 + row) * p.kv_heads + kvh) * D + seg * 16 + e;
          qv[e] = row < g ? load_q(p.q, p.q_kind, qi) * p.scale_log2 : 0.f;
          amax = fmaxf(amax, fabsf(qv[e]));
        }
        if constexpr (KV8) {
          // per-head e4m3 scale: the 8 threads of a row agree on the row maximum
          amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
          amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
          amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
          const float sc = amax > 0.f ? amax / 448.f : 1.f;
          const float inv = 1.f / sc;
          if (seg == 0) sm.qscale[row] = sc;
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(qv[4 * e] * inv, qv[4 * e + 1] * inv), __NV_SATFINITE, __NV_E4M3);
            const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(qv[4 * e + 2] * inv, qv[4 * e + 3] * inv), __NV_SATFINITE, __NV_E4M3);
            w[e] = lo | (hi << 16);
          }
          // one 16-byte chunk per thread: chunk index seg of the 128-byte row
          *reinterpret_cast<uint4*>(sm.q + row * 128 + ((seg ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = Q16 == 0 ? pack_bf16x2(qv[2 * e], qv[2 * e + 1]) : pack_f16x2(qv[2 * e], qv[2 * e + 1]);
          const int sub = seg / 4, c0 = (seg % 4) * 2;  // 64 d per sub-tile = 8 chunks; this thread owns two
          uint8_t* rowp = sm.q + sub * SUBQ + row * 128;
          *reinterpret_cast<uint4*>(rowp + (((c0 + 0) ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(rowp + (((c0 + 1) ^ (row & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
        }
        // P tiles: rows >= g stay zero for the whole unit (the softmax only writes rows < g)
        if (n_unit == 0) {
          for (int i = wt; i < 2 * Smem::QP_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sm.p[0])[i] = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async_shared();
      }
    }
    __syncthreads();  // Q^T is in shared memory (the issuer may start)

    if (warp == 0) {
      // =============================================== TMA producer ===============================================
      if (lane == 0) {
        auto issue_tile = [&](int bhk_, int key0) {
          const uint32_t st = n_prod % NST, ph = (n_prod / NST) & 1;
          mbar_wait(&sm.k_empty[st], ph ^ 1, 2000 + st);
          mbar_expect_tx(&sm.k_full[st], Smem::TILE_BYTES);
#pragma unroll
          for (int s2 = 0; s2 < NSUB; ++s2)
            tma_load_3d(sm.k[st] + s2 * SUBK, &map_k, &sm.k_full[st], s2 * (KV8 ? 128 : 64), key0, bhk_);
          mbar_wait(&sm.v_empty[st], ph ^ 1, 2010 + st);
          mbar_expect_tx(&sm.v_full[st], Smem::TILE_BYTES);
#pragma unroll
          for (int s2 = 0; s2 < NSUB; ++s2)
            tma_load_3d(sm.v[st] + s2 * SUBK, &map_v, &sm.v_full[st], s2 * (KV8 ? 128 : 64), key0, bhk_);
          n_prod++;
        };
        for (int t = pre_issued; t < ntiles; ++t) issue_tile(bhk, k0 + t * TC_TILE);
        // run ahead into the next unit: at most NST tiles, whose slots free up as THIS unit's last tiles are consumed
        pre_issued = 0;
        if (unit_next < total_units) {
          const int split_n = unit_next % p.splits;
          const int bhk_n = (unit_next / p.splits) / zchunks;
          const int k0_n = split_n * per, k1_n = min(p.n, k0_n + per);
          const int nt_n = k1_n > k0_n ? (k1_n - k0_n + TC_TILE - 1) / TC_TILE : 0;
          const int ahead = min(nt_n, NST);
          for (int t = 0; t < ahead; ++t) issue_tile(bhk_n, k0_n + t * TC_TILE);
          pre_issued = ahead;
        }
      }
    } else if (warp == 1) {
      // ================================================ MMA issuer ================================================
      constexpr uint32_t idesc_s = KV8 ? umma_idesc_e4m3(128, TC_NH, 0, 0) : umma_idesc_bf16(128, TC_NH, 0, 0, Q16 == 0 ? 1 : 0);
      constexpr uint32_t idesc_o = KV8 ? umma_idesc_e4m3(128, TC_NH, 1, 0) : umma_idesc_bf16(128, TC_NH, 1, 0, Q16 == 0 ? 1 : 0);
      constexpr uint64_t kmaj = umma_smem_desc_hi_lo(16, 1024, UMMA_LAYOUT_SW128);
      constexpr uint64_t mnmaj = umma_smem_desc_hi_lo(SUBK, 1024, UMMA_LAYOUT_SW128);
      const uint64_t q_desc = umma_desc(kmaj, smem_u32(sm.q));
      auto issue_s = [&](int t) {
        const uint32_t n = n_tile + t, st = n % NST, ph = (n / NST) & 1, sb = n & 1;
        mbar_wait(&sm.k_full[st], ph, 2100 + st);
        if (n >= 2) mbar_wait(&sm.s_free[sb], ((n >> 1) - 1) & 1, 2110 + sb);  // the softmax has pulled S(n-2) into registers
        tc_fence_after();
        const uint64_t kd = umma_desc(kmaj, smem_u32(sm.k[st]));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < D / KSTEP; ++kk) {
            // 128-byte swizzled rows: 64 (bf16) / 128 (fp8) elements per sub-tile; K steps advance 32 bytes inside one
            const uint32_t per_sub = 128 / (KSTEP * EB);  // MMAs per sub-tile: 4
            const uint32_t offk = (kk / per_sub) * SUBK + (kk % per_sub) * 32;
            const uint32_t offq = (kk / per_sub) * SUBQ + (kk % per_sub) * 32;
            if constexpr (KV8) umma_ss_f8(s_tm[sb], umma_desc_add(kd, offk), umma_desc_add(q_desc, offq), idesc_s, kk > 0);
            else umma_ss(s_tm[sb], umma_desc_add(kd, offk), umma_desc_add(q_desc, offq), idesc_s, kk > 0);
          }
          umma_commit(&sm.s_full[sb]);
          umma_commit(&sm.k_empty[st]);
        }
        __syncwarp();
      };
      if (ntiles > 0) issue_s(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) issue_s(t + 1);
        const uint32_t n = n_tile + t, st = n % NST, ph = (n / NST) & 1, pb = n & 1;
        mbar_wait(&sm.v_full[st], ph, 2120 + st);
        mbar_wait(&sm.p_ready[pb], (n >> 1) & 1, 2130 + pb);
        tc_fence_after();
        const uint64_t vd = umma_desc(mnmaj, smem_u32(sm.v[st]));
        const uint64_t pd = umma_desc(kmaj, smem_u32(sm.p[pb]));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < TC_TILE / KSTEP; ++kk) {
            // A = V^T read MN-major: KSTEP key rows of 128 bytes per instruction; B = P [16 heads x 128 keys] K-major
            const uint32_t offv = kk * KSTEP * 128;
            const uint32_t per_sub = 128 / (KSTEP * EB);
            const uint32_t offp = (kk / per_sub) * SUBQ + (kk % per_sub) * 32;
            if constexpr (KV8) umma_ss_f8(o_tm, umma_desc_add(vd, offv), umma_desc_add(pd, offp), idesc_o, (t > 0 || kk > 0) ? 1u : 0u);
            else umma_ss(o_tm, umma_desc_add(vd, offv), umma_desc_add(pd, offp), idesc_o, (t > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&sm.v_empty[st]);
          umma_commit(&sm.p_free[pb]);
          if (t == ntiles - 1) umma_commit(&sm.o_done);
        }
        __syncwarp();
      }
    } else {
      // ============================================ softmax warpgroup =============================================
      const int wt = tid - 64;                                   // key row inside the tile
      const uint32_t lane_off = uint32_t((warp % 4) * 32) << 16;  // TMEM lane quadrant of this warp
      const int r = (warp % 4) * 32 + lane;                      // TMEM lane == key row == d index (epilogue)
      (void)wt;
      float m_run[GM], l_part[GM];
#pragma unroll
      for (int h = 0; h < GM; ++h) {
        m_run[h] = -INFINITY;
        l_part[h] = 0.f;
      }
      const float* ksb = p.k_scale ? p.k_scale + (size_t)bhk * p.n_scale_blocks : nullptr;
      const float* vsb = p.v_scale ? p.v_scale + (size_t)bhk * p.n_scale_blocks : nullptr;
      // Lazy maximum: probabilities may exceed 1 by up to 2^RAISE_SLACK before the (expensive) rescale path runs.  An fp8
      // P must stay below the e4m3 maximum (448) and its small values above the subnormals (2^-9): slack 4, scale 2^4.
      constexpr float RAISE_SLACK = KV8 ? 4.f : 8.f;
      constexpr float P_SCALE = KV8 ? 16.f : 1.f;
      // V block scales ride on the probabilities RELATIVE to the largest scale of the unit (so an fp8 P never underflows
      // because of a small dequantisation scale); the reference scale itself is applied once in the epilogue.
      float vs_ref = 1.f;
      if (vsb != nullptr && ntiles > 0) {
        vs_ref = 0.f;
        for (int blk = k0 / p.scale_block; blk <= (k1 - 1) / p.scale_block; ++blk) vs_ref = fmaxf(vs_ref, vsb[blk]);
        if (!(vs_ref > 0.f)) vs_ref = 1.f;
      }
      const float inv_vs_ref = 1.f / vs_ref;
      for (int t = 0; t < ntiles; ++t) {
        const uint32_t n = n_tile + t, sb = n & 1, pb = n & 1;
        const int key0 = k0 + t * TC_TILE;
        const float ks = ksb ? ksb[key0 / p.scale_block] : 1.f;
        const float vs = vsb ? vsb[key0 / p.scale_block] * inv_vs_ref : 1.f;
        mbar_wait(&sm.s_full[sb], (n >> 1) & 1, 2200 + sb);
        tc_fence_after();
        uint32_t sr[GM];
        tmem_ld_n<GM>(s_tm[sb] + lane_off, sr);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(&sm.s_free[sb]);
        const bool live = key0 + r < k1;
        float sv[GM];
        bool raise = false;
#pragma unroll
        for (int h = 0; h < GM; ++h) {
          float x = __uint_as_float(sr[h]) * ks;
          if constexpr (KV8) x *= sm.qscale[h];
          sv[h] = (live && h < g) ? x : -INFINITY;
          raise = raise || (sv[h] > m_run[h] + RAISE_SLACK);
        }
        if (wg_vote_any(raise)) {
          // rare: raise the running maxima (all 128 threads agree on them), rescale l and the O^T accumulator
#pragma unroll
          for (int h = 0; h < GM; ++h) {
            float mx = sv[h];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) sm.red[warp % 4][h] = mx;
          }
          wg_sync();
          float f[GM];
#pragma unroll
          for (int h = 0; h < GM; ++h) {
            const float mx = fmaxf(fmaxf(sm.red[0][h], sm.red[1][h]), fmaxf(sm.red[2][h], sm.red[3][h]));
            const float m_new = fmaxf(m_run[h], mx);
            f[h] = (m_run[h] == -INFINITY) ? 0.f : fast_exp2(m_run[h] - m_new);
            l_part[h] *= f[h];
            m_run[h] = m_new;
          }
          if (t > 0) {
            // every P V issued so far must have landed before O^T is rescaled: the last one committed p_free of the
            // previous tile's buffer
            const uint32_t np = n - 1;
            mbar_wait(&sm.p_free[np & 1], (np >> 1) & 1, 2210);
            tc_fence_after();
            uint32_t orr[GM];
            tmem_ld_n<GM>(o_tm + lane_off, orr);
            tc_wait_ld();
#pragma unroll
            for (int h = 0; h < GM; ++h) orr[h] = __float_as_uint(__uint_as_float(orr[h]) * f[h]);
            tmem_st_n<GM>(o_tm + lane_off, orr);
            tc_wait_st();
            tc_fence_before();
          }
          wg_sync();  // sm.red may be reused; O^T is consistent before anyone hands out P of this tile
        }
        // P buffer of this tile: free once the P V that read it two tiles ago has completed
        if (n >= 2) mbar_wait(&sm.p_free[pb], ((n >> 1) - 1) & 1, 2220 + pb);
        uint8_t* pt = sm.p[pb];
#pragma unroll
        for (int h = 0; h < GM; ++h) {
          if (h < g) {
            const float pj = sv[h] == -INFINITY ? 0.f : fast_exp2(sv[h] - m_run[h]);
            l_part[h] += pj;
            const float pv = pj * vs * P_SCALE;
            if constexpr (KV8) {
              // row h, key r: one byte; 16-byte chunk (r / 16) of the 128-byte row, swizzled with the row index
              const uint32_t byte = __nv_cvt_float_to_fp8(pv, __NV_SATFINITE, __NV_E4M3);
              pt[h * 128 + ((((r >> 4) ^ (h & 7)) << 4) | (r & 15))] = (uint8_t)byte;
            } else {
              const int sub = r >> 6, rr = r & 63;  // 64 keys per 128-byte row
              uint16_t hv;
              if (Q16 == 0) hv = (uint16_t)(pack_bf16x2(pv, 0.f) & 0xffffu);
              else hv = (uint16_t)(pack_f16x2(pv, 0.f) & 0xffffu);
              *reinterpret_cast<uint16_t*>(pt + sub * SUBQ + h * 128 + ((((rr >> 3) ^ (h & 7)) << 4) | ((rr & 7) << 1))) = hv;
            }
          }
        }
        fence_proxy_async_shared();
        mbar_arrive(&sm.p_ready[pb]);
      }

      // ---- epilogue: O^T (lane = d) / l -> partial rows, or the split scratch -------------------------------------------
      // row sums: every thread holds the sum over ITS keys; reduce over the 128 threads
#pragma unroll
      for (int h = 0; h < GM; ++h) {
        float sum = l_part[h];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) sm.red[warp % 4][h] = sum;
      }
      wg_sync();
      float l_tot[GM];
#pragma unroll
      for (int h = 0; h < GM; ++h) l_tot[h] = (sm.red[0][h] + sm.red[1][h]) + (sm.red[2][h] + sm.red[3][h]);
      uint32_t orr[GM];
      if (ntiles > 0) {
        mbar_wait(&sm.o_done, n_unit & 1, 2230);
        tc_fence_after();
        tmem_ld_n<GM>(o_tm + lane_off, orr);
        tc_wait_ld();
        tc_fence_before();
      } else {
#pragma unroll
        for (int h = 0; h < GM; ++h) orr[h] = 0u;
      }
      if (p.splits == 1) {
#pragma unroll
        for (int h = 0; h < GM; ++h) {
          if (h < g) {
            const int head = (g0 + h) * p.kv_heads + kvh;
            float* row = my_partial + ((size_t)b * p.heads + head) * row_stride;
            const float inv = l_tot[h] > 0.f ? vs_ref / (l_tot[h] * P_SCALE) : 0.f;
            row[r] = __uint_as_float(orr[h]) * inv;
            if (r == 0) {
              row[D] = l_tot[h] > 0.f ? m_run[h] + log2f(l_tot[h]) : -INFINITY;
              row[D + 1] = l_tot[h] > 0.f ? 1.f : 0.f;
            }
          }
        }
      } else {
        float* out = p.scratch + (((size_t)bhk * p.splits + split) * g_total + g0) * row_stride;
#pragma unroll
        for (int h = 0; h < GM; ++h) {
          if (h < g) {
            out[h * row_stride + r] = __uint_as_float(orr[h]) * (vs_ref / P_SCALE);
            if (r == 0) {
              out[h * row_stride + D] = m_run[h];
              out[h * row_stride + D + 1] = l_tot[h];
            }
          }
        }
      }
    }
    n_tile += ntiles;
    if (ntiles > 0) n_unit++;  // o_done completes once per unit that ran at least one tile

    if (p.splits > 1) {
      // the CTA that completes the last split of the group merges the splits (same protocol as the CUDA-core kernel)
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const uint32_t done = atomicAdd(&p.group_done[grp], 1u);
        sm.last = (done == (uint32_t)p.splits - 1) ? 1u : 0u;
        if (sm.last) p.group_done[grp] = 0;
      }
      __syncthreads();
      if (sm.last) {
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
          for (int c = tid; c < D; c += TC_THREADS) {
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
  }
  if (total_units == 0) {  // this rank holds no keys: publish empty rows
    for (int i = blockIdx.x * TC_THREADS + tid; i < p.batch * p.heads; i += gridDim.x * TC_THREADS) {
      my_partial[(size_t)i * row_stride + D] = -INFINITY;
      my_partial[(size_t)i * row_stride + D + 1] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 64);
  td_cross_rank_merge<D>(p, cs, total_units);
}

template <bool KV8, int Q16>
const void* tc_kernel_ptr(bool small_group) {
  return small_group ? (const void*)tree_decode_tc_kernel<KV8, Q16, 4> : (const void*)tree_decode_tc_kernel<KV8, Q16, 16>;
}
const void* pick_tc(int kv_kind, bool small_group) {
  if (kv_kind == 2) return tc_kernel_ptr<true, 0>(small_group);
  return kv_kind == 0 ? tc_kernel_ptr<false, 0>(small_group) : tc_kernel_ptr<false, 1>(small_group);
}
size_t tc_smem(int kv_kind) { return (kv_kind == 2 ? sizeof(TcSmem<true>) : sizeof(TcSmem<false>)) + 1024; }

}  // namespace

int tree_decode_tc_max_ctas(int kv_kind, int num_sms) {
  const void* fn = pick_tc(kv_kind, false);
  const size_t smem = tc_smem(kv_kind);
  cuda_check(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "tree_decode_tc smem attr");
  int per_sm = 0;
  cuda_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, TC_THREADS, smem), "tree_decode_tc occupancy");
  return per_sm * num_sms;
}

void launch_tree_decode_tc(const CUtensorMap& map_k, const CUtensorMap& map_v, const TreeDecodeParams& p, int grid,
                           cudaStream_t stream) {
  const void* fn = pick_tc(p.kv_kind, p.heads / p.kv_heads <= 4);
  const size_t smem = tc_smem(p.kv_kind);
  cuda_check(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "tree_decode_tc smem attr");
  void* args[] = {(void*)&map_k, (void*)&map_v, (void*)&p};
  cuda_check(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(TC_THREADS), args, smem, stream), "tree_decode_tc launch");
}

}  // namespace rab
