// One-kernel (5-GEMM) ring flash-attention backward for sm_100a, head dim 128.
//
// KV-stationary, hop-ordered: a work item is one 128-key tile of one (batch, kv head) of ONE ring owner (hop 0 = this
// rank's own keys, later hops = the other owners' K/V slots in the gather buffer).  The LOCAL queries that can see the
// tile stream through in 64-row steps; per step
//
//      S^T  = K Q^T            (SS, N = 64)        -> P^T  = exp2(S^T c - lse)
//      dP^T = V dO^T           (SS, N = 64)        -> dS^T = P^T o (dP^T - delta)
//      dV  += P^T dO,  dK += dS^T Q                (TS, P^T | dS^T packed 16 bit in TMEM, accumulators resident in TMEM)
//      dQ^T = K^T dS^T         (SS, A = the K tile read MN-major, B = dS^T written to shared memory by the warpgroup)
//
// S and dP are computed once (the two-kernel backward in attn_bwd_sm100.cu computes them twice: 7 GEMMs).  What leaves
// the SM asynchronously:
//   * dQ^T of a step is drained TMEM -> registers -> shared memory ([q][d] fp32, one 32 x 32 box per warp and half
//     step) and added into this rank's fp32 dQ accumulator with cp.reduce.async.bulk.tensor (.add.f32, L2 reduction);
//   * dK / dV of an item are complete for this rank's queries when the item ends: with one rank they are written as
//     16 bit directly; in a ring they are added (same TMA reduction, 128B-swizzled boxes) into the OWNER's fp32
//     accumulators over NVLink — the only cross-GPU traffic of the backward besides the K/V slots themselves, and it
//     overlaps with the MMAs of the next item.  No Q / dO / statistics gather, no copy-engine pass, no second kernel.
//
// Reference: the Triton `_bwd_kernel` (triton_flash_attn.py:509-798) recomputes per hop inside a Python ring loop
// (ring_flash_attention_cuda.py:211-351) and ships k, v, dk, dv around the ring in 16 bit.
//
// Warp roles (384 threads, 1 CTA / SM, persistent):
//   warps 0-3 / 4-7 : warpgroup of stream 0 / 1 (even / odd steps): softmax algebra, dS^T to smem, dQ^T drain, epilogue
//   warp 8          : TMA producer (K/V tile per item; Q, dO, lse, delta per step; 3-stage ring)
//   warp 9          : logit issuer: S^T(j) into X_w (two steps ahead) and dP^T(j) into Y_w
//   warp 10         : acc issuer  : dV, dK and dQ^T(j) once the warpgroup has handed P^T | dS^T over
// TMEM (512 columns): stream w: X_w = [128 w, +64): S^T only (refilled as soon as it has been pulled into registers);
// Y_w = [128 w + 64, +64): dP^T -> P^T | dS^T (16 bit, 32 + 32 columns) -> dQ^T; dK = [256, 384), dV = [384, 512).
// The tcgen05.mma that read P^T | dS^T from Y_w and the one that overwrites it with dQ^T are issued by ONE thread in that
// order (the tensor pipe executes in order); every other reuse of a block is ordered by an mbarrier.
#include "attn_common.cuh"

namespace rab {
namespace {

constexpr int NTHREADS = 384;
constexpr int SUB128 = 128 * 128;  // 64-element-wide sub-tile, 128 rows
constexpr int SUB64 = 64 * 128;    // 64-element-wide sub-tile, 64 rows
constexpr int QST = 3;             // Q / dO stages
constexpr int D = 128;

struct FzSmem {
  static constexpr int KV_TILE = 2 * SUB128;  // 128 keys x 128 d
  static constexpr int Q_TILE = 2 * SUB64;    // 64 queries x 128 d
  alignas(1024) uint8_t k[KV_TILE];
  alignas(1024) uint8_t v[KV_TILE];
  alignas(1024) uint8_t q[QST][Q_TILE];
  alignas(1024) uint8_t dout[QST][Q_TILE];
  alignas(1024) uint8_t ds[2][128 * 128];   // dS^T tile of stream w (16 bit, 128B swizzle); doubles as drain staging "A"
  alignas(1024) uint8_t stg[2][128 * 128];  // drain staging "B" of warpgroup w (4 KB per warp)
  alignas(16) float lse2[QST][64];
  alignas(16) float delta[QST][64];
  uint64_t kv_full, kv_empty;
  uint64_t qd_full[QST], qd_empty[QST];
  uint64_t s_full[2], s_free[2], dp_full[2], pds_ready[2], dq_full[2], y_free[2];
  uint64_t acc_done, epi_done;
  uint32_t tmem_base;
};

struct FzItem {
  int owner, b, kvh, kt, key0;
  int klo, khi;
  bool k_tail;
};

__device__ __forceinline__ int fz_num_items(const AttnBwdFusedParams& p) {
  return p.hop_count * p.batch * p.kv_heads * ((p.n_k + 127) / 128);
}

// Item order: hop-major (the K/V of later hops arrive later), then (batch, kv head), then key tile ascending (under
// causal masking early key tiles are the heaviest).  CTAs that run at the same time therefore work on neighbouring
// key tiles of the SAME head and sweep the same Q / dO tiles and dQ accumulator rows at about the same time (L2
// locality of the streamed operands and of the fp32 reductions).  Waves of gridDim items alternate direction so that
// the static round-robin assignment balances the triangular work.
__device__ __forceinline__ void fz_decode(const AttnBwdFusedParams& p, int L, FzItem& it) {
  const int G = gridDim.x;
  const int total = fz_num_items(p);
  const int wave = L / G, lane = L % G;
  int Lp = L;
  if ((wave & 1) && (wave + 1) * G <= total) Lp = wave * G + (G - 1 - lane);
  const int nkt = (p.n_k + 127) / 128;
  const int bhk = p.batch * p.kv_heads;
  it.kt = Lp % nkt;
  const int r = Lp / nkt;
  const int bkv = r % bhk;
  it.owner = p.hop_owner[r / bhk];
  it.b = bkv / p.kv_heads;
  it.kvh = bkv % p.kv_heads;
  it.key0 = it.kt * 128;
  pos_range(p.pos, it.owner, it.key0, min(it.key0 + 128, p.n_k) - 1, it.klo, it.khi);
  it.k_tail = (it.key0 + 128) > p.n_k;
}

using FzScan = WarpTileScan<1, true>;

// streamed side = LOCAL query tiles of 64 rows; rep = index of the query head inside the GQA group
__device__ __forceinline__ void fz_init_scan(FzScan& sc, const AttnBwdFusedParams& p, const FzItem& it) {
  sc.pm = &p.pos;
  sc.hop_owner = p.self_owner;
  sc.hop_count = 1;
  sc.groups = p.heads / p.kv_heads;
  sc.n_stream = p.n_q;
  sc.tile = 64;
  sc.stream_off = p.q_pos_offset;
  sc.stat_off = 0;
  sc.mc = MaskCfg{p.causal, p.window, p.kmask_bits != nullptr};
  sc.st[0] = StatRange{it.klo, it.khi, true, it.k_tail};
}

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// warp 8: TMA producer
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fz_producer(FzSmem& sm, const AttnBwdFusedParams& p, const CUtensorMap* map_qd64,
                                            const CUtensorMap* map_kv) {
  const int lane = lane_id();
  uint32_t n_item = 0, n_tile = 0;
  uint32_t ready_mask = 1u << p.rank;
  const int total = fz_num_items(p);
  const size_t stat_half = (size_t)p.batch * p.heads * p.n_pad;
  for (int L = blockIdx.x; L < total; L += gridDim.x, ++n_item) {
    FzItem it;
    fz_decode(p, L, it);
    if (lane == 0) {
      if (!((ready_mask >> it.owner) & 1u)) {
        if (p.ready != nullptr) {
          spin_until_ge_gpu(p.ready + it.owner, p.ready_target, 1101);
          fence_proxy_async_global();
        }
        ready_mask |= 1u << it.owner;
      }
      mbar_wait(&sm.kv_empty, (n_item & 1) ^ 1, 1100);
      mbar_expect_tx(&sm.kv_full, 2 * FzSmem::KV_TILE);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        tma_load_4d(sm.k + s * SUB128, map_kv, &sm.kv_full, s * 64, it.key0, it.b * p.kv_heads + it.kvh,
                    it.owner * 2);
        tma_load_4d(sm.v + s * SUB128, map_kv, &sm.kv_full, s * 64, it.key0, it.b * p.kv_heads + it.kvh,
                    it.owner * 2 + 1);
      }
    }
    ready_mask = __shfl_sync(0xffffffffu, ready_mask, 0);
    FzScan scan;
    fz_init_scan(scan, p, it);
    ScanTile t;
    while (scan.next(lane, t)) {
      if (lane == 0) {
        const uint32_t st = n_tile % QST, ph = (n_tile / QST) & 1;
        const int h = t.rep * p.kv_heads + it.kvh;
        const int bh = it.b * p.heads + h;
        mbar_wait(&sm.qd_empty[st], ph ^ 1, 1110 + st);
        mbar_expect_tx(&sm.qd_full[st], 2 * FzSmem::Q_TILE + 2 * 64 * 4);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          tma_load_4d(sm.q[st] + s * SUB64, map_qd64, &sm.qd_full[st], s * 64, t.idx * 64, bh, 0);
          tma_load_4d(sm.dout[st] + s * SUB64, map_qd64, &sm.qd_full[st], s * 64, t.idx * 64, bh, 1);
        }
        const float* srow = p.stat + (size_t)bh * p.n_pad + (size_t)t.idx * 64;
        bulk_load_1d(sm.lse2[st], srow, 64 * 4, &sm.qd_full[st]);
        bulk_load_1d(sm.delta[st], srow + stat_half, 64 * 4, &sm.qd_full[st]);
      }
      n_tile++;
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// warp 9: logit issuer.  Per step j (stream w = j & 1):
//    S^T(j)  = K Q^T  into X_w   as soon as the warpgroup has pulled S^T(j-2) into registers        (s_free[w])
//    dP^T(j) = V dO^T into Y_w   as soon as the warpgroup has drained dQ^T(j-2) out of the block    (y_free[w])
// Issue order S(0) S(1) | dP(0) S(2) | dP(1) S(3) | ... follows the order in which those events happen, so every
// wait is a blocking in-order mbarrier wait and S^T runs two steps ahead of the warpgroup.
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ void fz_issue_logits(FzSmem& sm, const AttnBwdFusedParams& p, uint32_t tmem_in) {
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0, BF16 ? 1 : 0);
  constexpr uint64_t kmaj = umma_smem_desc_hi_lo(16, 1024, UMMA_LAYOUT_SW128);
  constexpr uint32_t STAGE16 = FzSmem::Q_TILE >> 4;
  const int lane = lane_id();
  const uint32_t tmem = warp_uniform(tmem_in);
  uint32_t n_item = 0, tile_base = 0;
  uint32_t c_s[2] = {0, 0};  // S^T issued per stream (cumulative): s_free parity
  uint32_t c_p[2] = {0, 0};  // dP^T issued per stream (cumulative): y_free parity
  const int total = fz_num_items(p);
  for (int L = blockIdx.x; L < total; L += gridDim.x, ++n_item) {
    FzItem it;
    fz_decode(p, L, it);
    FzScan scan;
    fz_init_scan(scan, p, it);
    const uint32_t ntiles = scan.count(lane);
    mbar_wait(&sm.kv_full, n_item & 1, 1200);
    tc_fence_after();
    const uint64_t k_desc = umma_desc(kmaj, smem_u32(sm.k)), v_desc = umma_desc(kmaj, smem_u32(sm.v));
    const uint64_t q_kdesc0 = umma_desc(kmaj, smem_u32(sm.q[0]));
    const uint64_t do_kdesc0 = umma_desc(kmaj, smem_u32(sm.dout[0]));

    auto issue_s = [&](uint32_t j) {
      const uint32_t w = j & 1u;
      const uint32_t g = tile_base + j;
      const uint32_t st = g % QST, ph = (g / QST) & 1;
      mbar_wait(&sm.qd_full[st], ph, 1210 + st);
      if (c_s[w] > 0) mbar_wait(&sm.s_free[w], (c_s[w] - 1u) & 1, 1220 + w);
      tc_fence_after();
      const uint32_t x_tm = tmem + w * 128u;
      const uint64_t qk = q_kdesc0 + uint64_t(st * STAGE16);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t offk = (kk / 4) * SUB128 + (kk % 4) * 32;
          const uint32_t offq = (kk / 4) * SUB64 + (kk % 4) * 32;
          umma_ss(x_tm, umma_desc_add(k_desc, offk), umma_desc_add(qk, offq), idesc_s, kk > 0);
        }
        umma_commit(&sm.s_full[w]);
      }
      __syncwarp();
      c_s[w]++;
    };
    auto issue_dp = [&](uint32_t j) {
      const uint32_t w = j & 1u;
      const uint32_t g = tile_base + j;
      const uint32_t st = g % QST, ph = (g / QST) & 1;
      mbar_wait(&sm.qd_full[st], ph, 1230 + st);
      if (c_p[w] > 0) mbar_wait(&sm.y_free[w], (c_p[w] - 1u) & 1, 1240 + w);
      tc_fence_after();
      const uint32_t y_tm = tmem + w * 128u + 64u;
      const uint64_t dok = do_kdesc0 + uint64_t(st * STAGE16);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t offk = (kk / 4) * SUB128 + (kk % 4) * 32;
          const uint32_t offq = (kk / 4) * SUB64 + (kk % 4) * 32;
          umma_ss(y_tm, umma_desc_add(v_desc, offk), umma_desc_add(dok, offq), idesc_s, kk > 0);
        }
        umma_commit(&sm.dp_full[w]);
      }
      __syncwarp();
      c_p[w]++;
    };

    if (ntiles > 0) issue_s(0);
    if (ntiles > 1) issue_s(1);
    for (uint32_t j = 0; j < ntiles; ++j) {
      issue_dp(j);
      if (j + 2 < ntiles) issue_s(j + 2);
    }
    umma_commit_w(&sm.kv_empty);  // second arrival comes from the acc issuer
    tile_base += ntiles;
  }
}

// ------------------------------------------------------------------------------------------------
// warp 10: acc issuer.  Once the warpgroup has packed P^T | dS^T into Y_w and written dS^T to shared memory:
//    dV += P^T dO, dK += dS^T Q (TS), then dQ^T(j) = K^T dS^T (SS) into the SAME block Y_w — the tensor pipe executes
//    this thread's instructions in order, so the overwrite follows the reads.
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ void fz_issue_acc(FzSmem& sm, const AttnBwdFusedParams& p, uint32_t tmem_in) {
  constexpr uint32_t idesc_acc = umma_idesc_bf16(128, D, 0, 1, BF16 ? 1 : 0);
  constexpr uint32_t idesc_dqt = umma_idesc_bf16(D, 64, 1, 1, BF16 ? 1 : 0);
  constexpr uint64_t mnmaj64 = umma_smem_desc_hi_lo(SUB64, 1024, UMMA_LAYOUT_SW128);
  constexpr uint64_t mnmaj128 = umma_smem_desc_hi_lo(SUB128, 1024, UMMA_LAYOUT_SW128);
  constexpr uint32_t STAGE16 = FzSmem::Q_TILE >> 4;
  const int lane = lane_id();
  const uint32_t tmem = warp_uniform(tmem_in);
  const uint32_t dk_tm = tmem + 256, dv_tm = tmem + 256 + D;
  uint32_t n_item = 0, tile_base = 0;
  uint32_t c_b[2] = {0, 0};  // batches issued per stream (cumulative): pds_ready parity
  const int total = fz_num_items(p);
  for (int L = blockIdx.x; L < total; L += gridDim.x, ++n_item) {
    FzItem it;
    fz_decode(p, L, it);
    FzScan scan;
    fz_init_scan(scan, p, it);
    const uint32_t ntiles = scan.count(lane);
    mbar_wait(&sm.kv_full, n_item & 1, 1300);
    tc_fence_after();
    const uint64_t k_mn = umma_desc(mnmaj128, smem_u32(sm.k));
    const uint64_t q_mndesc0 = umma_desc(mnmaj64, smem_u32(sm.q[0]));
    const uint64_t do_mndesc0 = umma_desc(mnmaj64, smem_u32(sm.dout[0]));
    const uint64_t ds_mn0 = umma_desc(mnmaj128, smem_u32(sm.ds[0]));
    for (uint32_t j = 0; j < ntiles; ++j) {
      const uint32_t w = j & 1u;
      const uint32_t st = (tile_base + j) % QST;
      mbar_wait(&sm.pds_ready[w], c_b[w] & 1, 1320 + w);
      if (j == 0) mbar_wait(&sm.epi_done, (n_item & 1) ^ 1, 1330);
      tc_fence_after();
      const uint32_t y_tm = tmem + w * 128u + 64u;
      const uint64_t qmn = q_mndesc0 + uint64_t(st * STAGE16), domn = do_mndesc0 + uint64_t(st * STAGE16);
      const uint64_t ds_mn = ds_mn0 + uint64_t(w * ((128 * 128) >> 4));
      const uint32_t acc0 = j > 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 64 / 16; ++kk)
          umma_ts(dv_tm, y_tm + kk * 8, umma_desc_add(domn, kk * 2048), idesc_acc, kk > 0 ? 1u : acc0);
#pragma unroll
        for (int kk = 0; kk < 64 / 16; ++kk)
          umma_ts(dk_tm, y_tm + 32 + kk * 8, umma_desc_add(qmn, kk * 2048), idesc_acc, kk > 0 ? 1u : acc0);
        umma_commit(&sm.qd_empty[st]);
        // dQ^T[d, q] = K^T[d, keys] dS^T[keys, q]: A = the K tile read MN-major (like V in the forward's P V), B = the
        // dS^T tile the warpgroup wrote to shared memory
#pragma unroll
        for (int kk = 0; kk < 128 / 16; ++kk)
          umma_ss(y_tm, umma_desc_add(k_mn, kk * 2048), umma_desc_add(ds_mn, kk * 2048), idesc_dqt, kk > 0);
        umma_commit(&sm.dq_full[w]);
      }
      __syncwarp();
      c_b[w]++;
    }
    if (ntiles == 0) mbar_wait(&sm.epi_done, (n_item & 1) ^ 1, 1331);
    umma_commit_w(&sm.acc_done);
    umma_commit_w(&sm.kv_empty);
    tile_base += ntiles;
  }
}

// ------------------------------------------------------------------------------------------------
// warps 0-7: the two warpgroups (thread = key row of the stationary tile = TMEM lane).  Software pipelined per stream:
//
//     A1(k)   : S^T(k) -> registers (frees X_w: s_free), P = exp2(S^T c - lse) with the mask applied, kept as fp32
//     A2(k)   : dP^T(k) -> registers, dS^T = P o (dP^T - delta); P^T | dS^T packed into Y_w, dS^T to shared memory,
//               arrive pds_ready  (warp 10 now issues dV, dK, dQ^T(k))
//     A1(k+1) : runs while the tensor core works on the batch of step k
//     drain(k): dQ^T(k) Y_w -> registers -> shared memory boxes -> TMA reduce-add; frees Y_w (y_free)
// ------------------------------------------------------------------------------------------------
template <bool BF16, bool RING, bool CLAMP>
__device__ __forceinline__ void fz_softmax(FzSmem& sm, const AttnBwdFusedParams& p, const int W, uint32_t tmem,
                                           const CUtensorMap* map_dq, const CUtensorMap* map_dkv) {
  const int wg_tid = threadIdx.x - 128 * W;
  const int wq = wg_tid / 32;  // warp inside the warpgroup = TMEM lane quadrant
  const uint32_t lane_off = uint32_t(wq * 32) << 16;
  const uint32_t x_tm = tmem + W * 128 + lane_off;
  const uint32_t y_tm = tmem + W * 128 + 64 + lane_off;
  const int lane = lane_id();
  uint32_t n_a1 = 0, n_done = 0, n_item = 0, tile_base = 0;  // A1 passes / completed steps of this stream (cumulative)
  // per-warp drain staging: A aliases this warp's 32 rows of the dS^T tile, B is private
  uint8_t* const buf_a = sm.ds[W] + wq * 4096;
  uint8_t* const buf_b = sm.stg[W] + wq * 4096;

  const float mul = CLAMP ? 1.f : p.scale * kLog2e;
  const float pre = CLAMP ? p.scale / p.softclamp : 0.f;
  const float post = CLAMP ? p.softclamp * kLog2e : 0.f;

  const int total = fz_num_items(p);
  for (int L = blockIdx.x; L < total; L += gridDim.x, ++n_item) {
    FzItem it;
    fz_decode(p, L, it);
    const int key = it.key0 + wg_tid;
    bool key_ok = key < p.n_k;
    const int pos_k = pos_of(p.pos, it.owner, min(key, p.n_k - 1));
    if (key_ok && p.kmask_bits != nullptr) {
      const uint32_t wbits = p.kmask_bits[((size_t)it.owner * p.batch + it.b) * p.kmask_words + (key >> 5)];
      key_ok = (wbits >> (key & 31)) & 1u;
    }

    FzScan scan;
    fz_init_scan(scan, p, it);
    uint32_t jn = 0;
    // next tile of THIS stream (tiles alternate between the two warpgroups); both walk the whole sequence
    auto next_mine = [&](ScanTile& t, uint32_t& jj) -> bool {
      while (scan.next(lane, t)) {
        jj = jn++;
        if ((jj & 1u) == (uint32_t)W) return true;
      }
      return false;
    };

    // ---- A1: logits -> probabilities (fp32, masked), kept in registers ------------------------------------------------
    auto pass_a1 = [&](const ScanTile& t, const uint32_t jj, uint32_t (&sr)[64], float (&ch)[CLAMP ? 64 : 1]) {
      const uint32_t stg = (tile_base + jj) % QST;
      mbar_wait(&sm.s_full[W], n_a1 & 1, 1400 + W);
      tc_fence_after();
      tmem_ld32(x_tm + 0, sr + 0);
      tmem_ld32(x_tm + 32, sr + 32);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(&sm.s_free[W]);  // X_w may take S^T of this stream's next step now
      n_a1++;
      const float4* l4 = reinterpret_cast<const float4*>(sm.lse2[stg]);
      if (!t.part[0] && !CLAMP) {
        const float2 mul2 = make_float2(mul, mul);
#pragma unroll
        for (int q4 = 0; q4 < 16; ++q4) {
          const float4 lv = l4[q4];
          const float2 a01 = ffma2(make_float2(__uint_as_float(sr[q4 * 4 + 0]), __uint_as_float(sr[q4 * 4 + 1])), mul2,
                                   make_float2(-lv.x, -lv.y));
          const float2 a23 = ffma2(make_float2(__uint_as_float(sr[q4 * 4 + 2]), __uint_as_float(sr[q4 * 4 + 3])), mul2,
                                   make_float2(-lv.z, -lv.w));
          sr[q4 * 4 + 0] = __float_as_uint(fast_exp2(a01.x));
          sr[q4 * 4 + 1] = __float_as_uint(fast_exp2(a01.y));
          sr[q4 * 4 + 2] = __float_as_uint(fast_exp2(a23.x));
          sr[q4 * 4 + 3] = __float_as_uint(fast_exp2(a23.y));
        }
      } else {
        const int c0 = t.idx * 64;
        const int split = p.pos.seg_len - c0;
        const int a0 = p.pos.base0[t.owner] + p.pos.stride * c0 + p.q_pos_offset;
        const int a1 = p.pos.base1[t.owner] + p.pos.stride * (c0 - p.pos.seg_len) + p.q_pos_offset;
        const int ncols = p.n_q - c0;
        const bool part = t.part[0];
#pragma unroll
        for (int q4 = 0; q4 < 16; ++q4) {
          const float4 lv = l4[q4];
          const float ls[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = q4 * 4 + e;
            const float sv = __uint_as_float(sr[j]);
            float pj;
            if (CLAMP) {
              const float th = fast_tanh(sv * pre);
              pj = fast_exp2(fmaf(th, post, -ls[e]));
              ch[CLAMP ? j : 0] = 1.f - th * th;
            } else {
              pj = fast_exp2(fmaf(sv, mul, -ls[e]));
            }
            bool keep = true;
            if (part) {
              const int pq = (j < split ? a0 : a1) + p.pos.stride * j;
              keep = key_ok && (j < ncols);
              if (p.causal) {
                keep = keep && (pos_k <= pq);
                if (p.window > 0) keep = keep && (pq - pos_k <= p.window);
              }
            }
            sr[j] = __float_as_uint(keep ? pj : 0.f);
          }
        }
      }
    };

    // ---- A2: dP^T -> dS^T; hand P^T | dS^T to the acc issuer ----------------------------------------------------------
    auto pass_a2 = [&](const uint32_t jj, uint32_t (&sr)[64], float (&ch)[CLAMP ? 64 : 1]) {
      const uint32_t stg = (tile_base + jj) % QST;
      const float4* d4 = reinterpret_cast<const float4*>(sm.delta[stg]);
      uint32_t dw[32];
      mbar_wait(&sm.dp_full[W], n_done & 1, 1404 + W);
      tc_fence_after();
      // Y_w holds dP^T (64 fp32 columns).  P^T (16 bit, 32 columns) may overwrite columns [0, 32) as soon as that half
      // of dP^T sits in registers: two sequential half passes keep the live set small
      // (the TMEM read port, shared with the other warpgroup's warp on this sub-partition, stays busy either way).
      uint32_t dpa[32], dpb[32];
      tmem_ld32(y_tm + 0, dpa);
      {
        uint32_t pw[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {  // pack P^T while dP^T is in flight
          const float a = __uint_as_float(sr[2 * i]), b2 = __uint_as_float(sr[2 * i + 1]);
          pw[i] = BF16 ? pack_bf16x2(a, b2) : pack_f16x2(a, b2);
        }
        tc_wait_ld();
        tmem_st32(y_tm, pw);
      }
      auto half = [&](const uint32_t (&dp)[32], const int h0) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 dv = d4[h0 * 8 + q4];
          const int j0 = h0 * 32 + q4 * 4;
          const float2 t01 = fadd2(make_float2(__uint_as_float(dp[q4 * 4 + 0]), __uint_as_float(dp[q4 * 4 + 1])),
                                   make_float2(-dv.x, -dv.y));
          const float2 t23 = fadd2(make_float2(__uint_as_float(dp[q4 * 4 + 2]), __uint_as_float(dp[q4 * 4 + 3])),
                                   make_float2(-dv.z, -dv.w));
          float2 e01 = fmul2(make_float2(__uint_as_float(sr[j0 + 0]), __uint_as_float(sr[j0 + 1])), t01);
          float2 e23 = fmul2(make_float2(__uint_as_float(sr[j0 + 2]), __uint_as_float(sr[j0 + 3])), t23);
          if (CLAMP) {
            e01 = fmul2(e01, make_float2(ch[CLAMP ? j0 + 0 : 0], ch[CLAMP ? j0 + 1 : 0]));
            e23 = fmul2(e23, make_float2(ch[CLAMP ? j0 + 2 : 0], ch[CLAMP ? j0 + 3 : 0]));
          }
          dw[h0 * 16 + q4 * 2] = BF16 ? pack_bf16x2(e01.x, e01.y) : pack_f16x2(e01.x, e01.y);
          dw[h0 * 16 + q4 * 2 + 1] = BF16 ? pack_bf16x2(e23.x, e23.y) : pack_f16x2(e23.x, e23.y);
        }
      };
      half(dpa, 0);
      tmem_ld32(y_tm + 32, dpb);
      tc_wait_ld();
      half(dpb, 1);
      // dS^T next to P^T over the consumed dP^T block (16-bit A operands of dK / dV) ...
      tmem_st32(y_tm + 32, dw);
      // ... and dS^T again as a shared-memory tile (B operand of dQ^T): row = key, 64 queries = 128 bytes, 128B swizzle
      // applied by hand (16-byte chunk c of row r sits at c ^ (r % 8)).  The rows of this warp alias its staging
      // buffer A: the TMA reductions that last read it must have finished reading.
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      {
        uint8_t* row = sm.ds[W] + wg_tid * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(row + ((c ^ (wg_tid & 7)) << 4)) =
              make_uint4(dw[4 * c], dw[4 * c + 1], dw[4 * c + 2], dw[4 * c + 3]);
      }
      fence_proxy_async_shared();
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&sm.pds_ready[W]);
    };

    // ---- drain: dQ^T (lane = d, 64 query columns) -> [q][32 d] fp32 box per warp -> TMA reduce-add --------------------
    auto drain = [&](const ScanTile& t) {
      mbar_wait(&sm.dq_full[W], n_done & 1, 1408 + W);
      tc_fence_after();
      const int head = t.rep * p.kv_heads + it.kvh;
      const int row0 = (it.b * p.heads + head) * p.n_pad + t.idx * 64;
      {
        uint32_t g0[32];
        tmem_ld32(y_tm + 0, g0);
        tc_wait_ld();
        float* dst = reinterpret_cast<float*>(buf_b) + lane;
#pragma unroll
        for (int qq = 0; qq < 32; ++qq) dst[qq * 32] = __uint_as_float(g0[qq]);
        fence_proxy_async_shared();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(map_dq, buf_b, wq * 32, row0);
          bulk_commit();
        }
      }
      {
        uint32_t g1[32];
        tmem_ld32(y_tm + 32, g1);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(&sm.y_free[W]);  // Y_w may take dP^T of this stream's next step now
        float* dst = reinterpret_cast<float*>(buf_a) + lane;
#pragma unroll
        for (int qq = 0; qq < 32; ++qq) dst[qq * 32] = __uint_as_float(g1[qq]);
        fence_proxy_async_shared();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(map_dq, buf_a, wq * 32, row0 + 32);
          bulk_commit();
        }
      }
    };

    ScanTile cur, nxt;
    uint32_t jcur = 0, jnxt = 0;
    uint32_t sr[64];              // probabilities of the step between its A1 and A2 passes
    float ch[CLAMP ? 64 : 1];     // softclamp chain-rule factors (1 - tanh^2)
    bool has = next_mine(cur, jcur);
    if (has) pass_a1(cur, jcur, sr, ch);
    while (has) {
      const bool hasn = next_mine(nxt, jnxt);
      pass_a2(jcur, sr, ch);                  // sr is dead afterwards ...
      if (hasn) pass_a1(nxt, jnxt, sr, ch);   // ... and refilled while the tensor core runs dV / dK / dQ^T of step jcur
      drain(cur);
      n_done++;
      cur = nxt;
      jcur = jnxt;
      has = hasn;
    }
    tile_base += jn;

    // epilogue: warpgroup 0 owns dK, warpgroup 1 owns dV
    mbar_wait(&sm.acc_done, n_item & 1, 1410 + W);
    tc_fence_after();
    {
      const bool any = jn > 0;
      const uint32_t acc_tm = tmem + 256 + (W == 0 ? 0 : D) + lane_off;
      const float osc = W == 0 ? p.scale : 1.f;  // dK carries the folded softmax scale
      if constexpr (RING) {
        // fp32 boxes of 32 rows x 32 columns (128B swizzle) added into the owner's accumulators over NVLink
        if (any) {
          const CUtensorMap* mp = map_dkv + it.owner;
          const int grow0 = ((W * p.batch + it.b) * p.kv_heads + it.kvh) * p.nk_pad + it.key0 + wq * 32;
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t acc[32];
            tmem_ld32(acc_tm + c * 32, acc);
            tc_wait_ld();
            uint8_t* buf = (c & 1) ? buf_a : buf_b;
            if (c >= 2) {
              if (lane == 0) bulk_wait_read<1>();
              __syncwarp();
            }
            uint8_t* row = buf + lane * 128;
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              float4 val;
              val.x = __uint_as_float(acc[4 * k4 + 0]) * osc;
              val.y = __uint_as_float(acc[4 * k4 + 1]) * osc;
              val.z = __uint_as_float(acc[4 * k4 + 2]) * osc;
              val.w = __uint_as_float(acc[4 * k4 + 3]) * osc;
              *reinterpret_cast<float4*>(row + ((k4 ^ (lane & 7)) << 4)) = val;
            }
            fence_proxy_async_shared();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_2d(mp, buf, c * 32, grow0);
              bulk_commit();
            }
          }
        }
      } else {
        const bool row_ok = key < p.n_k;
        uint16_t* out = reinterpret_cast<uint16_t*>(W == 0 ? p.dk : p.dv) +
                        (((size_t)it.b * p.n_k + (row_ok ? key : 0)) * p.kv_heads + it.kvh) * D;
#pragma unroll
        for (int c = 0; c < D; c += 32) {
          uint32_t acc[32];
          if (any) {
            tmem_ld32(acc_tm + c, acc);
            tc_wait_ld();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0u;
          }
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = __uint_as_float(acc[2 * i]) * osc, bq = __uint_as_float(acc[2 * i + 1]) * osc;
            w[i] = BF16 ? pack_bf16x2(a, bq) : pack_f16x2(a, bq);
          }
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(out + c);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
          }
        }
      }
    }
    tc_fence_before();
    mbar_arrive(&sm.epi_done);
  }
  // every reduction issued by this warp has been performed (also the ones that went over NVLink) before the grid ends
  if (lane == 0) bulk_wait<0>();
  __syncwarp();
}

template <bool BF16, bool RING, bool CLAMP>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap map_qd64, const __grid_constant__ CUtensorMap map_kv,
                      const __grid_constant__ CUtensorMap map_dq, const __grid_constant__ AttnBwdFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  FzSmem& sm = *reinterpret_cast<FzSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x / 32;
  if (threadIdx.x == 0) {
    mbar_init(&sm.kv_full, 1);
    mbar_init(&sm.kv_empty, 2);
    for (int i = 0; i < QST; ++i) {
      mbar_init(&sm.qd_full[i], 1);
      mbar_init(&sm.qd_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.s_full[i], 1);
      mbar_init(&sm.s_free[i], 128);
      mbar_init(&sm.dp_full[i], 1);
      mbar_init(&sm.pds_ready[i], 128);
      mbar_init(&sm.dq_full[i], 1);
      mbar_init(&sm.y_free[i], 128);
    }
    mbar_init(&sm.acc_done, 1);
    mbar_init(&sm.epi_done, 256);
    fence_mbar_init();
  }
  if (warp == 8 && lane_id() == 0) {
    tma_prefetch_desc(&map_qd64);
    tma_prefetch_desc(&map_kv);
    tma_prefetch_desc(&map_dq);
  }
  if (warp == 10) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  if (warp >= 8) {  // control warps on the highest warp ids: the scheduler favours them
    setmaxnreg_dec<120>();
    if (warp == 8) fz_producer(sm, p, &map_qd64, &map_kv);
    if (warp == 9) fz_issue_logits<BF16>(sm, p, tmem);
    if (warp == 10) fz_issue_acc<BF16>(sm, p, tmem);
  } else {
    setmaxnreg_inc<192>();
    fz_softmax<BF16, RING, CLAMP>(sm, p, warp < 4 ? 0 : 1, tmem, &map_dq, p.map_dkv);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc(tmem, 512);
}

// fp32 accumulator [rows_outer][n_pad][d] -> 16 bit [b][n][h][d] (rows_outer = b * h), scaled.  One thread = 8 elements.
template <bool BF16>
__global__ void acc_convert_kernel(const float* __restrict__ acc, uint16_t* __restrict__ out, int batch, int heads,
                                   int n, int n_pad, int d, float scale) {
  const int vec_per_row = d / 8;
  const long long total = (long long)batch * n * heads * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = i % vec_per_row;
    long long r = i / vec_per_row;
    const int h = r % heads;
    r /= heads;
    const int row = r % n;
    const int b = r / n;
    const float4* src =
        reinterpret_cast<const float4*>(acc + (((long long)b * heads + h) * n_pad + row) * d + c * 8);
    const float4 a = src[0], bq = src[1];
    uint4 w;
    if (BF16) {
      w.x = pack_bf16x2(a.x * scale, a.y * scale);
      w.y = pack_bf16x2(a.z * scale, a.w * scale);
      w.z = pack_bf16x2(bq.x * scale, bq.y * scale);
      w.w = pack_bf16x2(bq.z * scale, bq.w * scale);
    } else {
      w.x = pack_f16x2(a.x * scale, a.y * scale);
      w.y = pack_f16x2(a.z * scale, a.w * scale);
      w.z = pack_f16x2(bq.x * scale, bq.y * scale);
      w.w = pack_f16x2(bq.z * scale, bq.w * scale);
    }
    reinterpret_cast<uint4*>(out)[i] = w;
  }
}

}  // namespace

static_assert(sizeof(FzSmem) + 1024 <= 232448, "fused backward shared memory exceeds the 227 KB per-CTA limit");

size_t attn_bwd_fused_smem_bytes() { return sizeof(FzSmem) + 1024; }

void launch_attn_bwd_fused(const CUtensorMap& map_qd64, const CUtensorMap& map_kv, const CUtensorMap& map_dq,
                           const AttnBwdFusedParams& p, int num_sms, cudaStream_t stream) {
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnBwdFusedParams);
  const bool ring = p.ring_reduce != 0, clamp = p.softclamp > 0.f;
  Kern kern;
  if (clamp) {
    if (ring) kern = p.is_bf16 ? attn_bwd_fused_kernel<true, true, true> : attn_bwd_fused_kernel<false, true, true>;
    else kern = p.is_bf16 ? attn_bwd_fused_kernel<true, false, true> : attn_bwd_fused_kernel<false, false, true>;
  } else {
    if (ring) kern = p.is_bf16 ? attn_bwd_fused_kernel<true, true, false> : attn_bwd_fused_kernel<false, true, false>;
    else kern = p.is_bf16 ? attn_bwd_fused_kernel<true, false, false> : attn_bwd_fused_kernel<false, false, false>;
  }
  const size_t smem = sizeof(FzSmem) + 1024;
  cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "bwd_fused smem attr");
  const int items = p.hop_count * p.batch * p.kv_heads * ((p.n_k + 127) / 128);
  const int grid = items < num_sms ? items : num_sms;
  void* args[] = {(void*)&map_qd64, (void*)&map_kv, (void*)&map_dq, (void*)&p};
  cuda_check(cudaLaunchKernel((void*)kern, dim3(grid), dim3(NTHREADS), args, smem, stream), "bwd_fused launch");
}

void launch_acc_convert(const float* acc, void* out, int batch, int heads, int n, int n_pad, int d, float scale,
                        int is_bf16, cudaStream_t stream) {
  const long long vecs = (long long)batch * n * heads * (d / 8);
  if (vecs == 0) return;
  const int threads = 256;
  long long blocks = (vecs + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  auto kern = is_bf16 ? acc_convert_kernel<true> : acc_convert_kernel<false>;
  kern<<<(int)blocks, threads, 0, stream>>>(acc, reinterpret_cast<uint16_t*>(out), batch, heads, n, n_pad, d, scale);
  cuda_check(cudaGetLastError(), "acc_convert launch");
}

}  // namespace rab
