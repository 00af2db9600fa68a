// Descriptor validation kernel: runs ONE 128 x N x K tcgen05.mma tile with caller-supplied
// descriptor parameters so every operand flavour the attention kernels rely on can be checked against
// torch.matmul on the device (tests/test_umma_probe.py):
//   mode 0: D = A * B^T   A[128][K] K-major (smem), B[N][K] K-major (smem)          (S = Q K^T)
//   mode 1: D = A * B     A[128][K] K-major (smem), B[K][N] MN-major (smem)          (O = P V, SS form)
//   mode 2: D = A * B     A[128][K] from TMEM (packed 16-bit, written with tcgen05.st), B[K][N] MN-major
//                                                                                    (O = P V, TS form)
//   mode 3: D = A * B     A given transposed, At[K][128] (MN-major A, TMA-loaded like a V tile), B[K][64] MN-major
//                         written to shared memory BY THE THREADS with the 128B swizzle applied by hand and made
//                         visible to the tensor core with fence.proxy.async (dQ^T = K^T dS^T of a one-kernel backward)
// A and B tiles are brought in with 128B-swizzled TMA boxes of 64 elements x rows, exactly the way the
// attention kernels stage Q/K/V.
#include "kernels.h"
#include "ptx.cuh"

namespace rab {

struct ProbeSmem {
  alignas(1024) uint8_t a[2][128 * 128];  // two 64-element-wide sub-tiles of up to 128 rows
  alignas(1024) uint8_t b[2][128 * 128];
  uint64_t bar_load;
  uint64_t bar_mma;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  ProbeParams p, const uint16_t* __restrict__ a_raw, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  ProbeSmem& sm = *reinterpret_cast<ProbeSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x / 32;
  const int tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(&sm.bar_load, 1);
    mbar_init(&sm.bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t d_tmem = tmem;           // columns [0, N)
  const uint32_t a_tmem = tmem + 256;     // columns [256, 256 + K/2) for mode 2

  if (tid == 0) {
    // A: [128 rows][K] with K contiguous; one box per 64-wide k sub-tile.
    uint32_t bytes = 0;
    if (p.mode == 3) {
      // At: [K rows][128], one box of 64 (m) x K rows per 64-wide m sub-tile
      for (int s = 0; s < 2; ++s) {
        tma_load_2d(sm.a[s], &map_a, &sm.bar_load, s * 64, 0);
        bytes += p.k * 128;
      }
    } else if (p.mode != 2) {
      for (int s = 0; s < p.k / 64; ++s) {
        tma_load_2d(sm.a[s], &map_a, &sm.bar_load, s * 64, 0);
        bytes += 128 * 128;
      }
    }
    if (p.mode == 3) {
      // B is written by the threads below
    } else if (p.mode == 0) {
      // B: [N rows][K], box = 64 (k) x N rows
      for (int s = 0; s < p.k / 64; ++s) {
        tma_load_2d(sm.b[s], &map_b, &sm.bar_load, s * 64, 0);
        bytes += p.n * 128;
      }
    } else {
      // B: [K rows][N], box = 64 (n) x K rows, one per 64-wide n sub-tile
      for (int s = 0; s < p.n / 64; ++s) {
        tma_load_2d(sm.b[s], &map_b, &sm.bar_load, s * 64, 0);
        bytes += p.k * 128;
      }
    }
    mbar_expect_tx(&sm.bar_load, bytes);
  }

  if (p.mode == 2) {
    // each thread writes its row of A (K 16-bit values = K/2 packed words) into TMEM
    const uint32_t* row = reinterpret_cast<const uint32_t*>(a_raw + (size_t)tid * p.k);
    uint32_t regs[16];
    for (int c = 0; c < p.k / 2; c += 16) {
#pragma unroll
      for (int i = 0; i < 16; ++i) regs[i] = row[c + i];
      tmem_st16(a_tmem + (uint32_t(warp * 32) << 16) + c, regs);
    }
    tc_wait_st();
    tc_fence_before();
  }
  if (p.mode == 3) {
    // thread r owns row r of B[K][64] (64 16-bit values = eight 16-byte chunks); chunk c of row r lives at chunk
    // position c ^ (r % 8) of the 128-byte row: the layout a 128B-swizzled TMA box would have produced
    if (tid < p.k) {
      const uint4* src = reinterpret_cast<const uint4*>(a_raw + (size_t)tid * 64);
      uint8_t* row = sm.b[0] + tid * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(row + ((c ^ (tid & 7)) << 4)) = src[c];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    mbar_wait(&sm.bar_load, 0, 1);
    tc_fence_after();
    const uint64_t a_static = umma_smem_desc_hi_lo(p.a_lbo, p.a_sbo, UMMA_LAYOUT_SW128);
    const uint64_t b_static = umma_smem_desc_hi_lo(p.b_lbo, p.b_sbo, UMMA_LAYOUT_SW128);
    for (int kk = 0; kk < p.k / 16; ++kk) {
      uint64_t bdesc;
      if (p.mode == 0) {
        bdesc = umma_desc(b_static, smem_u32(sm.b[kk / 4]) + (kk % 4) * 32);
      } else {
        bdesc = umma_desc(b_static, smem_u32(sm.b[0]) + kk * p.b_kstep_bytes);
      }
      if (p.mode == 2) {
        umma_ts(d_tmem, a_tmem + kk * 8, bdesc, p.idesc, kk > 0);
      } else if (p.mode == 3) {
        const uint64_t adesc = umma_desc(a_static, smem_u32(sm.a[0]) + kk * p.b_kstep_bytes);
        umma_ss(d_tmem, adesc, bdesc, p.idesc, kk > 0);
      } else {
        const uint64_t adesc = umma_desc(a_static, smem_u32(sm.a[kk / 4]) + (kk % 4) * 32);
        umma_ss(d_tmem, adesc, bdesc, p.idesc, kk > 0);
      }
    }
    umma_commit(&sm.bar_mma);
  }
  __syncwarp();
  mbar_wait(&sm.bar_mma, 0, 2);
  tc_fence_after();

  // epilogue: thread t owns row t
  for (int c = 0; c < p.n; c += 32) {
    uint32_t regs[32];
    tmem_ld32(d_tmem + (uint32_t(warp * 32) << 16) + c, regs);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) out[(size_t)tid * p.n + c + i] = __uint_as_float(regs[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

void launch_umma_probe(const CUtensorMap& map_a, const CUtensorMap& map_b, const ProbeParams& p,
                       const void* a_raw, float* out, cudaStream_t stream) {
  const int smem = sizeof(ProbeSmem) + 1024;
  cuda_check(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
             "probe smem attr");
  umma_probe_kernel<<<1, 128, smem, stream>>>(map_a, map_b, p, reinterpret_cast<const uint16_t*>(a_raw), out);
  cuda_check(cudaGetLastError(), "probe launch");
}

// ------------------------------------------------------------------------------------------------
// Issue-rate micro-benchmark: `reps` back-to-back tcgen05.mma (M = 128, K = 16 per instruction) of one operand
// flavour, timed with clock64 from the first issue to the arrival of the final commit.  Used to calibrate the
// tile-time models in BASELINE.md (which shapes are tensor-bound, which are shared-memory-operand-bound).
//   mode 0: SS, B K-major      mode 1: SS, B MN-major      mode 2: TS (A in TMEM), B MN-major
//   mode 3: TS, B K-major      alt != 0: alternate between two accumulators (two independent streams)
// ------------------------------------------------------------------------------------------------
struct RateSmem {
  alignas(1024) uint8_t a[2 * 128 * 128];  // 128 rows x 128 k (two SW128 sub-tiles)
  alignas(1024) uint8_t b[2 * 256 * 128];  // up to 256 rows x 128 k, or 128 k x 256 n
  uint64_t bar;
  uint32_t tmem_base;
};

template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int reps, int alt, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  RateSmem& sm = *reinterpret_cast<RateSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x / 32;
  for (uint32_t i = threadIdx.x; i < sizeof(sm.a) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm.a)[i] = 0x3c003c00u;
  for (uint32_t i = threadIdx.x; i < sizeof(sm.b) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm.b)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&sm.bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  {  // A operand for the TS modes: columns [448, 512)
    uint32_t regs[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) regs[i] = 0x3c003c00u;
    tmem_st32(tmem + 448 + (uint32_t(warp * 32) << 16), regs);
    tmem_st32(tmem + 480 + (uint32_t(warp * 32) << 16), regs);
    tc_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {  // the whole warp runs the loop; one elected lane issues (the attention kernels' pattern)
    constexpr bool b_mn = (MODE == 1 || MODE == 2);
    constexpr bool a_tmem = (MODE >= 2);
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, b_mn ? 1 : 0, 1);
    const uint64_t a_desc0 = umma_desc(umma_smem_desc_hi_lo(16, 1024, UMMA_LAYOUT_SW128), smem_u32(sm.a));
    const uint64_t b_desc0 = b_mn ? umma_desc(umma_smem_desc_hi_lo(128 * 128, 1024, UMMA_LAYOUT_SW128), smem_u32(sm.b))
                                  : umma_desc(umma_smem_desc_hi_lo(16, 1024, UMMA_LAYOUT_SW128), smem_u32(sm.b));
    const uint32_t d0 = warp_uniform(tmem), d1 = warp_uniform(tmem + (N <= 128 ? 192 : 0));
    const uint32_t a_tm = warp_uniform(tmem + 448);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 8) {
      const uint32_t d = (alt && (r & 8)) ? d1 : d0;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          constexpr uint32_t dummy = 0;
          (void)dummy;
          const uint32_t a_off = (kk / 4) * (128 * 128) + (kk % 4) * 32;
          const uint32_t b_off = b_mn ? kk * 2048 : (kk / 4) * (N * 128) + (kk % 4) * 32;
          if (a_tmem) {
            umma_ts(d, a_tm + kk * 8, umma_desc_add(b_desc0, b_off), idesc, 1u);
          } else {
            umma_ss(d, umma_desc_add(a_desc0, a_off), umma_desc_add(b_desc0, b_off), idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    umma_commit_w(&sm.bar);
    mbar_wait(&sm.bar, 0, 3);
    const long long t2 = clock64();
    if (threadIdx.x == 0) {
      out[blockIdx.x * 2] = t2 - t0;
      out[blockIdx.x * 2 + 1] = t1 - t0;  // issue-side time
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int MODE, int N>
static void launch_rate_one(int reps, int alt, int ctas, long long* out, cudaStream_t stream) {
  const int smem = sizeof(RateSmem) + 1024;
  cuda_check(cudaFuncSetAttribute(umma_rate_kernel<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
             "rate smem attr");
  umma_rate_kernel<MODE, N><<<ctas, 128, smem, stream>>>(reps, alt, out);
  cuda_check(cudaGetLastError(), "rate launch");
}

template <int MODE>
static void launch_rate_mode(int n, int reps, int alt, int ctas, long long* out, cudaStream_t stream) {
  if (n == 64) launch_rate_one<MODE, 64>(reps, alt, ctas, out, stream);
  else if (n == 128) launch_rate_one<MODE, 128>(reps, alt, ctas, out, stream);
  else launch_rate_one<MODE, 256>(reps, alt, ctas, out, stream);
}

void launch_umma_rate(int mode, int n, int reps, int alt, int ctas, long long* out, cudaStream_t stream) {
  switch (mode) {
    case 0: launch_rate_mode<0>(n, reps, alt, ctas, out, stream); break;
    case 1: launch_rate_mode<1>(n, reps, alt, ctas, out, stream); break;
    case 2: launch_rate_mode<2>(n, reps, alt, ctas, out, stream); break;
    default: launch_rate_mode<3>(n, reps, alt, ctas, out, stream); break;
  }
}

}  // namespace rab
