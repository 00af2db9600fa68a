// Memory-bound helper kernels: K/V packing into the symmetric ring slot, the cross-device barrier on
// peer-mapped signal pads, and small conversion kernels used by the backward pass.
#include "kernels.h"
#include "ptx.cuh"

namespace rab {
namespace {

// k, v: [b, n, hk, d] (strided, unit stride on d) -> slot: [2][b*hk][n][d] contiguous 16-bit.
// One thread moves 16 bytes.
__global__ void pack_kv_kernel(const uint16_t* __restrict__ k, const uint16_t* __restrict__ v,
                               uint16_t* __restrict__ slot, int batch, int n, int kv_heads, int d, long long k_sb,
                               long long k_sn, long long k_sh, long long v_sb, long long v_sn, long long v_sh) {
  const int vec_per_row = d / 8;
  const long long per_tensor = (long long)batch * kv_heads * n * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < 2 * per_tensor;
       i += (long long)gridDim.x * blockDim.x) {
    const int which = i >= per_tensor;
    long long r = which ? i - per_tensor : i;
    const int c = r % vec_per_row;
    r /= vec_per_row;
    const int row = r % n;
    r /= n;
    const int h = r % kv_heads;
    const int b = r / kv_heads;
    const uint16_t* src = which ? v + b * v_sb + row * v_sn + h * v_sh : k + b * k_sb + row * k_sn + h * k_sh;
    const uint4 val = *reinterpret_cast<const uint4*>(src + c * 8);
    reinterpret_cast<uint4*>(slot)[i] = val;
  }
}

__global__ void device_barrier_kernel(const __grid_constant__ BarrierParams p) {
  const int peer = threadIdx.x;
  if (peer >= p.world) return;
  if (peer != p.rank) {
    __threadfence_system();
    st_release_sys(p.pads[peer] + p.rank, p.epoch);
    const uint32_t* mine = p.pads[p.rank] + peer;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - p.epoch) < 0) {
      __nanosleep(100);
      // ranks may reach their first barrier seconds apart (lazy module loading, allocator warm-up): be generous
      if (clock64() - t0 > 8 * RAB_WATCHDOG_CYCLES) {
        printf("[rab] device barrier watchdog: rank %d waiting for %d epoch %u (have %u)\n", p.rank, peer, p.epoch,
               ld_acquire_sys(mine));
        __trap();
      }
    }
  }
}

}  // namespace

void launch_pack_kv(const void* k, const void* v, void* slot, int batch, int n, int kv_heads, int d, long long k_sb,
                    long long k_sn, long long k_sh, long long v_sb, long long v_sn, long long v_sh,
                    cudaStream_t stream) {
  const long long vecs = 2ll * batch * kv_heads * n * (d / 8);
  if (vecs == 0) return;
  const int threads = 256;
  long long blocks = (vecs + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_kv_kernel<<<(int)blocks, threads, 0, stream>>>(
      reinterpret_cast<const uint16_t*>(k), reinterpret_cast<const uint16_t*>(v), reinterpret_cast<uint16_t*>(slot),
      batch, n, kv_heads, d, k_sb, k_sn, k_sh, v_sb, v_sn, v_sh);
  cuda_check(cudaGetLastError(), "pack_kv launch");
}

void launch_device_barrier(const BarrierParams& p, cudaStream_t stream) {
  device_barrier_kernel<<<1, 32, 0, stream>>>(p);
  cuda_check(cudaGetLastError(), "device_barrier launch");
}

}  // namespace rab
