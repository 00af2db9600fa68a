// Symmetric (peer-mapped) device memory for one NVSwitch box: every rank allocates the same-sized
// region with cudaMalloc, exports a CUDA IPC handle, and maps every peer's region into its own address
// space.  Kernels then address peer memory with ordinary global loads / TMA bulk copies over NVLink.
//
// The handle exchange itself (64 bytes per rank) rides on whatever torch.distributed backend is up;
// this file is transport-agnostic.
#include "symm.h"

#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstring>
#include <mutex>
#include <unordered_map>

#include "tmap.h"

namespace rab {

namespace {
std::mutex g_mu;
std::unordered_map<void*, size_t> g_allocs;   // base -> bytes (owned by this process)
std::unordered_map<void*, int> g_imports;     // imported peer base -> refcount
}  // namespace

void* symm_alloc(size_t bytes, unsigned char handle_out[kIpcHandleBytes]) {
  void* base = nullptr;
  // round to 2 MiB so the region is its own allocation (IPC handles map whole allocations)
  const size_t rounded = (bytes + (2u << 20) - 1) & ~size_t((2u << 20) - 1);
  cuda_check(cudaMalloc(&base, rounded), "symm_alloc cudaMalloc");
  cuda_check(cudaMemset(base, 0, rounded), "symm_alloc memset");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, base);
  if (e != cudaSuccess) {
    // single-process use still works without IPC; report an all-zero handle
    (void)cudaGetLastError();
    std::memset(&h, 0, sizeof(h));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == kIpcHandleBytes, "IPC handle size");
  std::memcpy(handle_out, &h, kIpcHandleBytes);
  std::lock_guard<std::mutex> lk(g_mu);
  g_allocs[base] = rounded;
  return base;
}

void symm_free(void* base) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_allocs.erase(base);
  }
  cudaFree(base);
}

void* symm_open(const unsigned char handle[kIpcHandleBytes]) {
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, kIpcHandleBytes);
  void* p = nullptr;
  cuda_check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "symm_open cudaIpcOpenMemHandle");
  std::lock_guard<std::mutex> lk(g_mu);
  g_imports[p]++;
  return p;
}

void stream_write_value32(void* addr, unsigned int value, cudaStream_t stream) {
  static PFN_cuStreamWriteValue32_v11070 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr)
      throw std::runtime_error("[ring_attention_b200] cannot resolve cuStreamWriteValue32");
    fn = reinterpret_cast<PFN_cuStreamWriteValue32_v11070>(p);
  });
  CUresult r = fn(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), value,
                  CU_STREAM_WRITE_VALUE_DEFAULT);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("[ring_attention_b200] cuStreamWriteValue32 failed with code " + std::to_string((int)r));
}

void symm_close(void* p) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_imports.find(p);
    if (it == g_imports.end()) return;
    if (--it->second > 0) return;
    g_imports.erase(it);
  }
  cudaIpcCloseMemHandle(p);
}

}  // namespace rab
