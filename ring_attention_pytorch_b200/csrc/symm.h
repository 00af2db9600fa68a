#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace rab {

constexpr int kIpcHandleBytes = 64;

// Allocate `bytes` of zeroed device memory suitable for peer mapping; writes the IPC handle.
void* symm_alloc(size_t bytes, unsigned char handle_out[kIpcHandleBytes]);
void symm_free(void* base);
// Map a peer's region (handle produced by symm_alloc in another process on the same box).
void* symm_open(const unsigned char handle[kIpcHandleBytes]);
void symm_close(void* p);

// cuStreamWriteValue32 on `stream` (resolved through the runtime, like the tensor-map encoder): *addr = value in
// stream order, performed by the stream front end — no kernel, no copy engine.
void stream_write_value32(void* addr, unsigned int value, cudaStream_t stream);

}  // namespace rab
