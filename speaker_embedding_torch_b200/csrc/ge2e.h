// Fused GE2E loss forward + backward (see ge2e.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace spk {
size_t ge2e_workspace_bytes(int N, int M);
// dE == nullptr -> forward only (loss).  w, b are device pointers to the 0-dim loss parameters.
int ge2e_fused(const float* E, int N, int M, int D, const float* w, const float* b, float* loss, float* dE, float* dw,
               float* db, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace spk
