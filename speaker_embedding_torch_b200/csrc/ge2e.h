// Fused GE2E loss forward + backward (see ge2e.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace spk {
size_t ge2e_workspace_bytes(int N, int M);
// dE == nullptr -> forward only (loss).  w, b are device pointers to the 0-dim loss parameters.
int ge2e_fused(const float* E, int N, int M, int D, const float* w, const float* b, float* loss, float* dE, float* dw,
               float* db, void* ws, size_t ws_bytes, cudaStream_t st);
void ge2e_set_tile_v2(int on);   // 0: first (16-row) row-tile stage, 1: re-tiled stage (default)
void ge2e_set_pdl(int on);       // programmatic dependent launch between the three small-N stages (default on)
// tensor-core composition for large N (ge2e_tc.cu); ge2e_fused dispatches to it for N >= GE2E_TC_MIN_SPEAKERS
constexpr int GE2E_TC_MIN_SPEAKERS = 256;
size_t ge2e_tc_workspace_bytes(int N, int M);
int ge2e_tc(const float* E, int N, int M, const float* w, const float* b, float* loss, float* dE, float* dw, float* db,
            void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace spk
