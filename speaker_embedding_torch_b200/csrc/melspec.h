// Mel front-end (see melspec.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spk {
// returns the number of frames (>= 1) or a negative SPK_E* code
int mel_spectrogram(const float* audio, int batch, int64_t samples, int n_fft, int hop, int win, const float* basis,
                    const int32_t* ranges, int n_mels, void* out, int out_fp16, cudaStream_t st);
}  // namespace spk
