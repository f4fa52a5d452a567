// Mel front-end on the device: the reference's `mel_spectrogram` (meldataset.py:73-96, called from
// Inference.py:59-85 and Pattern_Generator.py:93-106) -- reflect padding, STFT (periodic Hann window, center = False),
// magnitude sqrt(re^2 + im^2 + 1e-9), triangular mel filters, log(clamp(., 1e-5)) -- as ONE kernel: audio in, log-mel
// [B, n_mels, frames] out; the spectrogram never exists in global memory.
//
// Per frame (256 threads, frames are looped over by persistent CTAs):
//   load     n_fft windowed samples (reflect index at the edges) packed as N/2 complex numbers z[m] = x[2m] + i x[2m+1],
//            written to shared memory in bit-reversed order
//   FFT      N/2-point complex radix-2 DIT in shared memory, twiddles from a table built once per CTA (sincospi)
//   unpack   X[k] = (Z[k] + conj Z[N/2-k]) / 2 - i W_N^k (Z[k] - conj Z[N/2-k]) / 2,  k = 0 .. N/2   (real-input FFT)
//   mel      thread m < n_mels: sum over the filter's non-zero bin range [lo_m, hi_m)   (the filters are sparse)
// HBM traffic per frame: hop new samples in, n_mels values out -- bandwidth-trivial; the kernel is bound by its
// shared-memory butterflies (~5 N log2 N / 2 flops per frame).
#include "common.cuh"
#include "melspec.h"

namespace spk {

template <int NFFT>
__global__ void __launch_bounds__(256) melspec_kernel(const float* __restrict__ audio, int64_t samples, int pad, int hop,
                                                      int win, int frames, int64_t total_frames,
                                                      const float* __restrict__ basis, const int2* __restrict__ ranges,
                                                      int n_mels, float* __restrict__ out_f32, __half* __restrict__ out_f16) {
  constexpr int H = NFFT / 2;               // complex FFT size
  constexpr int LOGH = NFFT == 512 ? 8 : (NFFT == 1024 ? 9 : 10);
  __shared__ float2 z[H];
  __shared__ float2 tw[H / 2];              // W_H^k,  k < H/2
  __shared__ float2 tw2[H / 2 + 1];         // W_NFFT^k, k <= H/2
  __shared__ float window[NFFT];
  __shared__ float mag[H + 1];
  const int tid = threadIdx.x;
  for (int k = tid; k < H / 2; k += 256) {
    float s, c;
    sincospif(-2.f * k / H, &s, &c);
    tw[k] = make_float2(c, s);
  }
  for (int k = tid; k <= H / 2; k += 256) {
    float s, c;
    sincospif(-2.f * k / NFFT, &s, &c);
    tw2[k] = make_float2(c, s);
  }
  const int wl = (NFFT - win) / 2;          // torch.stft centres a short window inside n_fft
  for (int n = tid; n < NFFT; n += 256) {
    const int j = n - wl;
    window[n] = (j >= 0 && j < win) ? 0.5f - 0.5f * cospif(2.f * j / win) : 0.f;     // periodic Hann
  }
  __syncthreads();
  const int64_t period = 2 * (samples - 1);
  for (int64_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
    const int64_t b = f / frames;
    const int fr = static_cast<int>(f % frames);
    const float* src = audio + b * samples;
    const int64_t start = static_cast<int64_t>(fr) * hop - pad;
    for (int m = tid; m < H; m += 256) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int64_t i = start + 2 * m + e;
        if (i < 0 || i >= samples) {        // reflect (no edge repeat), as F.pad(mode='reflect')
          if (period == 0) {
            i = 0;
          } else {
            i %= period;
            if (i < 0) i += period;
            if (i >= samples) i = period - i;
          }
        }
        v[e] = __ldg(src + i) * window[2 * m + e];
      }
      z[__brev(static_cast<unsigned>(m)) >> (32 - LOGH)] = make_float2(v[0], v[1]);
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < LOGH; ++s) {
      const int half = 1 << s;
      for (int t = tid; t < H / 2; t += 256) {
        const int grp = t >> s, pos = t & (half - 1);
        const int i0 = (grp << (s + 1)) + pos, i1 = i0 + half;
        const float2 w = tw[pos << (LOGH - 1 - s)];
        const float2 a = z[i0], bq = z[i1];
        const float2 wb = make_float2(w.x * bq.x - w.y * bq.y, w.x * bq.y + w.y * bq.x);
        z[i0] = make_float2(a.x + wb.x, a.y + wb.y);
        z[i1] = make_float2(a.x - wb.x, a.y - wb.y);
      }
      __syncthreads();
    }
    for (int k = tid; k <= H; k += 256) {
      const float2 a = z[k & (H - 1)], bq = z[(H - k) & (H - 1)];
      const float er = 0.5f * (a.x + bq.x), ei = 0.5f * (a.y - bq.y);     // even part:  (Z[k] + conj Z[H-k]) / 2
      const float orr = 0.5f * (a.y + bq.y), oi = -0.5f * (a.x - bq.x);   // odd part:   (Z[k] - conj Z[H-k]) / (2i)
      const float2 w = k <= H / 2 ? tw2[k] : make_float2(-tw2[H - k].x, tw2[H - k].y);   // W^k = -conj W^(H-k)
      const float re = er + w.x * orr - w.y * oi, im = ei + w.x * oi + w.y * orr;
      mag[k] = sqrtf(re * re + im * im + 1e-9f);
    }
    __syncthreads();
    if (tid < n_mels) {
      const int2 r = ranges[tid];
      const float* brow = basis + static_cast<int64_t>(tid) * (H + 1);
      float acc = 0.f;
      for (int k = r.x; k < r.y; ++k) acc = fmaf(__ldg(brow + k), mag[k], acc);
      const float v = logf(fmaxf(acc, 1e-5f));
      const int64_t o = (b * n_mels + tid) * frames + fr;
      if (out_f32 != nullptr) out_f32[o] = v;
      else out_f16[o] = __float2half_rn(v);
    }
    __syncthreads();
  }
}

int mel_spectrogram(const float* audio, int batch, int64_t samples, int n_fft, int hop, int win, const float* basis,
                    const int32_t* ranges, int n_mels, void* out, int out_fp16, cudaStream_t st) {
  SPK_CHECK(audio && basis && ranges && out, "mel_spectrogram: null argument");
  SPK_CHECK(n_fft == 512 || n_fft == 1024 || n_fft == 2048, "mel_spectrogram: n_fft %d not in {512, 1024, 2048}", n_fft);
  SPK_CHECK(hop >= 1 && win >= 1 && win <= n_fft && n_mels >= 1 && n_mels <= 256, "mel_spectrogram: bad hop / win / n_mels");
  const int pad = (n_fft - hop) / 2;
  SPK_CHECK(batch >= 1 && samples > pad, "mel_spectrogram: %lld samples cannot be reflect-padded by %d",
            static_cast<long long>(samples), pad);
  const int64_t padded = samples + 2 * pad;
  SPK_CHECK(padded >= n_fft, "mel_spectrogram: signal shorter than one frame");
  const int frames = static_cast<int>(1 + (padded - n_fft) / hop);
  const int64_t total = static_cast<int64_t>(batch) * frames;
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = static_cast<int>(std::min<int64_t>(total, static_cast<int64_t>(sms) * 8));
  // algorithmic work: 2.5 N log2 N (real FFT) + 2 * bins * ~2 filters per frame; bytes: hop samples in, n_mels out
  ProfScope prof("mel_spectrogram", total * (2.5 * n_fft * 10 + 4.0 * (n_fft / 2 + 1)),
                 total * (4.0 * hop + (out_fp16 ? 2.0 : 4.0) * n_mels), st);
  float* o32 = out_fp16 ? nullptr : reinterpret_cast<float*>(out);
  __half* o16 = out_fp16 ? reinterpret_cast<__half*>(out) : nullptr;
  const int2* rg = reinterpret_cast<const int2*>(ranges);
  if (n_fft == 512)
    melspec_kernel<512><<<grid, 256, 0, st>>>(audio, samples, pad, hop, win, frames, total, basis, rg, n_mels, o32, o16);
  else if (n_fft == 1024)
    melspec_kernel<1024><<<grid, 256, 0, st>>>(audio, samples, pad, hop, win, frames, total, basis, rg, n_mels, o32, o16);
  else
    melspec_kernel<2048><<<grid, 256, 0, st>>>(audio, samples, pad, hop, win, frames, total, basis, rg, n_mels, o32, o16);
  SPK_CUDA(cudaGetLastError());
  return frames;
}

}  // namespace spk
