// Fused multi-tensor optimiser step: global gradient norm, clip (Train.py:154-159) and RAdam
// (Radam.py:25-90) or AdamW update in two launches over a pointer table (no per-tensor Python loop).
#include "optim.h"

namespace spk {

// Squared gradient norm, DETERMINISTIC: block b writes its partial sum to partial[b] (fixed per-thread, shuffle and
// shared-memory order); the update kernel adds the partials in a fixed order.  Bit-identical on every data-parallel
// rank, so the clip coefficient -- and with it the weights -- cannot drift apart (an atomicAdd of the block sums made
// the norm depend on the arrival order; tests/nccl_worker.py caught the ranks diverging in the last bit).
constexpr int NORM_BLOCKS = 296;
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const __grid_constant__ spk_optim_tensors t, float gs,
                                                          float* __restrict__ partial) {
  __shared__ float red[8];
  float acc = 0.f;
  // Block b owns the elements [b * per, (b + 1) * per) of the concatenation of the table's tensors: it touches one or two
  // tensors instead of walking all of them (43 dependent round trips to memory per block: 33 us for 9.8 MB).  The
  // mapping is fixed, so the partial sums -- and the norm -- stay bit-reproducible.
  int64_t total = 0;
  for (int i = 0; i < t.count; ++i) total += t.numel[i];
  const int64_t per = (total + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * per, hi = lo + per < total ? lo + per : total;
  int64_t base = 0;
  for (int i = 0; i < t.count && base < hi; ++i) {
    const int64_t n = t.numel[i];
    const int64_t a = (lo > base ? lo : base) - base, b = (hi < base + n ? hi : base + n) - base;
    if (a < b) {
      const float* g = t.grad[i];
#pragma unroll 4
      for (int64_t j = a + threadIdx.x; j < b; j += blockDim.x) {
        const float v = g[j] * gs;
        acc = fmaf(v, v, acc);
      }
    }
    base += n;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

struct StepScalars {
  int kind;          // 0 RAdam, 1 AdamW
  int rectified;     // RAdam: N_sma >= 5
  float lr, beta1, beta2, eps, wd;
  float step_size;   // RAdam: Radam.py:70-76 ; AdamW: 1 / (1 - beta1^t)
  float inv_sqrt_bc2;  // AdamW: 1 / sqrt(1 - beta2^t)
  float max_norm, gs;
};

__global__ void __launch_bounds__(256) optim_update_kernel(const __grid_constant__ spk_optim_tensors t,
                                                           const StepScalars s, float* __restrict__ scratch,
                                                           int npartial) {
  float coef = s.gs;
  if (s.max_norm > 0.f) {
    // every block adds the same partials in the same order: warp 0, lane-strided, then the shuffle tree
    __shared__ float total_s;
    if (threadIdx.x < 32) {
      float acc = 0.f;
      for (int i = threadIdx.x; i < npartial; i += 32) acc += scratch[1 + i];
      acc = warp_sum(acc);
      if (threadIdx.x == 0) {
        total_s = acc;
        if (blockIdx.x == 0) scratch[0] = acc;      // squared global norm, for grad_norm()
      }
    }
    __syncthreads();
    coef *= fminf(1.f, s.max_norm / (sqrtf(total_s) + 1e-6f));
  }
  int64_t total = 0;
  for (int i = 0; i < t.count; ++i) total += t.numel[i];
  const int64_t per = (total + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * per, hi = lo + per < total ? lo + per : total;
  int64_t base = 0;
  for (int i = 0; i < t.count && base < hi; ++i) {
    const int64_t n = t.numel[i];
    const int64_t a = (lo > base ? lo : base) - base, b = (hi < base + n ? hi : base + n) - base;
    base += n;
    if (a >= b) continue;
    float* p = t.param[i];
    const float* g = t.grad[i];
    float* m = t.exp_avg[i];
    float* v = t.exp_avg_sq[i];
#pragma unroll 4
    for (int64_t j = a + threadIdx.x; j < b; j += blockDim.x) {
      const float gj = g[j] * coef;
      float pj = p[j];
      const float vj = s.beta2 * v[j] + (1.f - s.beta2) * gj * gj;
      const float mj = s.beta1 * m[j] + (1.f - s.beta1) * gj;
      v[j] = vj;
      m[j] = mj;
      if (s.kind == 0) {
        if (s.wd != 0.f) pj += -s.wd * s.lr * pj;
        if (s.rectified) pj += -s.step_size * s.lr * mj / (sqrtf(vj) + s.eps);
        else pj += -s.step_size * s.lr * mj;
      } else {
        pj *= 1.f - s.lr * s.wd;
        pj -= s.lr * s.step_size * mj / (sqrtf(vj) * s.inv_sqrt_bc2 + s.eps);
      }
      p[j] = pj;
    }
  }
}

int optim_step(const spk_optim_tensors& t, int kind, int64_t step, float lr, float beta1, float beta2, float eps,
               float wd, float max_norm, float grad_scale, float* norm_scratch, int phase, int chunk, int nchunks,
               cudaStream_t st) {
  SPK_CHECK(t.count >= 1 && t.count <= 64, "optim: tensor count %d out of range", t.count);
  SPK_CHECK(phase >= 0 && phase <= 2 && nchunks >= 1 && nchunks <= SPK_OPTIM_MAX_CHUNKS && chunk >= 0 && chunk < nchunks,
            "optim: bad phase / chunk (%d, %d of %d)", phase, chunk, nchunks);
  if (phase == 1) {      // norm partials of this chunk only
    if (max_norm > 0.f) {
      ProfScope prof("optim_norm", 0, 0, st);
      grad_sqnorm_kernel<<<NORM_BLOCKS, 256, 0, st>>>(t, grad_scale, norm_scratch + 1 + chunk * NORM_BLOCKS);
      SPK_CUDA(cudaGetLastError());
    }
    return 0;
  }
  SPK_CHECK(step >= 1, "optim: step is 1-based");
  SPK_CHECK(kind == 0 || kind == 1, "optim: kind must be 0 (RAdam) or 1 (AdamW)");
  StepScalars s;
  s.kind = kind; s.lr = lr; s.beta1 = beta1; s.beta2 = beta2; s.eps = eps; s.wd = wd;
  s.max_norm = max_norm; s.gs = grad_scale;
  const double b1t = pow((double)beta1, (double)step), b2t = pow((double)beta2, (double)step);
  if (kind == 0) {
    const double n_max = 2.0 / (1.0 - beta2) - 1.0;
    const double n_sma = n_max - 2.0 * step * b2t / (1.0 - b2t);
    s.rectified = n_sma >= 5.0;
    if (s.rectified)
      s.step_size = (float)(sqrt((1.0 - b2t) * (n_sma - 4.0) / (n_max - 4.0) * (n_sma - 2.0) / n_sma * n_max /
                                 (n_max - 2.0)) / (1.0 - b1t));
    else
      s.step_size = (float)(1.0 / (1.0 - b1t));
    s.inv_sqrt_bc2 = 1.f;
  } else {
    s.rectified = 1;
    s.step_size = (float)(1.0 / (1.0 - b1t));
    s.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - b2t));
  }
  double numel = 0;
  for (int i = 0; i < t.count; ++i) numel += (double)t.numel[i];
  ProfScope prof("optim_step", 0, numel * 4.0 * (max_norm > 0.f ? 8 : 7), st);
  if (max_norm > 0.f && phase == 0) {
    grad_sqnorm_kernel<<<NORM_BLOCKS, 256, 0, st>>>(t, grad_scale, norm_scratch + 1 + chunk * NORM_BLOCKS);
    SPK_CUDA(cudaGetLastError());
  }
  optim_update_kernel<<<296, 256, 0, st>>>(t, s, norm_scratch, nchunks * NORM_BLOCKS);
  SPK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace spk
