// Shared host/device helpers: error reporting, split-fp16 tensors, Philox dropout masks.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>

namespace spk {

// ----------------------------------------------------------------------------- errors
// Thread-local last-error text, exposed through spk_last_error() (api.cu).
char* last_error_buf();
void set_error(const char* fmt, ...);

bool prof_enabled();
void prof_set(bool on);
int prof_begin(const char* tag, double flops, double bytes, cudaStream_t st);
void prof_end(int idx, cudaStream_t st);
int prof_report(char* buf, size_t cap);
struct ProfScope {
  int idx = -1;
  cudaStream_t st;
  ProfScope(const char* tag, double flops, double bytes, cudaStream_t s) : st(s) {
    if (prof_enabled()) idx = prof_begin(tag, flops, bytes, s);
  }
  ~ProfScope() {
    if (idx >= 0) prof_end(idx, st);
  }
};

// One-time setup per (call site, device): cudaFuncSetAttribute / occupancy queries are per device, so a process that
// touches a second GPU must repeat them there.  Thread-safe; the body's error code is returned and not latched.
struct PerDeviceOnce {
  std::mutex mu;
  bool done[64] = {};
  template <class F>
  int run(F fn) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return 0;
    const int r = fn();
    if (r == 0) done[dev] = true;
    return r;
  }
};

#define SPK_EINVAL (-22)
#define SPK_ENOMEM (-12)
#define SPK_EIO (-5)

#define SPK_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      spk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SPK_EIO;                                                                       \
    }                                                                                       \
  } while (0)

#define SPK_CHECK(cond, ...)       \
  do {                             \
    if (!(cond)) {                 \
      spk::set_error(__VA_ARGS__); \
      return SPK_EINVAL;           \
    }                              \
  } while (0)

#define SPK_TRY(expr)        \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

// ----------------------------------------------------------------------------- split-fp16 tensors
// Every activation / gradient that feeds a tensor-core GEMM is stored as 1..3 fp16 "planes":
//   planes = 1  x ~= hi                (11 significant bits; inference)
//   planes = 2  x ~= hi + lo           (22 bits ~ fp32: Ah*Bh + Ah*Bl + Al*Bh, 3 MMAs; training, forward and backward)
//   planes = 3  x ~= hi + mid + lo     (33 bits; six MMAs -- kept for cross-checks, it buys nothing over two planes)
// fp16 rather than bf16 because TWO fp16 planes already carry an fp32 mantissa (two bf16 planes carry 16 bits, which
// flips ReLU gates and costs gradient parity, so the bf16 training forward needed three planes and six MMAs per
// product -- DESIGN.md "precision").  The price is range: values saturate at +-65504 instead of overflowing to inf
// (activations of a post-LN encoder stay orders of magnitude below that), tiny values keep an ABSOLUTE precision of
// 2^-25 (the fp16 subnormal step), and gradients are carried scaled by a power of two chosen per backward pass from
// max |dL/d dvec| (encoder.cu: grad_scale_kernel) so that they sit in the middle of the fp16 range.
// Plane p of a tensor with `plane_stride` elements starts at base + p * plane_stride; a tensor written with 3 planes can
// be read with 2.
typedef __half elem_t;

// two fp32 -> one 32-bit word of two fp16 (a in the low half), round to nearest, saturating to +-65504 (never inf)
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float lo_to_f(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u & 0xFFFFu))); }
__device__ __forceinline__ float hi_to_f(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u >> 16))); }
__device__ __forceinline__ elem_t to_elem(float x) { return __ushort_as_half(static_cast<unsigned short>(pack2(x, 0.f) & 0xFFFFu)); }

// Load 8 consecutive elements (16 B per plane, must be 16-B aligned) of a split tensor as fp32.
__device__ __forceinline__ void load8_split(const elem_t* base, size_t plane_stride, int planes, size_t off,
                                            float (&v)[8]) {
  uint4 h = *reinterpret_cast<const uint4*>(base + off);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = lo_to_f(hw[i]);
    v[2 * i + 1] = hi_to_f(hw[i]);
  }
  for (int p = 1; p < planes; ++p) {
    uint4 l = *reinterpret_cast<const uint4*>(base + p * plane_stride + off);
    const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] += lo_to_f(lw[i]);
      v[2 * i + 1] += hi_to_f(lw[i]);
    }
  }
}
// Store 8 consecutive fp32 values as split planes (16-B aligned): plane p holds fp16(residual after planes < p).
__device__ __forceinline__ void store8_split(elem_t* base, size_t plane_stride, int planes, size_t off,
                                             const float (&v)[8]) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = v[i];
  for (int p = 0; p < planes; ++p) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      w[i] = pack2(r[2 * i], r[2 * i + 1]);
      r[2 * i] -= lo_to_f(w[i]);
      r[2 * i + 1] -= hi_to_f(w[i]);
    }
    *reinterpret_cast<uint4*>(base + p * plane_stride + off) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
__device__ __forceinline__ float load1_split(const elem_t* base, size_t plane_stride, int planes, size_t off) {
  float v = __half2float(base[off]);
  for (int p = 1; p < planes; ++p) v += __half2float(base[p * plane_stride + off]);
  return v;
}
__device__ __forceinline__ void store1_split(elem_t* base, size_t plane_stride, int planes, size_t off, float x) {
  for (int p = 0; p < planes; ++p) {
    const elem_t q = to_elem(x);
    base[p * plane_stride + off] = q;
    x -= __half2float(q);
  }
}

// ----------------------------------------------------------------------------- Philox4x32-7 dropout
// Counter-based RNG for dropout: the keep-mask of element `idx` at dropout site `site` is a pure
// function of (seed, site, idx), so backward regenerates it instead of storing 13 masks.
// One Philox4x32 call (7 rounds, the fewest that pass BigCrush) yields 128 bits = eight 16-bit uniforms,
// i.e. the masks of 8 consecutive elements; the drop probability is quantised to p_q = round(p * 2^16) / 2^16
// and the keep scale is 1 / (1 - p_q), so the mask stays exactly unbiased.  (The epilogues are
// instruction-issue bound in training, ncu r01: the 10-round, 32-bit-per-element version cost 3x more.)
struct Philox4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ Philox4 philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
// Keep-scale (0 or 1/(1-p_q)) for the 8 consecutive elements idx8*8 .. idx8*8+7.
__device__ __forceinline__ void dropout_scale8(uint64_t seed, uint32_t site, uint64_t idx8, uint32_t thresh,
                                               float inv_keep, float (&s)[8]) {
  const Philox4 r = philox4x32_7(static_cast<uint32_t>(idx8), static_cast<uint32_t>(idx8 >> 32), site, 0x5eedu,
                                 static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s[2 * i] = (w[i] & 0xFFFFu) >= thresh ? inv_keep : 0.f;
    s[2 * i + 1] = (w[i] >> 16) >= thresh ? inv_keep : 0.f;
  }
}

struct DropCfg {
  uint64_t seed;
  uint32_t thresh;   // drop iff rand16 < thresh (thresh = round(p * 65536)); 0 disables
  float inv_keep;    // 1 / (1 - thresh / 65536)
};
inline DropCfg make_drop(uint64_t seed, float p, bool training) {
  DropCfg d;
  d.seed = seed;
  d.thresh = 0;
  d.inv_keep = 1.f;
  if (!training || p <= 0.f) return d;
  double t = static_cast<double>(p) * 65536.0 + 0.5;
  d.thresh = t >= 65535.0 ? 65535u : static_cast<uint32_t>(t);
  if (d.thresh == 0) d.thresh = 1;
  d.inv_keep = static_cast<float>(1.0 / (1.0 - d.thresh / 65536.0));
  return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace spk
