// Fused attention for the TRAINING path (multi-plane split-bf16 operands, dropout on the probabilities, a backward
// pass that recomputes instead of reading stored probabilities).  Replaces, per dense layer, the QK^T GEMM + softmax
// kernel + PV GEMM of the forward and the dV / dP GEMMs + softmax-backward kernel + dQ / dK GEMMs of the backward
// (8 launches over a materialised [B*H, T, T] score tensor in 2-3 planes) by two kernels whose scores never leave the
// SM.  Reference: torch scaled_dot_product_attention inside nn.TransformerEncoderLayer (Modules.py:25-36,53), scale
// 1/sqrt(64), dropout on the probabilities.
//
// Forward, per (slice, head) and 128-query tile:
//   TMA       Q tile, K, V (PL planes each) out of the packed qkv buffer [B*T, 768]
//   tcgen05   S = Q K^T into TMEM (6 plane products for PL = 3, 3 for PL = 2), fp32
//   8 warps   thread <-> query row (two warps per TMEM lane quarter, alternating 32-key chunks): row max, exp2, row
//             sum, Philox keep mask, split into bf16 planes and written BACK INTO TMEM over the scores they came from
//             (tcgen05.st): the probabilities are the A operand of the next MMA straight from tensor memory
//   tcgen05   O = P_drop V  (A from TMEM, V read MN-major from shared memory)
//   8 warps   O * (1 / ((1 - p) * row sum)) -> PL bf16 planes -> att[(b*T + q) * 256 + h*64 ...]
//   kept for the backward: (row max * log2e / 8, row sum) per query and one keep BIT per probability.
//
// TMEM plane layout of P for key step t (16 keys = one MMA K step = 8 columns of packed bf16 pairs, even key in the low
// half -- tools/probe/ts_probe.cu): plane 0 at S columns [16t, 16t+8), plane 1 at [16t+8, 16t+16), plane 2 (PL = 3)
// at P_LO + [8t, 8t+8); i.e. planes 0 and 1 overwrite exactly the 16 score columns they were computed from.
#include "gemm.h"
#include "ptx.cuh"

namespace spk {

__device__ __forceinline__ float at_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: A is [128 lanes x K] packed bf16 pairs in tensor memory (K-major by construction)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void pair_sync(int quarter) {   // the two warps that share a TMEM lane quarter
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}
// keep bits (bit i <-> element idx8 * 8 + i) of one Philox call; same function of (seed, site, element) as dropout_scale8
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint32_t site, uint64_t idx8, uint32_t thresh) {
  const Philox4 r = philox4x32_7(static_cast<uint32_t>(idx8), static_cast<uint32_t>(idx8 >> 32), site, 0x5eedu,
                                 static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bits |= ((w[i] & 0xFFFFu) >= thresh ? 1u : 0u) << (2 * i);
    bits |= ((w[i] >> 16) >= thresh ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}

constexpr int ATF_THREADS = 384;
constexpr int ATF_TMEM_PLO = 256, ATF_TMEM_O = 384;
constexpr float ATF_SC = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)

template <int PL>
struct AtfCfg {
  static constexpr int KV_ROWS = PL == 3 ? 192 : 256;            // max frames of the fused path
  static constexpr int Q_BYTES = 128 * 128, KV_BYTES = KV_ROWS * 128;
  static constexpr int OFF_K = PL * Q_BYTES, OFF_V = OFF_K + PL * KV_BYTES;
  static constexpr int OFF_X = OFF_V + PL * KV_BYTES;            // float xmax[2][128], xsum[2][128]
  static constexpr int OFF_BAR = OFF_X + 2048;
  static constexpr int SMEM = OFF_BAR + 128;
};

struct AttnTrainFwdArgs {
  CUtensorMap q_map[3], k_map[3], v_map[3];
  int B, H, T, Tp, Tk16, Tk64, mtiles, nC;
  __nv_bfloat16* out;      // [PL][B*T, out_ld], head h at column h*64
  int64_t out_ps, out_ld;
  float2* stats;           // [B*H*T] (row max * ATF_SC, row sum of exp2), may be null
  uint32_t* mbits;         // [B*H][nC][T] keep bits of 32-key chunk c of query q, null without dropout
  DropCfg drop;
  uint32_t site;
};

template <int PL>
__global__ void __launch_bounds__(ATF_THREADS, 1) attn_train_fwd_kernel(const __grid_constant__ AttnTrainFwdArgs a) {
  using Cfg = AtfCfg<PL>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sQ = sbase, sK = sbase + Cfg::OFF_K, sV = sbase + Cfg::OFF_V;
  float* xch = reinterpret_cast<float*>(smem_raw + Cfg::OFF_X);   // [0..255] max halves, [256..511] sum halves
  const uint32_t bar_kv = sbase + Cfg::OFF_BAR, bar_q = bar_kv + 8, bar_s = bar_kv + 16, bar_p = bar_kv + 24,
                 bar_o = bar_kv + 32, bar_oe = bar_kv + 40, bar_kvfree = bar_kv + 48, tmem_slot = bar_kv + 56;
  // Every waiter of an mbarrier observes EVERY phase of it, in order (a parity wait that skips a phase passes
  // spuriously): per-unit barriers are waited on once per unit, per-item barriers once per item.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((sbase & 1023u) != 0) __trap();     // the swizzled operand tiles need a 1024-byte aligned base

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int p = 0; p < PL; ++p) { tma_prefetch_desc(&a.q_map[p]); tma_prefetch_desc(&a.k_map[p]); tma_prefetch_desc(&a.v_map[p]); }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_kv, 1); mbar_init(bar_q, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 256); mbar_init(bar_o, 1);
    mbar_init(bar_oe, 256); mbar_init(bar_kvfree, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tS = tmem_base, tPlo = tmem_base + ATF_TMEM_PLO, tO = tmem_base + ATF_TMEM_O;
  const int items = a.B * a.H;
  const int kv_plane = a.Tk64 * 128;      // bytes actually loaded per K / V plane

  if (warp == 0) {
    if (lane == 0) {
      uint32_t u = 0, it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int h = item % a.H, b = item / a.H;
        if (it > 0) mbar_wait(bar_kvfree, (it - 1) & 1u, 0x600u);   // last PV of the previous item: K / V are free
        mbar_arrive_expect_tx(bar_kv, 2u * PL * kv_plane);
#pragma unroll
        for (int p = 0; p < PL; ++p) {
          tma_load_4d(sK + p * Cfg::KV_BYTES, &a.k_map[p], bar_kv, 0, 0, h, b);
          tma_load_4d(sV + p * Cfg::KV_BYTES, &a.v_map[p], bar_kv, 0, 0, h, b);
        }
        for (int mt = 0; mt < a.mtiles; ++mt, ++u) {
          if (u > 0) mbar_wait(bar_s, (u - 1) & 1u, 0x601u);     // previous scores issued and retired: Q is free
          mbar_arrive_expect_tx(bar_q, PL * Cfg::Q_BYTES);
#pragma unroll
          for (int p = 0; p < PL; ++p) tma_load_4d(sQ + p * Cfg::Q_BYTES, &a.q_map[p], bar_q, 0, mt * 128, h, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, a.Tk16, false, false);
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
      constexpr int NCOMBO = PL == 3 ? 6 : 3;
      // plane products, smallest terms first (plane 0 = hi)
      constexpr int PA3[6] = {1, 0, 2, 0, 1, 0}, PB3[6] = {1, 2, 0, 1, 0, 0};
      constexpr int PA2[3] = {1, 0, 0}, PB2[3] = {0, 1, 0};
      uint32_t u = 0, it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        for (int mt = 0; mt < a.mtiles; ++mt, ++u) {
          mbar_wait(bar_q, u & 1u, 0x610u);
          if (mt == 0) mbar_wait(bar_kv, it & 1u, 0x611u);
          tc_fence_after();
          uint32_t acc = 0;
#pragma unroll
          for (int cb = 0; cb < NCOMBO; ++cb) {
            const int pa = PL == 3 ? PA3[cb] : PA2[cb], pb = PL == 3 ? PB3[cb] : PB2[cb];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(tS, umma_smem_desc(sQ + pa * Cfg::Q_BYTES + k * 32, 16, 1024),
                        umma_smem_desc(sK + pb * Cfg::KV_BYTES + k * 32, 16, 1024), idesc_s, acc);
              acc = 1;
            }
          }
          umma_commit(bar_s);
          mbar_wait(bar_p, u & 1u, 0x612u);                       // probabilities are in TMEM
          if (u > 0) mbar_wait(bar_oe, (u - 1) & 1u, 0x613u);     // previous output tile drained
          tc_fence_after();
          acc = 0;
          for (int t = 0; t < a.Tk16 / 16; ++t) {
#pragma unroll
            for (int cb = 0; cb < NCOMBO; ++cb) {
              const int pa = PL == 3 ? PA3[cb] : PA2[cb], pb = PL == 3 ? PB3[cb] : PB2[cb];
              const uint32_t ta = pa == 0 ? tS + 16 * t : (pa == 1 ? tS + 16 * t + 8 : tPlo + 8 * t);
              umma_bf16_ts(tO, ta, umma_smem_desc(sV + pb * Cfg::KV_BYTES + (t >> 2) * 8192 + (t & 3) * 2048, 8192, 1024),
                           idesc_o, acc);
              acc = 1;
            }
          }
          umma_commit(bar_o);
          if (mt == a.mtiles - 1) umma_commit(bar_kvfree);
        }
      }
    }
  } else if (warp >= 4) {
    const int w = (warp - 4) & 3, half = (warp - 4) >> 2;
    const int r = w * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(w * 32) << 16;
    const bool use_drop = a.drop.thresh != 0;
    const float inv_keep = use_drop ? a.drop.inv_keep : 1.f;
    uint32_t u = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int h = item % a.H, b = item / a.H;
      for (int mt = 0; mt < a.mtiles; ++mt, ++u) {
        const uint32_t ph = u & 1u;
        const int q = mt * 128 + r;
        const bool active = mt * 128 + w * 32 < a.T;              // warp-uniform
        mbar_wait(bar_s, ph, 0x620u);
        tc_fence_after();
        uint32_t sreg[32];
        float mx = -INFINITY;
        if (active) {
          for (int c = half; c * 32 < a.T; c += 2) {
            tmem_ld_32x32(tS + t_lane + c * 32, sreg);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < a.T) mx = fmaxf(mx, __uint_as_float(sreg[i]));
          }
        }
        xch[half * 128 + r] = mx;
        pair_sync(w);
        mx = fmaxf(mx, xch[(half ^ 1) * 128 + r]);
        const float mxs = mx * ATF_SC;
        float sum = 0.f;
        if (active) {
          const uint64_t row_idx8 = ((static_cast<uint64_t>(item) * a.T + q) * a.Tp) >> 3;   // Tp % 8 == 0
          for (int c = half; c * 32 < a.Tk16; c += 2) {
            tmem_ld_32x32(tS + t_lane + c * 32, sreg);
            uint32_t keep = 0xFFFFFFFFu;
            if (use_drop) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                keep = (g == 0 ? 0u : keep) | (dropout_keep8(a.drop.seed, a.site, row_idx8 + c * 4 + g, a.drop.thresh) << (8 * g));
            }
            tmem_ld_wait();
            uint32_t o01[32];
            uint32_t o2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int key = c * 32 + 2 * j;
              float e0 = key < a.T ? at_exp2(__uint_as_float(sreg[2 * j]) * ATF_SC - mxs) : 0.f;
              float e1 = key + 1 < a.T ? at_exp2(__uint_as_float(sreg[2 * j + 1]) * ATF_SC - mxs) : 0.f;
              sum += e0 + e1;
              e0 = ((keep >> (2 * j)) & 1u) ? e0 : 0.f;
              e1 = ((keep >> (2 * j + 1)) & 1u) ? e1 : 0.f;
              // key step t = 2c + (j >> 3): planes 0 / 1 at 16t + (j & 7) and 16t + 8 + (j & 7)
              const int slot = (j >> 3) * 16 + (j & 7);
              const __nv_bfloat162 hi = __floats2bfloat162_rn(e0, e1);
              o01[slot] = *reinterpret_cast<const uint32_t*>(&hi);
              e0 -= __bfloat162float(hi.x);
              e1 -= __bfloat162float(hi.y);
              const __nv_bfloat162 mid = __floats2bfloat162_rn(e0, e1);
              o01[slot + 8] = *reinterpret_cast<const uint32_t*>(&mid);
              if (PL == 3) {
                e0 -= __bfloat162float(mid.x);
                e1 -= __bfloat162float(mid.y);
                o2[j] = pack_bf16x2(e0, e1);
              }
            }
            tmem_st_32x32(tS + t_lane + c * 32, o01);
            if (PL == 3) tmem_st_32x16(tPlo + t_lane + c * 16, o2);
            if (a.mbits != nullptr && q < a.T)
              a.mbits[(static_cast<int64_t>(item) * a.nC + c) * a.T + q] = keep;
          }
          tmem_st_wait();
        }
        xch[256 + half * 128 + r] = sum;
        tc_fence_before();
        mbar_arrive(bar_p);
        pair_sync(w);
        sum += xch[256 + (half ^ 1) * 128 + r];
        // ---- epilogue: O * inv_keep / sum -> PL planes
        mbar_wait(bar_o, ph, 0x621u);
        tc_fence_after();
        if (active) {
          tmem_ld_32x32(tO + t_lane + half * 32, sreg);
          tmem_ld_wait();
          if (q < a.T) {
            const float sc = inv_keep / sum;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(sreg[i]) * sc;
            __nv_bfloat16* dst = a.out + (static_cast<int64_t>(b) * a.T + q) * a.out_ld + h * 64 + half * 32;
#pragma unroll
            for (int p = 0; p < PL; ++p) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint32_t wv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __nv_bfloat162 qv = __floats2bfloat162_rn(v[g * 8 + 2 * i], v[g * 8 + 2 * i + 1]);
                  wv[i] = *reinterpret_cast<const uint32_t*>(&qv);
                  v[g * 8 + 2 * i] -= __bfloat162float(qv.x);
                  v[g * 8 + 2 * i + 1] -= __bfloat162float(qv.y);
                }
                *reinterpret_cast<uint4*>(dst + p * a.out_ps + g * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
              }
            }
            if (half == 0 && a.stats != nullptr) a.stats[static_cast<int64_t>(item) * a.T + q] = make_float2(mxs, sum);
          }
        }
        tc_fence_before();
        mbar_arrive(bar_oe);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attn_train_max_frames(int planes) { return planes == 3 ? AtfCfg<3>::KV_ROWS : AtfCfg<2>::KV_ROWS; }

template <int PL>
static int attn_train_fwd_launch(AttnTrainFwdArgs& a, const __nv_bfloat16* qkv, int64_t qkv_ps, cudaStream_t st) {
  using Cfg = AtfCfg<PL>;
  const int64_t ld = 3 * 64 * a.H;
  const int64_t dims[4] = {64, a.T, a.H, a.B};
  const int64_t strides[3] = {ld, 64, static_cast<int64_t>(a.T) * ld};
  for (int p = 0; p < PL; ++p) {
    const __nv_bfloat16* base = qkv + p * qkv_ps;
    SPK_TRY(encode_map_4d(&a.q_map[p], base, dims, strides, 128));
    SPK_TRY(encode_map_4d(&a.k_map[p], base + 64 * a.H, dims, strides, a.Tk64));
    SPK_TRY(encode_map_4d(&a.v_map[p], base + 2 * 64 * a.H, dims, strides, a.Tk64));
  }
  static PerDeviceOnce once;
  SPK_TRY(once.run([]() -> int {
    SPK_CUDA(cudaFuncSetAttribute(attn_train_fwd_kernel<PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    return 0;
  }));
  const int items = a.B * a.H;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  attn_train_fwd_kernel<PL><<<grid, ATF_THREADS, Cfg::SMEM, st>>>(a);
  SPK_CUDA(cudaGetLastError());
  return 0;
}

// qkv: [planes][B*T, 768] split tensor; out: [planes][B*T, 256]; stats [B*H*T] float2; mbits [B*H][ceil(Tp/32)][T]
int attn_train_fwd(const void* qkv, int64_t qkv_ps, int planes, void* out, int64_t out_ps, int64_t out_ld, float* stats,
                   uint32_t* mbits, DropCfg drop, uint32_t site, int B, int H, int T, int Tp, cudaStream_t st) {
  SPK_CHECK(planes == 2 || planes == 3, "attn_train_fwd: planes must be 2 or 3");
  SPK_CHECK(T >= 1 && T <= attn_train_max_frames(planes) && H >= 1, "attn_train_fwd: T=%d outside [1, %d]", T,
            attn_train_max_frames(planes));
  SPK_CHECK(drop.thresh == 0 || mbits != nullptr, "attn_train_fwd: dropout needs the keep-bit buffer");
  AttnTrainFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.H = H; a.T = T; a.Tp = Tp;
  a.Tk16 = (T + 15) / 16 * 16;
  a.Tk64 = (T + 63) / 64 * 64;
  a.mtiles = (T + 127) / 128;
  a.nC = (Tp + 31) / 32;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.out_ps = out_ps; a.out_ld = out_ld;
  a.stats = reinterpret_cast<float2*>(stats);
  a.mbits = drop.thresh != 0 ? mbits : nullptr;
  a.drop = drop; a.site = site;
  // algorithmic work: QK^T and PV (bf16 dense count); bytes: Q, K, V in, O out
  ProfScope prof("attn_train_fwd", 4.0 * B * H * T * T * 64, 4.0 * B * T * 64 * H * 2.0 * planes, st);
  if (planes == 3) return attn_train_fwd_launch<3>(a, reinterpret_cast<const __nv_bfloat16*>(qkv), qkv_ps, st);
  return attn_train_fwd_launch<2>(a, reinterpret_cast<const __nv_bfloat16*>(qkv), qkv_ps, st);
}


// =====================================================================================================================
// Backward.  Two bf16 planes everywhere (gradients are smooth in their inputs, DESIGN.md "precision").  Per (slice,
// head) the keys are walked in 128-row tiles i, the queries in 128-row tiles j; everything is computed TRANSPOSED (keys
// on the TMEM lanes) so that the two products contracted over queries take their A operand straight from tensor memory:
//
//   tcgen05   S^T  = K_i Q_j^T ,  dP^T = V_i dO_j^T                    -> TMEM (fp32)
//   8 warps   thread <-> key row: P^T = exp2(S^T * c - m_q) / l_q (row statistics saved by the forward), keep bit from
//             the forward's bit mask (32 x 32 bit transpose by warp shuffles), dS^T = P^T (keep dP^T/(1-p) - delta_q) / 8;
//             P_drop^T and dS^T go back into TMEM over S^T / dP^T as bf16 planes; dS^T also into shared memory in the
//             MN-major operand layout (probe: tools/probe/ts_probe.cu)
//   tcgen05   dV_i += P_drop^T dO_j  and  dK_i += dS^T Q_j   (A from TMEM, B MN-major from shared memory)
//             dQ_j += dS K_i                                  (A = dS from shared memory, MN-major)
//   8 warps   drain dV_i, dK_i after the last j, dQ_j after the last i -> dqkv planes; their column sums (the in-proj
//             bias gradient) stay in registers for the whole kernel (a CTA only ever sees one head) and are added to
//             global memory once per warp at the end.
// delta_q = sum_d dO_qd O_qd comes from attn_delta_kernel.  T <= 192: Q and dO of the whole slice stay resident.
constexpr int ATB_THREADS = 384;
constexpr int ATB_MAXT = 192;
constexpr int ATB_QROWS = ATB_MAXT * 128;                 // bytes per plane of Q / dO
constexpr int ATB_OFF_DO = 2 * ATB_QROWS;                 // 49152
constexpr int ATB_OFF_K = 4 * ATB_QROWS;                  // 98304
constexpr int ATB_OFF_V = ATB_OFF_K + 32768;              // 131072
constexpr int ATB_OFF_DS = ATB_OFF_V + 32768;             // 163840
constexpr int ATB_OFF_STAT = ATB_OFF_DS + 65536;          // 229376: float2 (m, 1/l) [192]
constexpr int ATB_OFF_DELTA = ATB_OFF_STAT + ATB_MAXT * 8;   // float [192]
constexpr int ATB_OFF_BAR = ATB_OFF_DELTA + ATB_MAXT * 4;    // 231680
constexpr int ATB_SMEM = ATB_OFF_BAR + 128;               // 231808 <= 232448
constexpr int ATB_T_ST = 0, ATB_T_DP = 128, ATB_T_DV = 256, ATB_T_DK = 320, ATB_T_DQ = 384;

struct AttnTrainBwdArgs {
  CUtensorMap q_map[2], k_map[2], v_map[2], do_map[2];
  int B, H, T, Tk16, Tk64, tiles, nC;
  const float2* stats;      // [B*H*T] from the forward
  const float* delta;       // [B*H*T]
  const uint32_t* mbits;    // [B*H][nC][T] or null (no dropout)
  __nv_bfloat16* dqkv;      // [2][B*T, 768]
  int64_t dqkv_ps;
  float* dbias;             // [768] += column sums of dQ | dK | dV
  float inv_keep;
};

// lane r holds row r of a 32 x 32 bit matrix (bit c = column c); afterwards lane c holds column c (bit r = row r)
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t o = __shfl_xor_sync(0xffffffffu, x, s);
    x = (lane & s) ? (((o >> s) & m) | (x & ~m)) : ((x & m) | ((o & m) << s));
  }
  return x;
}
// v[x] of lane r = element (r, x); returns for lane c the sum over r of element (r, c)
__device__ __forceinline__ float column_sums32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
    for (int k = 0; k < s; ++k) {
      const float send = (lane & s) ? v[k] : v[k + s];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[k] = ((lane & s) ? v[k + s] : v[k]) + recv;
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(ATB_THREADS, 1) attn_train_bwd_kernel(const __grid_constant__ AttnTrainBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  const uint32_t sQ = sbase, sdO = sbase + ATB_OFF_DO, sK = sbase + ATB_OFF_K, sV = sbase + ATB_OFF_V, sdS = sbase + ATB_OFF_DS;
  float2* sStat = reinterpret_cast<float2*>(smem_raw + ATB_OFF_STAT);
  float* sDelta = reinterpret_cast<float*>(smem_raw + ATB_OFF_DELTA);
  const uint32_t bar_qdo = sbase + ATB_OFF_BAR, bar_kv = bar_qdo + 8, bar_m1 = bar_qdo + 16, bar_p = bar_qdo + 24,
                 bar_tile = bar_qdo + 32, bar_kvd = bar_qdo + 40, bar_dqd = bar_qdo + 48, tmem_slot = bar_qdo + 56;
  // Every waiter of an mbarrier observes EVERY phase of it, in order: bar_m1 / bar_p advance once per (i, j) unit,
  // bar_kv / bar_tile / bar_kvd once per key tile, bar_qdo / bar_dqd once per slice.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((sbase & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      tma_prefetch_desc(&a.q_map[p]); tma_prefetch_desc(&a.k_map[p]); tma_prefetch_desc(&a.v_map[p]); tma_prefetch_desc(&a.do_map[p]);
    }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_qdo, 1); mbar_init(bar_kv, 1); mbar_init(bar_m1, 1); mbar_init(bar_p, 256); mbar_init(bar_tile, 1);
    mbar_init(bar_kvd, 256); mbar_init(bar_dqd, 256);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const int items = a.B * a.H;
  const int tiles = a.tiles;                              // key tiles == query tiles == ceil(T / 128)
  auto nq_of = [&](int j) { const int n = a.Tk16 - 128 * j; return n < 128 ? n : 128; };   // multiple of 16

  if (warp == 0) {
    if (lane == 0) {
      uint32_t kt = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int h = item % a.H, b = item / a.H;
        for (int i = 0; i < tiles; ++i, ++kt) {
          if (kt > 0) mbar_wait(bar_tile, (kt - 1) & 1u, 0x700u);   // every MMA that read the buffers loaded next has retired
          if (i == 0) {
            mbar_arrive_expect_tx(bar_qdo, 4u * a.Tk64 * 128);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              tma_load_4d(sQ + p * ATB_QROWS, &a.q_map[p], bar_qdo, 0, 0, h, b);
              tma_load_4d(sdO + p * ATB_QROWS, &a.do_map[p], bar_qdo, 0, 0, h, b);
            }
          }
          mbar_arrive_expect_tx(bar_kv, 4u * 16384);
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            tma_load_4d(sK + p * 16384, &a.k_map[p], bar_kv, 0, i * 128, h, b);
            tma_load_4d(sV + p * 16384, &a.v_map[p], bar_kv, 0, i * 128, h, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr int PA[3] = {1, 0, 0}, PB[3] = {0, 1, 0};          // lo*hi, hi*lo, hi*hi
      const uint32_t tST = tmem_base + ATB_T_ST, tDP = tmem_base + ATB_T_DP, tdV = tmem_base + ATB_T_DV,
                     tdK = tmem_base + ATB_T_DK;
      const uint32_t idesc_acc = umma_idesc_bf16(128, 64, false, true);      // A from TMEM, B MN-major
      const uint32_t idesc_dq = umma_idesc_bf16(128, 64, true, true);        // A and B MN-major from shared memory
      uint32_t u = 0, kt = 0, it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        mbar_wait(bar_qdo, it & 1u, 0x710u);
        for (int i = 0; i < tiles; ++i, ++kt) {
          mbar_wait(bar_kv, kt & 1u, 0x711u);
          const int nk = (nq_of(i)) / 16;                          // key steps of this tile that hold frames
          for (int j = 0; j < tiles; ++j, ++u) {
            const int nq = nq_of(j);
            const uint32_t idesc_s = umma_idesc_bf16(128, nq, false, false);
            tc_fence_after();
            // ---- S^T = K_i Q_j^T, dP^T = V_i dO_j^T
#pragma unroll
            for (int which = 0; which < 2; ++which) {
              const uint32_t sa = which ? sV : sK, sb = (which ? sdO : sQ) + j * 16384, td = which ? tDP : tST;
              uint32_t acc = 0;
#pragma unroll
              for (int cb = 0; cb < 3; ++cb) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(td, umma_smem_desc(sa + PA[cb] * 16384 + k * 32, 16, 1024),
                            umma_smem_desc(sb + PB[cb] * ATB_QROWS + k * 32, 16, 1024), idesc_s, acc);
                  acc = 1;
                }
              }
            }
            umma_commit(bar_m1);
            mbar_wait(bar_p, u & 1u, 0x712u);
            if (j == 0 && kt > 0) mbar_wait(bar_kvd, (kt - 1) & 1u, 0x713u);            // dV / dK of the previous key tile drained
            if (i == 0 && j == 0 && it > 0) mbar_wait(bar_dqd, (it - 1) & 1u, 0x714u);  // dQ of the previous item drained
            tc_fence_after();
            // ---- dV_i += P_drop^T dO_j ; dK_i += dS^T Q_j     (contraction over the nq queries of tile j)
#pragma unroll
            for (int which = 0; which < 2; ++which) {
              const uint32_t ta = which ? tDP : tST, sb = (which ? sQ : sdO) + j * 16384, td = which ? tdK : tdV;
              for (int t = 0; t < nq / 16; ++t) {
#pragma unroll
                for (int cb = 0; cb < 3; ++cb) {
                  umma_bf16_ts(td, ta + 16 * t + 8 * PA[cb], umma_smem_desc(sb + PB[cb] * ATB_QROWS + t * 2048, 8192, 1024),
                               idesc_acc, (j > 0 || t > 0 || cb > 0) ? 1u : 0u);
                }
              }
            }
            // ---- dQ_j += dS_ij K_i     (contraction over the keys of tile i)
            const uint32_t tdQ = tmem_base + ATB_T_DQ + 64 * j;
            for (int t = 0; t < nk; ++t) {
#pragma unroll
              for (int cb = 0; cb < 3; ++cb) {
                umma_bf16(tdQ, umma_smem_desc(sdS + PA[cb] * 32768 + (t >> 2) * 16384 + (t & 3) * 2048, 8192, 1024),
                          umma_smem_desc(sK + PB[cb] * 16384 + t * 2048, 8192, 1024), idesc_dq,
                          (i > 0 || t > 0 || cb > 0) ? 1u : 0u);
              }
            }
            if (j == tiles - 1) umma_commit(bar_tile);       // dV_i / dK_i (and, after the last tile, dQ) are complete
          }
        }
      }
    }
  } else if (warp >= 4) {
    const int w = (warp - 4) & 3, half = (warp - 4) >> 2;
    const int rr = w * 32 + lane;
    const int cw = threadIdx.x - 128;                              // 0..255 among the compute threads
    const uint32_t t_lane = static_cast<uint32_t>(w * 32) << 16;
    const uint32_t tST = tmem_base + ATB_T_ST + t_lane, tDP = tmem_base + ATB_T_DP + t_lane;
    const int h = static_cast<int>(blockIdx.x) % a.H;              // gridDim.x % H == 0: one head per CTA
    float acc_dq = 0.f, acc_dk = 0.f, acc_dv = 0.f;                // bias-gradient column (half * 32 + lane) of this head
    uint32_t u = 0, kt = 0, it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int b = item / a.H;
      // ---- per-query statistics of this slice and head
      asm volatile("bar.sync 5, 256;" ::: "memory");               // everybody is done with the previous item's statistics
      if (cw < ATB_MAXT) {
        float2 st = make_float2(0.f, 0.f);
        float dl = 0.f;
        if (cw < a.T) {
          st = a.stats[static_cast<int64_t>(item) * a.T + cw];
          st.y = 1.f / st.y;
          dl = a.delta[static_cast<int64_t>(item) * a.T + cw];
        }
        sStat[cw] = st;
        sDelta[cw] = dl;
      }
      asm volatile("bar.sync 5, 256;" ::: "memory");
      for (int i = 0; i < tiles; ++i, ++kt) {
        const int key = 128 * i + rr;
        const bool kvalid = key < a.T;
        const bool warp_keys = 128 * i + w * 32 < a.Tk16;          // warp-uniform: some key step of this warp is contracted
        for (int j = 0; j < tiles; ++j, ++u) {
          const int nq = nq_of(j);
          mbar_wait(bar_m1, u & 1u, 0x720u);
          tc_fence_after();
          if (warp_keys) {
            for (int c = half; c * 32 < nq; c += 2) {
              const int q0 = 128 * j + 32 * c;
              uint32_t s_[32], d_[32];
              tmem_ld_32x32(tST + c * 32, s_);
              tmem_ld_32x32(tDP + c * 32, d_);
              uint32_t word = 0xFFFFFFFFu;
              if (a.mbits != nullptr) {
                const int kc = (128 * i + w * 32) >> 5;
                uint32_t mine = 0;
                if (q0 + lane < a.T && kc < a.nC) mine = __ldg(a.mbits + (static_cast<int64_t>(item) * a.nC + kc) * a.T + q0 + lane);
                word = transpose32(mine, lane);                    // bit x = keep(query q0 + x, this key)
              }
              tmem_ld_wait();
              uint32_t op[32], ods[32];
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) {
                float pd[2], ds[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int x = 2 * jj + e;
                  const int qi = q0 + x >= ATB_MAXT ? 0 : q0 + x;       // past the slice: value unused (ok == false)
                  const float2 st = sStat[qi];
                  const float dl = sDelta[qi];
                  const bool ok = kvalid && (q0 + x < a.T);
                  const bool kp = (word >> x) & 1u;
                  const float p = at_exp2(__uint_as_float(s_[x]) * ATF_SC - st.x) * st.y;
                  const float dpm = kp ? __uint_as_float(d_[x]) * a.inv_keep : 0.f;
                  pd[e] = (ok && kp) ? p * a.inv_keep : 0.f;
                  ds[e] = ok ? p * (dpm - dl) * 0.125f : 0.f;
                }
                const int slot = (jj >> 3) * 16 + (jj & 7);
                const __nv_bfloat162 ph = __floats2bfloat162_rn(pd[0], pd[1]);
                op[slot] = *reinterpret_cast<const uint32_t*>(&ph);
                op[slot + 8] = pack_bf16x2(pd[0] - __bfloat162float(ph.x), pd[1] - __bfloat162float(ph.y));
                const __nv_bfloat162 dh = __floats2bfloat162_rn(ds[0], ds[1]);
                ods[slot] = *reinterpret_cast<const uint32_t*>(&dh);
                ods[slot + 8] = pack_bf16x2(ds[0] - __bfloat162float(dh.x), ds[1] - __bfloat162float(dh.y));
              }
              tmem_st_32x32(tST + c * 32, op);
              tmem_st_32x32(tDP + c * 32, ods);
              // dS^T -> shared memory, MN-major operand layout: row = key, 64 queries per 128-byte row
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int qq = 32 * c + 8 * g;                     // query inside the tile
                const uint32_t dst = sdS + (rr >> 6) * 16384 + (qq >> 6) * 8192 + (rr & 63) * 128 +
                                     ((((qq & 63) >> 3) ^ (rr & 7)) << 4);
                // slots of queries 8g .. 8g+7: key step (g >> 1), pairs (g & 1) * 4 .. +3
                const int s0 = (g >> 1) * 16 + (g & 1) * 4;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(ods[s0]), "r"(ods[s0 + 1]),
                             "r"(ods[s0 + 2]), "r"(ods[s0 + 3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 32768), "r"(ods[s0 + 8]), "r"(ods[s0 + 9]),
                             "r"(ods[s0 + 10]), "r"(ods[s0 + 11]) : "memory");
              }
            }
            tmem_st_wait();
          }
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(bar_p);
        }
        // ---- dV_i, dK_i complete (last query tile's MMAs retired)
        mbar_wait(bar_tile, kt & 1u, 0x721u);
        tc_fence_after();
        if (warp_keys) {
#pragma unroll
          for (int which = 0; which < 2; ++which) {                // 0: dV, 1: dK
            uint32_t acc[32];
            tmem_ld_32x32(tmem_base + t_lane + (which ? ATB_T_DK : ATB_T_DV) + half * 32, acc);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int x = 0; x < 32; ++x) v[x] = kvalid ? __uint_as_float(acc[x]) : 0.f;
            if (kvalid) {
              __nv_bfloat16* dst = a.dqkv + (static_cast<int64_t>(b) * a.T + key) * (3 * 64 * a.H) + (which ? 1 : 2) * 64 * a.H +
                                   h * 64 + half * 32;
#pragma unroll
              for (int p = 0; p < 2; ++p) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  uint32_t wv[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float x0 = v[g * 8 + 2 * e], x1 = v[g * 8 + 2 * e + 1];
                    if (p == 0) {
                      wv[e] = pack_bf16x2(x0, x1);
                    } else {
                      const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                      wv[e] = pack_bf16x2(x0 - __bfloat162float(hh.x), x1 - __bfloat162float(hh.y));
                    }
                  }
                  *reinterpret_cast<uint4*>(dst + p * a.dqkv_ps + g * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                }
              }
            }
            const float cs = column_sums32(v, lane);
            if (which) acc_dk += cs; else acc_dv += cs;
          }
        }
        tc_fence_before();
        mbar_arrive(bar_kvd);
      }
      // ---- dQ of the whole slice (thread <-> query row)
      for (int j = 0; j < tiles; ++j) {
        const int q = 128 * j + rr;
        if (128 * j + w * 32 < a.Tk16) {                           // warp-uniform
          uint32_t acc[32];
          tmem_ld_32x32(tmem_base + t_lane + ATB_T_DQ + 64 * j + half * 32, acc);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int x = 0; x < 32; ++x) v[x] = q < a.T ? __uint_as_float(acc[x]) : 0.f;
          if (q < a.T) {
            __nv_bfloat16* dst = a.dqkv + (static_cast<int64_t>(b) * a.T + q) * (3 * 64 * a.H) + h * 64 + half * 32;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint32_t wv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float x0 = v[g * 8 + 2 * e], x1 = v[g * 8 + 2 * e + 1];
                  if (p == 0) {
                    wv[e] = pack_bf16x2(x0, x1);
                  } else {
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                    wv[e] = pack_bf16x2(x0 - __bfloat162float(hh.x), x1 - __bfloat162float(hh.y));
                  }
                }
                *reinterpret_cast<uint4*>(dst + p * a.dqkv_ps + g * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
              }
            }
          }
          acc_dq += column_sums32(v, lane);
        }
      }
      tc_fence_before();
      mbar_arrive(bar_dqd);
    }
    // ---- in-proj bias gradient: one atomic per warp and tensor for the whole kernel
    if (blockIdx.x < items) {
      const int col = h * 64 + half * 32 + lane;
      atomicAdd(a.dbias + col, acc_dq);
      atomicAdd(a.dbias + 64 * a.H + col, acc_dk);
      atomicAdd(a.dbias + 2 * 64 * a.H + col, acc_dv);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[(b*H + h)*T + t] = sum_d dO[b*T + t, h*64 + d] * O[b*T + t, h*64 + d]      (two planes of each; H = 4)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ dout, int64_t do_ps,
                                                         const __nv_bfloat16* __restrict__ out, int64_t o_ps,
                                                         float* __restrict__ delta, int64_t tokens, int T) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t tok = warp; tok < tokens; tok += nwarps) {
    float x[8], y[8];
    load8_split(dout, do_ps, 2, tok * 256 + lane * 8, x);
    load8_split(out, o_ps, 2, tok * 256 + lane * 8, y);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(x[i], y[i], s);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if ((lane & 7) == 0) {
      const int64_t b = tok / T, t = tok % T;
      delta[(b * 4 + (lane >> 3)) * T + t] = s;
    }
  }
}

// qkv: forward stash [>=2 planes][B*T, 768]; out: forward attention output [>=2 planes][B*T, 256]; dout: its gradient
// [2 planes][B*T, 256]; dqkv: [2 planes][B*T, 768] (every element of the Q | K | V columns is written); dbias [768] +=.
int attn_train_bwd(const void* qkv, int64_t qkv_ps, const void* out, int64_t out_ps, const void* dout, int64_t dout_ps,
                   const float* stats, const uint32_t* mbits, float* delta, void* dqkv, int64_t dqkv_ps, float* dbias,
                   DropCfg drop, int B, int H, int T, int Tp, cudaStream_t st) {
  SPK_CHECK(H == 4, "attn_train_bwd: 4 heads of 64");
  SPK_CHECK(T >= 1 && T <= ATB_MAXT, "attn_train_bwd: T=%d outside [1, %d]", T, ATB_MAXT);
  SPK_CHECK(drop.thresh == 0 || mbits != nullptr, "attn_train_bwd: dropout needs the forward's keep bits");
  const int64_t tokens = static_cast<int64_t>(B) * T;
  {
    ProfScope prof("attn_delta", 0, 2.0 * tokens * 256 * 2 * 2 + 4.0 * tokens * H, st);
    const int blocks = static_cast<int>(std::min<int64_t>((tokens + 7) / 8, 148 * 8));
    attn_delta_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dout), dout_ps,
                                              reinterpret_cast<const __nv_bfloat16*>(out), out_ps, delta, tokens, T);
    SPK_CUDA(cudaGetLastError());
  }
  AttnTrainBwdArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.H = H; a.T = T;
  a.Tk16 = (T + 15) / 16 * 16;
  a.Tk64 = (T + 63) / 64 * 64;
  a.tiles = (T + 127) / 128;
  a.nC = (Tp + 31) / 32;
  a.stats = reinterpret_cast<const float2*>(stats);
  a.delta = delta;
  a.mbits = drop.thresh != 0 ? mbits : nullptr;
  a.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  a.dqkv_ps = dqkv_ps;
  a.dbias = dbias;
  a.inv_keep = drop.thresh != 0 ? drop.inv_keep : 1.f;
  const int64_t ld = 3 * 64 * H;
  const int64_t dims[4] = {64, T, H, B};
  const int64_t strides[3] = {ld, 64, static_cast<int64_t>(T) * ld};
  const int64_t ostrides[3] = {64 * H, 64, static_cast<int64_t>(T) * 64 * H};
  for (int p = 0; p < 2; ++p) {
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(qkv) + p * qkv_ps;
    SPK_TRY(encode_map_4d(&a.q_map[p], base, dims, strides, a.Tk64));
    SPK_TRY(encode_map_4d(&a.k_map[p], base + 64 * H, dims, strides, 128));
    SPK_TRY(encode_map_4d(&a.v_map[p], base + 2 * 64 * H, dims, strides, 128));
    SPK_TRY(encode_map_4d(&a.do_map[p], reinterpret_cast<const __nv_bfloat16*>(dout) + p * dout_ps, dims, ostrides, a.Tk64));
  }
  static PerDeviceOnce once;
  SPK_TRY(once.run([]() -> int {
    SPK_CUDA(cudaFuncSetAttribute(attn_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB_SMEM));
    return 0;
  }));
  const int items = B * H;
  int grid = items < device_sm_count() ? items : device_sm_count();
  grid -= grid % H;                       // a CTA must only ever see one head (register-resident bias-gradient sums)
  SPK_CHECK(grid >= H, "attn_train_bwd: no CTAs");
  // algorithmic work: dV, dP, dQ, dK (the recomputed S is not credited); bytes: Q, K, V, O, dO in, dQ, dK, dV out
  ProfScope prof("attn_train_bwd", 8.0 * B * H * T * T * 64, (5.0 + 3.0) * B * T * 64 * H * 2.0 * 2, st);
  attn_train_bwd_kernel<<<grid, ATB_THREADS, ATB_SMEM, st>>>(a);
  SPK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace spk
