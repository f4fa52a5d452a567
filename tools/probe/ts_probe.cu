// Probe of two tcgen05 operand paths the fused training attention relies on (run once on a B200; prints mismatch counts):
//   1. A operand from TMEM (tcgen05.mma [d], [a_tmem], b_desc): bf16 A written by tcgen05.st.32x32b from registers,
//      thread <-> row (TMEM lane), two K elements per 32-bit column -- which half holds the even element?
//   2. A operand from shared memory, MN-major, in the 128-byte-swizzled layout written by threads (row = k index,
//      64 M elements per 128-byte row, 16-byte chunks XOR-swizzled with (k & 7), 64-row k blocks of 8 KB per 64-wide
//      M chunk) -- the layout TMA produces for the weight-gradient GEMMs, here produced by st.shared.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../speaker_embedding_torch_b200/csrc ts_probe.cu -o ts_probe
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx.cuh"

using namespace spk;

__device__ __forceinline__ float av(int r, int k) { return static_cast<float>(((r * 7 + k * 3) % 17) - 8); }
__device__ __forceinline__ float bv(int n, int k) { return static_cast<float>(((n * 5 + k * 11) % 13) - 6); }

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ uint32_t pack2h(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// mode 0: TS, even K element in the low half; mode 1: TS, even K element in the high half; mode 2: SS with MN-major A;
// mode 3: SS, A bf16 (K-major) x B fp16 (mixed operand formats in one kind::f16 MMA); mode 4: TS, A fp16 in TMEM x B bf16
__global__ void __launch_bounds__(128, 1) probe_kernel(int mode, int* mism, float* sample) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB = sbase;                 // [2 k-blocks][64 n rows][128 B]  K-major
  const uint32_t sA = sbase + 16384;         // mode 2: [2 k-blocks][2 M chunks][64 k rows][128 B]  MN-major
  const uint32_t bar = sbase + 16384 + 32768, tmem_slot = bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
  const int K = (mode == 2) ? 128 : 64;
  const float want_scale = mode == 3 ? 0.5f : (mode == 4 ? 0.25f : 1.f);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t t_lane = static_cast<uint32_t>(warp * 32) << 16;

  // B[n][k], K-major swizzled: thread n < 64 writes its row(s)
  if (r < 64) {
    for (int kb = 0; kb < K / 64; ++kb)
      for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
        for (int j = 0; j < 4; ++j)
          w[j] = (mode == 3) ? pack2h(0.5f * bv(r, kb * 64 + c * 8 + 2 * j), 0.5f * bv(r, kb * 64 + c * 8 + 2 * j + 1))
                             : pack2(bv(r, kb * 64 + c * 8 + 2 * j), bv(r, kb * 64 + c * 8 + 2 * j + 1));
        const uint32_t dst = sB + kb * 8192 + r * 128 + ((c ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      }
  }
  if (mode < 2 || mode == 4) {
    uint32_t regs[32];
    for (int c = 0; c < 32; ++c)
      regs[c] = (mode == 4) ? pack2h(0.25f * av(r, 2 * c), 0.25f * av(r, 2 * c + 1))
                            : ((mode == 0) ? pack2(av(r, 2 * c), av(r, 2 * c + 1)) : pack2(av(r, 2 * c + 1), av(r, 2 * c)));
    tmem_st_32x32(tmem_base + t_lane + 64, regs);
    tmem_st_wait();
  } else if (mode == 3) {
    // A[r][k] bf16, K-major swizzled, one 64-wide k block: thread r writes its row
    for (int c = 0; c < 8; ++c) {
      uint32_t w[4];
      for (int j = 0; j < 4; ++j) w[j] = pack2(av(r, c * 8 + 2 * j), av(r, c * 8 + 2 * j + 1));
      const uint32_t dst = sA + r * 128 + ((c ^ (r & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    }
  } else {
    // A^T: thread kk (= k index, 128 of them) writes the 128 M values of its k row
    const int kk = r;
    for (int mc = 0; mc < 2; ++mc)
      for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
        for (int j = 0; j < 4; ++j) w[j] = pack2(av(mc * 64 + c * 8 + 2 * j, kk), av(mc * 64 + c * 8 + 2 * j + 1, kk));
        const uint32_t dst = sA + (kk >> 6) * 16384 + mc * 8192 + (kk & 63) * 128 + ((c ^ (kk & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    if (mode < 2 || mode == 4) {
      // mode 4: A format field [7,10) = 0 (F16), B stays BF16
      const uint32_t idesc = umma_idesc_bf16(128, 64, false, false) & ~(mode == 4 ? (7u << 7) : 0u);
      for (int k = 0; k < 4; ++k)
        umma_bf16_ts(tmem_base, tmem_base + 64 + k * 8, umma_smem_desc(sB + k * 32, 16, 1024), idesc, k > 0);
    } else if (mode == 3) {
      const uint32_t idesc = umma_idesc_bf16(128, 64, false, false) & ~(7u << 10);   // B format = F16
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base, umma_smem_desc(sA + k * 32, 16, 1024), umma_smem_desc(sB + k * 32, 16, 1024), idesc, k > 0);
    } else {
      const uint32_t idesc = umma_idesc_bf16(128, 64, true, false);
      for (int kb = 0; kb < 2; ++kb)
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, umma_smem_desc(sA + kb * 16384 + k * 2048, 8192, 1024),
                    umma_smem_desc(sB + kb * 8192 + k * 32, 16, 1024), idesc, (kb | k) > 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0, 0x777u);
  tc_fence_after();
  int bad = 0;
  for (int c = 0; c < 2; ++c) {
    uint32_t d[32];
    tmem_ld_32x32(tmem_base + t_lane + c * 32, d);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) {
      const int n = c * 32 + i;
      float want = 0.f;
      for (int k = 0; k < K; ++k) want += av(r, k) * bv(n, k);
      want *= want_scale;
      if (__uint_as_float(d[i]) != want) ++bad;
      if (r == 5 && n == 9) { sample[0] = __uint_as_float(d[i]); sample[1] = want; }
    }
  }
  atomicAdd(mism, bad);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

int main() {
  int* mism; float* sample;
  cudaMalloc(&mism, 4); cudaMalloc(&sample, 8);
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[5] = {"TS  A-in-TMEM, even k in LOW  half", "TS  A-in-TMEM, even k in HIGH half", "SS  MN-major A written by threads  ",
                          "SS  A bf16 x B fp16 (mixed formats)  ", "TS  A fp16 in TMEM x B bf16        "};
  for (int mode = 0; mode < 5; ++mode) {
    cudaMemset(mism, 0, 4);
    probe_kernel<<<1, 128, smem>>>(mode, mism, sample);
    cudaError_t e = cudaDeviceSynchronize();
    int h = -1; float hs[2] = {0, 0};
    cudaMemcpy(&h, mism, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs, sample, 8, cudaMemcpyDeviceToHost);
    printf("probe mode %d (%s): %s, mismatches %d of 8192, sample got %.1f want %.1f\n", mode, names[mode],
           cudaGetErrorString(e), h, hs[0], hs[1]);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
