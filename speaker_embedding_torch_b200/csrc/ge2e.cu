// Fused GE2E loss: one kernel, three stream-ordered launches (stages), computes centroids, row/centroid normalisation, the
// N*M x N cosine-similarity matrix, w*S - b, log-softmax cross-entropy, and its gradient
// (dE, dw, db) without ever materialising the [N*M, N] logits or the reference's two
// [N*M, N, D] expanded operands (/root/reference/Modules.py:121-156; math: SURVEY.md App. B).
//
//   phase 1  (block per speaker)   c_k = mean_m E ; ehat scale 1/max(|E_i|,eps) ; chat_k ; zero dchat
//   phase 2  (block per 16 rows)   pass A: S tile -> online (max, sum) -> lse, loss
//                                  pass B: recompute S tile -> G = (w/NM)(softmax - onehot)
//                                          dehat += G * Chat   (registers)
//                                          dchat += G^T * Ehat (coalesced fp32 RED)
//                                  dE_i  = (dehat - (dehat.ehat) ehat) / |E_i|
//   phase 3  (block per speaker)   dE_i += (1/M) (dchat_j - (dchat_j.chat_j) chat_j) / |c_j|
//
// fp32 SIMT math throughout (bit-for-bit deterministic except the fp32 atomics of dchat/loss).
// D is fixed at 256 (= Embedding_Size of the reference hyper-parameters): thread <-> column.
#include "common.cuh"
#include "ge2e.h"
#include "ptx.cuh"

namespace spk {

constexpr int GD = 256;        // embedding size
constexpr int TR = 16;         // rows per block tile
constexpr int TC = 64;         // centroids per tile
constexpr int CPAD = GD + 4;   // padded centroid row (floats), keeps 16-B alignment, breaks bank aliasing

struct Ge2eSmem {
  float e[TR][GD];         // normalised rows
  float c[TC][CPAD];       // normalised centroids tile
  float g[TR][TC];         // G tile, row-major (for dchat)
  float gt[TC][TR];        // G tile, transposed (for dehat)
  float red[8][TR];
  float rowv[TR];
};

// Programmatic dependent launch (pdl_trigger / pdl_wait, ptx.cuh): the three stages are launched back to back with the
// stream-serialisation attribute; a stage touches its predecessor's results only after pdl_wait().
__device__ __forceinline__ float block_sum_256(float v, float* red8) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red8[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red8[w];
  return s;
}

__global__ void __launch_bounds__(256, 2)
ge2e_fused_kernel(const float* __restrict__ E, int N, int M, const float* __restrict__ w_ptr,
                  const float* __restrict__ b_ptr, float* __restrict__ loss, float* __restrict__ dE,
                  float* __restrict__ dw, float* __restrict__ db, float* __restrict__ chat,
                  float* __restrict__ cinv, float* __restrict__ einv, float* __restrict__ dchat, int need_grad,
                  float eps, int phase) {
  // `phase` selects one of the three stages; the launcher enqueues them back to back on one stream (stream order is the
  // grid-wide barrier: a cooperative launch with two grid.sync() cost 32 us at the training size, mostly in the syncs)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Ge2eSmem& sm = *reinterpret_cast<Ge2eSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t NM = static_cast<int64_t>(N) * M;
  const float w = __ldg(w_ptr), b = __ldg(b_ptr);
  float* red8 = &sm.red[0][0];

  // ------------------------------------------------------------------ phase 1
  if (phase == 1) {
  pdl_trigger();
  if (blockIdx.x == 0 && tid == 0) {
    loss[0] = 0.f;
    if (need_grad) { dw[0] = 0.f; db[0] = 0.f; }
  }
  for (int k = blockIdx.x; k < N; k += gridDim.x) {
    // warp per row: row norm + partial centroid (lane owns columns lane*8 .. +7)
    float cpart[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int m = warp; m < M; m += 8) {
      const float* row = E + (static_cast<int64_t>(k) * M + m) * GD + lane * 8;
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(row));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(row + 4));
      const float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { q += v[i] * v[i]; cpart[i] += v[i]; }
      q = warp_sum(q);
      if (lane == 0) einv[static_cast<int64_t>(k) * M + m] = 1.f / fmaxf(sqrtf(q), eps);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) sm.e[warp][lane * 8 + i] = cpart[i];
    __syncthreads();
    float cv = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) cv += sm.e[wv][tid];
    cv *= 1.f / static_cast<float>(M);
    const float nrm = fmaxf(sqrtf(block_sum_256(cv * cv, red8)), eps);
    chat[static_cast<int64_t>(k) * GD + tid] = cv / nrm;
    if (need_grad) dchat[static_cast<int64_t>(k) * GD + tid] = 0.f;
    if (tid == 0) cinv[k] = 1.f / nrm;
    __syncthreads();
  }
  return;
  }

  // ------------------------------------------------------------------ phase 2
  if (phase == 2) {
  const int r_loc = tid >> 4;      // 0..15 : row within tile (S-tile compute mapping)
  const int cgp = tid & 15;        // columns cgp, cgp+16, cgp+32, cgp+48 of the centroid tile
  const int64_t row_tiles = (NM + TR - 1) / TR;
  const int col_tiles = (N + TC - 1) / TC;
  const float inv_nm = 1.f / static_cast<float>(NM);

  for (int64_t rt = blockIdx.x; rt < row_tiles; rt += gridDim.x) {
    const int64_t row0 = rt * TR;
    __syncthreads();
    // load + normalise 16 rows: warp handles rows warp, warp+8
    for (int r = warp; r < TR; r += 8) {
      const int64_t gi = row0 + r;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (gi < NM) {
        const float sc = einv[gi];
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(E + gi * GD + lane * 8));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(E + gi * GD + lane * 8 + 4));
        v[0] = a0.x * sc; v[1] = a0.y * sc; v[2] = a0.z * sc; v[3] = a0.w * sc;
        v[4] = a1.x * sc; v[5] = a1.y * sc; v[6] = a1.z * sc; v[7] = a1.w * sc;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sm.e[r][lane * 8 + i] = v[i];
    }
    const int64_t my_row = row0 + r_loc;
    const int my_label = (my_row < NM) ? static_cast<int>(my_row / M) : -1;
    float run_max = -INFINITY, run_sum = 0.f, z_true = 0.f;
    float lse = 0.f;
    float dacc[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) dacc[r] = 0.f;
    float dw_part = 0.f, db_part = 0.f;

    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int pass = 0; pass < (need_grad ? 2 : 1); ++pass) {
      for (int ct = 0; ct < col_tiles; ++ct) {
        const int c0 = ct * TC;
        // N <= 64 (the training configuration): one centroid tile -- the second sweep reuses the tile in shared
        // memory and the S values still held in registers instead of recomputing them.
        const bool reuse = (pass == 1 && col_tiles == 1);
        if (!reuse) {
          __syncthreads();
          // centroid tile -> smem (coalesced float4)
          for (int i = tid; i < TC * (GD / 4); i += 256) {
            const int cr = i / (GD / 4), c4 = i % (GD / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + cr < N) v = __ldg(reinterpret_cast<const float4*>(chat + static_cast<int64_t>(c0 + cr) * GD) + c4);
            *reinterpret_cast<float4*>(&sm.c[cr][c4 * 4]) = v;
          }
          __syncthreads();
          // S tile: 4 dot products of length 256 per thread
          s[0] = s[1] = s[2] = s[3] = 0.f;
#pragma unroll 4
          for (int d = 0; d < GD; d += 4) {
            const float4 ev = *reinterpret_cast<const float4*>(&sm.e[r_loc][d]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 cvv = *reinterpret_cast<const float4*>(&sm.c[cgp + 16 * j][d]);
              s[j] = fmaf(ev.x, cvv.x, s[j]);
              s[j] = fmaf(ev.y, cvv.y, s[j]);
              s[j] = fmaf(ev.z, cvv.z, s[j]);
              s[j] = fmaf(ev.w, cvv.w, s[j]);
            }
          }
        }
        if (pass == 0) {
          float tmax = -INFINITY;
          float z[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = c0 + cgp + 16 * j;
            z[j] = (col < N) ? w * s[j] - b : -INFINITY;
            if (col == my_label) z_true = z[j];
            tmax = fmaxf(tmax, z[j]);
          }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
          const float new_max = fmaxf(run_max, tmax);
          float part = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) part += (z[j] == -INFINITY) ? 0.f : expf(z[j] - new_max);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          run_sum = run_sum * ((run_max == -INFINITY) ? 0.f : expf(run_max - new_max)) + part;
          run_max = new_max;
        } else {
          // G tile
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cl = cgp + 16 * j, col = c0 + cl;
            float gval = 0.f;
            if (col < N && my_row < NM) {
              const float p = expf(w * s[j] - b - lse);
              const float pm = (p - (col == my_label ? 1.f : 0.f)) * inv_nm;
              dw_part += pm * s[j];
              db_part -= pm;
              gval = w * pm;
            }
            sm.g[r_loc][cl] = gval;
            sm.gt[cl][r_loc] = gval;
          }
          __syncthreads();
          // dehat[r][tid] += sum_c G[r][c] * chat[c][tid]
#pragma unroll 4
          for (int c = 0; c < TC; ++c) {
            const float cvl = sm.c[c][tid];
#pragma unroll
            for (int r4 = 0; r4 < TR; r4 += 4) {
              const float4 gv = *reinterpret_cast<const float4*>(&sm.gt[c][r4]);
              dacc[r4 + 0] = fmaf(gv.x, cvl, dacc[r4 + 0]);
              dacc[r4 + 1] = fmaf(gv.y, cvl, dacc[r4 + 1]);
              dacc[r4 + 2] = fmaf(gv.z, cvl, dacc[r4 + 2]);
              dacc[r4 + 3] = fmaf(gv.w, cvl, dacc[r4 + 3]);
            }
          }
          // dchat[c][tid] += sum_r G[r][c] * ehat[r][tid]   (16 centroids at a time to bound registers)
#pragma unroll 1
          for (int cb = 0; cb < TC; cb += 16) {
            if (c0 + cb >= N) break;
            float a2[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) a2[i] = 0.f;
#pragma unroll
            for (int r = 0; r < TR; ++r) {
              const float evl = sm.e[r][tid];
#pragma unroll
              for (int c4 = 0; c4 < 16; c4 += 4) {
                const float4 gv = *reinterpret_cast<const float4*>(&sm.g[r][cb + c4]);
                a2[c4 + 0] = fmaf(gv.x, evl, a2[c4 + 0]);
                a2[c4 + 1] = fmaf(gv.y, evl, a2[c4 + 1]);
                a2[c4 + 2] = fmaf(gv.z, evl, a2[c4 + 2]);
                a2[c4 + 3] = fmaf(gv.w, evl, a2[c4 + 3]);
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + cb + i < N) atomicAdd(dchat + static_cast<int64_t>(c0 + cb + i) * GD + tid, a2[i]);
          }
        }
      }
      if (pass == 0) {
        lse = run_max + logf(run_sum);
        // loss contribution: one thread per row (cgp == 0 holds the reduced values; z_true lives in one lane)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) z_true += __shfl_xor_sync(0xffffffffu, z_true, o);
        float lrow = (cgp == 0 && my_row < NM) ? (lse - z_true) : 0.f;
        const float lsum = block_sum_256(lrow, red8);
        if (tid == 0) atomicAdd(loss, lsum * inv_nm);
      }
    }

    if (need_grad) {
      const float dws = block_sum_256(dw_part, red8);
      const float dbs = block_sum_256(db_part, red8);
      if (tid == 0) { atomicAdd(dw, dws); atomicAdd(db, dbs); }
      // project out the radial component and write the row part of dE
      __syncthreads();
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        float dot = warp_sum(dacc[r] * sm.e[r][tid]);
        if (lane == 0) sm.red[warp][r] = dot;
      }
      __syncthreads();
      if (tid < TR) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) s += sm.red[wv][tid];
        sm.rowv[tid] = s;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        const int64_t gi = row0 + r;
        if (gi < NM) dE[gi * GD + tid] = (dacc[r] - sm.rowv[r] * sm.e[r][tid]) * einv[gi];
      }
    }
  }
  return;
  }

  // ------------------------------------------------------------------ phase 3
  pdl_wait();
  for (int k = blockIdx.x; k < N; k += gridDim.x) {
    const float dc = dchat[static_cast<int64_t>(k) * GD + tid];
    const float ch = chat[static_cast<int64_t>(k) * GD + tid];
    const float dot = block_sum_256(dc * ch, red8);
    const float u = (dc - dot * ch) * cinv[k] * (1.f / static_cast<float>(M));
    for (int m = 0; m < M; ++m) dE[(static_cast<int64_t>(k) * M + m) * GD + tid] += u;
  }
}

// ------------------------------------------------------------------------------------------------ row-tile stage, v2
// Same mathematics as phase 2 above, re-tiled for the shared-memory pipe (ncu of the first version: L1/shared 64 % busy on
// the 60 SMs it used, 20 us of the 32 us call at N = 64):
//   * 8 rows per block instead of 16 -> 120 blocks at the training size (148 SMs), two blocks per SM;
//   * the centroid tile arrives by bulk copy (cp.async.bulk -> mbarrier), one 1 KB row per copy, while the block
//     normalises its rows;
//   * S tile: 2 x 2 outputs per thread, the warp covers 8 rows x 16 centroids, K split over the two warp halves of the
//     block: 4 shared-memory wavefronts per 16 FMA instructions (was 9);
//   * whole logit rows stay in shared memory (N < 256), warp <-> row for the log-sum-exp and G;
//   * one sweep over the centroids feeds both dEhat (8 row accumulators, thread <-> column) and dChat (row values of the
//     column in registers, one RED per centroid): 3 wavefronts per 16 FMA instructions (was 5 + 5).
constexpr int RT = 8;            // rows per tile
constexpr int EPAD = GD + 4;     // padded row: the four rows a warp reads at one d land in different banks
constexpr int NCAP = 256;        // this path serves N < GE2E_TC_MIN_SPEAKERS
constexpr int TILE_THREADS = 512;
static_assert(NCAP >= GE2E_TC_MIN_SPEAKERS - 1 && NCAP % TC == 0, "the row cache must hold every N of the SIMT path");

struct Ge2eTileSmem {
  float c[TC][CPAD];             // normalised centroid tile (bulk-copy destination)
  float e[RT][EPAD];             // normalised rows
  float s[RT][NCAP];             // cosine similarities of the whole row
  float gt[NCAP][RT];            // G^T: the 8 row values of one centroid are two 16-byte broadcasts
  float part[3][RT][TC];         // K quarters 1..3 of the S tile
  float red[8][RT];
  float scal[3][8];
  float rowv[RT];
  float einv[RT];
  unsigned long long bar;
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(TILE_THREADS, 2)
ge2e_rows_tile_kernel(const float* __restrict__ E, int N, int M, const float* __restrict__ w_ptr,
                      const float* __restrict__ b_ptr, float* __restrict__ loss, float* __restrict__ dE,
                      float* __restrict__ dw, float* __restrict__ db, const float* __restrict__ chat,
                      const float* __restrict__ einv, float* __restrict__ dchat, int need_grad) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Ge2eTileSmem& sm = *reinterpret_cast<Ge2eTileSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t NM = static_cast<int64_t>(N) * M;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * RT;
  const int col_tiles = (N + TC - 1) / TC;
  const float w = __ldg(w_ptr), b = __ldg(b_ptr);
  const float inv_nm = 1.f / static_cast<float>(NM);
  const uint32_t bar = smem_u32(&sm.bar);
  uint32_t parity = 0;

  pdl_trigger();
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();               // the centroid stage is complete: chat, einv, zeroed accumulators
  // warp 0 requests centroid tile `ct`; every thread of the block has finished reading the previous tile (caller syncs)
  auto request_tile = [&](int ct) {
    if (warp == 0) {
      const int c0 = ct * TC;
      const int rows = min(TC, N - c0);
      if (lane == 0) {
        fence_proxy_async();
        mbar_arrive_expect_tx(bar, static_cast<uint32_t>(rows) * GD * 4u);
      }
      __syncwarp();
      for (int r = lane; r < rows; r += 32)
        bulk_g2s(smem_u32(&sm.c[r][0]), chat + static_cast<int64_t>(c0 + r) * GD, GD * 4u, bar);
    }
  };
  request_tile(0);

  if (warp < RT) {  // this block's rows, normalised: warp <-> row (warps 0..7)
    const int64_t gi = row0 + warp;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    float sc = 0.f;
    if (gi < NM) {
      sc = __ldg(einv + gi);
      a0 = __ldg(reinterpret_cast<const float4*>(E + gi * GD + lane * 8));
      a1 = __ldg(reinterpret_cast<const float4*>(E + gi * GD + lane * 8 + 4));
    }
    a0.x *= sc; a0.y *= sc; a0.z *= sc; a0.w *= sc;
    a1.x *= sc; a1.y *= sc; a1.z *= sc; a1.w *= sc;
    *reinterpret_cast<float4*>(&sm.e[warp][lane * 8]) = a0;
    *reinterpret_cast<float4*>(&sm.e[warp][lane * 8 + 4]) = a1;
    if (lane == 0) sm.einv[warp] = sc;
  }
  __syncthreads();

  // ---- pass A: S = Ehat Chat^T for the whole rows
  const int kq = warp >> 2, wq = warp & 3, lr = lane >> 3, lc = lane & 7;      // kq: K quarter of the 16 warps
  const int ra = lr, rb = lr + 4, ca = 16 * wq + lc, cb = ca + 8;
  for (int ct = 0; ct < col_tiles; ++ct) {
    mbar_wait(bar, parity, 0x6e01);
    parity ^= 1;
    float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
    const int d0 = kq * (GD / 4);
#pragma unroll 8
    for (int d = d0; d < d0 + GD / 4; d += 4) {
      const float4 ea = *reinterpret_cast<const float4*>(&sm.e[ra][d]);
      const float4 eb = *reinterpret_cast<const float4*>(&sm.e[rb][d]);
      const float4 va = *reinterpret_cast<const float4*>(&sm.c[ca][d]);
      const float4 vb = *reinterpret_cast<const float4*>(&sm.c[cb][d]);
      s00 = fmaf(ea.x, va.x, s00); s00 = fmaf(ea.y, va.y, s00); s00 = fmaf(ea.z, va.z, s00); s00 = fmaf(ea.w, va.w, s00);
      s01 = fmaf(ea.x, vb.x, s01); s01 = fmaf(ea.y, vb.y, s01); s01 = fmaf(ea.z, vb.z, s01); s01 = fmaf(ea.w, vb.w, s01);
      s10 = fmaf(eb.x, va.x, s10); s10 = fmaf(eb.y, va.y, s10); s10 = fmaf(eb.z, va.z, s10); s10 = fmaf(eb.w, va.w, s10);
      s11 = fmaf(eb.x, vb.x, s11); s11 = fmaf(eb.y, vb.y, s11); s11 = fmaf(eb.z, vb.z, s11); s11 = fmaf(eb.w, vb.w, s11);
    }
    if (kq != 0) {
      float(*pt)[TC] = sm.part[kq - 1];
      pt[ra][ca] = s00; pt[ra][cb] = s01; pt[rb][ca] = s10; pt[rb][cb] = s11;
    }
    __syncthreads();          // every read of this centroid tile is done; partial sums visible
    if (ct + 1 < col_tiles) request_tile(ct + 1);
    if (kq == 0) {
      const int c0 = ct * TC;
      sm.s[ra][c0 + ca] = s00 + (sm.part[0][ra][ca] + sm.part[1][ra][ca] + sm.part[2][ra][ca]);
      sm.s[ra][c0 + cb] = s01 + (sm.part[0][ra][cb] + sm.part[1][ra][cb] + sm.part[2][ra][cb]);
      sm.s[rb][c0 + ca] = s10 + (sm.part[0][rb][ca] + sm.part[1][rb][ca] + sm.part[2][rb][ca]);
      sm.s[rb][c0 + cb] = s11 + (sm.part[0][rb][cb] + sm.part[1][rb][cb] + sm.part[2][rb][cb]);
    }
    __syncthreads();          // part may be rewritten by the next tile; s complete after the last one
  }
  // the centroid tile needed first by pass B: tile 0 (still resident when there is only one)
  if (need_grad && col_tiles > 1) request_tile(0);

  // ---- log-sum-exp, loss, dw, db, G: warp <-> row (warps 0..7)
  if (warp < RT) {
    const int64_t gi = row0 + warp;
    const bool live = gi < NM;
    const int label = live ? static_cast<int>(gi / M) : -1;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, w * sm.s[warp][j] - b);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) sum += expf(w * sm.s[warp][j] - b - mx);
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    float lrow = 0.f, dw_part = 0.f, db_part = 0.f;
    if (live && lane == 0) lrow = lse - (w * sm.s[warp][label] - b);
    if (need_grad) {
      for (int j = lane; j < N; j += 32) {
        float gval = 0.f;
        if (live) {
          const float sv = sm.s[warp][j];
          const float p = expf(w * sv - b - lse);
          const float pm = (p - (j == label ? 1.f : 0.f)) * inv_nm;
          dw_part += pm * sv;
          db_part -= pm;
          gval = w * pm;
        }
        sm.gt[j][warp] = gval;
      }
      dw_part = warp_sum(dw_part);
      db_part = warp_sum(db_part);
    }
    if (lane == 0) { sm.scal[0][warp] = lrow; sm.scal[1][warp] = dw_part; sm.scal[2][warp] = db_part; }
  }
  __syncthreads();
  if (tid < 3 && (tid == 0 || need_grad)) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += sm.scal[tid][i];
    atomicAdd(tid == 0 ? loss : (tid == 1 ? dw : db), tid == 0 ? a * inv_nm : a);
  }
  if (!need_grad) return;

  // ---- pass B: dEhat = G Chat (registers, thread <-> column) and dChat += G^T Ehat (RED per centroid)
  // the two halves of the block take alternate centroids of the sweep; thread <-> column `col` in both
  const int half = tid >> 8, col = tid & (GD - 1);
  float er[RT], dacc[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) { er[r] = sm.e[r][col]; dacc[r] = 0.f; }
  for (int ct = 0; ct < col_tiles; ++ct) {
    const int c0 = ct * TC;
    const int cols = min(TC, N - c0);
    if (col_tiles > 1) {
      mbar_wait(bar, parity, 0x6e02);
      parity ^= 1;
    }
    // every block starts its sweep at another centroid: the REDs of the 120 blocks then land on different lines (and
    // L2 slices) at any one time instead of queueing on the same 1 KB row
    const int rot = static_cast<int>(blockIdx.x % static_cast<unsigned>(cols));
#pragma unroll 4
    for (int i = half; i < cols; i += 2) {
      const int c = (i + rot >= cols) ? i + rot - cols : i + rot;
      const float4 g0 = *reinterpret_cast<const float4*>(&sm.gt[c0 + c][0]);
      const float4 g1 = *reinterpret_cast<const float4*>(&sm.gt[c0 + c][4]);
      const float cv = sm.c[c][col];
      dacc[0] = fmaf(g0.x, cv, dacc[0]); dacc[1] = fmaf(g0.y, cv, dacc[1]);
      dacc[2] = fmaf(g0.z, cv, dacc[2]); dacc[3] = fmaf(g0.w, cv, dacc[3]);
      dacc[4] = fmaf(g1.x, cv, dacc[4]); dacc[5] = fmaf(g1.y, cv, dacc[5]);
      dacc[6] = fmaf(g1.z, cv, dacc[6]); dacc[7] = fmaf(g1.w, cv, dacc[7]);
      float a = g0.x * er[0];
      a = fmaf(g0.y, er[1], a); a = fmaf(g0.z, er[2], a); a = fmaf(g0.w, er[3], a);
      a = fmaf(g1.x, er[4], a); a = fmaf(g1.y, er[5], a); a = fmaf(g1.z, er[6], a); a = fmaf(g1.w, er[7], a);
      atomicAdd(dchat + static_cast<int64_t>(c0 + c) * GD + col, a);
    }
    if (ct + 1 < col_tiles) {
      __syncthreads();
      request_tile(ct + 1);
    }
  }
  // the upper half hands its accumulators over (S is dead: its buffer carries them)
  if (half == 1) {
#pragma unroll
    for (int r = 0; r < RT; ++r) sm.s[r][col] = dacc[r];
  }
  __syncthreads();
  // project out the radial component, write the row part of dE
  if (half == 0) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      dacc[r] += sm.s[r][col];
      const float dot = warp_sum(dacc[r] * er[r]);
      if (lane == 0) sm.red[warp][r] = dot;
    }
  }
  __syncthreads();
  if (tid < RT) {
    float a = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) a += sm.red[wv][tid];
    sm.rowv[tid] = a;
  }
  __syncthreads();
  if (half == 0) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int64_t gi = row0 + r;
      if (gi < NM) dE[gi * GD + col] = (dacc[r] - sm.rowv[r] * er[r]) * sm.einv[r];
    }
  }
}

static int g_ge2e_tile_v2 = 1;     // spk_set_option("ge2e_row_tile_v2", 0/1): 0 = the first 16-row stage (A/B, tests)
void ge2e_set_tile_v2(int on) { g_ge2e_tile_v2 = on != 0; }
static int g_ge2e_pdl = 1;         // spk_set_option("ge2e_dependent_launch", 0/1): programmatic dependent launch of stages 2, 3
void ge2e_set_pdl(int on) { g_ge2e_pdl = on != 0; }

static size_t ge2e_simt_workspace_bytes(int N, int M) {
  const size_t NM = static_cast<size_t>(N) * M;
  return (2 * static_cast<size_t>(N) * GD + N + NM) * sizeof(float) + 256;
}
size_t ge2e_workspace_bytes(int N, int M) {
  return N >= GE2E_TC_MIN_SPEAKERS ? ge2e_tc_workspace_bytes(N, M) : ge2e_simt_workspace_bytes(N, M);
}

int ge2e_fused(const float* E, int N, int M, int D, const float* w, const float* b, float* loss, float* dE, float* dw,
               float* db, void* ws, size_t ws_bytes, cudaStream_t st) {
  SPK_CHECK(D == GD, "ge2e: embedding size %d not supported by this build (256)", D);
  SPK_CHECK(N >= 1 && M >= 1, "ge2e: need at least one speaker and one utterance");
  SPK_CHECK((reinterpret_cast<uintptr_t>(E) & 15) == 0, "ge2e: embeddings must be 16-byte aligned");
  if (N >= GE2E_TC_MIN_SPEAKERS)   // intensity 0.75 N FLOP/B: tensor-core composition (ge2e_tc.cu)
    return ge2e_tc(E, N, M, w, b, loss, dE, dw, db, ws, ws_bytes, st);
  if (ws_bytes < ge2e_simt_workspace_bytes(N, M)) {
    set_error("ge2e: workspace too small (%zu < %zu)", ws_bytes, ge2e_simt_workspace_bytes(N, M));
    return SPK_ENOMEM;
  }
  const int need_grad = (dE != nullptr);
  float* chat = reinterpret_cast<float*>(ws);
  float* dchat = chat + static_cast<size_t>(N) * GD;
  float* cinv = dchat + static_cast<size_t>(N) * GD;
  float* einv = cinv + N;

  static PerDeviceOnce once;
  const int smem = static_cast<int>(sizeof(Ge2eSmem));
  const int smem2 = static_cast<int>(sizeof(Ge2eTileSmem));
  SPK_TRY(once.run([&]() -> int {
    SPK_CUDA(cudaFuncSetAttribute(ge2e_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SPK_CUDA(cudaFuncSetAttribute(ge2e_rows_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    return 0;
  }));
  const long long NM = 1LL * N * M;
  const int row_tiles = static_cast<int>((NM + TR - 1) / TR);
  const float eps = 1e-8f;
  ProfScope prof(need_grad ? "ge2e_fused_fwd_bwd" : "ge2e_fused_fwd", (need_grad ? 6.0 : 2.0) * NM * N * GD,
                 (need_grad ? 2.0 : 1.0) * NM * GD * 4.0, st);
  // three launches in stream order: centroids | row tiles (loss, dE row part, dC) | centroid part of dE
  ge2e_fused_kernel<<<N, 256, smem, st>>>(E, N, M, w, b, loss, dE, dw, db, chat, cinv, einv, dchat, need_grad, eps, 1);
  SPK_CUDA(cudaGetLastError());
  cudaLaunchAttribute pdl_attr[1];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cfg.attrs = pdl_attr;
  const bool pdl = g_ge2e_tile_v2 && g_ge2e_pdl;
  cfg.numAttrs = pdl ? 1 : 0;
  const float* chat_c = chat;
  const float* einv_c = einv;
  if (g_ge2e_tile_v2) {
    cfg.gridDim = dim3(static_cast<unsigned>((NM + RT - 1) / RT));
    cfg.blockDim = dim3(TILE_THREADS);
    cfg.dynamicSmemBytes = smem2;
    SPK_CUDA(cudaLaunchKernelEx(&cfg, ge2e_rows_tile_kernel, E, N, M, w, b, loss, dE, dw, db, chat_c, einv_c, dchat,
                                need_grad));
  } else {
    ge2e_fused_kernel<<<row_tiles, 256, smem, st>>>(E, N, M, w, b, loss, dE, dw, db, chat, cinv, einv, dchat, need_grad, eps, 2);
    SPK_CUDA(cudaGetLastError());
  }
  if (need_grad) {
    cfg.gridDim = dim3(N);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    SPK_CUDA(cudaLaunchKernelEx(&cfg, ge2e_fused_kernel, E, N, M, w, b, loss, dE, dw, db, chat, cinv, einv, dchat, need_grad,
                                eps, 3));
  }
  return 0;
}

}  // namespace spk
