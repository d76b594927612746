// bf16 dual GEMM on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), sm_100a only.
//
//   D1[m,n] = sum_k A1[m,k] B1[n,k]        D2[m,n] = sum_k A2[m,k] B2[n,k]
//
// All four operands are K-major bf16 (row-major (rows, K)); both accumulators are fp32 in TMEM and
// meet in ONE epilogue.  This is the LRT hot op (LBBNN-GP-MF-LRT.py:172-175: two torch.mm + the
// sqrt/eps FMA) and, with the operands re-bound, both backward GEMM pairs:
//   forward   A = (x, x^2)        B = (M, V)          epilogue: act = D1 + b_mu + sqrt(D2 + s_b^2) eps
//   dX        A = (dE, dS)        B = (M^T, V^T)      epilogue: dx = D1 + 2 x D2, relu mask, next dE/dS
//   dW        A = (dE^T, dS^T)    B = (x^T, x^2^T)    epilogue: dM = D1, dV = D2 (fp32, for finalize)
//
// Kernel shape: persistent, one CTA per SM, 320 threads =
//   warp 0      TMA producer   (cp.async.bulk.tensor.2d, SWIZZLE_128B, 4 boxes of 128x64 bf16 per stage)
//   warp 1      MMA issuer     (one elected lane: tcgen05.mma.cta_group::1.kind::f16, M=128 N=128 K=16)
//   warps 2..9  epilogue       (tcgen05.ld 32x32b.x16 -> registers -> math -> global; two warps per TMEM
//                               lane quarter, one per 64-column half; r01: with 4 warps the fwd epilogue
//                               took ~2x the MMA time of a tile)
// smem: 3 stages x 64 KB ring (full/empty mbarriers); TMEM: 512 columns = 2 accumulator stages x
// (D1: 128 cols, D2: 128 cols), so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace lbbnn {
namespace {

constexpr int BM = 128, BN = 128, BK = 64;      // CTA tile; BK*2B = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kStages = 3;
constexpr int kTileBytes = BM * BK * 2;         // 16 KB per operand tile
constexpr int kStageBytes = 4 * kTileBytes;     // A1, A2, B1, B2
constexpr int kEpiWarps = 8;                    // 2 per TMEM lane quarter: each takes one 64-column half of the tile
constexpr int kTcThreads = 64 + kEpiWarps * 32; // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kTmemCols = 512;
constexpr int EW = 16;                          // epilogue chunk: columns per tcgen05.ld
constexpr int kBiasBytes = 2 * 2 * BN * 4;      // [accumulator stage][b_mu | sigma_b^2][BN]
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kBiasBytes;

using namespace tc;

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- epilogues ------------------------------------------------------------------------------------
struct TcEpi {
  int mode;                 // LBBNN_TC_EPI_*
  int flags;                // LBBNN_FLAG_RELU / LBBNN_FLAG_MASK_DX
  // RAW (dW): fp32 outputs
  float *d1, *d2;
  // FWD
  const float *bias_mu, *bias_rho;
  Noise noise;
  __nv_bfloat16 *o_bf, *o2_bf, *oT_bf, *o2T_bf;         // act, act^2 (M,N) and their transposes (N,M)
  float *o_f32, *dsf;                                    // optional fp32 copy of act; eps/(2 sd) (M,N)
  // DX
  const __nv_bfloat16* x_bf;                             // (M,N): the input this dx belongs to
  const float* dsf_prev;                                 // (M,N): ds factor of the layer that produced x
  __nv_bfloat16 *de, *ds, *deT, *dsT;                    // next (previous-layer) dE, dS and transposes
};

// one thread = one output row (TMEM lane), EW consecutive columns; sbias = this tile's b_mu[BN] | sigma_b^2[BN]
__device__ __forceinline__ void epilogue_chunk(const TcEpi& e, const Noise& nz, int64_t M, int64_t N, int64_t row, int64_t col0,
                                               const float* __restrict__ sbias, int cl, const float d1[EW],
                                               const float d2[EW]) {
  if (row >= M) return;
  const bool fullc = col0 + EW - 1 < N;
  if (e.mode == LBBNN_TC_EPI_RAW) {
    float* p1 = e.d1 + row * N + col0;
    float* p2 = e.d2 + row * N + col0;
    if (fullc && (N % 4 == 0)) {
#pragma unroll
      for (int j = 0; j < EW; j += 4) {
        *reinterpret_cast<float4*>(p1 + j) = make_float4(d1[j], d1[j + 1], d1[j + 2], d1[j + 3]);
        *reinterpret_cast<float4*>(p2 + j) = make_float4(d2[j], d2[j + 1], d2[j + 2], d2[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j)
        if (col0 + j < N) { p1[j] = d1[j]; p2[j] = d2[j]; }
    }
    return;
  }
  float o[EW], o2[EW];
  if (e.mode == LBBNN_TC_EPI_FWD) {
    // act = D1 + b_mu + sqrt(D2 + sigma_b^2) eps  (LRT:172-175); ds factor = eps / (2 sd)
    float ep[EW];
    if (nz.ptr) {
#pragma unroll
      for (int j = 0; j < EW; ++j) ep[j] = (col0 + j < N) ? __ldg(nz.ptr + row * N + col0 + j) : 0.f;
    } else if (N % 4 == 0) {   // quads of the flat (M,N) index are aligned with this row's columns
#pragma unroll
      for (int j = 0; j < EW; j += 4) philox_normal4(nz.seed, nz.stream, ((uint64_t)row * (uint64_t)N + (uint64_t)(col0 + j)) >> 2, ep + j);
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j) ep[j] = philox_normal1(nz.seed, nz.stream, (uint64_t)row * (uint64_t)N + (uint64_t)(col0 + j));
    }
#pragma unroll
    for (int j = 0; j < EW; ++j) {
      const int64_t n = col0 + j;
      float v = 0.f, f = 0.f;
      if (n < N) {
        const float sd = sqrtf(fmaxf(d2[j], 0.f) + sbias[BN + cl + j]);
        v = d1[j] + sbias[cl + j] + sd * ep[j];
        f = ep[j] / (2.0f * sd);
        if (e.flags & LBBNN_FLAG_RELU) v = fmaxf(v, 0.f);
      }
      o[j] = v;
      o2[j] = f;
    }
    const bool v4 = fullc && (N % 4 == 0);
    if (e.o_f32 || e.dsf) {
      if (v4) {
#pragma unroll
        for (int j = 0; j < EW; j += 4) {
          if (e.o_f32) *reinterpret_cast<float4*>(e.o_f32 + row * N + col0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          if (e.dsf) *reinterpret_cast<float4*>(e.dsf + row * N + col0 + j) = make_float4(o2[j], o2[j + 1], o2[j + 2], o2[j + 3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < EW; ++j)
          if (col0 + j < N) {
            if (e.o_f32) e.o_f32[row * N + col0 + j] = o[j];
            if (e.dsf) e.dsf[row * N + col0 + j] = o2[j];
          }
      }
    }
    // row-major bf16: act, act^2
    if (fullc && (N % 8 == 0)) {
#pragma unroll
      for (int j = 0; j < EW; j += 8) {
        uint4 a, b;
        a.x = pack_bf16(o[j], o[j + 1]); a.y = pack_bf16(o[j + 2], o[j + 3]); a.z = pack_bf16(o[j + 4], o[j + 5]); a.w = pack_bf16(o[j + 6], o[j + 7]);
        b.x = pack_bf16(o[j] * o[j], o[j + 1] * o[j + 1]); b.y = pack_bf16(o[j + 2] * o[j + 2], o[j + 3] * o[j + 3]);
        b.z = pack_bf16(o[j + 4] * o[j + 4], o[j + 5] * o[j + 5]); b.w = pack_bf16(o[j + 6] * o[j + 6], o[j + 7] * o[j + 7]);
        if (e.o_bf) *reinterpret_cast<uint4*>(e.o_bf + row * N + col0 + j) = a;
        if (e.o2_bf) *reinterpret_cast<uint4*>(e.o2_bf + row * N + col0 + j) = b;
      }
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j)
        if (col0 + j < N) {
          if (e.o_bf) e.o_bf[row * N + col0 + j] = __float2bfloat16_rn(o[j]);
          if (e.o2_bf) e.o2_bf[row * N + col0 + j] = __float2bfloat16_rn(o[j] * o[j]);
        }
    }
    // transposed bf16 (N,M): lanes of a warp are consecutive rows -> coalesced 64 B per column
    if (e.oT_bf) {
#pragma unroll
      for (int j = 0; j < EW; ++j)
        if (col0 + j < N) {
          e.oT_bf[(col0 + j) * M + row] = __float2bfloat16_rn(o[j]);
          e.o2T_bf[(col0 + j) * M + row] = __float2bfloat16_rn(o[j] * o[j]);
        }
    }
    return;
  }
  // LBBNN_TC_EPI_DX: dx = D1 + 2 x D2; through the relu that produced x; dS_prev = dx * dsf_prev
  float xv[EW], fv[EW];
  if (fullc && (N % 8 == 0)) {
#pragma unroll
    for (int j = 0; j < EW; j += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(e.x_bf + row * N + col0 + j);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f2 = __bfloat1622float2(h[t]);
        xv[j + 2 * t] = f2.x; xv[j + 2 * t + 1] = f2.y;
      }
    }
#pragma unroll
    for (int j = 0; j < EW; j += 4) {
      const float4 f4 = e.dsf_prev ? *reinterpret_cast<const float4*>(e.dsf_prev + row * N + col0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      fv[j] = f4.x; fv[j + 1] = f4.y; fv[j + 2] = f4.z; fv[j + 3] = f4.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < EW; ++j) {
      const bool ok = col0 + j < N;
      xv[j] = ok ? __bfloat162float(e.x_bf[row * N + col0 + j]) : 0.f;
      fv[j] = (ok && e.dsf_prev) ? e.dsf_prev[row * N + col0 + j] : 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < EW; ++j) {
    float g = fmaf(2.0f * xv[j], d2[j], d1[j]);
    if ((e.flags & LBBNN_FLAG_MASK_DX) && !(xv[j] > 0.f)) g = 0.f;
    o[j] = g;
    o2[j] = g * fv[j];
  }
  if (fullc && (N % 8 == 0)) {
#pragma unroll
    for (int j = 0; j < EW; j += 8) {
      uint4 a, b;
      a.x = pack_bf16(o[j], o[j + 1]); a.y = pack_bf16(o[j + 2], o[j + 3]); a.z = pack_bf16(o[j + 4], o[j + 5]); a.w = pack_bf16(o[j + 6], o[j + 7]);
      b.x = pack_bf16(o2[j], o2[j + 1]); b.y = pack_bf16(o2[j + 2], o2[j + 3]); b.z = pack_bf16(o2[j + 4], o2[j + 5]); b.w = pack_bf16(o2[j + 6], o2[j + 7]);
      *reinterpret_cast<uint4*>(e.de + row * N + col0 + j) = a;
      *reinterpret_cast<uint4*>(e.ds + row * N + col0 + j) = b;
    }
  } else {
#pragma unroll
    for (int j = 0; j < EW; ++j)
      if (col0 + j < N) {
        e.de[row * N + col0 + j] = __float2bfloat16_rn(o[j]);
        e.ds[row * N + col0 + j] = __float2bfloat16_rn(o2[j]);
      }
  }
  if (e.deT) {
#pragma unroll
    for (int j = 0; j < EW; ++j)
      if (col0 + j < N) {
        e.deT[(col0 + j) * M + row] = __float2bfloat16_rn(o[j]);
        e.dsT[(col0 + j) * M + row] = __float2bfloat16_rn(o2[j]);
      }
  }
}

// ---- tile order -----------------------------------------------------------------------------------
// The persistent CTAs walk t = blockIdx.x, +gridDim.x, ...: at any moment the SMs work on ~148 consecutive tile indices.
// Column-major order (all row blocks of one column block, then the next) makes every wave stream ALL of A: at the wide
// shape A1 + A2 are 134 MB > L2 and were re-read from HBM for each of the 32 column blocks (~4.3 GB per GEMM).  Tiles are
// therefore numbered inside groups of kGroupM row blocks: a wave covers a compact (kGroupM x ~9) patch, the group's A rows
// (32-64 MB) stay in L2 while B streams through once per group.
constexpr int kGroupM = 16;
__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int& mb, int& nb) {
  const int per_group = kGroupM * num_n;
  const int g = t / per_group, r = t - g * per_group;
  const int m_first = g * kGroupM;
  const int gm = min(kGroupM, num_m - m_first);       // ragged last group
  nb = r / gm;
  mb = m_first + (r - nb * gm);
}

// ---- the kernel -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
tc_dual_gemm_bf16(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, const TcEpi epi,
                  int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]        MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]      epilogue -> MMA
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  float* sbias_all = reinterpret_cast<float*>(smem + kStages * kStageBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n, num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmB2);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int mb, nb;
        tile_coords(t, num_m, num_n, mb, nb);
        const int m0 = mb * BM, n0 = nb * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(sa + 0 * kTileBytes, &tmA1, &full_bar[stage], kb * BK, m0);
          tma_load_2d(sa + 1 * kTileBytes, &tmA2, &full_bar[stage], kb * BK, m0);
          tma_load_2d(sa + 2 * kTileBytes, &tmB1, &full_bar[stage], kb * BK, n0);
          tma_load_2d(sa + 3 * kTileBytes, &tmB2, &full_bar[stage], kb * BK, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);   // epilogue drained this accumulator stage
        tc_fence_after();
        const uint32_t d1 = tmem_base + as * 256, d2 = d1 + 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t a1 = umma_desc_kmajor_sw128(sa + 0 * kTileBytes), a2 = umma_desc_kmajor_sw128(sa + 1 * kTileBytes);
          const uint64_t b1 = umma_desc_kmajor_sw128(sa + 2 * kTileBytes), b2 = umma_desc_kmajor_sw128(sa + 3 * kTileBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);   // advance the start address inside the swizzle row
            const uint32_t acc = (kb | k) ? 1u : 0u;
            umma_bf16(d1, a1 + koff, b1 + koff, acc);
            umma_bf16(d2, a2 + koff, b2 + koff, acc);
          }
          umma_commit(&empty_bar[stage]);                  // frees the smem stage once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);                       // accumulators complete -> epilogue
      }
    }
  } else {
    // ===== epilogue warps (2..9): TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    Noise nz = epi.noise;
    nz.resolve();
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;                      // 0..255 among the epilogue threads
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      int mb, nb;
      tile_coords(t, num_m, num_n, mb, nb);
      const int m0 = mb * BM, n0 = nb * BN;
      float* sbias = sbias_all + as * 2 * BN;
      if (epi.mode == LBBNN_TC_EPI_FWD) {                 // this tile's bias terms, once per column
        const int c = et & (BN - 1);
        const int64_t n = n0 + c;
        float v = 0.f;
        if (n < N) {
          if (et < BN) v = __ldg(epi.bias_mu + n);
          else { const float sb = sigma_of(__ldg(epi.bias_rho + n)); v = sb * sb; }
        }
        sbias[(et < BN ? 0 : BN) + c] = v;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      }
      mbar_wait(&tfull_bar[as], (it >> 1) & 1);
      tc_fence_after();
      const int64_t row = m0 + q * 32 + lane;
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256 + half * 64;
#pragma unroll 1
      for (int c = 0; c < 64 / EW; ++c) {
        float v1[EW], v2[EW];
        tmem_ld16(tbase + c * EW, v1);
        tmem_ld16(tbase + 128 + c * EW, v2);
        const int cl = half * 64 + c * EW;
        epilogue_chunk(epi, nz, M, N, row, n0 + cl, sbias, cl, v1, v2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- the CTA-pair kernel ------------------------------------------------------------------------------
// Same GEMM pair on 256 x 128 tiles owned by a cluster of two CTAs (one TPC): each CTA stages ITS 128 rows of A1, A2 and
// HALF (64 rows) of the B1, B2 tiles; the leader's tcgen05.mma.cta_group::2 (M = 256, N = 128) reads both halves, so a stage
// is 48 KB per SM instead of 64 KB for the same MMA work (25 % less L2 -> SM operand traffic, what the 1-CTA kernel waits
// on: tensor pipe 54-60 % active in profiles/r01_ncu_tc_dual_gemm.json) and FOUR stages fit.  Each CTA keeps its 128 rows
// of both accumulators in its own TMEM (2 stages x 256 columns) and runs the unchanged epilogue on them.
//   full barrier   (leader's): 2 arrivals (each CTA's producer, with its own expect_tx) + the bytes of both CTAs' loads
//   empty barrier  (per CTA):  the leader's tcgen05.commit multicast to both CTAs frees the stage in both
//   tmem full      (per CTA):  commit multicast after the tile's last MMA
//   tmem empty     (leader's): 2 x kEpiWarps arrivals, the peer's epilogue warps arrive remotely
constexpr int kStages2 = 4;
constexpr int kBHalfBytes = (BN / 2) * BK * 2;                 // 8 KB: this CTA's 64 rows of a B tile
constexpr int kStageBytes2 = 2 * kTileBytes + 2 * kBHalfBytes;  // 48 KB
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + 1024 + 256 + kBiasBytes;
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
constexpr int kGroupM2 = kGroupM / 2;                          // in 256-row blocks

__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc2), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tile_coords2(int t, int num_m2, int num_n, int group, int& mb2, int& nb) {
  const int per_group = group * num_n;
  const int g = t / per_group, r = t - g * per_group;
  const int m_first = g * group;
  const int gm = min(group, num_m2 - m_first);
  nb = r / gm;
  mb2 = m_first + (r - nb * gm);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
tc_dual_gemm_bf16_pair(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                       const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, const TcEpi epi,
                       int M, int N, int K, int group) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages2 * kStageBytes2);
  uint64_t* full_bar = bars;                       // [kStages2]  (the leader's is the live one)
  uint64_t* empty_bar = bars + kStages2;           // [kStages2]
  uint64_t* tfull_bar = bars + 2 * kStages2;       // [2]
  uint64_t* tempty_bar = bars + 2 * kStages2 + 2;  // [2]         (the leader's is the live one)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * kStages2 + 4);
  float* sbias_all = reinterpret_cast<float*>(smem + kStages2 * kStageBytes2 + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_m2 = (M + 2 * BM - 1) / (2 * BM), num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m2 * num_n, num_kb = (K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmB2);
    for (int s = 0; s < kStages2; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 2 * kEpiWarps); }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc_pair(tmem_holder, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows, own half of the B rows; bytes credited to the leader's full barrier =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mb2, nb;
        tile_coords2(t, num_m2, num_n, group, mb2, nb);
        const int m0 = mb2 * 2 * BM + (int)rank * BM, n0h = nb * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes2;
          const uint32_t lead_full = mapa_u32(&full_bar[stage], 0);
          mbar_expect_tx_cluster(lead_full, kStageBytes2);
          tma_load_2d_pair(sa, &tmA1, lead_full, kb * BK, m0);
          tma_load_2d_pair(sa + kTileBytes, &tmA2, lead_full, kb * BK, m0);
          tma_load_2d_pair(sa + 2 * kTileBytes, &tmB1, lead_full, kb * BK, n0h);
          tma_load_2d_pair(sa + 2 * kTileBytes + kBHalfBytes, &tmB2, lead_full, kb * BK, n0h);
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);   // both CTAs' epilogues drained this accumulator stage
        tc_fence_after();
        const uint32_t d1 = tmem_base + as * 256, d2 = d1 + 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes2);
          const uint64_t a1 = umma_desc_kmajor_sw128(sa), a2 = umma_desc_kmajor_sw128(sa + kTileBytes);
          const uint64_t b1 = umma_desc_kmajor_sw128(sa + 2 * kTileBytes);
          const uint64_t b2 = umma_desc_kmajor_sw128(sa + 2 * kTileBytes + kBHalfBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
            const uint32_t acc = (kb | k) ? 1u : 0u;
            umma_bf16_pair(d1, a1 + koff, b1 + koff, acc);
            umma_bf16_pair(d2, a2 + koff, b2 + koff, acc);
          }
          umma_commit_pair(&empty_bar[stage], 3);          // frees this stage in BOTH CTAs once the MMAs retire
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull_bar[as], 3);               // accumulators complete -> both epilogues
      }
    }
  } else {
    // ===== epilogue warps (both CTAs, each on its own 128 accumulator rows) =====
    Noise nz = epi.noise;
    nz.resolve();
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      const int as = it & 1;
      int mb2, nb;
      tile_coords2(t, num_m2, num_n, group, mb2, nb);
      const int m0 = mb2 * 2 * BM + (int)rank * BM, n0 = nb * BN;
      float* sbias = sbias_all + as * 2 * BN;
      if (epi.mode == LBBNN_TC_EPI_FWD) {
        const int c = et & (BN - 1);
        const int64_t n = n0 + c;
        float v = 0.f;
        if (n < N) {
          if (et < BN) v = __ldg(epi.bias_mu + n);
          else { const float sb = sigma_of(__ldg(epi.bias_rho + n)); v = sb * sb; }
        }
        sbias[(et < BN ? 0 : BN) + c] = v;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      }
      mbar_wait(&tfull_bar[as], (it >> 1) & 1);
      tc_fence_after();
      const int64_t row = m0 + q * 32 + lane;
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256 + half * 64;
#pragma unroll 1
      for (int c = 0; c < 64 / EW; ++c) {
        float v1[EW], v2[EW];
        tmem_ld16(tbase + c * EW, v1);
        tmem_ld16(tbase + 128 + c * EW, v2);
        const int cl = half * 64 + c * EW;
        epilogue_chunk(epi, nz, M, N, row, n0 + cl, sbias, cl, v1, v2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[as], 0));
    }
  }
  __syncwarp();                 // warps 0 and 1 ran single-lane loops: reconverge before the aligned cluster barrier
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ---- host: tensor maps --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// (rows, K) row-major bf16, box = 64 (K) x 128 (rows), 128B swizzle, OOB -> zeros
int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int box_rows = BM) {
  EncodeTiledFn enc = get_encode();
  LBBNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  LBBNN_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (K * 2) % 16 == 0,
                "TMA operand must be 16B aligned with a 16B-multiple row pitch (K %% 8 == 0), K=%lld", (long long)K);
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LBBNN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld", (int)r, (long long)rows, (long long)K);
  return LBBNN_OK;
}

int launch_tc(const void* A1, const void* A2, const void* B1, const void* B2, int64_t M, int64_t N, int64_t K, const TcEpi& epi,
              cudaStream_t st) {
  LBBNN_REQUIRE(A1 && A2 && B1 && B2 && M > 0 && N > 0 && K > 0, "bad GEMM operands");
  LBBNN_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "GEMM dims must fit int32");
  CUtensorMap mA1, mA2, mB1, mB2;
  if (int rc = make_map(&mA1, A1, M, K)) return rc;
  if (int rc = make_map(&mA2, A2, M, K)) return rc;
  // CTA pairs (256-row tiles) for problems that fill the GPU with them; LBBNN_TC_PAIR=0 forces the 1-CTA kernel
  // (2 = also for small problems: the tests use it to run odd shapes through the pair kernel; read per call, host only)
  const char* pe = getenv("LBBNN_TC_PAIR");
  const int pair_mode = pe ? atoi(pe) : 1;
  const int64_t tiles2 = ceil_div(M, 2 * BM) * ceil_div(N, BN);
  if (pair_mode && M > BM && (pair_mode == 2 || tiles2 >= sm_count() / 2)) {
    if (int rc = make_map(&mB1, B1, N, K, BN / 2)) return rc;
    if (int rc = make_map(&mB2, B2, N, K, BN / 2)) return rc;
    static bool attr2_set = false;
    if (!attr2_set) {
      LBBNN_CUDA(cudaFuncSetAttribute(tc_dual_gemm_bf16_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
      attr2_set = true;
    }
    const int clusters = (int)(tiles2 < sm_count() / 2 ? tiles2 : sm_count() / 2);
    const char* ge = getenv("LBBNN_TC_GROUP");      // 256-row blocks per tile group (experiments)
    const int group = ge && atoi(ge) > 0 ? atoi(ge) : kGroupM2;
    tc_dual_gemm_bf16_pair<<<2 * clusters, kTcThreads, kSmemBytes2, st>>>(mA1, mA2, mB1, mB2, epi, (int)M, (int)N, (int)K, group);
    return check_launch("tc_dual_gemm_bf16_pair");
  }
  if (int rc = make_map(&mB1, B1, N, K)) return rc;
  if (int rc = make_map(&mB2, B2, N, K)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    LBBNN_CUDA(cudaFuncSetAttribute(tc_dual_gemm_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int64_t tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  tc_dual_gemm_bf16<<<grid, kTcThreads, kSmemBytes, st>>>(mA1, mA2, mB1, mB2, epi, (int)M, (int)N, (int)K);
  return check_launch("tc_dual_gemm_bf16");
}

}  // namespace
}  // namespace lbbnn

using namespace lbbnn;

extern "C" int lbbnn_tc_dual_gemm_raw(const void* A1, const void* A2, const void* B1, const void* B2, int64_t M, int64_t N,
                                      int64_t K, float* D1, float* D2, lbbnn_stream s) {
  LBBNN_REQUIRE(D1 && D2, "NULL output");
  TcEpi e = {};
  e.mode = LBBNN_TC_EPI_RAW;
  e.d1 = D1; e.d2 = D2;
  return launch_tc(A1, A2, B1, B2, M, N, K, e, (cudaStream_t)s);
}

extern "C" int lbbnn_tc_lrt_fwd(const void* x_bf, const void* x2_bf, const void* M_bf, const void* V_bf, int64_t batch,
                                int64_t in_features, int64_t out_features, const float* bias_mu, const float* bias_rho,
                                const lbbnn_noise* nz, int flags, void* act_bf, void* act2_bf, void* actT_bf, void* act2T_bf,
                                float* ds_factor, float* act_f32, lbbnn_stream s) {
  LBBNN_REQUIRE(bias_mu && bias_rho, "NULL bias");
  LBBNN_REQUIRE((actT_bf == nullptr) == (act2T_bf == nullptr), "transposed outputs come in pairs");
  TcEpi e = {};
  e.mode = LBBNN_TC_EPI_FWD; e.flags = flags;
  e.bias_mu = bias_mu; e.bias_rho = bias_rho; e.noise = make_noise(nz);
  e.o_bf = (__nv_bfloat16*)act_bf; e.o2_bf = (__nv_bfloat16*)act2_bf; e.oT_bf = (__nv_bfloat16*)actT_bf;
  e.o2T_bf = (__nv_bfloat16*)act2T_bf; e.dsf = ds_factor; e.o_f32 = act_f32;
  return launch_tc(x_bf, x2_bf, M_bf, V_bf, batch, out_features, in_features, e, (cudaStream_t)s);
}

extern "C" int lbbnn_tc_lrt_bwd_input(const void* dE_bf, const void* dS_bf, const void* MT_bf, const void* VT_bf, int64_t batch,
                                      int64_t in_features, int64_t out_features, const void* x_bf, const float* ds_factor_prev,
                                      int flags, void* dE_prev_bf, void* dS_prev_bf, void* dE_prevT_bf, void* dS_prevT_bf,
                                      lbbnn_stream s) {
  LBBNN_REQUIRE(x_bf && dE_prev_bf && dS_prev_bf, "NULL argument");
  LBBNN_REQUIRE((dE_prevT_bf == nullptr) == (dS_prevT_bf == nullptr), "transposed outputs come in pairs");
  TcEpi e = {};
  e.mode = LBBNN_TC_EPI_DX; e.flags = flags;
  e.x_bf = (const __nv_bfloat16*)x_bf; e.dsf_prev = ds_factor_prev;
  e.de = (__nv_bfloat16*)dE_prev_bf; e.ds = (__nv_bfloat16*)dS_prev_bf; e.deT = (__nv_bfloat16*)dE_prevT_bf;
  e.dsT = (__nv_bfloat16*)dS_prevT_bf;
  // contraction over the out features: A = (dE, dS) (batch, out), B = (M^T, V^T) (in, out)
  return launch_tc(dE_bf, dS_bf, MT_bf, VT_bf, batch, in_features, out_features, e, (cudaStream_t)s);
}
