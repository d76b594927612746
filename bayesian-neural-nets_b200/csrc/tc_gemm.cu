// bf16 dual GEMM on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), sm_100a only.
//
//   D1[m,n] = sum_k A1[m,k] B1[n,k]        D2[m,n] = sum_k A2[m,k] B2[n,k]
//
// Operands are bf16, each either K-major (row-major (rows, K)) or MN-major (row-major (K, rows): the transposed view of
// a tensor that is stored with the contraction index as its ROW index); both accumulators are fp32 in TMEM and
// meet in ONE epilogue.  This is the LRT hot op (LBBNN-GP-MF-LRT.py:172-175: two torch.mm + the
// sqrt/eps FMA) and, with the operands re-bound, both backward GEMM pairs:
//   forward   A = (x, x^2)   K-major    B = (M, V)     K-major    epilogue: act = D1 + b_mu + sqrt(D2 + s_b^2) eps
//   dX        A = (dE, dS)   K-major    B = (M, V)     MN-major   epilogue: dx = D1 + 2 x D2, relu mask, next dE/dS, bias sums
//   dW        A = (dE, dS)   MN-major   B = (x, x^2)   MN-major   epilogue: dM = D1, dV = D2 -> chain rule + KL + Adam
// The MN-major forms read the SAME row-major tensors the other GEMMs use (TMA boxes of 64 contraction rows x 64 elements,
// tcgen05 "MN-major" canonical SWIZZLE_128B layout), so no transposed copy of an activation, gradient or weight moment is
// ever written.  (K-major transposed operands are still accepted: the r01 entry points and the tests use them.)
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
#include "lrt_chain.cuh"
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
// epilogue scratch after the barriers: the forward's bias terms, or (fused dW update) 4 KB per epilogue warp through which a
// 32-row x 16-column piece of both accumulators is re-dealt to the lanes row-contiguously (see dw_adam_tile)
constexpr int kEpiScratchBytes = kEpiWarps * 4096;
static_assert(kEpiScratchBytes >= kBiasBytes, "scratch holds the bias terms too");
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiScratchBytes;
// only the fused-update kernels ask for the large scratch (the others keep the shared-memory carve-out, and so L1, as it was)
constexpr int smem_for(int full, int mode) { return mode == LBBNN_TC_EPI_DW_ADAM ? full : full - kEpiScratchBytes + kBiasBytes; }

using namespace tc;

// kind::f16 instruction descriptor: D=f32, A=B=bf16, a_major / b_major (0 = K-major, 1 = MN-major) at bits 15 / 16,
// N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// MN-major operand tile in shared memory: TMA boxes of (64 contraction rows) x (64 elements = 128 B), SWIZZLE_128B, one
// box per 64 elements of the M / N extent, boxes kMnBoxBytes apart.  In tcgen05's canonical MN-major SWIZZLE_128B layout
// ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units that is SBO = 1024 B (8 contraction rows) and LBO = one box; one MMA
// (K = 16) spans two 8-row groups, so consecutive MMAs start 2048 B apart.
constexpr int kMnBoxBytes = 64 * 128;           // 8 KB
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(kMnBoxBytes >> 4) << 16;      // LBO: next 64 elements along M / N
  d |= (uint64_t)(1024 >> 4) << 32;             // SBO: next 8 rows along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, int mn) {
  return mn ? umma_desc_mn_sw128(smem_addr) : umma_desc_kmajor_sw128(smem_addr);
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
  float* colsum_part;                                    // optional [ceil(M/32)][2N]: per-32-row sums of dE | dS (fp32)
  // DW_ADAM: (M,N) = (out,in); chain rule + KL gradient + Adam on the accumulators, parameters updated in place
  float *p_mu, *p_rho, *p_lam;
  float *m_mu, *m_rho, *m_lam, *v_mu, *v_rho, *v_lam;
  const float* coef;
  lbbnn_priors pri;
  int var_mode;
  float klg, beta1, beta2, adam_eps;
  // optional: the NEXT step's bf16 operands M, V (out,in) and per-warp KL partials from the updated parameters
  __nv_bfloat16 *next_M, *next_V;
  double* next_kl_part;                                  // [gridDim.x * kEpiWarps]
  // operand majors
  int a_mn, b_mn;
  // TMA L2 prefetch distance in K blocks (0 = off): the producer asks L2 for the boxes it will load `l2_prefetch` K blocks
  // later, so the loads that miss L2 (~20 % at the wide shape) do not expose HBM latency to a 4-stage ring
  int l2_prefetch;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }

// ---- fused dW epilogue: chain rule + KL gradient + Adam on the accumulators --------------------------------------------------
// 4 consecutive weights of one output row (N % 4 == 0): nine fp32 tensors stream through (mu, rho, lambda + both Adam moments)
__device__ __forceinline__ void dw_adam_quad(const TcEpi& e, const chain::Consts& cc, const chain::KlC& kc, float& kl_acc, int64_t off,
                                             const float4 dM, const float4 dV) {
  const float4 mu4 = ld4(e.p_mu + off), rho4 = ld4(e.p_rho + off), lam4 = ld4(e.p_lam + off);
  const float4 mm4 = ld4(e.m_mu + off), mr4 = ld4(e.m_rho + off), ml4 = ld4(e.m_lam + off);
  const float4 vm4 = ld4(e.v_mu + off), vr4 = ld4(e.v_rho + off), vl4 = ld4(e.v_lam + off);
  float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rho[4] = {rho4.x, rho4.y, rho4.z, rho4.w}, lam[4] = {lam4.x, lam4.y, lam4.z, lam4.w};
  float mm[4] = {mm4.x, mm4.y, mm4.z, mm4.w}, mr[4] = {mr4.x, mr4.y, mr4.z, mr4.w}, ml[4] = {ml4.x, ml4.y, ml4.z, ml4.w};
  float vm[4] = {vm4.x, vm4.y, vm4.z, vm4.w}, vr[4] = {vr4.x, vr4.y, vr4.z, vr4.w}, vl[4] = {vl4.x, vl4.y, vl4.z, vl4.w};
  const float d1[4] = {dM.x, dM.y, dM.z, dM.w}, d2[4] = {dV.x, dV.y, dV.z, dV.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float gm, gr, gl;
    chain::grads(cc, mu[t], rho[t], lam[t], d1[t], d2[t], gm, gr, gl);
    chain::adam(cc, mu[t], mm[t], vm[t], gm);
    chain::adam(cc, rho[t], mr[t], vr[t], gr);
    chain::adam(cc, lam[t], ml[t], vl[t], gl);
  }
  st4(e.p_mu + off, mu[0], mu[1], mu[2], mu[3]); st4(e.p_rho + off, rho[0], rho[1], rho[2], rho[3]);
  st4(e.p_lam + off, lam[0], lam[1], lam[2], lam[3]);
  st4(e.m_mu + off, mm[0], mm[1], mm[2], mm[3]); st4(e.m_rho + off, mr[0], mr[1], mr[2], mr[3]);
  st4(e.m_lam + off, ml[0], ml[1], ml[2], ml[3]);
  st4(e.v_mu + off, vm[0], vm[1], vm[2], vm[3]); st4(e.v_rho + off, vr[0], vr[1], vr[2], vr[3]);
  st4(e.v_lam + off, vl[0], vl[1], vl[2], vl[3]);
  if (e.next_M) {   // the next forward's operands + KL term, from the values just written (replaces that step's prologue pass)
    float M[4], V[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float k_;
      chain::next_moments(kc, cc.var_mode, mu[t], rho[t], lam[t], M[t], V[t], k_);
      kl_acc += k_;
    }
    uint2 pm, pv;
    pm.x = pack_bf16(M[0], M[1]); pm.y = pack_bf16(M[2], M[3]);
    pv.x = pack_bf16(V[0], V[1]); pv.y = pack_bf16(V[2], V[3]);
    *reinterpret_cast<uint2*>(e.next_M + off) = pm;
    *reinterpret_cast<uint2*>(e.next_V + off) = pv;
  }
}

// One warp's 32 rows x 64 columns of both accumulators.  tcgen05.ld hands every lane ONE ROW (16 columns per load), but the
// update streams nine fp32 tensors from and to memory: row-per-lane addressing would scatter each warp access over 32 lines,
// 16 bytes each (first cut: 1.26 ms per 4096^2 layer, 0.9 TB/s).  So each 32 x 16 piece goes through 4 KB of shared memory
// (row = 8 quads of dM | dV, quad index XOR-swizzled by the row so the row-per-lane stores are conflict-free) and comes back
// with 4 lanes per row: a warp access covers 8 rows x 64 contiguous bytes, whole sectors (0.56-0.63 ms; raw GEMM 0.42-0.48,
// GEMM + separate update pass 0.69-0.74 on the same boxes).  Measured and dropped (profiles/r02_ab_wide_calls.txt): a 2-deep
// register pipeline of the nine loads (spills at the 168-register cap of a 10-warp CTA), an L2 prefetch of the next tile's
// parameters during the MMA wait, streaming (evict-first) hints -- each 5-15 % SLOWER: the fused kernel moves
// 1.2 GB of optimiser state next to ~0.75 GB of operand traffic and is bound by HBM / the power cap, not by load latency.
__device__ __forceinline__ void dw_adam_tile(const TcEpi& e, const chain::Consts& cc, const chain::KlC& kc, float& kl_acc, int64_t M,
                                             int64_t N, int64_t row0, int64_t col0, uint32_t tbase, float* __restrict__ stg,
                                             int lane) {
#pragma unroll 1
  for (int c = 0; c < 64 / EW; ++c) {
    float v1[EW], v2[EW];
    tmem_ld16(tbase + c * EW, v1);
    tmem_ld16(tbase + 128 + c * EW, v2);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      st4(stg + lane * 32 + ((jj ^ (lane & 7)) << 2), v1[4 * jj], v1[4 * jj + 1], v1[4 * jj + 2], v1[4 * jj + 3]);
      st4(stg + lane * 32 + (((4 + jj) ^ (lane & 7)) << 2), v2[4 * jj], v2[4 * jj + 1], v2[4 * jj + 2], v2[4 * jj + 3]);
    }
    __syncwarp();
    const int qd = lane & 3;
    const int64_t col = col0 + c * EW + qd * 4;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = it * 8 + (lane >> 2);
      const float4 dM = ld4(stg + rr * 32 + ((qd ^ (rr & 7)) << 2));
      const float4 dV = ld4(stg + rr * 32 + (((4 + qd) ^ (rr & 7)) << 2));
      const int64_t row = row0 + rr;
      if (row < M && col < N) dw_adam_quad(e, cc, kc, kl_acc, row * N + col, dM, dV);
    }
    __syncwarp();
  }
}

// sum of 32 per-lane values over the 32 lanes of a warp: lane l ends with the total of vals[l] (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum32(float (&vals)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = lane & s;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? vals[i] : vals[i + s];
      const float keep = up ? vals[i + s] : vals[i];
      vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return vals[0];
}

// one thread = one output row (TMEM lane), EW consecutive columns; sbias = this tile's b_mu[BN] | sigma_b^2[BN]
template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const TcEpi& e, const Noise& nz, const chain::Consts& cc, int64_t M, int64_t N,
                                               int64_t row, int64_t col0, const float* __restrict__ sbias, int cl,
                                               const float d1[EW], const float d2[EW]) {
  const bool row_ok = row < M;
  if (!row_ok && !(MODE == LBBNN_TC_EPI_DX && e.colsum_part)) return;      // (the bias sums below shuffle across the warp)
  const bool fullc = col0 + EW - 1 < N;
  if constexpr (MODE == LBBNN_TC_EPI_RAW) {
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
  if constexpr (MODE == LBBNN_TC_EPI_FWD) {
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
  if constexpr (MODE == LBBNN_TC_EPI_DX) {
  // dx = D1 + 2 x D2; through the relu that produced x; dS_prev = dx * dsf_prev
  float xv[EW], fv[EW];
  if (!row_ok) {
#pragma unroll
    for (int j = 0; j < EW; ++j) { xv[j] = 0.f; fv[j] = 0.f; }
  } else if (fullc && (N % 8 == 0)) {
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
    if (!row_ok || col0 + j >= N) g = 0.f;
    o[j] = g;
    o2[j] = g * fv[j];
  }
  if (e.colsum_part) {          // bias gradients of the layer below: sums over this warp's 32 rows, fp32, fixed order
    float vals[32];
#pragma unroll
    for (int j = 0; j < EW; ++j) { vals[j] = o[j]; vals[EW + j] = o2[j]; }
    const int lane = threadIdx.x & 31;
    const float tot = warp_transpose_sum32(vals, lane);
    const int64_t c = col0 + (lane & (EW - 1));
    const int64_t row0 = row - lane;                      // first row of this warp's 32-row group (always < M)
    if (row0 < M && c < N) e.colsum_part[(row0 >> 5) * (2 * N) + (lane < EW ? 0 : N) + c] = tot;
  }
  if (!row_ok) return;
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
template <int MODE, int AMN, int BMN>
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
          if constexpr (!AMN) {
            tma_load_2d(sa + 0 * kTileBytes, &tmA1, &full_bar[stage], kb * BK, m0);
            tma_load_2d(sa + 1 * kTileBytes, &tmA2, &full_bar[stage], kb * BK, m0);
          } else {   // (K, M) row-major: two boxes of 64 rows of K x 64 elements of M per operand
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              tma_load_2d(sa + 0 * kTileBytes + h * kMnBoxBytes, &tmA1, &full_bar[stage], m0 + 64 * h, kb * BK);
              tma_load_2d(sa + 1 * kTileBytes + h * kMnBoxBytes, &tmA2, &full_bar[stage], m0 + 64 * h, kb * BK);
            }
          }
          if constexpr (!BMN) {
            tma_load_2d(sa + 2 * kTileBytes, &tmB1, &full_bar[stage], kb * BK, n0);
            tma_load_2d(sa + 3 * kTileBytes, &tmB2, &full_bar[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              tma_load_2d(sa + 2 * kTileBytes + h * kMnBoxBytes, &tmB1, &full_bar[stage], n0 + 64 * h, kb * BK);
              tma_load_2d(sa + 3 * kTileBytes + h * kMnBoxBytes, &tmB2, &full_bar[stage], n0 + 64 * h, kb * BK);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int it = 0;
      constexpr uint32_t idesc = make_idesc(BM, BN, AMN, BMN);
      constexpr uint64_t ka = AMN ? (uint64_t)(2048 >> 4) : (uint64_t)((UMMA_K * 2) >> 4);   // descriptor advance per K = 16 MMA
      constexpr uint64_t kbs = BMN ? (uint64_t)(2048 >> 4) : (uint64_t)((UMMA_K * 2) >> 4);
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);   // epilogue drained this accumulator stage
        tc_fence_after();
        const uint32_t d1 = tmem_base + as * 256, d2 = d1 + 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t a1 = umma_desc(sa + 0 * kTileBytes, AMN), a2 = umma_desc(sa + 1 * kTileBytes, AMN);
          const uint64_t b1 = umma_desc(sa + 2 * kTileBytes, BMN), b2 = umma_desc(sa + 3 * kTileBytes, BMN);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: advance the start address inside the swizzle row; MN-major: by two 8-row groups
            const uint32_t acc = (kb | k) ? 1u : 0u;
            umma_bf16(d1, a1 + k * ka, b1 + k * kbs, idesc, acc);
            umma_bf16(d2, a2 + k * ka, b2 + k * kbs, idesc, acc);
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
    chain::Consts cc = {};
    chain::KlC kc = {};
    float kl_acc = 0.f;
    if constexpr (MODE == LBBNN_TC_EPI_DW_ADAM) {
      cc = chain::make_consts(epi.pri, epi.var_mode, epi.klg, epi.beta1, epi.beta2, epi.adam_eps, epi.coef);
      kc = chain::make_klc(epi.pri);
    }
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;                      // 0..255 among the epilogue threads
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      int mb, nb;
      tile_coords(t, num_m, num_n, mb, nb);
      const int m0 = mb * BM, n0 = nb * BN;
      float* sbias = sbias_all + as * 2 * BN;
      if constexpr (MODE == LBBNN_TC_EPI_FWD) {                 // this tile's bias terms, once per column
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
      if constexpr (MODE == LBBNN_TC_EPI_DW_ADAM) {
        dw_adam_tile(epi, cc, kc, kl_acc, M, N, m0 + q * 32, n0 + half * 64, tbase, sbias_all + (warp - 2) * 1024, lane);
      } else {
#pragma unroll 1
        for (int c = 0; c < 64 / EW; ++c) {
          float v1[EW], v2[EW];
          tmem_ld16(tbase + c * EW, v1);
          tmem_ld16(tbase + 128 + c * EW, v2);
          const int cl = half * 64 + c * EW;
          epilogue_chunk<MODE>(epi, nz, cc, M, N, row, n0 + cl, sbias, cl, v1, v2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    if constexpr (MODE == LBBNN_TC_EPI_DW_ADAM) {
      if (epi.next_kl_part) {     // this warp's share of the next step's KL (fixed tile -> CTA -> warp assignment: deterministic)
        const float w = warp_sum(kl_acc);
        if (lane == 0) epi.next_kl_part[(int64_t)blockIdx.x * kEpiWarps + (warp - 2)] = (double)w;
      }
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
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + 1024 + 256 + kEpiScratchBytes;
static_assert(kSmemBytes <= 232448 && kSmemBytes2 <= 232448, "over the 227 KB shared-memory limit");
constexpr int kGroupM2 = kGroupM / 2;                          // in 256-row blocks
constexpr int kL2Prefetch = 0;                                 // default TMA L2 prefetch distance (K blocks); see TcEpi::l2_prefetch

__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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

template <int MODE, int AMN, int BMN>
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
          if (epi.l2_prefetch > 0) {       // the boxes of K block kb + distance (this tile, or the head of this cluster's next one)
            int pk = kb + epi.l2_prefetch, pm0 = m0, pn0 = n0h;
            bool ok = true;
            if (pk >= num_kb) {
              pk -= num_kb;
              const int tn = t + num_clusters;
              ok = tn < num_tiles && pk < num_kb;
              if (ok) {
                int pmb2, pnb;
                tile_coords2(tn, num_m2, num_n, group, pmb2, pnb);
                pm0 = pmb2 * 2 * BM + (int)rank * BM; pn0 = pnb * BN + (int)rank * (BN / 2);
              }
            }
            if (ok) {
              if constexpr (!AMN) { tma_prefetch_l2_2d(&tmA1, pk * BK, pm0); tma_prefetch_l2_2d(&tmA2, pk * BK, pm0); }
              else {
#pragma unroll
                for (int h = 0; h < 2; ++h) { tma_prefetch_l2_2d(&tmA1, pm0 + 64 * h, pk * BK); tma_prefetch_l2_2d(&tmA2, pm0 + 64 * h, pk * BK); }
              }
              if constexpr (!BMN) { tma_prefetch_l2_2d(&tmB1, pk * BK, pn0); tma_prefetch_l2_2d(&tmB2, pk * BK, pn0); }
              else { tma_prefetch_l2_2d(&tmB1, pn0, pk * BK); tma_prefetch_l2_2d(&tmB2, pn0, pk * BK); }
            }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes2;
          const uint32_t lead_full = mapa_u32(&full_bar[stage], 0);
          mbar_expect_tx_cluster(lead_full, kStageBytes2);
          if constexpr (!AMN) {
            tma_load_2d_pair(sa, &tmA1, lead_full, kb * BK, m0);
            tma_load_2d_pair(sa + kTileBytes, &tmA2, lead_full, kb * BK, m0);
          } else {   // (K, M) row-major: two boxes of 64 rows of K x 64 elements of M per operand
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              tma_load_2d_pair(sa + h * kMnBoxBytes, &tmA1, lead_full, m0 + 64 * h, kb * BK);
              tma_load_2d_pair(sa + kTileBytes + h * kMnBoxBytes, &tmA2, lead_full, m0 + 64 * h, kb * BK);
            }
          }
          if constexpr (!BMN) {
            tma_load_2d_pair(sa + 2 * kTileBytes, &tmB1, lead_full, kb * BK, n0h);
            tma_load_2d_pair(sa + 2 * kTileBytes + kBHalfBytes, &tmB2, lead_full, kb * BK, n0h);
          } else {   // this CTA's 64 elements of N: one box
            tma_load_2d_pair(sa + 2 * kTileBytes, &tmB1, lead_full, n0h, kb * BK);
            tma_load_2d_pair(sa + 2 * kTileBytes + kBHalfBytes, &tmB2, lead_full, n0h, kb * BK);
          }
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0; int it = 0;
      constexpr uint32_t idesc = make_idesc(2 * BM, BN, AMN, BMN);
      constexpr uint64_t ka = AMN ? (uint64_t)(2048 >> 4) : (uint64_t)((UMMA_K * 2) >> 4);   // descriptor advance per K = 16 MMA
      constexpr uint64_t kbs = BMN ? (uint64_t)(2048 >> 4) : (uint64_t)((UMMA_K * 2) >> 4);
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);   // both CTAs' epilogues drained this accumulator stage
        tc_fence_after();
        const uint32_t d1 = tmem_base + as * 256, d2 = d1 + 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes2);
          const uint64_t a1 = umma_desc(sa, AMN), a2 = umma_desc(sa + kTileBytes, AMN);
          const uint64_t b1 = umma_desc(sa + 2 * kTileBytes, BMN);
          const uint64_t b2 = umma_desc(sa + 2 * kTileBytes + kBHalfBytes, BMN);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t acc = (kb | k) ? 1u : 0u;
            umma_bf16_pair(d1, a1 + k * ka, b1 + k * kbs, idesc, acc);
            umma_bf16_pair(d2, a2 + k * ka, b2 + k * kbs, idesc, acc);
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
    chain::Consts cc = {};
    chain::KlC kc = {};
    float kl_acc = 0.f;
    if constexpr (MODE == LBBNN_TC_EPI_DW_ADAM) {
      cc = chain::make_consts(epi.pri, epi.var_mode, epi.klg, epi.beta1, epi.beta2, epi.adam_eps, epi.coef);
      kc = chain::make_klc(epi.pri);
    }
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      const int as = it & 1;
      int mb2, nb;
      tile_coords2(t, num_m2, num_n, group, mb2, nb);
      const int m0 = mb2 * 2 * BM + (int)rank * BM, n0 = nb * BN;
      float* sbias = sbias_all + as * 2 * BN;
      if constexpr (MODE == LBBNN_TC_EPI_FWD) {
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
      if constexpr (MODE == LBBNN_TC_EPI_DW_ADAM) {
        dw_adam_tile(epi, cc, kc, kl_acc, M, N, m0 + q * 32, n0 + half * 64, tbase, sbias_all + (warp - 2) * 1024, lane);
      } else {
#pragma unroll 1
        for (int c = 0; c < 64 / EW; ++c) {
          float v1[EW], v2[EW];
          tmem_ld16(tbase + c * EW, v1);
          tmem_ld16(tbase + 128 + c * EW, v2);
          const int cl = half * 64 + c * EW;
          epilogue_chunk<MODE>(epi, nz, cc, M, N, row, n0 + cl, sbias, cl, v1, v2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[as], 0));
    }
    if constexpr (MODE == LBBNN_TC_EPI_DW_ADAM) {
      if (epi.next_kl_part) {
        const float w = warp_sum(kl_acc);
        if (lane == 0) epi.next_kl_part[(int64_t)blockIdx.x * kEpiWarps + (warp - 2)] = (double)w;
      }
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

// MN-major operand: the tensor is (K, rows) row-major bf16; box = 64 (rows) x 64 (K), 128B swizzle, OOB -> zeros
int make_map_mn(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K) {
  EncodeTiledFn enc = get_encode();
  LBBNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  LBBNN_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (rows * 2) % 16 == 0,
                "MN-major TMA operand must be 16B aligned with a 16B-multiple row pitch (rows %% 8 == 0), rows=%lld", (long long)rows);
  cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)K};
  cuuint64_t strides[1] = {(cuuint64_t)rows * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LBBNN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (MN-major) failed (%d) rows=%lld K=%lld", (int)r, (long long)rows, (long long)K);
  return LBBNN_OK;
}

int launch_tc(const void* A1, const void* A2, const void* B1, const void* B2, int64_t M, int64_t N, int64_t K, const TcEpi& epi_in,
              cudaStream_t st) {
  LBBNN_REQUIRE(A1 && A2 && B1 && B2 && M > 0 && N > 0 && K > 0, "bad GEMM operands");
  TcEpi epi = epi_in;
  {
    const char* pf = getenv("LBBNN_TC_L2_PREFETCH");     // K blocks ahead; default kL2Prefetch
    epi.l2_prefetch = pf ? atoi(pf) : kL2Prefetch;
  }
  LBBNN_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "GEMM dims must fit int32");
  CUtensorMap mA1, mA2, mB1, mB2;
  if (epi.a_mn) {
    if (int rc = make_map_mn(&mA1, A1, M, K)) return rc;
    if (int rc = make_map_mn(&mA2, A2, M, K)) return rc;
  } else {
    if (int rc = make_map(&mA1, A1, M, K)) return rc;
    if (int rc = make_map(&mA2, A2, M, K)) return rc;
  }
  // CTA pairs (256-row tiles) for problems that fill the GPU with them; LBBNN_TC_PAIR=0 forces the 1-CTA kernel
  // (2 = also for small problems: the tests use it to run odd shapes through the pair kernel; read per call, host only)
  const char* pe = getenv("LBBNN_TC_PAIR");
  const int pair_mode = pe ? atoi(pe) : 1;
  const int64_t tiles2 = ceil_div(M, 2 * BM) * ceil_div(N, BN);
  if (pair_mode && M > BM && (pair_mode == 2 || tiles2 >= sm_count() / 2)) {
    if (epi.b_mn) {
      if (int rc = make_map_mn(&mB1, B1, N, K)) return rc;
      if (int rc = make_map_mn(&mB2, B2, N, K)) return rc;
    } else {
      if (int rc = make_map(&mB1, B1, N, K, BN / 2)) return rc;
      if (int rc = make_map(&mB2, B2, N, K, BN / 2)) return rc;
    }
    const int clusters = (int)(tiles2 < sm_count() / 2 ? tiles2 : sm_count() / 2);
    const char* ge = getenv("LBBNN_TC_GROUP");      // 256-row blocks per tile group (experiments)
    const int group = ge && atoi(ge) > 0 ? atoi(ge) : kGroupM2;
    const int key = epi.mode * 4 + epi.a_mn * 2 + epi.b_mn;
    switch (key) {
#define LBBNN_PAIR_CASE(MODE_, A_, B_)                                                                                                \
  case (MODE_) * 4 + (A_) * 2 + (B_): {                                                                                              \
    static bool attr_set_ = false;                                                                                                   \
    if (!attr_set_) {                                                                                                                \
      LBBNN_CUDA(cudaFuncSetAttribute(tc_dual_gemm_bf16_pair<MODE_, A_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                      smem_for(kSmemBytes2, MODE_)));                                                               \
      attr_set_ = true;                                                                                                              \
    }                                                                                                                                \
    tc_dual_gemm_bf16_pair<MODE_, A_, B_><<<2 * clusters, kTcThreads, smem_for(kSmemBytes2, MODE_), st>>>(mA1, mA2, mB1, mB2, epi, \
                                                                                                           (int)M, (int)N, (int)K, \
                                                                                                           group);                 \
    break;                                                                                                                           \
  }
      LBBNN_PAIR_CASE(LBBNN_TC_EPI_RAW, 0, 0)
      LBBNN_PAIR_CASE(LBBNN_TC_EPI_RAW, 0, 1)
      LBBNN_PAIR_CASE(LBBNN_TC_EPI_RAW, 1, 0)
      LBBNN_PAIR_CASE(LBBNN_TC_EPI_RAW, 1, 1)
      LBBNN_PAIR_CASE(LBBNN_TC_EPI_FWD, 0, 0)
      LBBNN_PAIR_CASE(LBBNN_TC_EPI_DX, 0, 0)
      LBBNN_PAIR_CASE(LBBNN_TC_EPI_DX, 0, 1)
      LBBNN_PAIR_CASE(LBBNN_TC_EPI_DW_ADAM, 1, 1)
#undef LBBNN_PAIR_CASE
      default: LBBNN_REQUIRE(false, "no kernel for epilogue mode %d with operand majors (%d, %d)", epi.mode, epi.a_mn, epi.b_mn);
    }
    return check_launch("tc_dual_gemm_bf16_pair");
  }
  if (epi.b_mn) {
    if (int rc = make_map_mn(&mB1, B1, N, K)) return rc;
    if (int rc = make_map_mn(&mB2, B2, N, K)) return rc;
  } else {
    if (int rc = make_map(&mB1, B1, N, K)) return rc;
    if (int rc = make_map(&mB2, B2, N, K)) return rc;
  }
  const int64_t tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  const int key = epi.mode * 4 + epi.a_mn * 2 + epi.b_mn;
  switch (key) {
#define LBBNN_ONE_CASE(MODE_, A_, B_)                                                                                          \
  case (MODE_) * 4 + (A_) * 2 + (B_): {                                                                                       \
    static bool attr_set_ = false;                                                                                             \
    if (!attr_set_) {                                                                                                          \
      LBBNN_CUDA(cudaFuncSetAttribute(tc_dual_gemm_bf16<MODE_, A_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                      smem_for(kSmemBytes, MODE_)));                                                          \
      attr_set_ = true;                                                                                                        \
    }                                                                                                                          \
    tc_dual_gemm_bf16<MODE_, A_, B_><<<grid, kTcThreads, smem_for(kSmemBytes, MODE_), st>>>(mA1, mA2, mB1, mB2, epi, (int)M,   \
                                                                                             (int)N, (int)K);                   \
    break;                                                                                                                     \
  }
    LBBNN_ONE_CASE(LBBNN_TC_EPI_RAW, 0, 0)
    LBBNN_ONE_CASE(LBBNN_TC_EPI_RAW, 0, 1)
    LBBNN_ONE_CASE(LBBNN_TC_EPI_RAW, 1, 0)
    LBBNN_ONE_CASE(LBBNN_TC_EPI_RAW, 1, 1)
    LBBNN_ONE_CASE(LBBNN_TC_EPI_FWD, 0, 0)
    LBBNN_ONE_CASE(LBBNN_TC_EPI_DX, 0, 0)
    LBBNN_ONE_CASE(LBBNN_TC_EPI_DX, 0, 1)
    LBBNN_ONE_CASE(LBBNN_TC_EPI_DW_ADAM, 1, 1)
#undef LBBNN_ONE_CASE
    default: LBBNN_REQUIRE(false, "no kernel for epilogue mode %d with operand majors (%d, %d)", epi.mode, epi.a_mn, epi.b_mn);
  }
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

// ---- r02: operands in place (MN-major), fused dW update, bias sums from the dX epilogue ---------------------------------
namespace lbbnn {
namespace {
// out[c] = sum over the per-32-row partial sums written by the dX epilogue, fixed order
__global__ void __launch_bounds__(256) colsum_part_reduce(const float* __restrict__ part, int n_part, int64_t cols2,
                                                          float* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols2) return;
  float s = 0.f;
  for (int r = 0; r < n_part; ++r) s += part[(int64_t)r * cols2 + c];
  out[c] = s;
}
}  // namespace
}  // namespace lbbnn

extern "C" int lbbnn_tc_dual_gemm_raw_ex(const void* A1, const void* A2, const void* B1, const void* B2, int64_t M, int64_t N,
                                         int64_t K, int a_mn, int b_mn, float* D1, float* D2, lbbnn_stream s) {
  LBBNN_REQUIRE(D1 && D2, "NULL output");
  TcEpi e = {};
  e.mode = LBBNN_TC_EPI_RAW;
  e.d1 = D1; e.d2 = D2;
  e.a_mn = a_mn ? 1 : 0; e.b_mn = b_mn ? 1 : 0;
  return launch_tc(A1, A2, B1, B2, M, N, K, e, (cudaStream_t)s);
}

extern "C" size_t lbbnn_tc_colsum_part_floats(int64_t batch, int64_t in_features) {
  return (size_t)(ceil_div(batch, 32) * 2 * in_features);
}

extern "C" int lbbnn_tc_colsum_reduce(const float* part, int64_t batch, int64_t in_features, float* colsum, lbbnn_stream s) {
  LBBNN_REQUIRE(part && colsum && batch > 0 && in_features > 0, "bad argument");
  const int64_t cols2 = 2 * in_features;
  colsum_part_reduce<<<(unsigned)ceil_div(cols2, 256), 256, 0, (cudaStream_t)s>>>(part, (int)ceil_div(batch, 32), cols2, colsum);
  return check_launch("colsum_part_reduce");
}

extern "C" int lbbnn_tc_lrt_bwd_input_mn(const void* dE_bf, const void* dS_bf, const void* M_bf, const void* V_bf, int64_t batch,
                                         int64_t in_features, int64_t out_features, const void* x_bf, const float* ds_factor_prev,
                                         int flags, void* dE_prev_bf, void* dS_prev_bf, float* colsum_part, lbbnn_stream s) {
  LBBNN_REQUIRE(x_bf && dE_prev_bf && dS_prev_bf, "NULL argument");
  TcEpi e = {};
  e.mode = LBBNN_TC_EPI_DX; e.flags = flags;
  e.x_bf = (const __nv_bfloat16*)x_bf; e.dsf_prev = ds_factor_prev;
  e.de = (__nv_bfloat16*)dE_prev_bf; e.ds = (__nv_bfloat16*)dS_prev_bf;
  e.colsum_part = colsum_part;
  // contraction over the out features: A = (dE, dS) (batch, out) K-major; B = (M, V) (out, in) = MN-major (N = in)
  e.a_mn = 0; e.b_mn = 1;
  return launch_tc(dE_bf, dS_bf, M_bf, V_bf, batch, in_features, out_features, e, (cudaStream_t)s);
}

extern "C" size_t lbbnn_tc_lrt_dw_adam_kl_parts(void) { return (size_t)sm_count() * kEpiWarps; }

extern "C" int lbbnn_tc_lrt_dw_adam_next(const void* dE_bf, const void* dS_bf, const void* x_bf, const void* x2_bf,
                                         const lbbnn_layer* L, int64_t batch, const lbbnn_priors* pri, int var_mode,
                                         float kl_grad, const lbbnn_adam_layer_state* adam, void* next_M_bf, void* next_V_bf,
                                         double* next_kl_part, lbbnn_stream s);

extern "C" int lbbnn_tc_lrt_dw_adam(const void* dE_bf, const void* dS_bf, const void* x_bf, const void* x2_bf,
                                    const lbbnn_layer* L, int64_t batch, const lbbnn_priors* pri, int var_mode,
                                    float kl_grad, const lbbnn_adam_layer_state* adam, lbbnn_stream s) {
  return lbbnn_tc_lrt_dw_adam_next(dE_bf, dS_bf, x_bf, x2_bf, L, batch, pri, var_mode, kl_grad, adam, nullptr, nullptr, nullptr, s);
}

extern "C" int lbbnn_tc_lrt_dw_adam_next(const void* dE_bf, const void* dS_bf, const void* x_bf, const void* x2_bf,
                                         const lbbnn_layer* L, int64_t batch, const lbbnn_priors* pri, int var_mode,
                                         float kl_grad, const lbbnn_adam_layer_state* adam, void* next_M_bf, void* next_V_bf,
                                         double* next_kl_part, lbbnn_stream s) {
  LBBNN_REQUIRE((next_M_bf == nullptr) == (next_V_bf == nullptr), "next_M / next_V come in pairs");
  LBBNN_REQUIRE(next_kl_part == nullptr || next_M_bf != nullptr, "next_kl_part needs next_M / next_V");
  LBBNN_REQUIRE(L && pri && adam && adam->coef, "NULL argument");
  LBBNN_REQUIRE(L->weight_mu && L->weight_rho && L->lambdal && L->in_features > 0 && L->out_features > 0, "bad layer");
  LBBNN_REQUIRE(L->z == nullptr && L->z_kl == nullptr, "the fused update is for LRT layers (no multiplicative z)");
  LBBNN_REQUIRE(L->in_features % 8 == 0 && L->out_features % 8 == 0, "fused dW update needs in/out features divisible by 8");
  for (int i = 0; i < 3; ++i) LBBNN_REQUIRE(adam->exp_avg[i] && adam->exp_avg_sq[i], "NULL Adam state %d", i);
  TcEpi e = {};
  e.mode = LBBNN_TC_EPI_DW_ADAM;
  e.p_mu = const_cast<float*>(L->weight_mu); e.p_rho = const_cast<float*>(L->weight_rho); e.p_lam = const_cast<float*>(L->lambdal);
  e.m_mu = adam->exp_avg[0]; e.m_rho = adam->exp_avg[1]; e.m_lam = adam->exp_avg[2];
  e.v_mu = adam->exp_avg_sq[0]; e.v_rho = adam->exp_avg_sq[1]; e.v_lam = adam->exp_avg_sq[2];
  const float* ptrs[9] = {e.p_mu, e.p_rho, e.p_lam, e.m_mu, e.m_rho, e.m_lam, e.v_mu, e.v_rho, e.v_lam};
  for (int i = 0; i < 9; ++i) LBBNN_REQUIRE((reinterpret_cast<uintptr_t>(ptrs[i]) & 15) == 0, "parameters / Adam state must be 16B aligned");
  e.coef = adam->coef; e.pri = *pri; e.var_mode = var_mode; e.klg = kl_grad;
  e.beta1 = adam->beta1; e.beta2 = adam->beta2; e.adam_eps = adam->eps;
  e.next_M = (__nv_bfloat16*)next_M_bf; e.next_V = (__nv_bfloat16*)next_V_bf; e.next_kl_part = next_kl_part;
  if (next_kl_part)   // CTAs without a tile (small problems) leave their slots untouched: clear all of them first
    LBBNN_CUDA(cudaMemsetAsync(next_kl_part, 0, lbbnn_tc_lrt_dw_adam_kl_parts() * sizeof(double), (cudaStream_t)s));
  // dM = dE^T x, dV = dS^T x^2 (out, in), contraction over the batch: both operands read in place as MN-major
  e.a_mn = 1; e.b_mn = 1;
  return launch_tc(dE_bf, dS_bf, x_bf, x2_bf, L->out_features, L->in_features, batch, e, (cudaStream_t)s);
}
