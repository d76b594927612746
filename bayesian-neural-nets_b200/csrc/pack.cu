// Operand staging for the bf16 tensor-core path: fp32 -> bf16 operand pairs (+ transposes), and the
// column sums over the batch that feed the bias gradients.
#include "common.cuh"

#include <algorithm>

namespace lbbnn {
namespace {

// 32x32 tile per block (32x8 threads): row-major outputs straight from registers, transposed
// outputs through a padded smem tile so both directions are coalesced.
__global__ void __launch_bounds__(256) bf16_pack_kernel(const float* __restrict__ a, const float* __restrict__ b, int op,
                                                        int64_t rows, int64_t cols, __nv_bfloat16* __restrict__ o1,
                                                        __nv_bfloat16* __restrict__ o2, __nv_bfloat16* __restrict__ o1T,
                                                        __nv_bfloat16* __restrict__ o2T) {
  __shared__ float t1[32][33], t2[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + i * 8, c = c0 + tx;
    float v1 = 0.f, v2 = 0.f;
    if (r < rows && c < cols) {
      v1 = a[r * cols + c];
      v2 = op == LBBNN_PACK_PAIR ? b[r * cols + c] : (op == LBBNN_PACK_SQUARE ? v1 * v1 : v1 * b[r * cols + c]);
      if (o1) o1[r * cols + c] = __float2bfloat16_rn(v1);
      if (o2) o2[r * cols + c] = __float2bfloat16_rn(v2);
    }
    t1[ty + i * 8][tx] = v1;
    t2[ty + i * 8][tx] = v2;
  }
  if (o1T == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t c = c0 + ty + i * 8, r = r0 + tx;   // output (cols, rows): row index c, column index r
    if (c < cols && r < rows) {
      o1T[c * rows + r] = __float2bfloat16_rn(t1[tx][ty + i * 8]);
      o2T[c * rows + r] = __float2bfloat16_rn(t2[tx][ty + i * 8]);
    }
  }
}

// Same job on 64x64 tiles for shapes with cols % 4 == 0: float4 loads, 8-byte bf16 stores in both directions (the
// transposes go through bf16 shared-memory tiles).  The 32x32 kernel above stays for ragged shapes.
__device__ __forceinline__ uint2 pack4v(float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

__global__ void __launch_bounds__(256) bf16_pack64_kernel(const float* __restrict__ a, const float* __restrict__ b, int op,
                                                          int64_t rows, int64_t cols, __nv_bfloat16* __restrict__ o1,
                                                          __nv_bfloat16* __restrict__ o2, __nv_bfloat16* __restrict__ o1T,
                                                          __nv_bfloat16* __restrict__ o2T) {
  __shared__ __align__(8) __nv_bfloat16 t1[64][68], t2[64][68];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t c0 = (int64_t)blockIdx.x * 64, r0 = (int64_t)blockIdx.y * 64;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = ty + 16 * i;
    const int64_t r = r0 + rl, c = c0 + tx * 4;
    uint2 p1 = make_uint2(0u, 0u), p2 = make_uint2(0u, 0u);
    if (r < rows && c < cols) {
      const int64_t e = r * cols + c;
      const float4 v = __ldg(reinterpret_cast<const float4*>(a + e));
      float4 w;
      if (op == LBBNN_PACK_SQUARE) {
        w = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
      } else {
        w = __ldg(reinterpret_cast<const float4*>(b + e));
        if (op == LBBNN_PACK_SCALE) { w.x *= v.x; w.y *= v.y; w.z *= v.z; w.w *= v.w; }
      }
      p1 = pack4v(v.x, v.y, v.z, v.w);
      p2 = pack4v(w.x, w.y, w.z, w.w);
      if (o1) *reinterpret_cast<uint2*>(o1 + e) = p1;
      if (o2) *reinterpret_cast<uint2*>(o2 + e) = p2;
    }
    *reinterpret_cast<uint2*>(&t1[rl][tx * 4]) = p1;
    *reinterpret_cast<uint2*>(&t2[rl][tx * 4]) = p2;
  }
  if (o1T == nullptr) return;
  __syncthreads();
  const bool vec = (rows % 4 == 0);
  auto bits = [](const __nv_bfloat16 lo, const __nv_bfloat16 hi) {
    return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
  };
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cl = ty + 16 * i, rl = tx * 4;
    const int64_t c = c0 + cl, r = r0 + rl;
    if (c >= cols || r >= rows) continue;
    const __nv_bfloat16 u[4] = {t1[rl][cl], t1[rl + 1][cl], t1[rl + 2][cl], t1[rl + 3][cl]};
    const __nv_bfloat16 w[4] = {t2[rl][cl], t2[rl + 1][cl], t2[rl + 2][cl], t2[rl + 3][cl]};
    if (vec) {
      *reinterpret_cast<uint2*>(o1T + c * rows + r) = make_uint2(bits(u[0], u[1]), bits(u[2], u[3]));
      *reinterpret_cast<uint2*>(o2T + c * rows + r) = make_uint2(bits(w[0], w[1]), bits(w[2], w[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r + j < rows) { o1T[c * rows + r + j] = u[j]; o2T[c * rows + r + j] = w[j]; }
    }
  }
}

// column sums in two deterministic stages: stage 1 = grid (cols/32, S row slices), each block sums its
// slice with 8 row-lanes per column and a fixed-order smem reduction -> partial[s][2][cols];
// stage 2 sums the S partials in order.
template <bool BF16>
__global__ void __launch_bounds__(256) colsum2_stage1(const void* __restrict__ a_, const void* __restrict__ b_, int64_t rows,
                                                      int64_t cols, int64_t rows_per_slice, float* __restrict__ partial) {
  __shared__ float s1[8][33], s2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  const int64_t rbeg = (int64_t)blockIdx.y * rows_per_slice, rend = min(rows, rbeg + rows_per_slice);
  float acc1 = 0.f, acc2 = 0.f;
  if (c < cols) {
#pragma unroll 4
    for (int64_t r = rbeg + ty; r < rend; r += 8) {
      if (BF16) {
        acc1 += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a_)[r * cols + c]);
        acc2 += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(b_)[r * cols + c]);
      } else {
        const float v = __ldg(reinterpret_cast<const float*>(a_) + r * cols + c);
        acc1 += v;
        acc2 += b_ ? v * __ldg(reinterpret_cast<const float*>(b_) + r * cols + c) : 0.f;
      }
    }
  }
  s1[ty][tx] = acc1;
  s2[ty][tx] = acc2;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r1 += s1[i][tx]; r2 += s2[i][tx]; }
    partial[((int64_t)blockIdx.y * 2 + 0) * cols + c] = r1;
    partial[((int64_t)blockIdx.y * 2 + 1) * cols + c] = r2;
  }
}

__global__ void __launch_bounds__(256) colsum2_stage2(const float* __restrict__ partial, int slices, int64_t cols,
                                                      float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over 2*cols
  if (i >= 2 * cols) return;
  const int64_t which = i / cols, c = i % cols;
  float acc = 0.f;
  for (int s = 0; s < slices; ++s) acc += partial[((int64_t)s * 2 + which) * cols + c];
  out[i] = acc;
}

// Input gradient of a layer with FEW outputs (the classifier: out <= 12, 0.04 % of the step's FLOPs) fused with the staging
// of the previous layer's backward operands -- what lbbnn_tc_lrt_bwd_input's epilogue does for the wide layers:
//   dx = g M + 2 x .* (gS V),  gS = g .* ds;  through the relu that produced x;  dE = dx, dS = dx .* ds_prev
// written as bf16 (batch, in) and transposed (in, batch), plus this tile's column sums of the fp32 dE, dS (the previous
// layer's bias gradients).  64 batch rows x 64 input columns per block; M, V and the tile's g, gS sit in shared memory.
constexpr int kSmallMaxOut = 12;      // static shared memory: two 64x65 transpose tiles + M, V, g, gS stay under 48 KB

struct BwdSmallArgs {
  const float *g, *ds, *M, *V;            // (B, O), (B, O), (O, K), (O, K)
  const __nv_bfloat16* x;                 // (B, K)
  const float* ds_prev;                   // (B, K)
  int64_t B, K;
  int O, mask;
  __nv_bfloat16 *de, *dse, *deT, *dsT;
  float* partial;                         // [gridDim.y][2][K]
};

__device__ __forceinline__ uint2 pack4(const float v[4]) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
  return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

__global__ void __launch_bounds__(256, 3) lrt_bwd_input_small_kernel(const BwdSmallArgs a) {
  __shared__ float tE[64][65], tS[64][65];
  __shared__ __align__(16) float sM[kSmallMaxOut][64], sV[kSmallMaxOut][64];
  __shared__ float sG[64][kSmallMaxOut + 1], sGS[64][kSmallMaxOut + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t c0 = (int64_t)blockIdx.x * 64, r0 = (int64_t)blockIdx.y * 64;
  // this thread's x and ds_prev quads (4 rows): issued first so that HBM latency hides behind the shared-memory fill and the
  // contraction below
  uint2 xq[4];
  float4 fq[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + 16 * i, c = c0 + tx * 4;
    xq[i] = make_uint2(0u, 0u);
    fq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < a.B && c < a.K) {
      xq[i] = __ldg(reinterpret_cast<const uint2*>(a.x + r * a.K + c));
      if (a.ds_prev) fq[i] = __ldg(reinterpret_cast<const float4*>(a.ds_prev + r * a.K + c));
    }
  }
  for (int e = tid; e < a.O * 64; e += 256) {
    const int o = e >> 6, c = e & 63;
    const bool ok = c0 + c < a.K;
    sM[o][c] = ok ? __ldg(a.M + (int64_t)o * a.K + c0 + c) : 0.f;
    sV[o][c] = ok ? __ldg(a.V + (int64_t)o * a.K + c0 + c) : 0.f;
  }
  for (int e = tid; e < 64 * a.O; e += 256) {
    const int rl = e / a.O, o = e % a.O;
    const int64_t r = r0 + rl;
    const float gv = r < a.B ? __ldg(a.g + r * a.O + o) : 0.f;
    sG[rl][o] = gv;
    sGS[rl][o] = r < a.B ? gv * __ldg(a.ds + r * a.O + o) : 0.f;
  }
  __syncthreads();
  // the two contractions over the layer's few outputs for this thread's 4 rows x 4 columns: M, V quads read once per
  // output, the rows' g, gS are warp-wide broadcasts
  float e1[4][4], e2[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) e1[i][j] = e2[i][j] = 0.f;
  for (int o = 0; o < a.O; ++o) {
    const float4 m = *reinterpret_cast<const float4*>(&sM[o][tx * 4]);
    const float4 v = *reinterpret_cast<const float4*>(&sV[o][tx * 4]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float gv = sG[ty + 16 * i][o], gs = sGS[ty + 16 * i][o];
      e1[i][0] = fmaf(gv, m.x, e1[i][0]); e1[i][1] = fmaf(gv, m.y, e1[i][1]);
      e1[i][2] = fmaf(gv, m.z, e1[i][2]); e1[i][3] = fmaf(gv, m.w, e1[i][3]);
      e2[i][0] = fmaf(gs, v.x, e2[i][0]); e2[i][1] = fmaf(gs, v.y, e2[i][1]);
      e2[i][2] = fmaf(gs, v.z, e2[i][2]); e2[i][3] = fmaf(gs, v.w, e2[i][3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rl = ty + 16 * i;
    const int64_t r = r0 + rl, c = c0 + tx * 4;
    float dE[4] = {0.f, 0.f, 0.f, 0.f}, dS[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < a.B && c < a.K) {                  // K % 4 == 0: whole quad inside the row
      const int64_t e = r * a.K + c;
      const float2 x01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xq[i].x));
      const float2 x23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xq[i].y));
      const float xv[4] = {x01.x, x01.y, x23.x, x23.y};
      const float fv[4] = {fq[i].x, fq[i].y, fq[i].z, fq[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float gx = fmaf(2.0f * xv[j], e2[i][j], e1[i][j]);
        if (a.mask && !(xv[j] > 0.f)) gx = 0.f;
        dE[j] = gx;
        dS[j] = gx * fv[j];
      }
      *reinterpret_cast<uint2*>(a.de + e) = pack4(dE);
      *reinterpret_cast<uint2*>(a.dse + e) = pack4(dS);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { tE[rl][tx * 4 + j] = dE[j]; tS[rl][tx * 4 + j] = dS[j]; }
  }
  __syncthreads();
  const bool vec = (a.B % 4 == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cl = ty + 16 * i, rl = tx * 4;             // output row = input column, 4 consecutive batch rows
    const int64_t c = c0 + cl, r = r0 + rl;
    const float e[4] = {tE[rl][cl], tE[rl + 1][cl], tE[rl + 2][cl], tE[rl + 3][cl]};
    const float q[4] = {tS[rl][cl], tS[rl + 1][cl], tS[rl + 2][cl], tS[rl + 3][cl]};
    if (c < a.K && r < a.B && a.deT) {
      if (vec) {
        *reinterpret_cast<uint2*>(a.deT + c * a.B + r) = pack4(e);
        *reinterpret_cast<uint2*>(a.dsT + c * a.B + r) = pack4(q);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (r + j < a.B) {
            a.deT[c * a.B + r + j] = __float2bfloat16_rn(e[j]);
            a.dsT[c * a.B + r + j] = __float2bfloat16_rn(q[j]);
          }
      }
    }
    // column sums over the tile's 64 rows: 16 lanes (tx) of a half-warp hold 4 rows each; fixed-order butterfly
    float s1 = (e[0] + e[1]) + (e[2] + e[3]), s2 = (q[0] + q[1]) + (q[2] + q[3]);     // out-of-range rows hold zeros
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (tx == 0 && c < a.K && a.partial) {
      a.partial[((int64_t)blockIdx.y * 2 + 0) * a.K + c] = s1;
      a.partial[((int64_t)blockIdx.y * 2 + 1) * a.K + c] = s2;
    }
  }
}

// ---- rows-by-few dual dot products on the CUDA cores -------------------------------------------------------------------
//   s1[r][o] = sum_c big1[r][c] small1[o][c]      s2[r][o] = sum_c big2[r][c] small2[o][c]        o < O <= 12
// The classifier head at the wide shape (out = 10) as a GEMM on 128-wide tensor-core tiles is 64 (forward) or 32 (dW) CTAs
// each streaming megabytes of activations: 81 + 69 us for 1.3 GFLOP.  Here every SM streams rows: the small matrices sit in
// shared memory as bf16 (chunks of the contraction), each warp takes two rows per pass, lanes stride over the contraction
// with 16-byte loads, fp32 accumulation, a butterfly reduction, and lane o applies the epilogue of output o.
// Measured (r01): 138 / 136 us -- the 160 KB shared-memory fill per CTA and 8 warps per SM (227 registers) leave too few
// bytes in flight -- so the trainer keeps the tensor-core calls by default (LRTTensorCoreTrainer(small_head=True) selects
// these); an mma.sync formulation with the outputs padded to 16 would make it bandwidth-bound (~30 us).
//   FWD : big = (x, x^2) (batch, in), small = (M, V) (out, in): act = s1 + b_mu + sqrt(s2 + sigma_b^2) eps, ds = eps / (2 sd)
//   RAWT: big = (x^T, x^2^T) (in, batch), small = (dE^T, dS^T) (out, batch): dM[o][r] = s1, dV[o][r] = s2   (fp32, (out, in))
constexpr int kSkinnyMaxO = 12;
constexpr int kSkinnyThreads = 256;

struct SkinnyArgs {
  const __nv_bfloat16 *big1, *big2, *small1, *small2;
  int64_t R, C;
  int O, chunk, mode;                  // chunk: contraction elements per shared-memory fill (multiple of 256)
  // FWD
  const float *bias_mu, *bias_rho;
  Noise noise;
  int relu;
  float *act, *ds;
  // RAWT
  float *d1, *d2;
};
enum { SKINNY_FWD = 0, SKINNY_RAWT = 1 };

__device__ __forceinline__ void bf16x8_to_f32(const uint4 u, float f[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

template <int O>
__device__ __forceinline__ void skinny_rows(const SkinnyArgs& a, const __nv_bfloat16* s1, const __nv_bfloat16* s2, int64_t c0,
                                            int clen, bool first_chunk, bool last_chunk, const Noise& nz) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_g = (int64_t)blockIdx.x * (kSkinnyThreads / 32) + (threadIdx.x >> 5);
  const int64_t warps = (int64_t)gridDim.x * (kSkinnyThreads / 32);
  for (int64_t r = warp_g * 2; r < a.R; r += warps * 2) {
    const bool two = r + 1 < a.R;
    float acc[2][2][O];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int o = 0; o < O; ++o) acc[i][0][o] = acc[i][1][o] = 0.f;
    const __nv_bfloat16* p1 = a.big1 + r * a.C + c0;
    const __nv_bfloat16* p2 = a.big2 + r * a.C + c0;
    for (int k = lane * 8; k < clen; k += 256) {
      float x[2][2][8];
      bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(p1 + k)), x[0][0]);
      bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(p2 + k)), x[0][1]);
      if (two) {
        bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(p1 + a.C + k)), x[1][0]);
        bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(p2 + a.C + k)), x[1][1]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[1][0][j] = x[1][1][j] = 0.f;
      }
#pragma unroll
      for (int o = 0; o < O; ++o) {
        float m[8], v[8];
        bf16x8_to_f32(*reinterpret_cast<const uint4*>(s1 + (int64_t)o * a.chunk + k), m);
        bf16x8_to_f32(*reinterpret_cast<const uint4*>(s2 + (int64_t)o * a.chunk + k), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][0][o] = fmaf(x[0][0][j], m[j], acc[0][0][o]);
          acc[0][1][o] = fmaf(x[0][1][j], v[j], acc[0][1][o]);
          acc[1][0][o] = fmaf(x[1][0][j], m[j], acc[1][0][o]);
          acc[1][1][o] = fmaf(x[1][1][j], v[j], acc[1][1][o]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int o = 0; o < O; ++o) acc[i][w][o] = warp_sum(acc[i][w][o]);
    // lane (i * 16 + o) finishes output o of row r + i
    const int i = lane >> 4, ol = lane & 15;
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int o = 0; o < O; ++o)
      if (ol == o) { t1 = i ? acc[1][0][o] : acc[0][0][o]; t2 = i ? acc[1][1][o] : acc[0][1][o]; }
    const int64_t row = r + i;
    if (ol < O && row < a.R) {
      if (a.mode == SKINNY_RAWT) {
        float* q1 = a.d1 + (int64_t)ol * a.R + row;
        float* q2 = a.d2 + (int64_t)ol * a.R + row;
        *q1 = first_chunk ? t1 : *q1 + t1;        // chunks of the contraction are added in order by the same thread
        *q2 = first_chunk ? t2 : *q2 + t2;
      } else if (last_chunk) {                    // FWD runs as a single chunk (checked on the host)
        const float sb = sigma_of(__ldg(a.bias_rho + ol));
        const int64_t idx = row * O + ol;
        const float ep = nz.ptr ? __ldg(nz.ptr + idx) : philox_normal1(nz.seed, nz.stream, (uint64_t)idx);
        const float sd = sqrtf(fmaxf(t2, 0.f) + sb * sb);
        float v = t1 + __ldg(a.bias_mu + ol) + sd * ep;
        if (a.relu) v = fmaxf(v, 0.f);
        a.act[idx] = v;
        if (a.ds) a.ds[idx] = ep / (2.0f * sd);
      }
    }
  }
}

__global__ void __launch_bounds__(kSkinnyThreads, 1) skinny_dual_kernel(const SkinnyArgs a) {
  extern __shared__ __align__(16) uint8_t skinny_smem[];
  __nv_bfloat16* s1 = reinterpret_cast<__nv_bfloat16*>(skinny_smem);
  __nv_bfloat16* s2 = s1 + (int64_t)a.O * a.chunk;
  Noise nz = a.noise;
  nz.resolve();
  for (int64_t c0 = 0; c0 < a.C; c0 += a.chunk) {
    const int clen = (int)min((int64_t)a.chunk, a.C - c0);
    if (c0) __syncthreads();
    for (int e = threadIdx.x * 8; e < a.O * a.chunk; e += kSkinnyThreads * 8) {
      const int o = e / a.chunk, k = e - o * a.chunk;
      uint4 u = make_uint4(0u, 0u, 0u, 0u), w = u;
      if (k < clen) {
        u = __ldg(reinterpret_cast<const uint4*>(a.small1 + (int64_t)o * a.C + c0 + k));
        w = __ldg(reinterpret_cast<const uint4*>(a.small2 + (int64_t)o * a.C + c0 + k));
      }
      *reinterpret_cast<uint4*>(s1 + e) = u;
      *reinterpret_cast<uint4*>(s2 + e) = w;
    }
    __syncthreads();
    const bool first = c0 == 0, last = c0 + a.chunk >= a.C;
    switch (a.O) {
      case 1: skinny_rows<1>(a, s1, s2, c0, clen, first, last, nz); break;
      case 2: skinny_rows<2>(a, s1, s2, c0, clen, first, last, nz); break;
      case 3: skinny_rows<3>(a, s1, s2, c0, clen, first, last, nz); break;
      case 4: skinny_rows<4>(a, s1, s2, c0, clen, first, last, nz); break;
      case 5: skinny_rows<5>(a, s1, s2, c0, clen, first, last, nz); break;
      case 6: skinny_rows<6>(a, s1, s2, c0, clen, first, last, nz); break;
      case 7: skinny_rows<7>(a, s1, s2, c0, clen, first, last, nz); break;
      case 8: skinny_rows<8>(a, s1, s2, c0, clen, first, last, nz); break;
      case 9: skinny_rows<9>(a, s1, s2, c0, clen, first, last, nz); break;
      case 10: skinny_rows<10>(a, s1, s2, c0, clen, first, last, nz); break;
      case 11: skinny_rows<11>(a, s1, s2, c0, clen, first, last, nz); break;
      default: skinny_rows<12>(a, s1, s2, c0, clen, first, last, nz); break;
    }
  }
}

int launch_skinny(SkinnyArgs a, cudaStream_t st, const char* what) {
  LBBNN_REQUIRE(a.O >= 1 && a.O <= kSkinnyMaxO, "%s: 1 <= out <= %d (got %d)", what, kSkinnyMaxO, a.O);
  LBBNN_REQUIRE(a.C % 8 == 0, "%s: contraction length must be a multiple of 8 (got %lld)", what, (long long)a.C);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  LBBNN_REQUIRE(al(a.big1) && al(a.big2) && al(a.small1) && al(a.small2), "%s: operands must be 16-byte aligned", what);
  // the contraction in as few shared-memory fills as fit ~192 KB: 2 matrices x O rows x chunk x 2 B
  const int64_t cap = (192 * 1024) / (4 * (int64_t)a.O) / 256 * 256;
  int64_t chunks = ceil_div(a.C, cap);
  int64_t chunk = ceil_div(ceil_div(a.C, chunks), 256) * 256;
  LBBNN_REQUIRE(a.mode == SKINNY_RAWT || chunks == 1, "%s: in_features %lld x out %d does not fit one shared-memory fill", what,
                (long long)a.C, a.O);
  a.chunk = (int)chunk;
  const size_t smem = (size_t)4 * a.O * chunk;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    LBBNN_CUDA(cudaFuncSetAttribute(skinny_dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  const int64_t passes = ceil_div(a.R, 2 * (kSkinnyThreads / 32));
  const int grid = (int)std::min<int64_t>(sm_count(), std::max<int64_t>(1, passes));
  skinny_dual_kernel<<<grid, kSkinnyThreads, smem, st>>>(a);
  return check_launch(what);
}

int colsum_slices(int64_t rows) {
  int64_t s = rows / 128;
  return (int)(s < 1 ? 1 : (s > 64 ? 64 : s));
}

}  // namespace
}  // namespace lbbnn

using namespace lbbnn;

extern "C" int lbbnn_bf16_pack(const float* a, const float* b, int op, int64_t rows, int64_t cols, void* out1, void* out2,
                               void* out1T, void* out2T, lbbnn_stream s) {
  LBBNN_REQUIRE(a && rows > 0 && cols > 0, "bad input");
  LBBNN_REQUIRE(op == LBBNN_PACK_SQUARE || b, "second operand required");
  LBBNN_REQUIRE((out1T == nullptr) == (out2T == nullptr), "transposed outputs come in pairs");
  auto al = [](const void* p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & m) == 0; };
  if (cols % 4 == 0 && al(a, 15) && al(b, 15) && al(out1, 7) && al(out2, 7) && al(out1T, 7) && al(out2T, 7) &&
      ceil_div(rows, 64) <= 65535) {
    dim3 grid64((unsigned)ceil_div(cols, 64), (unsigned)ceil_div(rows, 64));
    bf16_pack64_kernel<<<grid64, 256, 0, (cudaStream_t)s>>>(a, b, op, rows, cols, (__nv_bfloat16*)out1, (__nv_bfloat16*)out2,
                                                           (__nv_bfloat16*)out1T, (__nv_bfloat16*)out2T);
    return check_launch("bf16_pack64");
  }
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
  LBBNN_REQUIRE(grid.y <= 65535, "too many rows for one launch (%lld)", (long long)rows);
  bf16_pack_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(a, b, op, rows, cols, (__nv_bfloat16*)out1, (__nv_bfloat16*)out2,
                                                     (__nv_bfloat16*)out1T, (__nv_bfloat16*)out2T);
  return check_launch("bf16_pack");
}

extern "C" size_t lbbnn_colsum2_workspace_bytes(int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  return (size_t)colsum_slices(rows) * 2 * cols * sizeof(float);
}

extern "C" int lbbnn_colsum2(const void* a, const void* b, int a_is_bf16, int64_t rows, int64_t cols, float* out,
                             void* ws, size_t ws_bytes, lbbnn_stream s) {
  LBBNN_REQUIRE(a && out && rows > 0 && cols > 0, "bad input");
  LBBNN_REQUIRE(!a_is_bf16 || b, "bf16 mode needs both tensors");
  LBBNN_REQUIRE(ws && ws_bytes >= lbbnn_colsum2_workspace_bytes(rows, cols), "colsum2 workspace too small");
  const int slices = colsum_slices(rows);
  const int64_t rps = ceil_div(rows, slices);
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)slices);
  if (a_is_bf16) colsum2_stage1<true><<<grid, 256, 0, (cudaStream_t)s>>>(a, b, rows, cols, rps, (float*)ws);
  else colsum2_stage1<false><<<grid, 256, 0, (cudaStream_t)s>>>(a, b, rows, cols, rps, (float*)ws);
  if (int rc = check_launch("colsum2_stage1")) return rc;
  colsum2_stage2<<<(unsigned)ceil_div(2 * cols, 256), 256, 0, (cudaStream_t)s>>>((const float*)ws, slices, cols, out);
  return check_launch("colsum2_stage2");
}

extern "C" size_t lbbnn_tc_lrt_bwd_input_small_workspace_bytes(int64_t batch, int64_t in_features) {
  if (batch <= 0 || in_features <= 0) return 0;
  return (size_t)ceil_div(batch, 64) * 2 * in_features * sizeof(float);
}

extern "C" int lbbnn_tc_lrt_bwd_input_small(const float* gact, const float* ds_factor, const float* M32, const float* V32,
                                            int64_t batch, int64_t in_features, int64_t out_features, const void* x_bf,
                                            const float* ds_prev, int flags, void* dE_bf, void* dS_bf, void* dET_bf,
                                            void* dST_bf, float* colsum, void* ws, size_t ws_bytes, lbbnn_stream s) {
  LBBNN_REQUIRE(gact && ds_factor && M32 && V32 && x_bf && dE_bf && dS_bf, "NULL argument");
  LBBNN_REQUIRE(batch > 0 && in_features > 0 && out_features > 0, "empty shape");
  LBBNN_REQUIRE(out_features <= kSmallMaxOut, "out_features %lld > %d: use lbbnn_tc_lrt_bwd_input", (long long)out_features,
                kSmallMaxOut);
  LBBNN_REQUIRE(in_features % 4 == 0, "in_features must be a multiple of 4");
  LBBNN_REQUIRE((dET_bf == nullptr) == (dST_bf == nullptr), "transposed outputs come in pairs");
  const size_t need = lbbnn_tc_lrt_bwd_input_small_workspace_bytes(batch, in_features);
  LBBNN_REQUIRE(colsum == nullptr || (ws && ws_bytes >= need), "workspace too small for the column-sum partials");
  dim3 grid((unsigned)ceil_div(in_features, 64), (unsigned)ceil_div(batch, 64));
  LBBNN_REQUIRE(grid.y <= 65535, "too many rows for one launch");
  BwdSmallArgs a;
  a.g = gact; a.ds = ds_factor; a.M = M32; a.V = V32; a.x = (const __nv_bfloat16*)x_bf; a.ds_prev = ds_prev;
  a.B = batch; a.K = in_features; a.O = (int)out_features; a.mask = (flags & LBBNN_FLAG_MASK_DX) ? 1 : 0;
  a.de = (__nv_bfloat16*)dE_bf; a.dse = (__nv_bfloat16*)dS_bf; a.deT = (__nv_bfloat16*)dET_bf; a.dsT = (__nv_bfloat16*)dST_bf;
  a.partial = colsum ? (float*)ws : nullptr;
  lrt_bwd_input_small_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(a);
  if (int rc = check_launch("lrt_bwd_input_small")) return rc;
  if (colsum) {
    colsum2_stage2<<<(unsigned)ceil_div(2 * in_features, 256), 256, 0, (cudaStream_t)s>>>((const float*)ws, (int)grid.y,
                                                                                          in_features, colsum);
    return check_launch("colsum2_stage2");
  }
  return 0;
}

extern "C" int lbbnn_tc_lrt_fwd_small(const void* x_bf, const void* x2_bf, const void* M_bf, const void* V_bf, int64_t batch,
                                      int64_t in_features, int64_t out_features, const float* bias_mu, const float* bias_rho,
                                      const lbbnn_noise* nz, int flags, float* act_f32, float* ds_factor, lbbnn_stream s) {
  LBBNN_REQUIRE(x_bf && x2_bf && M_bf && V_bf && bias_mu && bias_rho && act_f32 && nz, "NULL argument");
  LBBNN_REQUIRE(batch > 0 && in_features > 0, "empty shape");
  SkinnyArgs a = {};
  a.big1 = (const __nv_bfloat16*)x_bf; a.big2 = (const __nv_bfloat16*)x2_bf;
  a.small1 = (const __nv_bfloat16*)M_bf; a.small2 = (const __nv_bfloat16*)V_bf;
  a.R = batch; a.C = in_features; a.O = (int)out_features; a.mode = SKINNY_FWD;
  a.bias_mu = bias_mu; a.bias_rho = bias_rho; a.noise = make_noise(nz); a.relu = (flags & LBBNN_FLAG_RELU) ? 1 : 0;
  a.act = act_f32; a.ds = ds_factor;
  return launch_skinny(a, (cudaStream_t)s, "tc_lrt_fwd_small");
}

extern "C" int lbbnn_tc_dual_gemm_raw_small(const void* A1, const void* A2, const void* B1, const void* B2, int64_t M, int64_t N,
                                            int64_t K, float* D1, float* D2, lbbnn_stream s) {
  // D1 (M, N) = A1 (M, K) B1 (N, K)^T with M <= 12 rows: the rows of B are the streamed side
  LBBNN_REQUIRE(A1 && A2 && B1 && B2 && D1 && D2 && N > 0 && K > 0, "NULL argument");
  SkinnyArgs a = {};
  a.big1 = (const __nv_bfloat16*)B1; a.big2 = (const __nv_bfloat16*)B2;
  a.small1 = (const __nv_bfloat16*)A1; a.small2 = (const __nv_bfloat16*)A2;
  a.R = N; a.C = K; a.O = (int)M; a.mode = SKINNY_RAWT; a.d1 = D1; a.d2 = D2;
  return launch_skinny(a, (cudaStream_t)s, "tc_dual_gemm_raw_small");
}
