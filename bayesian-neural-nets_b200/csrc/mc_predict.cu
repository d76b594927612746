// Monte-Carlo posterior-predictive loop (test_ensemble, LBBNN-GP-MF.py:345-436), batched over weight samples.
//
// The reference runs one stochastic forward per (sample, test batch): fresh hard masks gamma ~ Bernoulli(alpha)
// (MF:113), weights w = gamma (mu + sigma eps) and biases (MF:232-234), F.linear + relu per layer (MF:255,268-271),
// then accumulates log_softmax outputs (MF:416) and row-normalised expit (MF:397-408) on the host.  Here SB samples
// go through each stage in ONE launch:
//   mc_sample    (blocks, SB)          mask + weights + bias of one layer for SB samples, vectorised and coalesced
//   sgemm_tn     (n-tiles, m-tiles, SB) out[s] = act(x[s] W_s^T + b_s): fp32 SIMT, 128 x BN x 16 tiles, 8 x BN/16 outputs
//                                       per thread, register-prefetched double-buffered shared memory
//   mc_accum     (row blocks)          per input row, the SB samples IN ORDER: log_softmax, the two fp64 accumulators
// Sample s draws from Philox streams keyed by its global index (first + s), exactly the streams of the one-sample
// entry points (lbbnn_mf_sample_predict), so results do not depend on SB or on how samples are sharded over GPUs.
#include <algorithm>

#include "common.cuh"

namespace lbbnn {
namespace {

constexpr int kThreads = 256;

// ---- sampling ---------------------------------------------------------------------------------------------------
struct McSampleArgs {
  const float *mu, *rho, *lam, *bias_mu, *bias_rho;
  int64_t n, n_bias;
  uint64_t seed, stream_base, stream_stride;   // stream of (sample, which) = stream_base + which + sample * stride
  const int64_t* first;                        // device: global index of sample 0 of this launch
  float *w, *bias;                             // (SB, n), (SB, n_bias)
  float* w_lo;                                 // optional: w then holds the TF32-rounded value and w_lo the remainder
  int prepared;                                // rho / lam / bias_rho already hold sigma / alpha / sigma_b (mc_prepare)
};

__global__ void __launch_bounds__(kThreads) mc_sample_kernel(const McSampleArgs a) {
  const int s = blockIdx.y;
  const uint64_t st = a.stream_base + (uint64_t)(*a.first + s) * a.stream_stride;
  float* w = a.w + (int64_t)s * a.n;
  float* wl = a.w_lo ? a.w_lo + (int64_t)s * a.n : nullptr;
  if (blockIdx.x == 0) {
    float* bo = a.bias + (int64_t)s * a.n_bias;
    for (int64_t i = threadIdx.x; i < a.n_bias; i += blockDim.x)
    {
      const float br = __ldg(a.bias_rho + i);
      bo[i] = fmaf(a.prepared ? br : sigma_of(br), philox_normal1(a.seed, st + 2, (uint64_t)i), __ldg(a.bias_mu + i));
    }
  }
  const bool vec = (a.n % 4 == 0) && aligned16(a.mu) && aligned16(a.rho) && aligned16(a.lam) && aligned16(a.w) &&
                   (!a.w_lo || aligned16(a.w_lo));
  const int64_t nq = ceil_div(a.n, 4);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = q * 4;
    float mu[4], rho[4], lam[4], u[4], ep[4], o[4];
    if (vec) {
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.mu + e0)), r4 = __ldg(reinterpret_cast<const float4*>(a.rho + e0));
      const float4 l4 = __ldg(reinterpret_cast<const float4*>(a.lam + e0));
      mu[0] = m4.x; mu[1] = m4.y; mu[2] = m4.z; mu[3] = m4.w; rho[0] = r4.x; rho[1] = r4.y; rho[2] = r4.z; rho[3] = r4.w;
      lam[0] = l4.x; lam[1] = l4.y; lam[2] = l4.z; lam[3] = l4.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = e0 + j < a.n;
        mu[j] = ok ? __ldg(a.mu + e0 + j) : 0.f; rho[j] = ok ? __ldg(a.rho + e0 + j) : 0.f; lam[j] = ok ? __ldg(a.lam + e0 + j) : 0.f;
      }
    }
    philox_uniform4(a.seed, st + 0, (uint64_t)q, u);
    philox_normal4(a.seed, st + 1, (uint64_t)q, ep);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float al = a.prepared ? lam[j] : alpha_of(lam[j]), sg = a.prepared ? rho[j] : sigma_of(rho[j]);
      const float g = u[j] < al ? 1.0f : 0.0f;                    // Bernoulli(alpha).sample() (MF:113)
      o[j] = g * fmaf(sg, ep[j], mu[j]);                          // gamma * (mu + sigma eps)   (MF:232-233)
    }
    float lo[4] = {0.f, 0.f, 0.f, 0.f};
    if (wl) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float full = o[j];
        tf32_split(full, o[j], lo[j]);
      }
    }
    if (vec) {
      *reinterpret_cast<float4*>(w + e0) = make_float4(o[0], o[1], o[2], o[3]);
      if (wl) *reinterpret_cast<float4*>(wl + e0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e0 + j < a.n) {
          w[e0 + j] = o[j];
          if (wl) wl[e0 + j] = lo[j];
        }
    }
  }
}

// sigma = log1p(exp rho), alpha = sigmoid(lambda) once per parameter set instead of once per weight sample
__global__ void __launch_bounds__(kThreads) mc_prepare_kernel(const float* __restrict__ rho, const float* __restrict__ lam,
                                                              const float* __restrict__ bias_rho, int64_t n, int64_t n_bias,
                                                              float* __restrict__ sigma, float* __restrict__ alpha,
                                                              float* __restrict__ bias_sigma) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    sigma[i] = sigma_of(__ldg(rho + i));
    alpha[i] = alpha_of(__ldg(lam + i));
    if (i < n_bias) bias_sigma[i] = sigma_of(__ldg(bias_rho + i));
  }
}

// ---- batched fp32 "TN" GEMM: out[s][m][n] = act(sum_k X[s][m][k] W[s][n][k] + bias[s][n]) ------------------------------
constexpr int BM = 128, BK = 16;

struct GemmArgs {
  const float *X, *W, *bias;
  float* out;
  int64_t xs, ws, bs, os;   // per-sample strides in floats (xs = 0: every sample reads the same input)
  int M, N, K, relu;
};

__device__ __forceinline__ float4 ld4g(const float* __restrict__ base, int row, int nrows, int col, int K, bool vec) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= nrows || col >= K) return r;
  const float* p = base + (int64_t)row * K + col;
  if (vec && col + 3 < K) return *reinterpret_cast<const float4*>(p);
  r.x = p[0];
  if (col + 1 < K) r.y = p[1];
  if (col + 2 < K) r.z = p[2];
  if (col + 3 < K) r.w = p[3];
  return r;
}

template <int BN>
__global__ void __launch_bounds__(kThreads, BN == 64 ? 3 : 2) sgemm_tn_batched_kernel(const GemmArgs a) {
  constexpr int TN = BN / 16;                  // columns per thread (8 or 4)
  constexpr int NB4 = BN * BK / 4 / kThreads;  // float4 loads of the W tile per thread (2 or 1)
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int s = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* X = a.X + (int64_t)s * a.xs;
  const float* W = a.W + (int64_t)s * a.ws;
  const bool vec = (a.K % 4 == 0) && aligned16(X) && aligned16(W);

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[NB4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * kThreads;
      ra[i] = ld4g(X, m0 + (idx & (BM - 1)), a.M, k0 + (idx / BM) * 4, a.K, vec);
    }
#pragma unroll
    for (int i = 0; i < NB4; ++i) {
      const int idx = tid + i * kThreads;
      rb[i] = ld4g(W, n0 + (idx % BN), a.N, k0 + (idx / BN) * 4, a.K, vec);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * kThreads;
      const int r = idx & (BM - 1), kq = (idx / BM) * 4;
      As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y; As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < NB4; ++i) {
      const int idx = tid + i * kThreads;
      const int c = idx % BN, kq = (idx / BN) * 4;
      Bs[buf][kq + 0][c] = rb[i].x; Bs[buf][kq + 1][c] = rb[i].y; Bs[buf][kq + 2][c] = rb[i].z; Bs[buf][kq + 3][c] = rb[i].w;
    }
  };

  const int ntile = (a.K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int t = 0; t < ntile; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntile) gload((t + 1) * BK);   // next tile's global loads in flight during this tile's math
    // rows ty*4..+3 and 64+ty*4..+3; columns tx*4..+3 (and BN/2 + tx*4..+3 for BN = 128): conflict-free LDS.128.
    // The fragments of step k+1 are loaded while step k's FMAs issue (explicit register double buffering).
    float4 fa0 = *reinterpret_cast<const float4*>(&As[buf][0][ty * 4]);
    float4 fa1 = *reinterpret_cast<const float4*>(&As[buf][0][64 + ty * 4]);
    float4 fb0 = *reinterpret_cast<const float4*>(&Bs[buf][0][tx * 4]);
    float4 fb1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (TN == 8) fb1 = *reinterpret_cast<const float4*>(&Bs[buf][0][BN / 2 + tx * 4]);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float av[8] = {fa0.x, fa0.y, fa0.z, fa0.w, fa1.x, fa1.y, fa1.z, fa1.w};
      float bv[TN];
      bv[0] = fb0.x; bv[1] = fb0.y; bv[2] = fb0.z; bv[3] = fb0.w;
      if constexpr (TN == 8) { bv[TN - 4] = fb1.x; bv[TN - 3] = fb1.y; bv[TN - 2] = fb1.z; bv[TN - 1] = fb1.w; }
      if (k + 1 < BK) {
        fa0 = *reinterpret_cast<const float4*>(&As[buf][k + 1][ty * 4]);
        fa1 = *reinterpret_cast<const float4*>(&As[buf][k + 1][64 + ty * 4]);
        fb0 = *reinterpret_cast<const float4*>(&Bs[buf][k + 1][tx * 4]);
        if constexpr (TN == 8) fb1 = *reinterpret_cast<const float4*>(&Bs[buf][k + 1][BN / 2 + tx * 4]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < ntile) {
      sstore(buf ^ 1);   // the other buffer was last read in iteration t-1, which ended with a barrier
      __syncthreads();
    }
  }

  const float* bias = a.bias + (int64_t)s * a.bs;
  float* out = a.out + (int64_t)s * a.os;
  const bool vst = (a.N % 4 == 0) && aligned16(out);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= a.M) continue;
#pragma unroll
    for (int h = 0; h < TN / 4; ++h) {
      const int c = n0 + h * (BN / 2) + tx * 4;
      if (c >= a.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = acc[i][h * 4 + j] + (c + j < a.N ? __ldg(bias + c + j) : 0.f);
        if (a.relu) v[j] = fmaxf(v[j], 0.f);
      }
      float* p = out + (int64_t)r * a.N + c;
      if (vst && c + 3 < a.N) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < a.N) p[j] = v[j];
      }
    }
  }
}

// ---- accumulation: per input row, the SB samples in order (deterministic fp64 sums) ---------------------------------
__global__ void __launch_bounds__(kThreads) mc_accumulate_batched_kernel(const float* __restrict__ logits, int64_t B, int64_t C,
                                                                         int n_samples, double* __restrict__ sum_logp,
                                                                         double* __restrict__ sum_prob, int64_t* counter) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (b < B) {
    for (int s = 0; s < n_samples; ++s) {
      const float* row = logits + ((int64_t)s * B + b) * C;
      float mx = -INFINITY;
      for (int64_t c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float se = 0.f;
      for (int64_t c = lane; c < C; c += 32) se += expf(row[c] - mx);
      se = warp_sum(se);
      const float lse = mx + logf(se);
      float ps = 0.f;
      for (int64_t c = lane; c < C; c += 32) ps += 1.0f / (1.0f + expf(-(row[c] - lse)));
      ps = warp_sum(ps);
      for (int64_t c = lane; c < C; c += 32) {
        const float lp = row[c] - lse;
        sum_logp[b * C + c] += (double)lp;
        sum_prob[b * C + c] += (double)((1.0f / (1.0f + expf(-lp))) / ps);
      }
    }
  }
  if (counter && blockIdx.x == 0 && threadIdx.x == 0) *counter += n_samples;
}

// ---- classifier head + accumulation in one kernel ---------------------------------------------------------------
// The last layer of the loop is a (batch, K) x (K, C) product with C = 10 classes: a 64-wide GEMM tile is 84 % padding and
// the logits only exist to be fed to the accumulators.  Here one warp owns one input row for ALL samples of the launch:
// per sample the block stages that sample's (C, K) weights in shared memory (cp.async, double-buffered), every warp
// forms its row's C dot products (lanes stride K in float4, C accumulators, xor-shuffle reduction), then log_softmax and
// the two fp64 accumulators exactly as mc_accumulate_batched_kernel (same expressions, samples in the same order, the
// running sums held in registers by lanes 0..C-1 and written once).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kHeadRows = 2;     // input rows per warp: every weight float4 read from shared memory feeds two rows
constexpr int kHeadThreads = 128;
constexpr int kHeadBlockRows = kHeadRows * kHeadThreads / 32;
constexpr int kHeadStages = 4;   // samples in flight (cp.async): weights and activations of sample s + 3 load during s
constexpr size_t kHeadSmemMax = 220 * 1024;
inline size_t head_smem_bytes(int64_t classes, int64_t K) {
  return (size_t)kHeadStages * (size_t)(classes + kHeadBlockRows) * (size_t)K * sizeof(float);
}

struct HeadArgs {
  const float *h, *W, *bias;   // (SB, B, K), (SB, C, K), (SB, C)
  int64_t hs;                  // floats between the activations of consecutive samples
  int B, K, n_samples;
  double *sum_logp, *sum_prob;
  int64_t* counter;
};

// v[c], c < 16, summed over the 32 lanes; lane L returns the total of class L >> 1 (16 shuffles instead of 5 per class):
// at offset 16 / 8 / 4 / 2 a lane keeps the half of its values selected by that bit of its index and adds the partner's
// copy of the same half, so each step halves the values per lane; offset 1 is a plain sum.
__device__ __forceinline__ float transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const bool up = lane & 16;
    const float keep = up ? v[c + 8] : v[c], send = up ? v[c] : v[c + 8];
    v[c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const bool up = lane & 8;
    const float keep = up ? v[c + 4] : v[c], send = up ? v[c] : v[c + 4];
    v[c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const bool up = lane & 4;
    const float keep = up ? v[c + 2] : v[c], send = up ? v[c] : v[c + 2];
    v[c] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool up = lane & 2;
    const float keep = up ? v[1] : v[0], send = up ? v[0] : v[1];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// reductions over the 16 class slots (lanes differing in bits 1..4; bit 0 holds a duplicate)
__device__ __forceinline__ float class_max(float v) {
#pragma unroll
  for (int o = 16; o > 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float class_sum(float v) {
#pragma unroll
  for (int o = 16; o > 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int C>
__global__ void __launch_bounds__(kHeadThreads) mc_head_accumulate_kernel(const HeadArgs a) {
  // kHeadStages buffers of [ C x K weights of one sample | this block's kHeadBlockRows x K activations of that sample ]
  extern __shared__ float4 sbuf[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K4 = a.K >> 2, CK4 = C * K4, stage4 = CK4 + kHeadBlockRows * K4;
  const int blk0 = blockIdx.x * kHeadBlockRows;                             // first row of this block
  const int rows_here = min(kHeadBlockRows, a.B - blk0);
  const int b0 = blk0 + warp * kHeadRows;                                   // this warp's rows b0 .. b0 + kHeadRows - 1
  const int cls = lane >> 1;                                                // the class this lane accumulates
  const bool owner = (lane & 1) == 0 && cls < C;
  auto stage = [&](int s) {                        // always commits a group so that group s == sample s
    if (s < a.n_samples) {
      float4* dst = sbuf + (s % kHeadStages) * stage4;
      const float4* wsrc = reinterpret_cast<const float4*>(a.W + (int64_t)s * C * a.K);
      for (int i = tid; i < CK4; i += kHeadThreads) cp_async16(dst + i, wsrc + i);
      const float4* hsrc = reinterpret_cast<const float4*>(a.h + (int64_t)s * a.hs + (int64_t)blk0 * a.K);
      for (int i = tid; i < rows_here * K4; i += kHeadThreads) cp_async16(dst + CK4 + i, hsrc + i);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < kHeadStages - 1; ++s) stage(s);
  double run_lp[kHeadRows], run_pr[kHeadRows];
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r) {
    const bool ok = owner && b0 + r < a.B;
    run_lp[r] = ok ? a.sum_logp[(int64_t)(b0 + r) * C + cls] : 0.0;
    run_pr[r] = ok ? a.sum_prob[(int64_t)(b0 + r) * C + cls] : 0.0;
  }
  for (int s = 0; s < a.n_samples; ++s) {
    stage(s + kHeadStages - 1);                    // into the buffer sample s - 1 used (released by the barrier below)
    cp_async_wait<kHeadStages - 1>();
    __syncthreads();                               // sample s's weights and activations have landed
    if (b0 < a.B) {
      const float4* w = sbuf + (s % kHeadStages) * stage4;
      const float4* hs = w + CK4 + warp * kHeadRows * K4;
      float acc[kHeadRows][16];
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[r][c] = 0.f;
      for (int k4 = lane; k4 < K4; k4 += 32) {
        float4 hv[kHeadRows];
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r)
          hv[r] = (b0 + r < a.B) ? hs[r * K4 + k4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float4 wv = w[c * K4 + k4];
#pragma unroll
          for (int r = 0; r < kHeadRows; ++r)
            acc[r][c] = fmaf(hv[r].x, wv.x, fmaf(hv[r].y, wv.y, fmaf(hv[r].z, wv.z, fmaf(hv[r].w, wv.w, acc[r][c]))));
        }
      }
      const float bias = cls < C ? __ldg(a.bias + (int64_t)s * C + cls) : 0.f;
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) {
        const float logit = transpose_reduce16(acc[r], lane) + bias;          // class cls of row b0 + r
        const float mx = class_max(cls < C ? logit : -INFINITY);
        const float se = class_sum(cls < C ? expf(logit - mx) : 0.f);
        const float lp = logit - (mx + logf(se));
        const float e = cls < C ? 1.0f / (1.0f + expf(-lp)) : 0.f;
        const float ps = class_sum(e);
        if (owner) {
          run_lp[r] += (double)lp;
          run_pr[r] += (double)(e / ps);
        }
      }
    }
    __syncthreads();                               // everyone is done with this buffer before it is refilled
  }
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r)
    if (owner && b0 + r < a.B) {
      a.sum_logp[(int64_t)(b0 + r) * C + cls] = run_lp[r];
      a.sum_prob[(int64_t)(b0 + r) * C + cls] = run_pr[r];
    }
  if (a.counter && blockIdx.x == 0 && tid == 0) *a.counter += a.n_samples;
}

template <int C>
int launch_head(const HeadArgs& a, cudaStream_t s) {
  const size_t smem = head_smem_bytes(C, a.K);
  LBBNN_REQUIRE(smem <= kHeadSmemMax, "fused head: %d stages of weights + activations need %zu bytes of shared memory",
                kHeadStages, smem);
  static bool attr_set = false;
  if (!attr_set) {
    LBBNN_CUDA(cudaFuncSetAttribute(mc_head_accumulate_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHeadSmemMax));
    attr_set = true;
  }
  mc_head_accumulate_kernel<C><<<(unsigned)ceil_div(a.B, kHeadBlockRows), kHeadThreads, smem, s>>>(a);
  return check_launch("mc_head_accumulate");
}

}  // namespace
}  // namespace lbbnn

using namespace lbbnn;

static int mc_sample_launch(const lbbnn_layer* L, int n_samples, const int64_t* first_sample_dev, uint64_t seed,
                            uint64_t stream_base, uint64_t stream_stride, float* w, float* w_lo, float* bias, int prepared,
                            lbbnn_stream s) {
  LBBNN_REQUIRE(L && L->weight_mu && L->weight_rho && L->lambdal && L->bias_mu && L->bias_rho, "layer has NULL parameters");
  LBBNN_REQUIRE(n_samples > 0 && n_samples <= 65535 && first_sample_dev && w && bias, "bad argument");
  McSampleArgs a;
  a.mu = L->weight_mu; a.rho = L->weight_rho; a.lam = L->lambdal; a.bias_mu = L->bias_mu; a.bias_rho = L->bias_rho;
  a.n = L->in_features * L->out_features; a.n_bias = L->out_features;
  a.seed = seed; a.stream_base = stream_base; a.stream_stride = stream_stride; a.first = first_sample_dev;
  a.w = w; a.bias = bias; a.w_lo = w_lo; a.prepared = prepared;
  int64_t blocks = ceil_div(ceil_div(a.n, 4), kThreads);
  const int64_t cap = std::max<int64_t>(1, 8LL * sm_count() / n_samples);
  if (blocks > cap) blocks = cap;
  mc_sample_kernel<<<dim3((unsigned)blocks, (unsigned)n_samples), kThreads, 0, (cudaStream_t)s>>>(a);
  return check_launch("mc_sample");
}

extern "C" int lbbnn_mc_sample(const lbbnn_layer* L, int n_samples, const int64_t* first_sample_dev, uint64_t seed,
                               uint64_t stream_base, uint64_t stream_stride, float* w, float* bias, lbbnn_stream s) {
  return mc_sample_launch(L, n_samples, first_sample_dev, seed, stream_base, stream_stride, w, nullptr, bias, 0, s);
}

extern "C" int lbbnn_mc_sample_split(const lbbnn_layer* L, int n_samples, const int64_t* first_sample_dev, uint64_t seed,
                                     uint64_t stream_base, uint64_t stream_stride, int prepared, float* w_hi, float* w_lo,
                                     float* bias, lbbnn_stream s) {
  return mc_sample_launch(L, n_samples, first_sample_dev, seed, stream_base, stream_stride, w_hi, w_lo, bias, prepared, s);
}

extern "C" int lbbnn_mc_prepare(const lbbnn_layer* L, float* sigma, float* alpha, float* bias_sigma, lbbnn_stream s) {
  LBBNN_REQUIRE(L && L->weight_rho && L->lambdal && L->bias_rho && sigma && alpha && bias_sigma, "NULL argument");
  const int64_t n = L->in_features * L->out_features;
  LBBNN_REQUIRE(n >= L->out_features, "bad shape");
  int64_t blocks = ceil_div(n, kThreads);
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  mc_prepare_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)s>>>(L->weight_rho, L->lambdal, L->bias_rho, n, L->out_features,
                                                                      sigma, alpha, bias_sigma);
  return check_launch("mc_prepare");
}

extern "C" int lbbnn_linear_f32_batched(const float* x, int64_t x_stride, const float* W, const float* bias, int n_samples,
                                        int64_t batch, int64_t in_features, int64_t out_features, int flags, float* out,
                                        lbbnn_stream s) {
  LBBNN_REQUIRE(x && W && bias && out && n_samples > 0 && n_samples <= 65535, "bad argument");
  LBBNN_REQUIRE(batch > 0 && in_features > 0 && out_features > 0 && batch < (1LL << 30) && in_features < (1LL << 30) &&
                    out_features < (1LL << 30), "bad shape");
  GemmArgs a;
  a.X = x; a.W = W; a.bias = bias; a.out = out;
  a.xs = x_stride; a.ws = in_features * out_features; a.bs = out_features; a.os = batch * out_features;
  a.M = (int)batch; a.N = (int)out_features; a.K = (int)in_features; a.relu = (flags & LBBNN_FLAG_RELU) ? 1 : 0;
  // 128-wide column tiles unless they would be mostly padding
  const int64_t t128 = ceil_div(out_features, 128) * 128, t64 = ceil_div(out_features, 64) * 64;
  if (t128 * 100 <= t64 * 112) {
    dim3 grid((unsigned)ceil_div(out_features, 128), (unsigned)ceil_div(batch, BM), (unsigned)n_samples);
    sgemm_tn_batched_kernel<128><<<grid, kThreads, 0, (cudaStream_t)s>>>(a);
  } else {
    dim3 grid((unsigned)ceil_div(out_features, 64), (unsigned)ceil_div(batch, BM), (unsigned)n_samples);
    sgemm_tn_batched_kernel<64><<<grid, kThreads, 0, (cudaStream_t)s>>>(a);
  }
  return check_launch("sgemm_tn_batched");
}

extern "C" int lbbnn_mc_accumulate_batched(const float* logits, int n_samples, int64_t batch, int64_t classes,
                                           double* sum_logp, double* sum_prob, int64_t* counter, lbbnn_stream s) {
  LBBNN_REQUIRE(logits && sum_logp && sum_prob && batch > 0 && classes > 0 && n_samples > 0, "bad argument");
  mc_accumulate_batched_kernel<<<(unsigned)ceil_div(batch, kThreads / 32), kThreads, 0, (cudaStream_t)s>>>(
      logits, batch, classes, n_samples, sum_logp, sum_prob, counter);
  return check_launch("mc_accumulate_batched");
}

extern "C" int lbbnn_mc_head_accumulate(const float* h, int64_t h_stride, const float* W, const float* bias, int n_samples,
                                        int64_t batch, int64_t in_features, int64_t classes, double* sum_logp,
                                        double* sum_prob, int64_t* counter, lbbnn_stream s) {
  LBBNN_REQUIRE(h && W && bias && sum_logp && sum_prob && n_samples > 0 && batch > 0 && batch < (1LL << 31), "bad argument");
  LBBNN_REQUIRE(classes >= 1 && classes <= 16, "fused head handles 1..16 classes, got %lld", (long long)classes);
  LBBNN_REQUIRE(head_smem_bytes(classes, in_features) <= kHeadSmemMax,
                "fused head: (classes + %d) * in_features = %lld floats per stage do not fit in shared memory %d times",
                kHeadBlockRows, (long long)((classes + kHeadBlockRows) * in_features), kHeadStages);
  LBBNN_REQUIRE(in_features > 0 && in_features % 4 == 0 && h_stride % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(W)) & 15) == 0,
                "in_features and the sample stride must be multiples of 4 floats, operands 16-byte aligned");
  HeadArgs a;
  a.h = h; a.W = W; a.bias = bias; a.hs = h_stride; a.B = (int)batch; a.K = (int)in_features; a.n_samples = n_samples;
  a.sum_logp = sum_logp; a.sum_prob = sum_prob; a.counter = counter;
  cudaStream_t st = (cudaStream_t)s;
  switch ((int)classes) {
#define LBBNN_HEAD_CASE(C) case C: return launch_head<C>(a, st);
    LBBNN_HEAD_CASE(1) LBBNN_HEAD_CASE(2) LBBNN_HEAD_CASE(3) LBBNN_HEAD_CASE(4) LBBNN_HEAD_CASE(5) LBBNN_HEAD_CASE(6)
    LBBNN_HEAD_CASE(7) LBBNN_HEAD_CASE(8) LBBNN_HEAD_CASE(9) LBBNN_HEAD_CASE(10) LBBNN_HEAD_CASE(11) LBBNN_HEAD_CASE(12)
    LBBNN_HEAD_CASE(13) LBBNN_HEAD_CASE(14) LBBNN_HEAD_CASE(15) LBBNN_HEAD_CASE(16)
#undef LBBNN_HEAD_CASE
  }
  return LBBNN_ERR_INVALID;
}
