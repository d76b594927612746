// LRT layer, fp32 SIMT path ("parity mode": matches the reference within 1e-5, TF32 never used).
//
// Replaces BayesianLinear.forward of LBBNN-GP-MF-LRT.py:166-196 (and the GEMM part of
// LBBNN-GP-MF-MNF.py:190-206) and autograd through it.
//
// Forward  = split-K dual GEMM with the parameter prologue and the KL reduction fused into the
//            weight-tile loader (each mu/rho/lambda element is read from HBM exactly once per
//            m-tile and never materialised as M,V), followed by a distributed epilogue that
//            sums the split partials in a fixed order (deterministic), adds the biases and applies
//            sqrt / eps / FMA (/ relu).
// Backward = dW kernel (dM = dE^T x, dV = dS^T x^2, contraction over the batch) whose epilogue
//            applies the chain rule to (mu, rho, lambda) and adds the closed-form KL gradient, and
//            a split-N dX kernel (dx = dE M + 2 x (dS V)) + epilogue (relu mask of the producer).
#include "common.cuh"

namespace lbbnn {
namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// guarded 4-wide loads (vector when the row is 16B aligned and fully in range, else element-wise)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load4(const float* __restrict__ base, int64_t row, int64_t col, int64_t nrows,
                                        int64_t col_end, int64_t ld, bool vec) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= nrows) return r;
  const float* p = base + row * ld + col;
  if (vec && col + 3 < col_end) {
    r = __ldg(reinterpret_cast<const float4*>(p));
  } else {
    if (col + 0 < col_end) r.x = __ldg(p + 0);
    if (col + 1 < col_end) r.y = __ldg(p + 1);
    if (col + 2 < col_end) r.z = __ldg(p + 2);
    if (col + 3 < col_end) r.w = __ldg(p + 3);
  }
  return r;
}

__device__ __forceinline__ void store4(float* __restrict__ base, int64_t row, int64_t col, int64_t nrows,
                                       int64_t ncols, int64_t ld, bool vec, float4 v, bool accumulate) {
  if (row >= nrows) return;
  float* p = base + row * ld + col;
  if (vec && col + 3 < ncols) {
    if (accumulate) {
      float4 o = *reinterpret_cast<float4*>(p);
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    *reinterpret_cast<float4*>(p) = v;
  } else {
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col + j < ncols) p[j] = accumulate ? p[j] + e[j] : e[j];
  }
}

// eps for elements (row, col..col+3) of a (rows, ld) tensor: injected or native Philox.
__device__ __forceinline__ float4 eps4(const Noise& nz, int64_t row, int64_t col, int64_t nrows, int64_t ld,
                                       bool vec) {
  if (nz.ptr) return load4(nz.ptr, row, col, nrows, ld, ld, vec);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= nrows) return r;
  const uint64_t e = (uint64_t)row * (uint64_t)ld + (uint64_t)col;
  if ((e & 3u) == 0 && col + 3 < ld) {
    float n[4];
    philox_normal4(nz.seed, nz.stream, e >> 2, n);
    r = make_float4(n[0], n[1], n[2], n[3]);
  } else {
    if (col + 0 < ld) r.x = philox_normal1(nz.seed, nz.stream, e + 0);
    if (col + 1 < ld) r.y = philox_normal1(nz.seed, nz.stream, e + 1);
    if (col + 2 < ld) r.z = philox_normal1(nz.seed, nz.stream, e + 2);
    if (col + 3 < ld) r.w = philox_normal1(nz.seed, nz.stream, e + 3);
  }
  return r;
}

// dS = G * eps / (2 std)  (SURVEY.md §3.5), zero where std is not positive (padding)
__device__ __forceinline__ float ds_of(float g, float e, float sd) { return sd > 0.f ? g * e / (2.0f * sd) : 0.f; }

// ================================================================================================
// forward: split-K partial dual GEMM with fused prologue + KL
// ================================================================================================
constexpr int F_BM = 128, F_BN = 64, F_BK = 16;

struct FwdArgs {
  const float *x, *mu, *rho, *lam, *z;
  int64_t B, K, N;
  int chunks_per_split, splits;
  float* part;      // [splits][2][B][N]
  double* kl_part;  // [splits * gridDim.x]
  int var_mode, want_kl, sample;
  lbbnn_priors pri;
};

__global__ void __launch_bounds__(kThreads) lrt_f32_fwd_partial(const FwdArgs a) {
  __shared__ __align__(16) float xs[F_BK][F_BM];
  __shared__ __align__(16) float xq[F_BK][F_BM];
  __shared__ __align__(16) float ms[F_BK][F_BN];
  __shared__ __align__(16) float vs[F_BK][F_BN];
  __shared__ float red[32];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx: 4 output features, ty: 8 batch rows
  const int64_t n0 = (int64_t)blockIdx.x * F_BN, m0 = (int64_t)blockIdx.y * F_BM;
  const int64_t kbeg = (int64_t)blockIdx.z * a.chunks_per_split * F_BK;
  const int64_t kend = min(a.K, kbeg + (int64_t)a.chunks_per_split * F_BK);
  const bool vec = (a.K % 4 == 0) && aligned16(a.x) && aligned16(a.mu) && aligned16(a.rho) && aligned16(a.lam);
  const bool do_kl = a.want_kl && blockIdx.y == 0;

  float accE[8][4], accS[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) accE[i][j] = accS[i][j] = 0.f;
  float kl = 0.f;

  float4 xr[2], pm, pr, pl;
  const int prow = tid >> 2, pkq = tid & 3;  // loader coordinates for the parameter tile

  auto gload = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int id = tid + i * kThreads;
      xr[i] = load4(a.x, m0 + (id >> 2), k0 + (id & 3) * 4, a.B, kend, a.K, vec);
    }
    pm = load4(a.mu, n0 + prow, k0 + pkq * 4, a.N, kend, a.K, vec);
    pr = load4(a.rho, n0 + prow, k0 + pkq * 4, a.N, kend, a.K, vec);
    pl = load4(a.lam, n0 + prow, k0 + pkq * 4, a.N, kend, a.K, vec);
  };
  auto sstore = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int id = tid + i * kThreads;
      const int r = id >> 2, kq = (id & 3) * 4;
      const float e[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xs[kq + j][r] = e[j];
        xq[kq + j][r] = e[j] * e[j];
      }
    }
    const float mu[4] = {pm.x, pm.y, pm.z, pm.w}, rho[4] = {pr.x, pr.y, pr.z, pr.w}, lam[4] = {pl.x, pl.y, pl.z, pl.w};
    const bool rowok = n0 + prow < a.N;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gk = k0 + pkq * 4 + j;
      float m = 0.f, v = 0.f;
      if (rowok && gk < kend) {
        const float sg = sigma_of(rho[j]), al = alpha_of(lam[j]);
        const float zk = a.z ? __ldg(a.z + gk) : 1.0f;
        const Moments mo = weight_moments(mu[j], sg, al, a.var_mode);
        m = mo.m * zk;
        v = a.sample ? mo.v : 0.f;
        if (do_kl) kl += kl_weight_elem(mu[j] * zk, sg, al, a.pri);
      }
      ms[pkq * 4 + j][prow] = m;
      vs[pkq * 4 + j][prow] = v;
    }
  };

  if (kbeg < kend) {
    gload(kbeg);
    for (int64_t k0 = kbeg; k0 < kend; k0 += F_BK) {
      __syncthreads();  // previous tile fully consumed
      sstore(k0);
      __syncthreads();
      if (k0 + F_BK < kend) gload(k0 + F_BK);  // prefetch next tile into registers
#pragma unroll
      for (int k = 0; k < F_BK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&xs[k][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&xs[k][ty * 8 + 4]);
        const float4 q0 = *reinterpret_cast<const float4*>(&xq[k][ty * 8]);
        const float4 q1 = *reinterpret_cast<const float4*>(&xq[k][ty * 8 + 4]);
        const float4 bm = *reinterpret_cast<const float4*>(&ms[k][tx * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&vs[k][tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        const float mv[4] = {bm.x, bm.y, bm.z, bm.w}, vv[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            accE[i][j] = fmaf(av[i], mv[j], accE[i][j]);
            accS[i][j] = fmaf(qv[i], vv[j], accS[i][j]);
          }
      }
    }
  }

  float* pe = a.part + ((int64_t)blockIdx.z * 2 + 0) * a.B * a.N;
  float* ps = a.part + ((int64_t)blockIdx.z * 2 + 1) * a.B * a.N;
  const bool vst = (a.N % 4 == 0) && aligned16(a.part);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + ty * 8 + i, gn = n0 + tx * 4;
    store4(pe, gm, gn, a.B, a.N, a.N, vst, make_float4(accE[i][0], accE[i][1], accE[i][2], accE[i][3]), false);
    if (a.sample)
      store4(ps, gm, gn, a.B, a.N, a.N, vst, make_float4(accS[i][0], accS[i][1], accS[i][2], accS[i][3]), false);
  }
  if (do_kl) {
    const float s = block_sum(kl, red);
    if (tid == 0) a.kl_part[(int64_t)blockIdx.z * gridDim.x + blockIdx.x] = (double)s;
  }
}

struct FwdEpiArgs {
  const float* part;
  int splits;
  int64_t B, N;
  const float *bias_mu, *bias_rho;
  Noise noise;
  int flags;
  float *act, *std_out, *kl_out;
  const double* kl_part;
  int n_kl_part;
  lbbnn_priors pri;
};

__global__ void __launch_bounds__(kThreads) lrt_f32_fwd_epilogue(const FwdEpiArgs a) {
  __shared__ double dred[32];
  Noise nz = a.noise;
  nz.resolve();
  const bool sample = a.flags & LBBNN_FLAG_SAMPLE;
  const int64_t total = a.B * a.N;
  const bool vec = (a.N % 4 == 0) && aligned16(a.part) && aligned16(a.act) &&
                   (a.std_out == nullptr || aligned16(a.std_out)) && (a.noise.ptr == nullptr || aligned16(a.noise.ptr));
  const int64_t nquads = ceil_div(total, 4);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = q * 4;
    float E[4] = {0.f, 0.f, 0.f, 0.f}, S[4] = {0.f, 0.f, 0.f, 0.f}, ep[4] = {0.f, 0.f, 0.f, 0.f};
    const bool full = vec && e0 + 3 < total;
    if (full) {
      for (int s = 0; s < a.splits; ++s) {
        const float4 pe = *reinterpret_cast<const float4*>(a.part + ((int64_t)s * 2 + 0) * total + e0);
        E[0] += pe.x; E[1] += pe.y; E[2] += pe.z; E[3] += pe.w;
        if (sample) {
          const float4 ps = *reinterpret_cast<const float4*>(a.part + ((int64_t)s * 2 + 1) * total + e0);
          S[0] += ps.x; S[1] += ps.y; S[2] += ps.z; S[3] += ps.w;
        }
      }
    } else {
      for (int s = 0; s < a.splits; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (e0 + j < total) {
            E[j] += a.part[((int64_t)s * 2 + 0) * total + e0 + j];
            if (sample) S[j] += a.part[((int64_t)s * 2 + 1) * total + e0 + j];
          }
    }
    if (sample) {
      if (nz.ptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (e0 + j < total) ep[j] = nz.ptr[e0 + j];
      } else {
        philox_normal4(nz.seed, nz.stream, (uint64_t)q, ep);
      }
    }
    float out[4], sd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      out[j] = sd[j] = 0.f;
      if (e0 + j < total) {
        const int64_t n = (e0 + j) % a.N;
        float v = E[j] + __ldg(a.bias_mu + n);
        if (sample) {
          const float sb = sigma_of(__ldg(a.bias_rho + n));
          sd[j] = sqrtf(S[j] + sb * sb);
          v = fmaf(sd[j], ep[j], v);
        }
        out[j] = (a.flags & LBBNN_FLAG_RELU) ? fmaxf(v, 0.f) : v;
      }
    }
    if (full) {
      *reinterpret_cast<float4*>(a.act + e0) = make_float4(out[0], out[1], out[2], out[3]);
      if (a.std_out && sample) *reinterpret_cast<float4*>(a.std_out + e0) = make_float4(sd[0], sd[1], sd[2], sd[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e0 + j < total) {
          a.act[e0 + j] = out[j];
          if (a.std_out && sample) a.std_out[e0 + j] = sd[j];
        }
    }
  }
  // KL finalisation: fixed-order sum of the per-CTA partials in double + the bias term (LRT:185-186)
  if ((a.flags & LBBNN_FLAG_KL) && blockIdx.x == 0) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < a.n_kl_part; i += blockDim.x) acc += a.kl_part[i];
    for (int64_t n = threadIdx.x; n < a.N; n += blockDim.x)
      acc += (double)kl_bias_elem(__ldg(a.bias_mu + n), sigma_of(__ldg(a.bias_rho + n)), a.pri);
    const double tot = block_sum(acc, dred);
    if (threadIdx.x == 0) *a.kl_out = (float)tot;
  }
}

// ================================================================================================
// backward wrt parameters: dM = dE^T x, dV = dS^T x^2 with fused chain rule + KL gradient
// ================================================================================================
constexpr int W_BN = 64, W_BK = 64, W_BB = 16;

struct BwdWArgs {
  const float *x, *g, *sd, *mu, *rho, *lam, *z, *bias_mu, *bias_rho;
  Noise noise;
  int64_t B, K, N;
  int var_mode, sample, accumulate;
  const float* klg_dev;
  float klg_host;
  lbbnn_priors pri;
  float *dmu, *drho, *dlam, *dbmu, *dbrho, *dz;
};

__global__ void __launch_bounds__(kThreads) lrt_f32_bwd_params(const BwdWArgs a) {
  __shared__ __align__(16) float ge[W_BB][W_BN];
  __shared__ __align__(16) float gs[W_BB][W_BN];
  __shared__ __align__(16) float xs[W_BB][W_BK];
  __shared__ __align__(16) float xq[W_BB][W_BK];

  const int tid = threadIdx.x;
  const int tk = tid & 15, tn = tid >> 4;  // thread tile: 4 n x 4 k, lanes run along k (coalesced epilogue)
  const int64_t n0 = (int64_t)blockIdx.x * W_BN, k0 = (int64_t)blockIdx.y * W_BK;
  const bool vecn = (a.N % 4 == 0) && aligned16(a.g) && aligned16(a.sd) && (a.noise.ptr == nullptr || aligned16(a.noise.ptr));
  const bool veck = (a.K % 4 == 0) && aligned16(a.x) && aligned16(a.mu) && aligned16(a.rho) && aligned16(a.lam) &&
                    aligned16(a.dmu) && aligned16(a.drho) && aligned16(a.dlam);
  const float klg = (a.klg_dev ? __ldg(a.klg_dev) : 1.0f) * a.klg_host;
  Noise nz = a.noise;
  nz.resolve();

  float accM[4][4], accV[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) accM[i][j] = accV[i][j] = 0.f;
  float be[4] = {0.f, 0.f, 0.f, 0.f}, bs[4] = {0.f, 0.f, 0.f, 0.f};  // bias column sums (k-tile 0 only)

  const int lb = tid >> 4, lq = (tid & 15) * 4;  // loader: row b, 4 columns
  float4 rg, rs, rx;

  auto gload = [&](int64_t b0) {
    const int64_t b = b0 + lb;
    const float4 g = load4(a.g, b, n0 + lq, a.B, a.N, a.N, vecn);
    rg = g;
    rs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.sample) {
      const float4 sd = load4(a.sd, b, n0 + lq, a.B, a.N, a.N, vecn);
      const float4 ep = eps4(nz, b, n0 + lq, a.B, a.N, vecn);
      rs = make_float4(ds_of(g.x, ep.x, sd.x), ds_of(g.y, ep.y, sd.y), ds_of(g.z, ep.z, sd.z), ds_of(g.w, ep.w, sd.w));
    }
    rx = load4(a.x, b, k0 + lq, a.B, a.K, a.K, veck);
  };

  gload(0);
  for (int64_t b0 = 0; b0 < a.B; b0 += W_BB) {
    __syncthreads();
    *reinterpret_cast<float4*>(&ge[lb][lq]) = rg;
    *reinterpret_cast<float4*>(&gs[lb][lq]) = rs;
    *reinterpret_cast<float4*>(&xs[lb][lq]) = rx;
    *reinterpret_cast<float4*>(&xq[lb][lq]) = make_float4(rx.x * rx.x, rx.y * rx.y, rx.z * rx.z, rx.w * rx.w);
    be[0] += rg.x; be[1] += rg.y; be[2] += rg.z; be[3] += rg.w;
    bs[0] += rs.x; bs[1] += rs.y; bs[2] += rs.z; bs[3] += rs.w;
    __syncthreads();
    if (b0 + W_BB < a.B) gload(b0 + W_BB);
#pragma unroll
    for (int b = 0; b < W_BB; ++b) {
      const float4 e4 = *reinterpret_cast<const float4*>(&ge[b][tn * 4]);
      const float4 s4 = *reinterpret_cast<const float4*>(&gs[b][tn * 4]);
      const float4 x4 = *reinterpret_cast<const float4*>(&xs[b][tk * 4]);
      const float4 q4 = *reinterpret_cast<const float4*>(&xq[b][tk * 4]);
      const float ev[4] = {e4.x, e4.y, e4.z, e4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
      const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          accM[i][j] = fmaf(ev[i], xv[j], accM[i][j]);
          accV[i][j] = fmaf(sv[i], qv[j], accV[i][j]);
        }
    }
  }

  // ---- epilogue: chain rule through M = alpha mu z, V(sigma, alpha[, mu]) + KL gradient ------------
  const lbbnn_priors P = a.pri;
  const float inv_sp2 = 1.0f / (P.sigma * P.sigma);
  float dzv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t n = n0 + tn * 4 + i, kc = k0 + tk * 4;
    if (n >= a.N || kc >= a.K) continue;
    const float4 m4 = load4(a.mu, n, kc, a.N, a.K, a.K, veck);
    const float4 r4 = load4(a.rho, n, kc, a.N, a.K, a.K, veck);
    const float4 l4 = load4(a.lam, n, kc, a.N, a.K, a.K, veck);
    const float mu[4] = {m4.x, m4.y, m4.z, m4.w}, rho[4] = {r4.x, r4.y, r4.z, r4.w}, lam[4] = {l4.x, l4.y, l4.z, l4.w};
    float gm[4], gr[4], gl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gm[j] = gr[j] = gl[j] = 0.f;
      if (kc + j >= a.K) continue;
      const float sg = sigma_of(rho[j]), al = alpha_of(lam[j]);
      const float zk = a.z ? __ldg(a.z + kc + j) : 1.0f;
      const float dM = accM[i][j] * zk, dV = accV[i][j];
      float dmu = al * dM, dsg, dal;
      if (a.var_mode == LBBNN_VAR_REFERENCE) {
        dsg = 2.0f * al * al * sg * dV;
        dal = mu[j] * dM + 2.0f * al * sg * sg * dV;
      } else {
        dmu += 2.0f * al * (1.0f - al) * mu[j] * dV;
        dsg = 2.0f * al * sg * dV;
        dal = mu[j] * dM + (sg * sg + (1.0f - 2.0f * al) * mu[j] * mu[j]) * dV;
      }
      if (klg != 0.f) {
        const float d = mu[j] * zk - P.mu;
        dmu += klg * al * d * inv_sp2 * zk;
        dsg += klg * al * (sg * inv_sp2 - 1.0f / sg);
        dal += klg * (logf(P.sigma / sg) - 0.5f + logf(al / P.alpha) + (sg * sg + d * d) * 0.5f * inv_sp2 -
                      logf((1.0f - al) / (1.0f - P.alpha)));
        dzv[j] += klg * al * d * inv_sp2 * mu[j];
      }
      dzv[j] += al * mu[j] * accM[i][j];
      gm[j] = dmu;
      gr[j] = dsg * dsigma_drho(rho[j]);
      gl[j] = dal * al * (1.0f - al);
    }
    store4(a.dmu, n, kc, a.N, a.K, a.K, veck, make_float4(gm[0], gm[1], gm[2], gm[3]), a.accumulate);
    store4(a.drho, n, kc, a.N, a.K, a.K, veck, make_float4(gr[0], gr[1], gr[2], gr[3]), a.accumulate);
    store4(a.dlam, n, kc, a.N, a.K, a.K, veck, make_float4(gl[0], gl[1], gl[2], gl[3]), a.accumulate);
  }
  if (a.dz) {  // MNF: dz_k = sum_n (alpha mu dM' + KL term); caller zeroes dz first
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (k0 + tk * 4 + j < a.K) atomicAdd(a.dz + k0 + tk * 4 + j, dzv[j]);
  }

  // ---- bias gradients (k-tile 0 CTAs): column sums over the batch -----------------------------------
  if (blockIdx.y == 0) {
    __syncthreads();
    *reinterpret_cast<float4*>(&ge[lb][lq]) = make_float4(be[0], be[1], be[2], be[3]);
    *reinterpret_cast<float4*>(&gs[lb][lq]) = make_float4(bs[0], bs[1], bs[2], bs[3]);
    __syncthreads();
    if (tid < W_BN && n0 + tid < a.N) {
      float se = 0.f, ss = 0.f;
#pragma unroll
      for (int b = 0; b < W_BB; ++b) { se += ge[b][tid]; ss += gs[b][tid]; }
      const int64_t n = n0 + tid;
      const float bm = __ldg(a.bias_mu + n), br = __ldg(a.bias_rho + n), sb = sigma_of(br);
      float dbm = se, dsb = 2.0f * sb * ss;
      if (klg != 0.f) {
        const float inv = 1.0f / (P.bias_sigma * P.bias_sigma);
        dbm += klg * (bm - P.bias_mu) * inv;
        dsb += klg * (sb * inv - 1.0f / sb);
      }
      const float dbr = dsb * dsigma_drho(br);
      a.dbmu[n] = a.accumulate ? a.dbmu[n] + dbm : dbm;
      a.dbrho[n] = a.accumulate ? a.dbrho[n] + dbr : dbr;
    }
  }
}

// ================================================================================================
// backward wrt the input: dx = dE M + 2 x (dS V), split over the out-feature contraction
// ================================================================================================
constexpr int X_BM = 128, X_BK = 64, X_BN = 16;

struct BwdXArgs {
  const float *x, *g, *sd, *mu, *rho, *lam, *z;
  Noise noise;
  int64_t B, K, N;
  int chunks_per_split, splits, var_mode, sample;
  float* part;  // [splits][B][K]
};

__global__ void __launch_bounds__(kThreads) lrt_f32_bwd_input_partial(const BwdXArgs a) {
  __shared__ __align__(16) float ge[X_BN][X_BM];
  __shared__ __align__(16) float gs[X_BN][X_BM];
  __shared__ __align__(16) float ms[X_BN][X_BK];
  __shared__ __align__(16) float vs[X_BN][X_BK];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // tx: 4 input features, ty: 8 batch rows
  const int64_t k0 = (int64_t)blockIdx.x * X_BK, m0 = (int64_t)blockIdx.y * X_BM;
  const int64_t nbeg = (int64_t)blockIdx.z * a.chunks_per_split * X_BN;
  const int64_t nend = min(a.N, nbeg + (int64_t)a.chunks_per_split * X_BN);
  const bool vecn = (a.N % 4 == 0) && aligned16(a.g) && aligned16(a.sd) && (a.noise.ptr == nullptr || aligned16(a.noise.ptr));
  const bool veck = (a.K % 4 == 0) && aligned16(a.x) && aligned16(a.mu) && aligned16(a.rho) && aligned16(a.lam) && aligned16(a.part);

  float accE[8][4], accS[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) accE[i][j] = accS[i][j] = 0.f;

  float4 rg[2], rs[2], pm, pr, pl;
  const int prow = tid >> 4, pkq = (tid & 15) * 4;
  Noise nz = a.noise;
  nz.resolve();

  auto gload = [&](int64_t nb) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int id = tid + i * kThreads;
      const int64_t b = m0 + (id >> 2), n = nb + (id & 3) * 4;
      const float4 g = load4(a.g, b, n, a.B, nend, a.N, vecn);
      rg[i] = g;
      rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.sample) {
        const float4 sd = load4(a.sd, b, n, a.B, nend, a.N, vecn);
        const float4 ep = eps4(nz, b, n, a.B, a.N, vecn);
        rs[i] = make_float4(ds_of(g.x, ep.x, sd.x), ds_of(g.y, ep.y, sd.y), ds_of(g.z, ep.z, sd.z), ds_of(g.w, ep.w, sd.w));
      }
    }
    pm = load4(a.mu, nb + prow, k0 + pkq, nend, a.K, a.K, veck);
    pr = load4(a.rho, nb + prow, k0 + pkq, nend, a.K, a.K, veck);
    pl = load4(a.lam, nb + prow, k0 + pkq, nend, a.K, a.K, veck);
  };
  auto sstore = [&](int64_t nb) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int id = tid + i * kThreads;
      const int r = id >> 2, nq = (id & 3) * 4;
      const float e[4] = {rg[i].x, rg[i].y, rg[i].z, rg[i].w}, s[4] = {rs[i].x, rs[i].y, rs[i].z, rs[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ge[nq + j][r] = e[j];
        gs[nq + j][r] = s[j];
      }
    }
    const float mu[4] = {pm.x, pm.y, pm.z, pm.w}, rho[4] = {pr.x, pr.y, pr.z, pr.w}, lam[4] = {pl.x, pl.y, pl.z, pl.w};
    float m[4], v[4];
    const bool rowok = nb + prow < nend;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m[j] = v[j] = 0.f;
      const int64_t gk = k0 + pkq + j;
      if (rowok && gk < a.K) {
        const float sg = sigma_of(rho[j]), al = alpha_of(lam[j]);
        const Moments mo = weight_moments(mu[j], sg, al, a.var_mode);
        m[j] = mo.m * (a.z ? __ldg(a.z + gk) : 1.0f);
        v[j] = mo.v;
      }
    }
    *reinterpret_cast<float4*>(&ms[prow][pkq]) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(&vs[prow][pkq]) = make_float4(v[0], v[1], v[2], v[3]);
  };

  if (nbeg < nend) {
    gload(nbeg);
    for (int64_t nb = nbeg; nb < nend; nb += X_BN) {
      __syncthreads();
      sstore(nb);
      __syncthreads();
      if (nb + X_BN < nend) gload(nb + X_BN);
#pragma unroll
      for (int n = 0; n < X_BN; ++n) {
        const float4 e0 = *reinterpret_cast<const float4*>(&ge[n][ty * 8]);
        const float4 e1 = *reinterpret_cast<const float4*>(&ge[n][ty * 8 + 4]);
        const float4 s0 = *reinterpret_cast<const float4*>(&gs[n][ty * 8]);
        const float4 s1 = *reinterpret_cast<const float4*>(&gs[n][ty * 8 + 4]);
        const float4 bm = *reinterpret_cast<const float4*>(&ms[n][tx * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&vs[n][tx * 4]);
        const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float mv[4] = {bm.x, bm.y, bm.z, bm.w}, vv[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            accE[i][j] = fmaf(ev[i], mv[j], accE[i][j]);
            accS[i][j] = fmaf(sv[i], vv[j], accS[i][j]);
          }
      }
    }
  }
  float* part = a.part + (int64_t)blockIdx.z * a.B * a.K;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t b = m0 + ty * 8 + i, kc = k0 + tx * 4;
    if (b >= a.B || kc >= a.K) continue;
    const float4 x4 = load4(a.x, b, kc, a.B, a.K, a.K, veck);
    const float4 o = make_float4(fmaf(2.0f * x4.x, accS[i][0], accE[i][0]), fmaf(2.0f * x4.y, accS[i][1], accE[i][1]),
                                 fmaf(2.0f * x4.z, accS[i][2], accE[i][2]), fmaf(2.0f * x4.w, accS[i][3], accE[i][3]));
    store4(part, b, kc, a.B, a.K, a.K, veck, o, false);
  }
}

__global__ void __launch_bounds__(kThreads) lrt_f32_bwd_input_epilogue(const float* __restrict__ part, int splits,
                                                                       int64_t total, const float* __restrict__ x,
                                                                       int mask, int accumulate, float* __restrict__ dx) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < splits; ++i) s += part[(int64_t)i * total + e];
    if (mask && !(x[e] > 0.f)) s = 0.f;
    dx[e] = accumulate ? dx[e] + s : s;
  }
}

// ---- split heuristics (shared by the workspace query and the launchers) ----------------------------
struct Split { int chunks_per_split, splits; };

Split pick_split(int64_t tiles, int64_t chunks) {
  int64_t target = 2LL * sm_count();
  int64_t want = tiles >= target ? 1 : ceil_div(target, tiles);
  if (want > chunks) want = chunks;
  if (want < 1) want = 1;
  Split s;
  s.chunks_per_split = (int)ceil_div(chunks, want);
  s.splits = (int)ceil_div(chunks, s.chunks_per_split);
  if (s.splits < 1) s.splits = 1;
  return s;
}
Split fwd_split(int64_t B, int64_t K, int64_t N) {
  return pick_split(ceil_div(B, F_BM) * ceil_div(N, F_BN), ceil_div(K, F_BK));
}
Split dx_split(int64_t B, int64_t K, int64_t N) {
  return pick_split(ceil_div(B, X_BM) * ceil_div(K, X_BK), ceil_div(N, X_BN));
}

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }
size_t fwd_part_bytes(int64_t B, int64_t K, int64_t N) { return align_up((size_t)fwd_split(B, K, N).splits * 2 * B * N * sizeof(float)); }
size_t kl_part_bytes(int64_t B, int64_t K, int64_t N) { return align_up((size_t)fwd_split(B, K, N).splits * ceil_div(N, F_BN) * sizeof(double)); }
size_t dx_part_bytes(int64_t B, int64_t K, int64_t N) { return align_up((size_t)dx_split(B, K, N).splits * B * K * sizeof(float)); }

int check_layer(const lbbnn_layer* L) {
  LBBNN_REQUIRE(L != nullptr, "layer is NULL");
  LBBNN_REQUIRE(L->in_features > 0 && L->out_features > 0, "bad layer shape (%lld,%lld)", (long long)L->out_features,
                (long long)L->in_features);
  LBBNN_REQUIRE(L->weight_mu && L->weight_rho && L->lambdal && L->bias_mu && L->bias_rho, "layer has NULL parameters");
  return LBBNN_OK;
}

}  // namespace
}  // namespace lbbnn

using namespace lbbnn;

extern "C" size_t lbbnn_lrt_f32_workspace_bytes(int64_t B, int64_t K, int64_t N) {
  if (B <= 0 || K <= 0 || N <= 0) return 0;
  size_t fwd = fwd_part_bytes(B, K, N) + kl_part_bytes(B, K, N);
  size_t bwd = dx_part_bytes(B, K, N);
  return (fwd > bwd ? fwd : bwd) + 256;
}

extern "C" int lbbnn_lrt_f32_fwd(const lbbnn_layer* L, const float* x, int64_t B, const lbbnn_noise* nz,
                                 const lbbnn_priors* pri, int var_mode, int flags, float* act,
                                 float* std_out, float* kl_out, void* ws, size_t ws_bytes, lbbnn_stream s) {
  if (int rc = check_layer(L)) return rc;
  LBBNN_REQUIRE(x && act && B > 0, "x/act NULL or empty batch");
  LBBNN_REQUIRE(pri != nullptr, "priors NULL");
  LBBNN_REQUIRE(var_mode == LBBNN_VAR_REFERENCE || var_mode == LBBNN_VAR_EXACT, "bad var_mode %d", var_mode);
  LBBNN_REQUIRE(!(flags & LBBNN_FLAG_KL) || kl_out, "FLAG_KL needs kl_out");
  const int64_t K = L->in_features, N = L->out_features;
  LBBNN_REQUIRE(ws && ws_bytes >= lbbnn_lrt_f32_workspace_bytes(B, K, N), "workspace too small (%zu < %zu)", ws_bytes,
                lbbnn_lrt_f32_workspace_bytes(B, K, N));
  const Split sp = fwd_split(B, K, N);
  cudaStream_t st = (cudaStream_t)s;

  FwdArgs fa;
  fa.x = x; fa.mu = L->weight_mu; fa.rho = L->weight_rho; fa.lam = L->lambdal; fa.z = L->z;
  fa.B = B; fa.K = K; fa.N = N;
  fa.chunks_per_split = sp.chunks_per_split; fa.splits = sp.splits;
  fa.part = (float*)ws;
  fa.kl_part = (double*)((char*)ws + fwd_part_bytes(B, K, N));
  fa.var_mode = var_mode; fa.want_kl = (flags & LBBNN_FLAG_KL) ? 1 : 0; fa.sample = (flags & LBBNN_FLAG_SAMPLE) ? 1 : 0;
  fa.pri = *pri;
  dim3 grid((unsigned)ceil_div(N, F_BN), (unsigned)ceil_div(B, F_BM), (unsigned)sp.splits);
  lrt_f32_fwd_partial<<<grid, kThreads, 0, st>>>(fa);
  if (int rc = check_launch("lrt_f32_fwd_partial")) return rc;

  FwdEpiArgs ea;
  ea.part = fa.part; ea.splits = sp.splits; ea.B = B; ea.N = N;
  ea.bias_mu = L->bias_mu; ea.bias_rho = L->bias_rho;
  ea.noise = make_noise(nz);
  ea.flags = flags; ea.act = act; ea.std_out = std_out; ea.kl_out = kl_out;
  ea.kl_part = fa.kl_part; ea.n_kl_part = sp.splits * (int)grid.x; ea.pri = *pri;
  int64_t blocks = ceil_div(ceil_div(B * N, 4), kThreads);
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  lrt_f32_fwd_epilogue<<<(unsigned)blocks, kThreads, 0, st>>>(ea);
  return check_launch("lrt_f32_fwd_epilogue");
}

extern "C" int lbbnn_lrt_f32_bwd_params(const lbbnn_layer* L, const float* x, int64_t B, const float* gact,
                                        const float* std_saved, const lbbnn_noise* nz,
                                        const lbbnn_priors* pri, int var_mode, int flags, const float* kl_grad_dev,
                                        float kl_grad_host, const lbbnn_layer_grads* G, void* ws, size_t ws_bytes,
                                        lbbnn_stream s) {
  (void)ws; (void)ws_bytes;
  if (int rc = check_layer(L)) return rc;
  LBBNN_REQUIRE(x && gact && B > 0 && pri && G, "NULL argument");
  LBBNN_REQUIRE(G->weight_mu && G->weight_rho && G->lambdal && G->bias_mu && G->bias_rho, "NULL gradient buffer");
  const bool sample = flags & LBBNN_FLAG_SAMPLE;
  LBBNN_REQUIRE(!sample || std_saved, "sample-branch backward needs the saved std");
  LBBNN_REQUIRE(G->z == nullptr || L->z != nullptr, "dz requested but the layer has no z");
  BwdWArgs a;
  a.x = x; a.g = gact; a.sd = std_saved; a.mu = L->weight_mu; a.rho = L->weight_rho; a.lam = L->lambdal; a.z = L->z;
  a.bias_mu = L->bias_mu; a.bias_rho = L->bias_rho;
  a.noise = make_noise(nz);
  a.B = B; a.K = L->in_features; a.N = L->out_features;
  a.var_mode = var_mode; a.sample = sample ? 1 : 0; a.accumulate = (flags & LBBNN_FLAG_ACCUMULATE) ? 1 : 0;
  a.klg_dev = kl_grad_dev; a.klg_host = kl_grad_host; a.pri = *pri;
  a.dmu = G->weight_mu; a.drho = G->weight_rho; a.dlam = G->lambdal; a.dbmu = G->bias_mu; a.dbrho = G->bias_rho; a.dz = G->z;
  dim3 grid((unsigned)ceil_div(a.N, W_BN), (unsigned)ceil_div(a.K, W_BK));
  lrt_f32_bwd_params<<<grid, kThreads, 0, (cudaStream_t)s>>>(a);
  return check_launch("lrt_f32_bwd_params");
}

extern "C" int lbbnn_lrt_f32_bwd_input(const lbbnn_layer* L, const float* x, int64_t B, const float* gact,
                                       const float* std_saved, const lbbnn_noise* nz,
                                       int var_mode, int flags, float* dx, void* ws, size_t ws_bytes, lbbnn_stream s) {
  if (int rc = check_layer(L)) return rc;
  LBBNN_REQUIRE(x && gact && dx && B > 0, "NULL argument");
  const bool sample = flags & LBBNN_FLAG_SAMPLE;
  LBBNN_REQUIRE(!sample || std_saved, "sample-branch backward needs the saved std");
  const int64_t K = L->in_features, N = L->out_features;
  LBBNN_REQUIRE(ws && ws_bytes >= lbbnn_lrt_f32_workspace_bytes(B, K, N), "workspace too small");
  const Split sp = dx_split(B, K, N);
  BwdXArgs a;
  a.x = x; a.g = gact; a.sd = std_saved; a.mu = L->weight_mu; a.rho = L->weight_rho; a.lam = L->lambdal; a.z = L->z;
  a.noise = make_noise(nz);
  a.B = B; a.K = K; a.N = N;
  a.chunks_per_split = sp.chunks_per_split; a.splits = sp.splits; a.var_mode = var_mode; a.sample = sample ? 1 : 0;
  a.part = (float*)ws;
  dim3 grid((unsigned)ceil_div(K, X_BK), (unsigned)ceil_div(B, X_BM), (unsigned)sp.splits);
  lrt_f32_bwd_input_partial<<<grid, kThreads, 0, (cudaStream_t)s>>>(a);
  if (int rc = check_launch("lrt_f32_bwd_input_partial")) return rc;
  int64_t blocks = ceil_div(B * K, kThreads);
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  lrt_f32_bwd_input_epilogue<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)s>>>(
      a.part, sp.splits, B * K, x, (flags & LBBNN_FLAG_MASK_DX) ? 1 : 0, (flags & LBBNN_FLAG_ACCUMULATE) ? 1 : 0, dx);
  return check_launch("lrt_f32_bwd_input_epilogue");
}
