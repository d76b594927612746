// Auxiliary terms of the MNF layer's KL branch (LBBNN-GP-MF-MNF.py:208-235) that are neither a flow nor the weight KL:
//
//   log_q0 = sum_i [ -0.5 log(pi) - 0.5 lv_i - 0.5 (z0_i - m_i)^2 / exp(lv_i) ]                       (MNF:212-214)
//   act_mu = r0_c @ W_mean^T, act_var = r0_c^2 @ W_var^T, W_mean = z2 * mu * alpha, W_var = sigma^2 alpha^2  (MNF:211,216-217)
//   a_r    = tanh(act_mu + sqrt(act_var) * eps_r)                                                     (MNF:218-219)
//   mean_r = outer(r0_b1, a_r).mean(-1) = r0_b1 * mean(a_r),  log_var_r = r0_b2 * mean(a_r)           (MNF:220-221)
//   log_rb = sum_i [ -0.5 log(pi) - 0.5 log_var_r_i - 0.5 (z_b[-1] - mean_r_i)^2 / exp(log_var_r_i) ] (MNF:225-227)
//
// (log pi, not log 2 pi, and z_b[-1] the LAST ELEMENT of the flowed vector: reference quirks, kept.)  The layer needs
// log_q0 - log_rb; one forward kernel produces it, one backward kernel produces every gradient, replacing ~50 forward
// and ~120 backward elementwise / reduction launches of the eager formulation.
//
// forward:  one warp per output row j: the two length-`in` dot products against M0 = alpha mu and V (z2 and r0_c folded
//           into the vector side), tanh; the last block to finish (ticket) reduces mean(a_r) in a fixed order and forms
//           the two sums over i.
// backward: one block per 32 input columns, 8 row groups: d act_mu[j], d act_var[j] are per-row scalars (every block
//           recomputes the O(in) reductions they hang off), so dM0 / dV are rank-1 and the column sums
//           sum_j d act_mu[j] M0[j,i], sum_j d act_var[j] V[j,i] come from one coalesced sweep over M0 and V.
#include <cstdlib>

#include "common.cuh"

namespace lbbnn {
namespace {

constexpr int kThreads = 256;
constexpr float kHalfLogPi = 0.57236494292470008707f;   // 0.5 * log(pi)

__global__ void __launch_bounds__(kThreads) mnf_aux_fwd_kernel(const lbbnn_mnf_aux a, float* __restrict__ out,
                                                               float* __restrict__ save, unsigned int* ticket) {
  __shared__ float scratch[32];
  __shared__ bool last;
  const int D = (int)a.in_features, O = (int)a.out_features;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * (kThreads / 32) + warp;
  if (j < O) {
    const float* m = a.M0 + (int64_t)j * D;
    const float* v = a.V + (int64_t)j * D;
    float d1 = 0.f, d2 = 0.f;
    for (int base = 0; base < D; base += 256) {      // 8 independent (M0, V) load pairs in flight per lane
      float mv[8], vv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + lane + 32 * k;
        mv[k] = i < D ? __ldg(m + i) : 0.f;
        vv[k] = i < D ? __ldg(v + i) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + lane + 32 * k;
        if (i < D) {
          const float c = __ldg(a.r0_c + i);
          d1 = fmaf(c * __ldg(a.z2 + i), mv[k], d1);
          d2 = fmaf(c * c, vv[k], d2);
        }
      }
    }
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    if (lane == 0) {
      save[j] = tanhf(d1 + sqrtf(d2) * __ldg(a.eps_r + j));   // a_r
      save[O + j] = d2;                                       // act_var
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  float s = 0.f;
  for (int r = threadIdx.x; r < O; r += kThreads) s += __ldcg(save + r);
  s = block_sum(s, scratch);
  __shared__ float amean_s;
  if (threadIdx.x == 0) amean_s = s / (float)O;
  __syncthreads();
  const float amean = amean_s;
  const float zb = __ldg(a.z_b + D - 1);
  float q0 = 0.f, rb = 0.f;
  for (int i = threadIdx.x; i < D; i += kThreads) {
    const float lv = __ldg(a.q0_log_var + i), dz = __ldg(a.z0 + i) - __ldg(a.q0_mean + i);
    q0 += -kHalfLogPi - 0.5f * lv - 0.5f * (dz * dz / expf(lv));
    const float mr = __ldg(a.r0_b1 + i) * amean, lvr = __ldg(a.r0_b2 + i) * amean, u = zb - mr;
    rb += -kHalfLogPi - 0.5f * lvr - 0.5f * (u * u / expf(lvr));
  }
  q0 = block_sum(q0, scratch);
  rb = block_sum(rb, scratch);
  if (threadIdx.x == 0) {
    out[0] = q0 - rb;
    out[1] = q0;
    out[2] = rb;
    save[2 * O] = amean;
    *ticket = 0;
  }
}

__global__ void __launch_bounds__(kThreads) mnf_aux_bwd_kernel(const lbbnn_mnf_aux a, const float* __restrict__ save,
                                                               const float* __restrict__ gout, const lbbnn_mnf_aux_grads g) {
  __shared__ float scratch[32];
  __shared__ float bc[2];
  __shared__ float ps[8][32], qs[8][32];
  const int D = (int)a.in_features, O = (int)a.out_features;
  const float gq = __ldg(gout), gr = -gq;           // d/d log_q0 = +g, d/d log_rb = -g
  const float amean = __ldg(save + 2 * O);
  const float zb = __ldg(a.z_b + D - 1);
  // O(in) reductions every block needs: d amean and d z_b[-1]
  float da = 0.f, dzb = 0.f;
  for (int i = threadIdx.x; i < D; i += kThreads) {
    const float b1 = __ldg(a.r0_b1 + i), b2 = __ldg(a.r0_b2 + i);
    const float lvr = b2 * amean, u = zb - b1 * amean, iv = 1.0f / expf(lvr);
    const float dmr = gr * u * iv;                          // d log_rb / d mean_r_i
    const float dlv = gr * (-0.5f + 0.5f * u * u * iv);     // d log_rb / d log_var_r_i
    da += dmr * b1 + dlv * b2;
    dzb -= dmr;
  }
  da = block_sum(da, scratch);
  dzb = block_sum(dzb, scratch);
  if (threadIdx.x == 0) { bc[0] = da; bc[1] = dzb; }
  __syncthreads();
  const float d_ar = bc[0] / (float)O;                      // d / d a_r[j], the same for every j
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  const bool ok = i < D;
  const float c = ok ? __ldg(a.r0_c + i) : 0.f, z2 = ok ? __ldg(a.z2 + i) : 0.f;
  const float cz = c * z2, cc = c * c;
  float p = 0.f, q = 0.f;
  // rows are split over gridDim.y (25 blocks of 32 columns alone left most of the GPU idle: 30 us for 5 MB); each row slice adds
  // its share of the two column sums to d_r0_c / d_z2 (zeroed by the host call), slice 0 also writes the per-column terms
  const int rs = (O + gridDim.y - 1) / gridDim.y, j_lo = blockIdx.y * rs, j_hi = min(O, j_lo + rs);
  for (int j0 = j_lo + ty; j0 < j_hi; j0 += 8 * 8) {      // 8 rows = 16 independent loads in flight per thread
    float mv[8], vv[8], dpre[8], dvar[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = j0 + 8 * k;
      const bool okj = ok && j < j_hi;
      mv[k] = okj ? __ldg(a.M0 + (int64_t)j * D + i) : 0.f;
      vv[k] = okj ? __ldg(a.V + (int64_t)j * D + i) : 0.f;
      dpre[k] = dvar[k] = 0.f;
      if (j < j_hi) {
        const float ar = __ldg(save + j), av = __ldg(save + O + j);
        dpre[k] = (1.0f - ar * ar) * d_ar;
        dvar[k] = dpre[k] * __ldg(a.eps_r + j) * (0.5f / sqrtf(av));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = j0 + 8 * k;
      if (ok && j < j_hi) {
        const int64_t e = (int64_t)j * D + i;
        p = fmaf(dpre[k], mv[k], p);
        q = fmaf(dvar[k], vv[k], q);
        g.dM0[e] = dpre[k] * cz;
        g.dV[e] = dvar[k] * cc;
      }
    }
  }
  ps[ty][tx] = p;
  qs[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && ok) {
#pragma unroll
    for (int r = 1; r < 8; ++r) { p += ps[r][tx]; q += qs[r][tx]; }
    atomicAdd(g.d_r0_c + i, p * z2 + 2.0f * c * q);
    atomicAdd(g.d_z2 + i, p * c);
    if (blockIdx.y != 0) return;
    const float lv = __ldg(a.q0_log_var + i), dz = __ldg(a.z0 + i) - __ldg(a.q0_mean + i), iv0 = 1.0f / expf(lv);
    g.d_q0_mean[i] = gq * dz * iv0;
    g.d_q0_log_var[i] = gq * (-0.5f + 0.5f * dz * dz * iv0);
    g.d_z0[i] = -gq * dz * iv0;
    const float b1 = __ldg(a.r0_b1 + i), b2 = __ldg(a.r0_b2 + i);
    const float lvr = b2 * amean, u = zb - b1 * amean, iv = 1.0f / expf(lvr);
    g.d_r0_b1[i] = gr * u * iv * amean;
    g.d_r0_b2[i] = gr * (-0.5f + 0.5f * u * u * iv) * amean;
    g.d_z_b[i] = (i == D - 1) ? bc[1] : 0.f;
  }
}

// ---- the glue of one layer call, fused (each of these replaces 3-8 elementwise launches of the eager formulation) --------------
// z0[r,i] = q0_mean[i] + sqrt(exp(q0_log_var[i])) eps[r,i]   (MNF:183-185), eps injected or Philox; + the eps_r draw (out,)
__global__ void __launch_bounds__(kThreads) mnf_draw_kernel(const float* __restrict__ q0_mean, const float* __restrict__ q0_log_var,
                                                            Noise nz, int R, int D, float* __restrict__ eps_out,
                                                            float* __restrict__ z0, Noise nzr, int O, float* __restrict__ eps_r) {
  nz.resolve();
  nzr.resolve();
  const int n = R * D;
  for (int idx = blockIdx.x * kThreads + threadIdx.x; idx < n + O; idx += gridDim.x * kThreads) {
    if (idx < n) {
      const int i = idx % D;
      const float e = nz.ptr ? __ldg(nz.ptr + idx) : philox_normal1(nz.seed, nz.stream, (uint64_t)idx);
      eps_out[idx] = e;
      z0[idx] = __ldg(q0_mean + i) + sqrtf(expf(__ldg(q0_log_var + i))) * e;
    } else if (eps_r) {
      const int j = idx - n;
      eps_r[j] = nzr.ptr ? __ldg(nzr.ptr + j) : philox_normal1(nzr.seed, nzr.stream, (uint64_t)j);
    }
  }
}

// gradients of q0_mean / q0_log_var: through z0 (all rows) + the auxiliary term's direct ones; the auxiliary term also reads
// z0's KL row directly (d_z0_aux)
__global__ void __launch_bounds__(kThreads) mnf_draw_bwd_kernel(const float* __restrict__ q0_log_var, const float* __restrict__ eps,
                                                                const float* __restrict__ dz0, int R, int D, int kl_row,
                                                                const float* __restrict__ aux_dmean, const float* __restrict__ aux_dlv,
                                                                const float* __restrict__ aux_dz0, float* __restrict__ d_mean,
                                                                float* __restrict__ d_lv) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= D) return;
  float dm = 0.f, de = 0.f;
  for (int r = 0; r < R; ++r) {
    float g = dz0[(int64_t)r * D + i];
    if (r == kl_row && aux_dz0) g += aux_dz0[i];
    dm += g;
    de = fmaf(g, eps[(int64_t)r * D + i], de);
  }
  const float sd = sqrtf(expf(__ldg(q0_log_var + i)));
  d_mean[i] = dm + (aux_dmean ? aux_dmean[i] : 0.f);
  d_lv[i] = 0.5f * sd * de + (aux_dlv ? aux_dlv[i] : 0.f);
}

// kl = kl_wb + (log_q0 - log_rb) - log_det_q - log_det_r   (MNF:235)
__global__ void mnf_kl_combine_kernel(const float* kl_wb, const float* aux, const float* ldq, const float* ldr, float* out) {
  out[0] = (kl_wb[0] + aux[0]) - ldq[0] - ldr[0];
}

// backward staging: dld = [0, g_scale * g] (log-det gradients of the activation row / the KL row; dld + 1 serves the r flow),
// and -- second form -- the z-flow's output gradient rows: row 0 = d z_k (activation path), row 1 = a + b (the KL row's
// direct terms; the weight KL's share is accumulated on top by lbbnn_lrt_f32_finalize)
__global__ void __launch_bounds__(kThreads) mnf_bwd_rows_kernel(const float* __restrict__ g, float g_scale, float* __restrict__ dld,
                                                                const float* __restrict__ dz_k, const float* __restrict__ a,
                                                                const float* __restrict__ b, int D, float* __restrict__ rows) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (dld && i < 2) dld[i] = i == 0 ? 0.f : g_scale * __ldg(g);
  if (rows && i < D) {
    rows[i] = dz_k ? dz_k[i] : 0.f;
    rows[D + i] = (a ? a[i] : 0.f) + (b ? b[i] : 0.f);
  }
}

}  // namespace
}  // namespace lbbnn

using namespace lbbnn;

static int check_aux(const lbbnn_mnf_aux* a) {
  LBBNN_REQUIRE(a && a->in_features > 0 && a->out_features > 0 && a->in_features < (1LL << 30) && a->out_features < (1LL << 30),
                "bad shape");
  LBBNN_REQUIRE(a->q0_mean && a->q0_log_var && a->z0 && a->r0_c && a->r0_b1 && a->r0_b2 && a->z2 && a->M0 && a->V && a->eps_r &&
                    a->z_b, "NULL argument");
  return LBBNN_OK;
}

extern "C" size_t lbbnn_mnf_aux_save_floats(int64_t out_features) { return (size_t)(2 * out_features + 1); }

extern "C" int lbbnn_mnf_aux_kl_fwd(const lbbnn_mnf_aux* aux, float* out3, float* save, unsigned int* ticket, lbbnn_stream s) {
  if (int rc = check_aux(aux)) return rc;
  LBBNN_REQUIRE(out3 && save && ticket, "NULL output");
  const unsigned blocks = (unsigned)ceil_div(aux->out_features, kThreads / 32);
  mnf_aux_fwd_kernel<<<blocks, kThreads, 0, (cudaStream_t)s>>>(*aux, out3, save, ticket);
  return check_launch("mnf_aux_fwd");
}

extern "C" int lbbnn_mnf_aux_kl_bwd(const lbbnn_mnf_aux* aux, const float* save, const float* gout,
                                    const lbbnn_mnf_aux_grads* grads, lbbnn_stream s) {
  if (int rc = check_aux(aux)) return rc;
  LBBNN_REQUIRE(save && gout && grads && grads->d_q0_mean && grads->d_q0_log_var && grads->d_z0 && grads->d_r0_c &&
                    grads->d_r0_b1 && grads->d_r0_b2 && grads->d_z2 && grads->d_z_b && grads->dM0 && grads->dV, "NULL output");
  LBBNN_CUDA(cudaMemsetAsync(grads->d_r0_c, 0, (size_t)aux->in_features * sizeof(float), (cudaStream_t)s));
  LBBNN_CUDA(cudaMemsetAsync(grads->d_z2, 0, (size_t)aux->in_features * sizeof(float), (cudaStream_t)s));
  const unsigned blocks = (unsigned)ceil_div(aux->in_features, 32);
  unsigned slices = (unsigned)(2 * sm_count() / blocks);                 // ~2 blocks per SM in total
  const unsigned max_slices = (unsigned)ceil_div(aux->out_features, 64);   // at least one 64-row pass per slice
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  if (const char* e = getenv("LBBNN_AUX_SLICES")) slices = (unsigned)(atoi(e) > 0 ? atoi(e) : 1);   // A/B knob
  mnf_aux_bwd_kernel<<<dim3(blocks, slices), kThreads, 0, (cudaStream_t)s>>>(*aux, save, gout, *grads);
  return check_launch("mnf_aux_bwd");
}

extern "C" int lbbnn_mnf_draw(const float* q0_mean, const float* q0_log_var, const lbbnn_noise* eps_z, int64_t rows,
                              int64_t in_features, float* eps_out, float* z0, const lbbnn_noise* eps_r_noise,
                              int64_t out_features, float* eps_r, lbbnn_stream s) {
  LBBNN_REQUIRE(q0_mean && q0_log_var && eps_z && eps_out && z0 && rows > 0 && in_features > 0, "bad argument");
  LBBNN_REQUIRE(rows * in_features < (1LL << 30) && out_features < (1LL << 30), "too large");
  LBBNN_REQUIRE(eps_r == nullptr || (eps_r_noise && out_features > 0), "eps_r needs its noise source");
  const int64_t n = rows * in_features + (eps_r ? out_features : 0);
  mnf_draw_kernel<<<(unsigned)ceil_div(n, kThreads), kThreads, 0, (cudaStream_t)s>>>(
      q0_mean, q0_log_var, make_noise(eps_z), (int)rows, (int)in_features, eps_out, z0, make_noise(eps_r_noise),
      eps_r ? (int)out_features : 0, eps_r);
  return check_launch("mnf_draw");
}

extern "C" int lbbnn_mnf_draw_bwd(const float* q0_log_var, const float* eps, const float* dz0, int64_t rows, int64_t in_features,
                                  int kl_row, const float* aux_d_q0_mean, const float* aux_d_q0_log_var, const float* aux_d_z0,
                                  float* d_q0_mean, float* d_q0_log_var, lbbnn_stream s) {
  LBBNN_REQUIRE(q0_log_var && eps && dz0 && d_q0_mean && d_q0_log_var && rows > 0 && in_features > 0, "bad argument");
  LBBNN_REQUIRE(kl_row < rows, "kl_row out of range");
  mnf_draw_bwd_kernel<<<(unsigned)ceil_div(in_features, kThreads), kThreads, 0, (cudaStream_t)s>>>(
      q0_log_var, eps, dz0, (int)rows, (int)in_features, kl_row, aux_d_q0_mean, aux_d_q0_log_var, aux_d_z0, d_q0_mean, d_q0_log_var);
  return check_launch("mnf_draw_bwd");
}

extern "C" int lbbnn_mnf_kl_combine(const float* kl_wb, const float* aux_out, const float* log_det_q, const float* log_det_r,
                                    float* kl_out, lbbnn_stream s) {
  LBBNN_REQUIRE(kl_wb && aux_out && log_det_q && log_det_r && kl_out, "NULL argument");
  mnf_kl_combine_kernel<<<1, 1, 0, (cudaStream_t)s>>>(kl_wb, aux_out, log_det_q, log_det_r, kl_out);
  return check_launch("mnf_kl_combine");
}

extern "C" int lbbnn_mnf_bwd_rows(const float* g, float g_scale, float* dld2, const float* dz_k, const float* a, const float* b,
                                  int64_t in_features, float* rows2, lbbnn_stream s) {
  LBBNN_REQUIRE((dld2 && g) || rows2, "nothing to do");
  LBBNN_REQUIRE(rows2 == nullptr || in_features > 0, "bad shape");
  const int64_t n = rows2 ? (in_features > 2 ? in_features : 2) : 2;
  mnf_bwd_rows_kernel<<<(unsigned)ceil_div(n, kThreads), kThreads, 0, (cudaStream_t)s>>>(g, g_scale, dld2, dz_k, a, b, (int)in_features, rows2);
  return check_launch("mnf_bwd_rows");
}
