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
    for (int i = lane; i < D; i += 32) {
      const float c = __ldg(a.r0_c + i);
      d1 = fmaf(c * __ldg(a.z2 + i), __ldg(m + i), d1);
      d2 = fmaf(c * c, __ldg(v + i), d2);
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
  for (int j = ty; j < O; j += 8) {
    const float ar = __ldg(save + j), av = __ldg(save + O + j);
    const float dpre = (1.0f - ar * ar) * d_ar;
    const float dvar = dpre * __ldg(a.eps_r + j) * (0.5f / sqrtf(av));
    if (ok) {
      const int64_t e = (int64_t)j * D + i;
      p = fmaf(dpre, __ldg(a.M0 + e), p);
      q = fmaf(dvar, __ldg(a.V + e), q);
      g.dM0[e] = dpre * cz;
      g.dV[e] = dvar * cc;
    }
  }
  ps[ty][tx] = p;
  qs[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && ok) {
#pragma unroll
    for (int r = 1; r < 8; ++r) { p += ps[r][tx]; q += qs[r][tx]; }
    const float lv = __ldg(a.q0_log_var + i), dz = __ldg(a.z0 + i) - __ldg(a.q0_mean + i), iv0 = 1.0f / expf(lv);
    g.d_q0_mean[i] = gq * dz * iv0;
    g.d_q0_log_var[i] = gq * (-0.5f + 0.5f * dz * dz * iv0);
    g.d_z0[i] = -gq * dz * iv0;
    const float b1 = __ldg(a.r0_b1 + i), b2 = __ldg(a.r0_b2 + i);
    const float lvr = b2 * amean, u = zb - b1 * amean, iv = 1.0f / expf(lvr);
    g.d_r0_b1[i] = gr * u * iv * amean;
    g.d_r0_b2[i] = gr * (-0.5f + 0.5f * u * u * iv) * amean;
    g.d_r0_c[i] = p * z2 + 2.0f * c * q;
    g.d_z2[i] = p * c;
    g.d_z_b[i] = (i == D - 1) ? bc[1] : 0.f;
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
  const unsigned blocks = (unsigned)ceil_div(aux->in_features, 32);
  mnf_aux_bwd_kernel<<<blocks, kThreads, 0, (cudaStream_t)s>>>(*aux, save, gout, *grads);
  return check_launch("mnf_aux_bwd");
}
