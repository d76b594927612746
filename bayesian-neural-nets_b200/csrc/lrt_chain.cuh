// Per-weight chain rule of the LRT layer + closed-form KL gradient + torch.optim.Adam's update, as device functions that a
// GEMM epilogue can apply to the (dM, dV) values it holds in registers (csrc/tc_gemm.cu: the fused dW epilogue).  Same
// expressions as lrt_f32_finalize<ADAM> in csrc/lrt_f32.cu (LRT:185-192 for the KL terms).
#pragma once
#include "common.cuh"

namespace lbbnn {
namespace chain {

struct Consts {
  float klg;                        // weight of the KL gradient (1 / NUM_BATCHES), 0 = none
  float inv_sp2, log_ps, logit_pa;  // 1 / sigma_p^2, log sigma_p, log(alpha_p / (1 - alpha_p))
  float mu_p;
  int var_mode;
  float b1, b2, eps, step_size, inv_bc2_sqrt;   // Adam: betas, eps, lr / (1 - b1^t), 1 / sqrt(1 - b2^t)
};

__device__ __forceinline__ Consts make_consts(const lbbnn_priors& p, int var_mode, float klg, float b1, float b2, float eps,
                                              const float* __restrict__ coef) {
  Consts c;
  c.klg = klg;
  c.inv_sp2 = 1.0f / (p.sigma * p.sigma);
  c.log_ps = logf(p.sigma);
  c.logit_pa = logf(p.alpha) - logf(1.0f - p.alpha);
  c.mu_p = p.mu;
  c.var_mode = var_mode;
  c.b1 = b1; c.b2 = b2; c.eps = eps;
  c.step_size = __ldg(coef);
  c.inv_bc2_sqrt = 1.0f / __ldg(coef + 1);
  return c;
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// (dM, dV) of one weight -> (d mu, d rho, d lambda), KL gradient included
__device__ __forceinline__ void grads(const Consts& c, float mu, float rho, float lam, float dM, float dV, float& gmu, float& grho,
                                      float& glam) {
  const float er = expf(rho), sg = log1pf(er);            // sigma = log1p(e^rho); e^rho reused for d sigma / d rho
  const float al = 1.0f / (1.0f + expf(-lam));
  float dmu = al * dM, dsg, dal;
  if (c.var_mode == LBBNN_VAR_REFERENCE) {                // V = sigma^2 alpha^2 (LRT:171)
    dsg = 2.0f * al * al * sg * dV;
    dal = mu * dM + 2.0f * al * sg * sg * dV;
  } else {                                                // V = alpha (sigma^2 + (1 - alpha) mu^2)
    dmu += 2.0f * al * (1.0f - al) * mu * dV;
    dsg = 2.0f * al * sg * dV;
    dal = mu * dM + (sg * sg + (1.0f - 2.0f * al) * mu * mu) * dV;
  }
  if (c.klg != 0.f) {
    const float d = mu - c.mu_p;
    dmu += c.klg * al * d * c.inv_sp2;
    dsg += c.klg * al * (sg * c.inv_sp2 - __frcp_rn(sg));
    // log(sigma_p / sigma) - 1/2 + log(alpha / alpha_p) - log((1 - alpha) / (1 - alpha_p)) + ...; log(alpha / (1 - alpha)) = lambda
    dal += c.klg * ((c.log_ps - logf(sg)) - 0.5f + (lam - c.logit_pa) + (sg * sg + d * d) * 0.5f * c.inv_sp2);
  }
  gmu = dmu;
  grho = dsg * (er / (1.0f + er));
  glam = dal * al * (1.0f - al);
}

// The NEXT step's operands and KL term from the just-updated parameters -- what lrt_bf16_prologue computes (same expressions:
// sigma = log1p(e^rho), alpha = 1 / (1 + e^-lambda), M = alpha mu, V per var_mode; KL of LRT:189-192 in the shared-exponential
// form of kl_weight_elem_shared in lrt_f32.cu).
struct KlC { float log_ps, inv_2ps2, log_pa, log_1mpa, mu_p; };
__device__ __forceinline__ KlC make_klc(const lbbnn_priors& p) {
  KlC c;
  c.log_ps = logf(p.sigma);
  c.inv_2ps2 = 0.5f / (p.sigma * p.sigma);
  c.log_pa = logf(p.alpha);
  c.log_1mpa = logf(1.0f - p.alpha);
  c.mu_p = p.mu;
  return c;
}
__device__ __forceinline__ void next_moments(const KlC& c, int var_mode, float mu, float rho, float lam, float& M, float& V, float& kl) {
  const float sg = log1pf(expf(rho));
  const float t = expf(-lam), al = 1.0f / (1.0f + t);
  M = mu * al;
  V = (var_mode == LBBNN_VAR_REFERENCE) ? (sg * sg) * (al * al) : al * (sg * sg + (1.0f - al) * mu * mu);
  const float d = mu - c.mu_p;
  const float log_al = -log1pf(t);
  const float slab = (c.log_ps - logf(sg)) - 0.5f + (log_al - c.log_pa) + (sg * sg + d * d) * c.inv_2ps2;
  kl = al * slab + (1.0f - al) * ((log_al - lam) - c.log_1mpa);
}

// torch.optim.Adam (no amsgrad / weight decay) on one element; same approximations as adam_quad in lrt_f32.cu
__device__ __forceinline__ void adam(const Consts& c, float& p, float& m, float& v, float g) {
  m = m + (g - m) * (1.0f - c.b1);
  v = c.b2 * v + (1.0f - c.b2) * g * g;
  p = p - c.step_size * __fdividef(m, sqrt_approx(v) * c.inv_bc2_sqrt + c.eps);
}

}  // namespace chain
}  // namespace lbbnn
