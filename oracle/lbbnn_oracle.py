"""CPU oracle for the variational-layer hot path of LarsELund/Bayesian-Neural-Nets.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module, and only as the
checker or the timed CPU baseline -- never as part of the product path.  The
product (bayesian-neural-nets_b200/lbbnn) has no CPU fallback and does not
import this file.

What it is: a plain-PyTorch, noise-explicit restatement of the reference
algorithm.  Every random draw the reference takes from the global torch RNG is
an ARGUMENT here (eps, u, masks, tau), so the same noise can be fed to the CUDA
kernels and results compared element by element.  dtype follows the inputs, so
the same code is the fp32 oracle and (with .double() inputs) the fp64 truth used
to judge both fp32 implementations.  Gradients come from autograd.

Pinning: the reference ships no tests or golden vectors, so this
restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF: its class
definitions were executed in the build container (tests/golden/ref_harness.py,
AST-sliced from /root/reference, driven by the same replayed noise) and the
results stored in tests/golden/*.npz by tests/golden/make_golden.py.
tests/test_oracle_golden.py checks every function below against those files.

Each function cites the reference file:line it restates (paths relative to the
reference root; LRT = LBBNN-GP-MF-LRT.py, MNF = LBBNN-GP-MF-MNF.py,
MF = LBBNN-GP-MF.py, MFsim = LBBNN-GP-MFsim_study.py,
MNFsim = LBBNN-GP-MF-MNFsim_study.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------------------


@dataclass
class Priors:
    """Fixed priors of the LRT / MNF layers.

    LRT:141-158 / MNF:145-160: N(0,1) weights and biases, Bernoulli(0.05) inclusion.
    MNFsim:157-174 uses mu=0.1, sigma=1.3, alpha=0.3, bias N(0,1.3): the values are
    fields here so both parametrisations share the code.
    """
    mu: float = 0.0
    sigma: float = 1.0
    alpha: float = 0.05
    bias_mu: float = 0.0
    bias_sigma: float = 1.0


def sigma_of(rho):
    """sigma = log1p(exp(rho)); no softplus threshold.  LRT:80-82 (MF:81-83, MNF:84-86)."""
    return torch.log1p(torch.exp(rho))


def alpha_of(lam):
    """alpha = 1/(1+exp(-lambda)).  LRT:167, MNF:191, MF:246,292."""
    return 1.0 / (1.0 + torch.exp(-lam))


# --------------------------------------------------------------------------------------
# LRT layer  (LBBNN-GP-MF-LRT.py:129-197)
# --------------------------------------------------------------------------------------


def lrt_weight_moments(weight_mu, weight_rho, lambdal, var_mode="reference"):
    """Mean and variance of the masked weight used by the local reparameterisation.

    reference: M = alpha*mu, V = sigma^2 * alpha^2          (LRT:170-171, MNF:195-196)
    exact    : V = alpha*(sigma^2 + (1-alpha)*mu^2)         (spike-and-slab variance; not in
               the reference -- BASELINE.json north_star formula, kept as an option)
    """
    a = alpha_of(lambdal)
    s = sigma_of(weight_rho)
    m = weight_mu * a
    if var_mode == "reference":
        v = s ** 2 * a ** 2
    elif var_mode == "exact":
        v = a * (s ** 2 + (1.0 - a) * weight_mu ** 2)
    else:
        raise ValueError(var_mode)
    return m, v


def lrt_forward(x, p, eps=None, sample=True, z=None, var_mode="reference"):
    """Activations of one LRT layer.

    sample branch  LRT:169-175: e_b + sqrt(var_b)*eps with e_b = x M^T + b_mu,
                                var_b = x^2 V^T + sigma_b^2
    mean branch    LRT:177-180: x M^T + b_mu
    MNF (z given)  MNF:197,205: the mean GEMM sees x*z, the variance GEMM plain x^2.
    `p` maps weight_mu, weight_rho, lambdal, bias_mu, bias_rho to tensors.
    """
    m, v = lrt_weight_moments(p["weight_mu"], p["weight_rho"], p["lambdal"], var_mode)
    xin = x if z is None else x * z
    e_b = torch.mm(xin, m.T) + p["bias_mu"]
    if not sample:
        return e_b
    var_b = torch.mm(x ** 2, v.T) + sigma_of(p["bias_rho"]) ** 2
    return e_b + torch.sqrt(var_b) * eps


def lrt_kl(p, priors: Priors = Priors(), z=None):
    """Closed-form KL of one layer.  Bias term LRT:185-186 (eq 4.5), weight term LRT:189-192
    (eq 4.4).  With z (MNF:230-233) the mean enters as (mu*z - mu_prior)."""
    sb = sigma_of(p["bias_rho"])
    kl_bias = (math.log(priors.bias_sigma) - torch.log(sb) - 0.5
               + (sb ** 2 + (p["bias_mu"] - priors.bias_mu) ** 2) / (2.0 * priors.bias_sigma ** 2)).sum()
    a = alpha_of(p["lambdal"])
    s = sigma_of(p["weight_rho"])
    mu = p["weight_mu"] if z is None else p["weight_mu"] * z
    # torch.log(sigma_prior / sigma) in the reference; written as one log of the ratio here too
    slab = (torch.log(priors.sigma / s) - 0.5 + torch.log(a / priors.alpha)
            + (s ** 2 + (mu - priors.mu) ** 2) / (2.0 * priors.sigma ** 2))
    spike = (1.0 - a) * torch.log((1.0 - a) / (1.0 - priors.alpha))
    return kl_bias + (a * slab + spike).sum()


def lrt_net_forward(x, layers, eps=None, sample=True, var_mode="reference"):
    """BayesianNetwork.forward, LRT:206-211: view(-1, in) -> [layer -> relu]* -> layer -> log_softmax."""
    h = x.reshape(-1, layers[0]["weight_mu"].shape[1])
    last = len(layers) - 1
    for i, p in enumerate(layers):
        h = lrt_forward(h, p, None if eps is None else eps[i], sample, var_mode=var_mode)
        h = F.relu(h) if i < last else F.log_softmax(h, dim=1)
    return h


def lrt_net_loss(x, y, layers, eps, num_batches, priors: Priors = Priors(), var_mode="reference"):
    """One training objective, LRT:222-224: nll(sum) + (kl_1+kl_2+kl_3)/NUM_BATCHES.
    Returns (loss, nll, kl, log-probs)."""
    logp = lrt_net_forward(x, layers, eps, True, var_mode)
    nll = F.nll_loss(logp, y, reduction="sum")
    kl = sum(lrt_kl(p, priors) for p in layers)
    return nll + kl / num_batches, nll, kl, logp


def init_lrt_params(rng, in_features, out_features, mu_range=0.2, dtype=torch.float32):
    """Parameter draw with the reference's ranges (LRT:137-151: mu~U(-.2,.2), rho~U(-5,-4),
    lambda~U(0,1); MNF:140 uses mu~U(-.01,.01)) from a numpy Generator so that fixtures and
    the GPU tests regenerate identical values on any machine."""
    def u(lo, hi, *shape):
        return torch.from_numpy(rng.uniform(lo, hi, size=shape)).to(dtype)
    return {
        "weight_mu": u(-mu_range, mu_range, out_features, in_features),
        "weight_rho": u(-5.0, -4.0, out_features, in_features),
        "lambdal": u(0.0, 1.0, out_features, in_features),
        "bias_mu": u(-0.2, 0.2, out_features),
        "bias_rho": u(-5.0, -4.0, out_features),
    }


# --------------------------------------------------------------------------------------
# MF (full weight sampling) layer  (LBBNN-GP-MF.py:74-319; sim-study variant MFsim:173-300)
# --------------------------------------------------------------------------------------
LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


class _StdGammaReparam(torch.autograd.Function):
    """A given standard-gamma draw g0 ~ Gamma(a, 1) with torch's implicit reparameterisation gradient
    dg0/da (torch/distributions/gamma.py:79-87 -> torch._standard_gamma / _standard_gamma_grad)."""

    @staticmethod
    def forward(ctx, a, g0):
        ctx.save_for_backward(a.detach(), g0)
        return g0.clone()

    @staticmethod
    def backward(ctx, grad):
        a, g0 = ctx.saved_tensors
        return grad * torch._standard_gamma_grad(a, g0), None


def gamma_rsample(a, b, g0):
    """tau = Gamma(a, b).rsample() for the injected standard-gamma draw g0 (MF:141)."""
    return _StdGammaReparam.apply(a, g0) / b


def relaxed_bernoulli_rsample(alpha, u, temperature=0.001):
    """RelaxedBernoulli(probs=alpha, T).rsample() for the injected uniform u (MF:115;
    torch relaxed_bernoulli.py:104-112 + SigmoidTransform's clipped sigmoid)."""
    fi = torch.finfo(alpha.dtype)
    p = alpha.clamp(min=fi.eps, max=1 - fi.eps)
    uu = u.clamp(min=fi.eps, max=1 - fi.eps)
    logits = (uu.log() - (-uu).log1p() + p.log() - (-p).log1p()) / temperature
    return torch.clamp(torch.sigmoid(logits), min=fi.tiny, max=1.0 - fi.eps)


def exact_bernoulli_sample(alpha, u):
    """Bernoulli(alpha).sample() as a mask of the injected uniform: gamma = [u < alpha] (MF:113)."""
    return (u < alpha).to(alpha.dtype)


def gaussian_log_prob_iid(w, mu, sigma):
    """MF:93-96."""
    return -LOG_SQRT_2PI - torch.log(sigma) - ((w - mu) ** 2) / (2 * sigma ** 2)


def gaussian_full_log_prob(w, gamma, mu, sigma):
    """MF:99-101: sum log(gamma * N(w; mu, sigma) + (1-gamma) + 1e-8)."""
    return torch.log(gamma * torch.exp(gaussian_log_prob_iid(w, mu, sigma)) + (1 - gamma) + 1e-8).sum()


def bernoulli_log_prob(gamma, alpha, exact):
    """MF:122-128."""
    g = torch.round(gamma.detach()) if exact else gamma
    return (g * torch.log(alpha + 1e-8) + (1 - g) * torch.log(1 - alpha + 1e-8)).sum()


def gauss_gamma_log_prob(w, gamma, a, b, tau, exact):
    """MF:140-151 with tau already drawn."""
    g = torch.round(gamma.detach()) if exact else gamma
    const = a * torch.log(b) + (a - 0.5) * tau - b * tau - torch.lgamma(a) - 0.5 * math.log(2 * math.pi)
    return (g * const - tau * w ** 2 + (1 - g) + 1e-8).sum()


def beta_binomial_log_prob(gamma, pa, pb, exact):
    """MF:162-173 (nine lgamma terms, two of which cancel)."""
    g = torch.round(gamma.detach()) if exact else gamma
    one = torch.ones_like(gamma)
    lg = torch.lgamma
    return (lg(one) + lg(g + one * pa) + lg(one * (1 + pb) - g) + lg(one * (pa + pb)) - lg(one * pa + g)
            - lg(one * 2 - g) - lg(one * (1 + pa + pb)) - lg(one * pa) - lg(one * pb)).sum()


def mf_forward(x, p, cgamma, noise, sample=True, medimean=False, alpha_stale=None, calc_log_probs=True,
               exact=(False, False, False, False), logprob_on_ws=False):
    """MF BayesianLinear.forward, MF:228-255.

    p: weight_mu, weight_rho, lambdal, bias_mu, bias_rho, weight_a, weight_b, bias_a, bias_b, pa, pb.
    noise: eps_w (out,in), eps_b (out,), g0_w (1,), g0_b (out,) -- the normal draws of MF:85-87 and the
    standard-gamma draws behind MF:141.  exact = (.exact of gamma, weight_prior, bias_prior, gamma_prior).
    logprob_on_ws: the sim-study variant evaluates the weight log-probs at the unmasked ws (MFsim:233,237).
    Returns (F.linear output, log_prior, log_variational_posterior).
    """
    sw, sb = sigma_of(p["weight_rho"]), sigma_of(p["bias_rho"])
    ws = None
    if sample:
        ws = p["weight_mu"] + sw * noise["eps_w"]
        weight = cgamma * ws
        bias = p["bias_mu"] + sb * noise["eps_b"]
    elif medimean:
        weight = cgamma * p["weight_mu"]
        bias = p["bias_mu"]
    else:
        weight = alpha_stale * p["weight_mu"]
        bias = p["bias_mu"]
    log_prior = log_q = 0
    if calc_log_probs:
        alpha = alpha_of(p["lambdal"])
        wlp = ws if (logprob_on_ws and ws is not None) else weight
        tau_w = gamma_rsample(p["weight_a"], p["weight_b"], noise["g0_w"])
        tau_b = gamma_rsample(p["bias_a"], p["bias_b"], noise["g0_b"])
        log_prior = (gauss_gamma_log_prob(wlp, cgamma, p["weight_a"], p["weight_b"], tau_w, exact[1])
                     + gauss_gamma_log_prob(bias, torch.ones_like(bias), p["bias_a"], p["bias_b"], tau_b, exact[2])
                     + beta_binomial_log_prob(cgamma, p["pa"], p["pb"], exact[3]))
        log_q = (gaussian_full_log_prob(wlp, cgamma, p["weight_mu"], sw)
                 + bernoulli_log_prob(cgamma, alpha, exact[0])
                 + gaussian_log_prob_iid(bias, p["bias_mu"], sb).sum())
    return F.linear(x, weight, bias), log_prior, log_q


def mf_net_elbo(x, y, layers, noises, us, num_batches, temperature=0.001, gamma_exact=False):
    """sample_elbo with SAMPLES=1, MF:285-319: gamma_k = rsample(alpha_k) per layer (relaxed unless
    gamma_exact), forward, loss = nll + (log_q - log_p)/NUM_BATCHES."""
    h = x.reshape(-1, layers[0]["weight_mu"].shape[1])
    log_p = log_q = 0
    gammas = []
    for i, (p, nz, u) in enumerate(zip(layers, noises, us)):
        alpha = alpha_of(p["lambdal"])
        g = exact_bernoulli_sample(alpha, u) if gamma_exact else relaxed_bernoulli_rsample(alpha, u, temperature)
        gammas.append(g)
        h, lp, lq = mf_forward(h, p, g, nz, exact=(gamma_exact, False, False, False))
        log_p, log_q = log_p + lp, log_q + lq
        h = F.relu(h) if i < len(layers) - 1 else F.log_softmax(h, dim=1)
    nll = F.nll_loss(h, y, reduction="sum")
    return nll + (log_q - log_p) / num_batches, nll, log_p, log_q, h, gammas


def mfsim_elbo(x, y, p, noise, u, num_batches, temperature=0.001, gamma_exact=False, gamma=None):
    """Simulation-study objective with SAMPLES=1, MFsim:272-300: one 20->1 MF layer whose log-probs are taken at the
    unmasked ws (MFsim:233,237), sigmoid output, BCELoss(sum); loss = nll + (log_q - log_p)/NUM_BATCHES.
    `gamma` overrides the relaxed draw (injected-gamma parity).  Returns (loss, nll, log_p, log_q, out)."""
    alpha = alpha_of(p["lambdal"])
    if gamma is None:
        gamma = exact_bernoulli_sample(alpha, u) if gamma_exact else relaxed_bernoulli_rsample(alpha, u, temperature)
    h, lp, lq = mf_forward(x, p, gamma, noise, exact=(gamma_exact, False, False, False), logprob_on_ws=True)
    out = torch.sigmoid(h)
    nll = F.binary_cross_entropy(out, y.unsqueeze(1).to(out.dtype), reduction="sum")
    return nll + (lq - lp) / num_batches, nll, lp, lq, out


def init_mf_params(rng, in_features, out_features, dtype=torch.float32, sim=False):
    """MF:192-220 ranges (sim study MFsim:182-191: mu~U(-.01,.01), lambda~U(-.5,.5))."""
    def u(lo, hi, *shape):
        return torch.from_numpy(rng.uniform(lo, hi, size=shape)).to(dtype)
    mr, lr = (0.01, (-0.5, 0.5)) if sim else (0.2, (0.0, 1.0))
    return {
        "weight_mu": u(-mr, mr, out_features, in_features), "weight_rho": u(-5, -4, out_features, in_features),
        "lambdal": u(lr[0], lr[1], out_features, in_features),
        "bias_mu": u(-0.2, 0.2, out_features), "bias_rho": u(-5, -4, out_features),
        "weight_a": u(1, 1.1, 1), "weight_b": u(1, 1.1, 1), "bias_a": u(1, 1.1, out_features),
        "bias_b": u(1, 1.1, out_features), "pa": u(1, 1.1, 1), "pb": u(1, 1.1, 1),
    }


# --------------------------------------------------------------------------------------
# Flows (flows2.py) and the MNF layer (LBBNN-GP-MF-MNF.py:133-239; sim variant MNFsim:145-251)
# --------------------------------------------------------------------------------------
def mlp_forward(h, linears, leaky=0.1):
    """flows2.MLP (flows2:176-185): Linear -> LeakyReLU(0.1) ..., activation after the last Linear dropped."""
    for i, (w, b) in enumerate(linears):
        h = F.linear(h, w, b)
        if i < len(linears) - 1:
            h = F.leaky_relu(h, leaky)
    return h


def rnvp_forward(z, mask, tp):
    """flows2.RNVP.forward/log_det (flows2:206-219).  tp: {"net": [(W,b)...], "t": (W,b), "s": (W,b)}.
    Note (1-gate)*shift is added on ALL dims, masked ones included (flows2:215)."""
    z1, z2 = (1 - mask) * z, mask * z
    y = mlp_forward(z2, tp["net"])
    shift, scale = F.linear(y, *tp["t"]), F.linear(y, *tp["s"])
    gate = torch.sigmoid(scale)
    x = (z1 * gate + (1 - gate) * shift) + z2
    return x, ((1 - mask) * gate.log()).sum(-1)


def iaf_forward(z, mask, tp):
    """flows2.MNF (flows2:225-241): h = tanh(f(m z)); out = m z + (1-m)(z sigma + (1-sigma) mu);
    log_det = ((1-m) log sigma).sum() over ALL dims (a scalar even for batched z, flows2:241)."""
    h = torch.tanh(F.linear(mask * z, *tp["f"]))
    mu, sigma = F.linear(h, *tp["g"]), torch.sigmoid(F.linear(h, *tp["k"]))
    return mask * z + (1 - mask) * (z * sigma + (1 - sigma) * mu), ((1 - mask) * sigma.log()).sum()


def propagate_flow(z, masks, transforms, kind="RNVP"):
    """flows2.PropagateFlow.forward (flows2:41-46): returns the transformed TENSOR and the summed log-dets."""
    logdet = 0
    f = rnvp_forward if kind == "RNVP" else iaf_forward
    for m, tp in zip(masks, transforms):
        z, ld = f(z, m, tp)
        logdet = logdet + ld
    return z, logdet


def mnf_sample_z(p, eps_z, masks, kind="RNVP"):
    """sample_z (MNF:182-187): z0 = q0_mean + exp(q0_log_var)^0.5 * eps (B,in); z_flow; returns the LAST ROW of
    the flowed batch, the squeezed log-dets, and z0 (what the layer stores in self.z)."""
    q0_std = p["q0_log_var"].exp().sqrt().repeat(eps_z.shape[0], 1)
    z0 = p["q0_mean"] + q0_std * eps_z
    zs, logdet = propagate_flow(z0, masks, p["z_flow"], kind)
    return zs[-1], (logdet.squeeze() if torch.is_tensor(logdet) else logdet), z0


def mnf_forward(x, p, noise, sample=True, calc_kl=True, priors: Priors = Priors(), kind="RNVP"):
    """MNF BayesianLinear.forward (MNF:190-239) with every draw injected:
    noise = {eps_z (B,in), z_masks [T x (B,in)], eps (B,out),                      # activation branch
             eps_z2 (1,in), z_masks2 [T x (1,in)], eps_r (out,), r_masks [T x (in,)]}  # KL branch
    Returns (activations, kl)."""
    z_k, _, _ = mnf_sample_z(p, noise["eps_z"], noise["z_masks"], kind)
    act = lrt_forward(x, p, noise.get("eps"), sample=sample, z=z_k)        # MNF:195-200 / 203-206
    if not calc_kl:
        return act, 0
    z2, log_det_q, z0 = mnf_sample_z(p, noise["eps_z2"], noise["z_masks2"], kind)   # MNF:210, self.z := (1,in)
    a, s = alpha_of(p["lambdal"]), sigma_of(p["weight_rho"])
    w_mean, w_var = z2 * p["weight_mu"] * a, s ** 2 * a ** 2
    log_q0 = (-0.5 * math.log(math.pi) - 0.5 * p["q0_log_var"]
              - 0.5 * ((z0 - p["q0_mean"]) ** 2 / p["q0_log_var"].exp())).sum()           # log pi, not log 2 pi
    log_q = -log_det_q + log_q0
    act_mu, act_var = p["r0_c"] @ w_mean.T, p["r0_c"] ** 2 @ w_var.T
    a_r = torch.tanh(act_mu + act_var.sqrt() * noise["eps_r"])
    mean_r = p["r0_b1"].outer(a_r).mean(-1)
    log_var_r = p["r0_b2"].outer(a_r).mean(-1)
    z_b, log_det_r = propagate_flow(z2, noise["r_masks"], p["r_flow"], kind)            # 1-D input
    log_rb = (-0.5 * math.log(math.pi) - 0.5 * log_var_r
              - 0.5 * ((z_b[-1] - mean_r) ** 2 / log_var_r.exp())).sum()                 # z_b[-1]: a SCALAR
    log_r = log_det_r + log_rb
    return act, lrt_kl(p, priors, z=z2) + log_q - log_r


def mnf_net_loss(x, y, layers, noises, num_batches, priors: Priors = Priors(), kind="RNVP"):
    """MNF BayesianNetwork.forward + train objective (MNF:244-275)."""
    h = x.reshape(-1, layers[0]["weight_mu"].shape[1])
    kl = 0
    for i, (p, nz) in enumerate(zip(layers, noises)):
        h, k = mnf_forward(h, p, nz, priors=priors, kind=kind)
        kl = kl + k
        h = F.relu(h) if i < len(layers) - 1 else F.log_softmax(h, dim=1)
    nll = F.nll_loss(h, y, reduction="sum")
    return nll + kl / num_batches, nll, kl, h


def init_flow_params(rng, dim, num_transforms=2, h_sizes=(75, 75, 75, 75), kind="RNVP", hidden=100, dtype=torch.float32):
    """nn.Linear default init ranges (U(+-1/sqrt(fan_in))) from a numpy Generator."""
    def lin(i, o):
        k = 1.0 / math.sqrt(i)
        return (torch.from_numpy(rng.uniform(-k, k, size=(o, i))).to(dtype), torch.from_numpy(rng.uniform(-k, k, size=(o,))).to(dtype))
    out = []
    for _ in range(num_transforms):
        if kind == "RNVP":
            sizes = [dim] + list(h_sizes)
            out.append({"net": [lin(a, b) for a, b in zip(sizes[:-1], sizes[1:])], "t": lin(sizes[-1], dim), "s": lin(sizes[-1], dim)})
        else:
            out.append({"f": lin(dim, hidden), "g": lin(hidden, dim), "k": lin(hidden, dim)})
    return out


def init_mnf_params(rng, in_features, out_features, num_transforms=2, h_sizes=(75, 75, 75, 75), kind="RNVP", dtype=torch.float32):
    """MNF:140-176: LRT params with mu~U(-.01,.01) + q0/r0 vectors + the two flows."""
    p = init_lrt_params(rng, in_features, out_features, mu_range=0.01, dtype=dtype)
    n = lambda s=0.1, off=0.0: torch.from_numpy(off + s * rng.standard_normal(size=(in_features,))).to(dtype)  # noqa: E731
    p.update({"q0_mean": n(), "q0_log_var": n(0.1, -9.0), "r0_c": n(), "r0_b1": n(), "r0_b2": n(),
              "z_flow": init_flow_params(rng, in_features, num_transforms, h_sizes, kind, dtype=dtype),
              "r_flow": init_flow_params(rng, in_features, num_transforms, h_sizes, kind, dtype=dtype)})
    return p


# --------------------------------------------------------------------------------------
# variational dropout (VD = variational_dropout.py), SURVEY.md §8(f) rank 4
# --------------------------------------------------------------------------------------
VD_KL_C = (1.16145124, -1.50204118, 0.58629921)       # VD:98


def vd_forward(x, theta, alpha, zeta):
    """BayesianLayer.forward (VD:63-68): theta is (n, m) = (in, out) -- the "NN" operand layout -- and alpha (m,) one
    dropout rate per output neuron: phi = x theta, delta = (x^2 theta^2) alpha, act = phi + sqrt(delta) zeta."""
    phi = torch.matmul(x, theta)
    delta = torch.matmul(x ** 2, theta ** 2) * alpha
    return phi + torch.sqrt(delta) * zeta


def vd_kl(alpha):
    """Per-layer term of loss_fn (VD:98-102): sum 0.5 log a + c1 a + c2 a^2 + c3 a^3 (added to the loss as written)."""
    c1, c2, c3 = VD_KL_C
    return (0.5 * torch.log(alpha) + c1 * alpha + c2 * alpha ** 2 + c3 * alpha ** 3).sum()


def vd_net_loss(x, y, layers, zetas, num_batches):
    """BNN.forward (VD:79-85) + loss_fn (VD:88-106): relu between layers, log_softmax at the end,
    loss = KL / num_batches + nll_loss(sum).  layers: [{"theta", "alpha"}]."""
    h = x.reshape(-1, layers[0]["theta"].shape[0])
    kl = 0
    for i, (p, z) in enumerate(zip(layers, zetas)):
        h = vd_forward(h, p["theta"], p["alpha"], z)
        h = F.relu(h) if i < len(layers) - 1 else F.log_softmax(h, dim=1)
        kl = kl + vd_kl(p["alpha"])
    nll = F.nll_loss(h, y, reduction="sum")
    return kl / num_batches + nll, nll, kl, h


def init_vd_params(rng, n, m, alpha=0.2, dtype=torch.float32):
    """VD:58-61: theta ~ U(-0.1, 0.1) (n, m); alpha = 0.2 for every output neuron."""
    return {"theta": torch.from_numpy(rng.uniform(-0.1, 0.1, size=(n, m))).to(dtype),
            "alpha": torch.full((m,), alpha, dtype=dtype)}
