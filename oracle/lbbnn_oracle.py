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
