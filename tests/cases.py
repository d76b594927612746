"""Seeded synthetic cases shared by tests/golden/make_golden.py (which runs the reference on them)
and the tests (which run the oracle and the CUDA path on them).  numpy Generators only, so the
values are identical on every machine; nothing here depends on torch's RNG."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "bayesian-neural-nets_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import lbbnn_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
MNIST_SIZES = [(784, 400), (400, 600), (600, 10)]
NUM_BATCHES = 600


def t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def lrt_layer_case(seed, batch, in_features, out_features, mu_range=0.2, spread_lambda=False):
    """One layer: params by the reference's init ranges, x~U[0,1), eps~N(0,1), random upstream grad."""
    rng = np.random.default_rng(seed)
    p = O.init_lrt_params(rng, in_features, out_features, mu_range)
    if spread_lambda:  # trained-like inclusion probabilities spanning (0,1)
        p["lambdal"] = t(rng.normal(0.0, 2.0, size=(out_features, in_features)))
    x = t(rng.uniform(0.0, 1.0, size=(batch, in_features)))
    eps = t(rng.standard_normal(size=(batch, out_features)))
    gout = t(rng.standard_normal(size=(batch, out_features)))
    return {"p": p, "x": x, "eps": eps, "gout": gout}


def lrt_net_case(seed, batch, sizes=MNIST_SIZES, classes=10):
    rng = np.random.default_rng(seed)
    layers = [O.init_lrt_params(rng, i, o) for i, o in sizes]
    x = t(rng.uniform(0.0, 1.0, size=(batch, sizes[0][0])))
    y = torch.from_numpy(rng.integers(0, classes, size=(batch,))).long()
    eps = [t(rng.standard_normal(size=(batch, o))) for _, o in sizes]
    return {"layers": layers, "x": x, "y": y, "eps": eps}


def grad_digest(g, stride=97):
    """Small, position-sensitive summary of a big gradient tensor: strided sample + moments."""
    flat = g.detach().reshape(-1).double()
    return {
        "sample": flat[::stride].float().numpy(),
        "sum": np.float64(flat.sum().item()),
        "abs": np.float64(flat.abs().sum().item()),
        "l2": np.float64(flat.pow(2).sum().sqrt().item()),
    }


def rel_err(a, b):
    """max|a-b| / max|b|  -- the tolerance definition of SURVEY.md §4."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    den = b.abs().max().item()
    if den == 0.0:
        den = 1.0
    return (a - b).abs().max().item() / den


def mf_layer_case(seed, batch, in_features, out_features, sim=False, spread_lambda=True):
    """MF layer: params (reference ranges), x, the layer's noise draws and a uniform for gamma."""
    rng = np.random.default_rng(seed)
    p = O.init_mf_params(rng, in_features, out_features, sim=sim)
    if spread_lambda:
        p["lambdal"] = t(rng.normal(0.0, 2.0, size=(out_features, in_features)))
    x = t(rng.uniform(0.0, 1.0, size=(batch, in_features)))
    noise = {"eps_w": t(rng.standard_normal(size=(out_features, in_features))),
             "eps_b": t(rng.standard_normal(size=(out_features,))),
             "g0_w": t(rng.gamma(1.05, size=(1,))), "g0_b": t(rng.gamma(1.05, size=(out_features,)))}
    u = t(rng.uniform(0.0, 1.0, size=(out_features, in_features)))
    gout = t(rng.standard_normal(size=(batch, out_features)))
    return {"p": p, "x": x, "noise": noise, "u": u, "gout": gout}


def mf_net_case(seed, batch, sizes=MNIST_SIZES, classes=10):
    rng = np.random.default_rng(seed)
    layers, noises, us = [], [], []
    for i, o in sizes:
        layers.append(O.init_mf_params(rng, i, o))
        noises.append({"eps_w": t(rng.standard_normal(size=(o, i))), "eps_b": t(rng.standard_normal(size=(o,))),
                       "g0_w": t(rng.gamma(1.05, size=(1,))), "g0_b": t(rng.gamma(1.05, size=(o,)))})
        us.append(t(rng.uniform(0.0, 1.0, size=(o, i))))
    x = t(rng.uniform(0.0, 1.0, size=(batch, sizes[0][0])))
    y = torch.from_numpy(rng.integers(0, classes, size=(batch,))).long()
    return {"layers": layers, "noises": noises, "us": us, "x": x, "y": y}


def _masks(rng, T, *shape):
    """RealNVP / IAF masks: bernoulli(0.5) as [u < 0.5] (flows2:209,234)."""
    return [t((rng.uniform(0.0, 1.0, size=shape) < 0.5).astype(np.float32)) for _ in range(T)]


def mnf_noise(rng, batch, in_features, out_features, T=2):
    return {"eps_z": t(rng.standard_normal(size=(batch, in_features))), "z_masks": _masks(rng, T, batch, in_features),
            "eps": t(rng.standard_normal(size=(batch, out_features))),
            "eps_z2": t(rng.standard_normal(size=(1, in_features))), "z_masks2": _masks(rng, T, 1, in_features),
            "eps_r": t(rng.standard_normal(size=(out_features,))), "r_masks": _masks(rng, T, in_features)}


def mnf_layer_case(seed, batch, in_features, out_features, T=2, h_sizes=(75, 75, 75, 75), kind="RNVP"):
    rng = np.random.default_rng(seed)
    p = O.init_mnf_params(rng, in_features, out_features, T, h_sizes, kind)
    # a wider q0 than the init (-9) so that z and the flows matter numerically
    p["q0_log_var"] = t(-2.0 + 0.3 * rng.standard_normal(size=(in_features,)))
    p["q0_mean"] = t(1.0 + 0.1 * rng.standard_normal(size=(in_features,)))
    x = t(rng.uniform(0.0, 1.0, size=(batch, in_features)))
    gout = t(rng.standard_normal(size=(batch, out_features)))
    return {"p": p, "x": x, "gout": gout, "noise": mnf_noise(rng, batch, in_features, out_features, T)}


def mnf_net_case(seed, batch, sizes=MNIST_SIZES, classes=10, T=2):
    rng = np.random.default_rng(seed)
    layers = [O.init_mnf_params(rng, i, o, T) for i, o in sizes]
    x = t(rng.uniform(0.0, 1.0, size=(batch, sizes[0][0])))
    y = torch.from_numpy(rng.integers(0, classes, size=(batch,))).long()
    noises = [mnf_noise(rng, batch, i, o, T) for i, o in sizes]
    return {"layers": layers, "noises": noises, "x": x, "y": y}


def flat_named(p, prefix=""):
    """Flatten a (possibly nested) MNF parameter dict into {name: tensor} with the reference's state_dict names."""
    out = {}
    for k, v in p.items():
        if k in ("z_flow", "r_flow"):
            for ti, tp in enumerate(v):
                if "net" in tp:
                    for li, (w, b) in enumerate(tp["net"]):
                        out[f"{k}.transforms.{ti}.network.{2 * li}.weight"] = w
                        out[f"{k}.transforms.{ti}.network.{2 * li}.bias"] = b
                    for nm in ("t", "s"):
                        out[f"{k}.transforms.{ti}.{nm}.weight"], out[f"{k}.transforms.{ti}.{nm}.bias"] = tp[nm]
                else:
                    for nm in ("f", "g", "k"):
                        out[f"{k}.transforms.{ti}.{nm}.weight"], out[f"{k}.transforms.{ti}.{nm}.bias"] = tp[nm]
        else:
            out[k] = v
    return out


def unflatten_like(p, named):
    """Inverse of flat_named: rebuild the nested dict `p` from {name: tensor}."""
    out = {}
    for k, v in p.items():
        if k in ("z_flow", "r_flow"):
            ts = []
            for ti, tp in enumerate(v):
                g = lambda nm: (named[f"{k}.transforms.{ti}.{nm}.weight"], named[f"{k}.transforms.{ti}.{nm}.bias"])  # noqa: E731
                if "net" in tp:
                    ts.append({"net": [g(f"network.{2 * li}") for li in range(len(tp["net"]))], "t": g("t"), "s": g("s")})
                else:
                    ts.append({"f": g("f"), "g": g("g"), "k": g("k")})
            out[k] = ts
        else:
            out[k] = named[k]
    return out


SIM_TRUE_WEIGHTS = np.array([-4., 0., 1., 0., 0., 0., 1., 0., 0., 0., 1.2, 0., 37.1, 0., 0., 50., -0.00005, 10., 3., 0.])  # MFsim:350


def sim_study_case(seed, n=2000, batch=400, steps=5):
    """BASELINE.json configs[0]: the reference's CSVs are not shipped, so X ~ N(0,1) (n,20) with
    columns 6:19 standardised as MFsim:48-51, y ~ Bernoulli(sigmoid(X true_weights)) with MFsim:350's true_weights;
    sim-study parameter init (MFsim:182-191); per-step noise draws of the layer."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal(size=(n, 20))
    X[:, 6:19] = (X[:, 6:19] - X[:, 6:19].mean(0)) / X[:, 6:19].std(0)
    y = (rng.uniform(size=n) < 1.0 / (1.0 + np.exp(-X @ SIM_TRUE_WEIGHTS))).astype(np.int64)
    p = O.init_mf_params(rng, 20, 1, sim=True)
    noises = [{"eps_w": t(rng.standard_normal(size=(1, 20))), "eps_b": t(rng.standard_normal(size=(1,))),
               "g0_w": t(rng.gamma(1.05, size=(1,))), "g0_b": t(rng.gamma(1.05, size=(1,)))} for _ in range(steps)]
    us = [t(rng.uniform(0.0, 1.0, size=(1, 20))) for _ in range(steps)]
    return {"X": t(X), "y": torch.from_numpy(y), "p": p, "noises": noises, "us": us, "batch": batch, "num_batches": n / batch}


VD_SIZES = [(784, 1200), (1200, 1200), (1200, 1200), (1200, 10)]      # BNN of variational_dropout.py:74-77


def vd_layer_case(seed, batch, n, m, spread_alpha=False):
    """One variational-dropout layer (VD:55-68): theta~U(-.1,.1) (n,m), alpha = 0.2 (or spread over (0.05, 1)),
    x ~ N(0,1) like the normalised MNIST input (VD:38-40), zeta ~ N(0,1), random upstream gradient."""
    rng = np.random.default_rng(seed)
    p = O.init_vd_params(rng, n, m)
    if spread_alpha:
        p["alpha"] = t(rng.uniform(0.05, 1.0, size=(m,)))
    x = t(rng.standard_normal(size=(batch, n)))
    zeta = t(rng.standard_normal(size=(batch, m)))
    gout = t(rng.standard_normal(size=(batch, m)))
    return {"p": p, "x": x, "zeta": zeta, "gout": gout}


def vd_net_case(seed, batch, sizes=VD_SIZES, classes=10):
    rng = np.random.default_rng(seed)
    layers = [O.init_vd_params(rng, n, m) for n, m in sizes]
    x = t(rng.standard_normal(size=(batch, sizes[0][0])))
    y = torch.from_numpy(rng.integers(0, classes, size=(batch,))).long()
    zetas = [t(rng.standard_normal(size=(batch, m))) for _, m in sizes]
    return {"layers": layers, "x": x, "y": y, "zetas": zetas}


def ensemble_case(seed, batch, samples, sizes=MNIST_SIZES, kind="lrt", T=2):
    """Inputs of the posterior-predictive loops (test_ensemble / outofsample): a network (reference init; lambdal spread so
    the inclusion probabilities span (0,1)), a test batch and the noise of every MC sample: LRT eps[l] (S, B, out); MNF
    additionally eps_z[l] (S, B, in) and z_masks[l] = T x (S, B, in) -- the reference pushes all B rows through the flow
    and keeps the last (MNF:187)."""
    rng = np.random.default_rng(seed)
    init = O.init_lrt_params if kind == "lrt" else (lambda r, i, o: O.init_mnf_params(r, i, o, T))
    layers = [init(rng, i, o) for i, o in sizes]
    for p in layers:
        p["lambdal"] = t(rng.normal(0.0, 2.0, size=tuple(p["lambdal"].shape)))
        if kind != "lrt":      # a q0 wide enough for z to matter
            p["q0_log_var"] = t(-2.0 + 0.3 * rng.standard_normal(size=tuple(p["q0_log_var"].shape)))
            p["q0_mean"] = t(1.0 + 0.1 * rng.standard_normal(size=tuple(p["q0_mean"].shape)))
    x = t(rng.uniform(0.0, 1.0, size=(batch, sizes[0][0])))
    eps = [t(rng.standard_normal(size=(samples, batch, o))) for _, o in sizes]
    out = {"layers": layers, "x": x, "eps": eps}
    if kind != "lrt":
        out["eps_z"] = [t(rng.standard_normal(size=(samples, batch, i))) for i, _ in sizes]
        out["z_masks"] = [[t((rng.uniform(0.0, 1.0, size=(samples, batch, i)) < 0.5).astype(np.float32)) for _ in range(T)]
                          for i, _ in sizes]
    return out
