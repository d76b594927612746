"""Pin the oracle restatement against outputs of the reference's own classes (tests/golden/*.npz,
produced by tests/golden/make_golden.py in the build container).  CPU only."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases as C
import lbbnn_oracle as O

TOL = 2e-6   # fp32 oracle vs fp32 reference: same op order up to log(a/b) vs log a - log b


def _npz(name):
    return np.load(os.path.join(C.GOLDEN, name))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_lrt_layer_matches_reference(tag):
    g = _npz("lrt_layer.npz")
    seed, b, i, o, spread = (int(v) for v in g[f"{tag}_meta"])
    case = C.lrt_layer_case(seed, b, i, o, spread_lambda=bool(spread))
    p = {k: v.clone().requires_grad_(True) for k, v in case["p"].items()}
    x = case["x"].clone().requires_grad_(True)
    act = O.lrt_forward(x, p, case["eps"], sample=True)
    kl = O.lrt_kl(p)
    ((act * case["gout"]).sum() + kl / C.NUM_BATCHES).backward()
    assert C.rel_err(act, g[f"{tag}_act"]) < TOL
    assert abs(kl.item() - float(g[f"{tag}_kl"])) / abs(float(g[f"{tag}_kl"])) < TOL
    assert C.rel_err(x.grad, g[f"{tag}_dx"]) < TOL
    for k in p:
        assert C.rel_err(p[k].grad, g[f"{tag}_d_{k}"]) < 5e-6, k
    with torch.no_grad():
        assert C.rel_err(O.lrt_forward(case["x"], case["p"], sample=False), g[f"{tag}_act_mean"]) < TOL
        assert C.rel_err(O.lrt_forward(case["x"], case["p"], case["eps"], sample=True),
                         g[f"{tag}_act_eval_sample"]) < TOL
        assert abs(O.lrt_kl(case["p"]).item() - float(g[f"{tag}_kl_eval"])) / abs(float(g[f"{tag}_kl_eval"])) < TOL


def test_lrt_mnist_net_matches_reference():
    g = _npz("lrt_net_mnist.npz")
    case = C.lrt_net_case(seed=0, batch=100)
    layers = [{k: v.clone().requires_grad_(True) for k, v in p.items()} for p in case["layers"]]
    loss, nll, kl, logp = O.lrt_net_loss(case["x"], case["y"], layers, case["eps"], C.NUM_BATCHES)
    loss.backward()
    assert C.rel_err(logp, g["logp"]) < TOL
    for name, val in (("nll", nll), ("kl", kl), ("loss", loss)):
        assert abs(val.item() - float(g[name])) / abs(float(g[name])) < TOL, name
    for li, p in enumerate(layers):
        for k, v in p.items():
            d = C.grad_digest(v.grad)
            assert C.rel_err(d["sample"], g[f"l{li}_{k}_sample"]) < 5e-6, (li, k)
            assert abs(d["l2"] - float(g[f"l{li}_{k}_l2"])) / float(g[f"l{li}_{k}_l2"]) < 5e-6, (li, k)
            if f"l{li}_{k}_full" in g:
                assert C.rel_err(v.grad, g[f"l{li}_{k}_full"]) < 5e-6, (li, k)
    with torch.no_grad():
        mean_logp = O.lrt_net_forward(case["x"], case["layers"], sample=False)
        samp = O.lrt_net_forward(case["x"], case["layers"], case["eps"], sample=True)
    assert C.rel_err(mean_logp, g["mean_logp"]) < TOL
    assert np.array_equal(mean_logp.argmax(1).numpy(), g["mean_argmax"])      # bit-exact predictions
    assert np.array_equal(samp.argmax(1).numpy(), g["eval_sample_argmax"])


def test_fp64_oracle_brackets_fp32():
    """The fp64 evaluation of the oracle is the truth both fp32 implementations are judged against."""
    case = C.lrt_layer_case(11, 9, 37, 23)
    a32 = O.lrt_forward(case["x"], case["p"], case["eps"])
    a64 = O.lrt_forward(case["x"].double(), {k: v.double() for k, v in case["p"].items()}, case["eps"].double())
    assert C.rel_err(a32, a64) < 1e-6


MF_NAMES = ["weight_mu", "weight_rho", "lambdal", "bias_mu", "bias_rho", "weight_a", "weight_b", "bias_a", "bias_b",
            "pa", "pb"]


@pytest.mark.parametrize("key", ["ma_rel", "ma_ex", "mb_rel", "mb_ex", "sa_rel", "sa_ex"])
def test_mf_layer_matches_reference(key):
    g = _npz("mf_layer.npz")
    seed, b, i, o, sim = (int(v) for v in g[f"{key}_meta"])
    relaxed = key.endswith("_rel")
    case = C.mf_layer_case(seed, b, i, o, sim=bool(sim))
    p = {k: v.clone().requires_grad_(True) for k, v in case["p"].items()}
    x = case["x"].clone().requires_grad_(True)
    alpha = O.alpha_of(p["lambdal"])
    cg = O.relaxed_bernoulli_rsample(alpha, case["u"]) if relaxed else O.exact_bernoulli_sample(alpha, case["u"])
    assert np.array_equal(cg.detach().numpy(), g[f"{key}_gamma"])           # inclusion masks bit-exact
    act, lp, lq = O.mf_forward(x, p, cg, case["noise"], exact=(not relaxed, False, False, False),
                               logprob_on_ws=bool(sim))
    ((act * case["gout"]).sum() + (lq - lp) / C.NUM_BATCHES).backward()
    assert C.rel_err(act, g[f"{key}_act"]) < TOL
    assert abs(lp.item() - float(g[f"{key}_log_prior"])) / abs(float(g[f"{key}_log_prior"])) < 5e-6
    assert abs(lq.item() - float(g[f"{key}_log_q"])) / abs(float(g[f"{key}_log_q"])) < 5e-6
    assert C.rel_err(x.grad, g[f"{key}_dx"]) < TOL
    for k in MF_NAMES:
        assert C.rel_err(p[k].grad, g[f"{key}_d_{k}"]) < 2e-5, k


def test_mf_means_match_reference():
    g = _npz("mf_layer.npz")
    for key, (seed, b, i, o, sim) in {"ma": (41, 7, 37, 23, False), "mb": (42, 33, 130, 10, False), "sa": (43, 9, 20, 1, True)}.items():
        case = C.mf_layer_case(seed, b, i, o, sim=sim)
        alpha = O.alpha_of(case["p"]["lambdal"])
        med, _, _ = O.mf_forward(case["x"], case["p"], (alpha > 0.5).float(), None, sample=False, medimean=True,
                                 calc_log_probs=False)
        jm, _, _ = O.mf_forward(case["x"], case["p"], None, None, sample=False, medimean=False, alpha_stale=alpha,
                                calc_log_probs=False)
        assert C.rel_err(med, g[f"{key}_medimean"]) < TOL and C.rel_err(jm, g[f"{key}_jointmean"]) < TOL


def test_mf_mnist_elbo_matches_reference():
    g = _npz("mf_net_mnist.npz")
    case = C.mf_net_case(seed=50, batch=100)
    layers = [{k: v.clone().requires_grad_(True) for k, v in p.items()} for p in case["layers"]]
    loss, nll, lp, lq, _, _ = O.mf_net_elbo(case["x"], case["y"], layers, case["noises"], case["us"], C.NUM_BATCHES)
    loss.backward()
    for name, val in (("loss", loss), ("nll", nll), ("log_prior", lp), ("log_q", lq)):
        assert abs(val.item() - float(g[name])) / abs(float(g[name])) < 5e-6, name
    for li, p in enumerate(layers):
        for k, v in p.items():
            d = C.grad_digest(v.grad)
            assert C.rel_err(d["sample"], g[f"l{li}_{k}_sample"]) < 2e-5, (li, k)
            if f"l{li}_{k}_full" in g:
                assert C.rel_err(v.grad, g[f"l{li}_{k}_full"]) < 2e-5, (li, k)


def _mnf_oracle(case, priors=O.Priors(), dtype=torch.float32, kind="RNVP"):
    named = {k: v.to(dtype).clone().requires_grad_(True) for k, v in C.flat_named(case["p"]).items()}
    p = C.unflatten_like(case["p"], named)
    x = case["x"].to(dtype).clone().requires_grad_(True)
    nz = {k: ([m.to(dtype) for m in v] if isinstance(v, list) else v.to(dtype)) for k, v in case["noise"].items()}
    act, kl = O.mnf_forward(x, p, nz, priors=priors, kind=kind)
    ((act * case["gout"].to(dtype)).sum() + kl / C.NUM_BATCHES).backward()
    return act, kl, x, named


@pytest.mark.parametrize("key", ["ma", "mb", "sa"])
def test_mnf_layer_matches_reference(key):
    g = _npz("mnf_layer.npz")
    seed, b, i, o, nh, hw = (int(v) for v in g[f"{key}_meta"])
    case = C.mnf_layer_case(seed, b, i, o, h_sizes=(hw,) * nh)
    pri = O.Priors(0.1, 1.3, 0.3, 0.0, 1.3) if key == "sa" else O.Priors()        # MNFsim:157-174
    act, kl, x, named = _mnf_oracle(case, pri)
    assert C.rel_err(act, g[f"{key}_act"]) < TOL
    assert abs(kl.item() - float(g[f"{key}_kl"])) / abs(float(g[f"{key}_kl"])) < 5e-6
    assert C.rel_err(x.grad, g[f"{key}_dx"]) < 5e-6
    for name, v in named.items():
        ref = g[f"{key}_d_{name}"]
        got = v.grad if v.numel() <= 4000 else torch.from_numpy(C.grad_digest(v.grad)["sample"])
        assert C.rel_err(got, ref) < 5e-5, name


@pytest.mark.parametrize("key", ["a", "b"])
def test_mnf_layer_with_iaf_flows_matches_reference(key):
    """Z_FLOW_TYPE = R_FLOW_TYPE = 'MNF' (flows2:225-241): the KL branch's log_det_q covers the KL row only (MNF:210)."""
    g = _npz("mnf_layer_iaf.npz")
    seed, b, i, o = (int(v) for v in g[f"{key}_meta"])
    case = C.mnf_layer_case(seed, b, i, o, kind="MNF")
    act, kl, x, named = _mnf_oracle(case, kind="MNF")
    assert C.rel_err(act, g[f"{key}_act"]) < TOL
    assert abs(kl.item() - float(g[f"{key}_kl"])) / abs(float(g[f"{key}_kl"])) < 5e-6
    assert C.rel_err(x.grad, g[f"{key}_dx"]) < 5e-6
    for name, v in named.items():
        ref = g[f"{key}_d_{name}"]
        got = v.grad if v.numel() <= 4000 else torch.from_numpy(C.grad_digest(v.grad)["sample"])
        assert C.rel_err(got, ref) < 5e-5, name


@pytest.mark.parametrize("kind", ["RNVP", "MNF"])
def test_flows_match_reference(kind):
    g = _npz("flows.npz")
    rng = np.random.default_rng(0)
    tmpl = O.init_flow_params(rng, 24, 2, (75, 75, 75, 75), kind)
    named = {k: torch.from_numpy(g[f"{kind}_p_{k}"]) for k in C.flat_named({"z_flow": tmpl})}
    tps = C.unflatten_like({"z_flow": tmpl}, named)["z_flow"]
    for tag in ("b", "v"):
        z = torch.from_numpy(g[f"{kind}_{tag}_z"]).requires_grad_(True)
        masks = [torch.from_numpy(m) for m in g[f"{kind}_{tag}_masks"]]
        zo, ld = O.propagate_flow(z, masks, tps, kind)
        ((zo * zo).sum() + ld.sum()).backward()
        assert C.rel_err(zo, g[f"{kind}_{tag}_out"]) < TOL
        assert C.rel_err(ld, g[f"{kind}_{tag}_logdet"]) < TOL
        assert C.rel_err(z.grad, g[f"{kind}_{tag}_dz"]) < 5e-6


def test_mnf_mnist_net_matches_reference():
    g = _npz("mnf_net_mnist.npz")
    case = C.mnf_net_case(seed=90, batch=100)
    named = [{k: v.clone().requires_grad_(True) for k, v in C.flat_named(p).items()} for p in case["layers"]]
    layers = [C.unflatten_like(p, n) for p, n in zip(case["layers"], named)]
    loss, nll, kl, logp = O.mnf_net_loss(case["x"], case["y"], layers, case["noises"], C.NUM_BATCHES)
    loss.backward()
    assert C.rel_err(logp, g["logp"]) < TOL
    assert abs(nll.item() - float(g["nll"])) / float(g["nll"]) < 5e-6 and abs(kl.item() - float(g["kl"])) / float(g["kl"]) < 5e-6
    for li, n in enumerate(named):
        for name, v in n.items():
            got = torch.from_numpy(C.grad_digest(v.grad)["sample"]) if v.numel() > 2000 else v.grad
            assert C.rel_err(got, g[f"l{li}_{name}"]) < 5e-5, (li, name)


def test_sim_study_network_matches_reference():
    """BASELINE.json configs[0]: sample_elbo of the 20 -> 1 simulation-study model (MFsim:250-300) on the first
    minibatch of the synthetic stand-in data, relaxed gamma drawn from the replayed uniform."""
    g = _npz("mfsim_net.npz")
    case = C.sim_study_case(seed=70)
    B = case["batch"]
    p = {k: v.clone().requires_grad_(True) for k, v in case["p"].items()}
    loss, nll, lp, lq, out = O.mfsim_elbo(case["X"][:B], case["y"][:B], p, case["noises"][0], case["us"][0], case["num_batches"])
    loss.backward()
    assert C.rel_err(out, g["out"]) < TOL
    for name, val in (("loss", loss), ("nll", nll), ("log_prior", lp), ("log_q", lq)):
        assert abs(val.item() - float(g[name])) <= 5e-6 * abs(float(g[name])), name
    for k, v in p.items():
        assert C.rel_err(v.grad, g["d_" + k]) < 5e-5, k


# ---- variational dropout (variational_dropout.py) -----------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_vd_layer_matches_reference(tag):
    g = _npz("vd.npz")
    seed, b, n, m, spread = (int(v) for v in g[f"{tag}_meta"])
    case = C.vd_layer_case(seed, b, n, m, spread_alpha=bool(spread))
    theta = case["p"]["theta"].clone().requires_grad_(True)
    alpha = case["p"]["alpha"].clone().requires_grad_(True)
    x = case["x"].clone().requires_grad_(True)
    act = O.vd_forward(x, theta, alpha, case["zeta"])
    (act * case["gout"]).sum().backward()
    assert C.rel_err(act, g[f"{tag}_act"]) < TOL
    assert C.rel_err(x.grad, g[f"{tag}_dx"]) < TOL
    assert C.rel_err(theta.grad, g[f"{tag}_d_theta"]) < TOL
    assert C.rel_err(alpha.grad, g[f"{tag}_d_alpha"]) < TOL


def test_vd_net_matches_reference():
    g = _npz("vd.npz")
    case = C.vd_net_case(seed=80, batch=100)
    layers = [{k: v.clone().requires_grad_(True) for k, v in p.items()} for p in case["layers"]]
    loss, nll, kl, logp = O.vd_net_loss(case["x"], case["y"], layers, case["zetas"], 600.0)
    loss.backward()
    assert C.rel_err(logp, g["net_logp"]) < TOL
    assert abs(loss.item() - float(g["net_loss"])) / abs(float(g["net_loss"])) < TOL
    for li, p in enumerate(layers):
        d = C.grad_digest(p["theta"].grad)
        assert C.rel_err(d["sample"], g[f"net_l{li}_theta_sample"]) < 5e-6, li
        assert abs(d["l2"] - float(g[f"net_l{li}_theta_l2"])) / float(g[f"net_l{li}_theta_l2"]) < 5e-6, li
        assert C.rel_err(p["alpha"].grad, g[f"net_l{li}_d_alpha"]) < 5e-6, li


def test_vd_training_steps_match_reference():
    """Two run_epoch training iterations (VD:160-166) with the script's AdamW(model.parameters(), lr=1e-4): the oracle's
    loss + torch AdamW over theta ONLY (alpha is not a registered parameter of the reference, VD:61) reproduces the reference's
    losses and parameters; alpha stays 0.2."""
    g = _npz("vd_train.npz")
    case = C.vd_net_case(seed=87, batch=100)
    layers = [{k: v.clone().requires_grad_(k == "theta") for k, v in p.items()} for p in case["layers"]]
    opt = torch.optim.AdamW([p["theta"] for p in layers], lr=1e-4)
    rng = np.random.default_rng(870)
    for step in range(int(g["n_steps"])):
        zetas = [C.t(rng.standard_normal(size=tuple(z.shape))) for z in case["zetas"]]
        opt.zero_grad()
        loss = O.vd_net_loss(case["x"], case["y"], layers, zetas, 600.0)[0]
        loss.backward()
        opt.step()
        assert abs(loss.item() - float(g[f"loss_{step}"])) / abs(float(g[f"loss_{step}"])) < TOL, step
    for li, p in enumerate(layers):
        d = C.grad_digest(p["theta"].detach())
        assert C.rel_err(d["sample"], g[f"theta_l{li}_sample"]) < 5e-6, li
        assert abs(d["l2"] - float(g[f"theta_l{li}_l2"])) / float(g[f"theta_l{li}_l2"]) < 5e-6, li
        assert np.array_equal(g[f"alpha_l{li}"], np.full_like(g[f"alpha_l{li}"], 0.2)), li


def test_ensemble_outputs_of_the_reference_networks():
    """tests/golden/ensemble.npz: the oracle's LRT / MNF forwards reproduce every MC sample's log-softmax output of the
    reference networks in eval mode (the inputs of test_ensemble / outofsample's statistics)."""
    g = _npz("ensemble.npz")
    S, B = 12, 52
    case = C.ensemble_case(seed=120, batch=B, samples=S, kind="lrt")
    for s in (0, 5, 11):
        logp = O.lrt_net_forward(case["x"], case["layers"], [e[s] for e in case["eps"]], True)
        assert C.rel_err(logp, g["lrt_outputs"][s]) < TOL
    assert C.rel_err(O.lrt_net_forward(case["x"], case["layers"], None, False), g["lrt_posterior_mean_logp"]) < TOL
    S = 6
    case = C.ensemble_case(seed=121, batch=B, samples=S, kind="mnf")
    for s in (0, 5):
        h = case["x"]
        for l, p in enumerate(case["layers"]):
            nz = {"eps_z": case["eps_z"][l][s], "z_masks": [m[s] for m in case["z_masks"][l]], "eps": case["eps"][l][s]}
            h, _ = O.mnf_forward(h, p, nz, sample=True, calc_kl=False)
            h = torch.relu(h) if l < 2 else torch.log_softmax(h, 1)
        assert C.rel_err(h, g["mnf_outputs"][s]) < TOL
