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
