"""GPU parity tests of the MNF path: the fused flow kernels (flows2.py RNVP / IAF-style 'MNF' transform) and the MNF
layer / network (LBBNN-GP-MF-MNF.py:133-260) vs the oracle and the reference's golden outputs, all noise injected.
fp32 tolerance 1e-5 (max|a-b|/max|b|) on activations / flow outputs, 5e-5 on gradients that pass through the deep
coupling stacks (same bound the oracle itself meets against the reference, tests/test_oracle_golden.py)."""
import os

import numpy as np
import pytest
import torch

import cases as C
import lbbnn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
GTOL = 5e-5


@pytest.fixture(scope="module")
def lb():
    import lbbnn
    return lbbnn


def _load_flow(flow, named, prefix):
    sd = {k[len(prefix):]: v for k, v in named.items() if k.startswith(prefix)}
    flow.load_state_dict(sd)
    return flow


def _flow_grads(flow, prefix):
    return {prefix + k: v.grad for k, v in flow.named_parameters()}


@pytest.mark.parametrize("kind", ["RNVP", "MNF"])
def test_flow_kernels_match_reference_golden(lb, kind):
    g = np.load(os.path.join(C.GOLDEN, "flows.npz"))
    rng = np.random.default_rng(0)
    tmpl = O.init_flow_params(rng, 24, 2, (75, 75, 75, 75), kind)
    named = {k: torch.from_numpy(g[f"{kind}_p_{k}"]) for k in C.flat_named({"z_flow": tmpl})}
    flow = _load_flow(lb.flows.PropagateFlow(kind, 24, 2), named, "z_flow.").cuda()
    for tag in ("b", "v"):
        z = torch.from_numpy(g[f"{kind}_{tag}_z"]).cuda().requires_grad_(True)
        masks = [torch.from_numpy(m).cuda() for m in g[f"{kind}_{tag}_masks"]]
        zo, ld = flow(z, masks)
        assert zo.shape == z.shape and tuple(ld.shape) == tuple(g[f"{kind}_{tag}_logdet"].shape)
        ((zo * zo).sum() + ld.sum()).backward()
        assert C.rel_err(zo.detach(), g[f"{kind}_{tag}_out"]) < TOL
        assert C.rel_err(ld.detach(), g[f"{kind}_{tag}_logdet"]) < TOL
        assert C.rel_err(z.grad, g[f"{kind}_{tag}_dz"]) < TOL


@pytest.mark.parametrize("kind,dim,rows,h_sizes", [("RNVP", 784, 3, (75, 75, 75, 75)), ("RNVP", 20, 1, (50,) * 5),
                                                    ("RNVP", 400, 2, (75, 75, 75, 75)), ("MNF", 600, 2, None),
                                                    ("RNVP", 37, 7, (33, 128))])
def test_flow_kernels_all_gradients_match_oracle(lb, kind, dim, rows, h_sizes):
    rng = np.random.default_rng(dim + rows)
    tmpl = O.init_flow_params(rng, dim, 2, h_sizes or (75,) * 4, kind)
    named = C.flat_named({"z_flow": tmpl})
    z0 = C.t(rng.standard_normal(size=(rows, dim)))
    masks = C._masks(rng, 2, rows, dim)
    gz = C.t(rng.standard_normal(size=(rows, dim)))
    gl = C.t(rng.standard_normal(size=(rows,)))
    # fp64 truth and fp32 oracle
    outs = {}
    for dt in (torch.float64, torch.float32):
        nm = {k: v.to(dt).clone().requires_grad_(True) for k, v in named.items()}
        tps = C.unflatten_like({"z_flow": tmpl}, nm)["z_flow"]
        z = z0.to(dt).clone().requires_grad_(True)
        zo, ld = O.propagate_flow(z, [m.to(dt) for m in masks], tps, kind)
        ldv = ld if ld.dim() else ld.expand(rows) / rows       # IAF kind: one scalar over everything
        ((zo * gz.to(dt)).sum() + (ldv * gl.to(dt)).sum()).backward()
        outs[dt] = (zo.detach(), ld.detach(), z.grad, {k: v.grad for k, v in nm.items()})
    kw = dict(h_sizes=h_sizes) if kind == "RNVP" else {}
    flow = _load_flow(lb.flows.PropagateFlow(kind, dim, 2, **kw), named, "z_flow.").cuda()
    z = z0.cuda().requires_grad_(True)
    zo, ld = flow(z, [m.cuda() for m in masks])
    ldv = ld if ld.dim() else ld.expand(rows) / rows
    ((zo * gz.cuda()).sum() + (ldv * gl.cuda()).sum()).backward()
    t64, t32 = outs[torch.float64], outs[torch.float32]
    assert C.rel_err(zo.detach(), t64[0]) < TOL and C.rel_err(ld.detach(), t64[1]) < TOL
    assert C.rel_err(z.grad, t64[2]) < GTOL
    got = _flow_grads(flow, "z_flow.")
    for k, ref in t64[3].items():
        # not worse than a few times the fp32 oracle's own distance from the fp64 truth
        bound = max(GTOL, 4 * C.rel_err(t32[3][k], ref))
        assert C.rel_err(got[k], ref) < bound, k


def test_flow_native_masks_reproducible_from_philox(lb):
    """Masks drawn inside the kernel equal [u < 0.5] of the exported Philox uniforms (bit-exact inclusion masks):
    the oracle fed with the exported masks reproduces the native forward."""
    rng = np.random.default_rng(5)
    dim, rows = 400, 4
    tmpl = O.init_flow_params(rng, dim, 2)
    named = C.flat_named({"z_flow": tmpl})
    lb.manual_seed(77)
    flow = _load_flow(lb.flows.PropagateFlow("RNVP", dim, 2), named, "z_flow.").cuda()
    z = C.t(rng.standard_normal(size=(rows, dim)))
    with torch.no_grad():
        zo, ld = flow(z.cuda())
        zo2, _ = flow(z.cuda())
    seed, stream = flow.last_noise_key
    assert not torch.equal(zo, zo2), "every call must draw fresh masks"
    masks = [(lb.philox_uniform(rows * dim, seed, stream + t, device="cuda") < 0.5).float().view(rows, dim) for t in range(2)]
    frac = torch.stack(masks).mean().item()
    assert 0.45 < frac < 0.55
    zo_inj, ld_inj = flow(z.cuda(), masks)     # injected path with the exported masks: bit-identical
    assert torch.equal(zo_inj.detach(), zo2) and ld_inj.shape == ld.shape
    ref, ref_ld = O.propagate_flow(z, [m.cpu() for m in masks], tmpl, "RNVP")
    assert C.rel_err(zo2, ref) < TOL


def _make_mnf_layer(lb, case, i, o, h_sizes, priors=None):
    kw = {}
    if priors is not None:
        kw = dict(mu_prior=priors.mu, sigma_prior=priors.sigma, alpha_prior=priors.alpha, bias_mu_prior=priors.bias_mu,
                  bias_sigma_prior=priors.bias_sigma)
    layer = lb.mnf.BayesianLinear(i, o, 2, h_sizes=h_sizes, **kw)
    named = C.flat_named(case["p"] if "p" in case else case)
    layer.load_state_dict(named)
    return layer.cuda()


def _cuda_noise(nz):
    return {k: ([m.cuda() for m in v] if isinstance(v, list) else v.cuda()) for k, v in nz.items()}


@pytest.mark.parametrize("key", ["ma", "mb", "sa"])
def test_mnf_layer_matches_oracle_and_reference(lb, key):
    g = np.load(os.path.join(C.GOLDEN, "mnf_layer.npz"))
    seed, b, i, o, nh, hw = (int(v) for v in g[f"{key}_meta"])
    case = C.mnf_layer_case(seed, b, i, o, h_sizes=(hw,) * nh)
    pri = O.Priors(0.1, 1.3, 0.3, 0.0, 1.3) if key == "sa" else O.Priors()        # MNFsim:157-174
    # fp64 truth
    named64 = {k: v.double().clone().requires_grad_(True) for k, v in C.flat_named(case["p"]).items()}
    p64 = C.unflatten_like(case["p"], named64)
    x64 = case["x"].double().clone().requires_grad_(True)
    nz64 = {k: ([m.double() for m in v] if isinstance(v, list) else v.double()) for k, v in case["noise"].items()}
    act64, kl64 = O.mnf_forward(x64, p64, nz64, priors=pri)
    ((act64 * case["gout"].double()).sum() + kl64 / C.NUM_BATCHES).backward()

    layer = _make_mnf_layer(lb, case, i, o, (hw,) * nh, pri)
    layer.train()
    x = case["x"].cuda().requires_grad_(True)
    act = layer(x, noise=_cuda_noise(case["noise"]))
    ((act * case["gout"].cuda()).sum() + layer.kl / C.NUM_BATCHES).backward()
    # vs the reference's own outputs
    assert C.rel_err(act.detach(), g[f"{key}_act"]) < TOL
    assert abs(layer.kl.item() - float(g[f"{key}_kl"])) / abs(float(g[f"{key}_kl"])) < TOL
    assert C.rel_err(layer.z.detach(), g[f"{key}_z"]) < TOL
    assert C.rel_err(x.grad, g[f"{key}_dx"]) < TOL
    # vs the fp64 oracle, every parameter
    assert C.rel_err(act.detach(), act64.detach()) < TOL
    got = dict(layer.named_parameters())
    assert set(got) == set(named64)
    for name, v in named64.items():
        assert got[name].grad is not None, name
        assert C.rel_err(got[name].grad, v.grad) < GTOL, name
        ref = g[f"{key}_d_{name}"]
        mine = got[name].grad if v.numel() <= 4000 else torch.from_numpy(C.grad_digest(got[name].grad.cpu())["sample"])
        assert C.rel_err(mine, ref) < GTOL, name


@pytest.mark.parametrize("key", ["a", "b"])
def test_mnf_layer_with_iaf_flows_matches_oracle_and_reference(lb, key):
    """z_flow_type = r_flow_type = 'MNF' (the IAF-style transform, flows2:225-241) as the layer's flows: the KL branch's
    log_det_q must be the KL row's own log-determinant although the activation row shares the flow launch."""
    g = np.load(os.path.join(C.GOLDEN, "mnf_layer_iaf.npz"))
    seed, b, i, o = (int(v) for v in g[f"{key}_meta"])
    case = C.mnf_layer_case(seed, b, i, o, kind="MNF")
    named64 = {k: v.double().clone().requires_grad_(True) for k, v in C.flat_named(case["p"]).items()}
    p64 = C.unflatten_like(case["p"], named64)
    x64 = case["x"].double().clone().requires_grad_(True)
    nz64 = {k: ([m.double() for m in v] if isinstance(v, list) else v.double()) for k, v in case["noise"].items()}
    act64, kl64 = O.mnf_forward(x64, p64, nz64, kind="MNF")
    ((act64 * case["gout"].double()).sum() + kl64 / C.NUM_BATCHES).backward()

    layer = lb.mnf.BayesianLinear(i, o, 2, z_flow_type="MNF", r_flow_type="MNF")
    layer.load_state_dict(C.flat_named(case["p"]))
    layer = layer.cuda().train()
    x = case["x"].cuda().requires_grad_(True)
    act = layer(x, noise=_cuda_noise(case["noise"]))
    ((act * case["gout"].cuda()).sum() + layer.kl / C.NUM_BATCHES).backward()
    assert C.rel_err(act.detach(), g[f"{key}_act"]) < TOL
    assert abs(layer.kl.item() - float(g[f"{key}_kl"])) / abs(float(g[f"{key}_kl"])) < TOL
    assert abs(layer.kl.item() - kl64.item()) / abs(kl64.item()) < TOL
    assert C.rel_err(x.grad, g[f"{key}_dx"]) < TOL
    got = dict(layer.named_parameters())
    assert set(got) == set(named64)
    for name, v in named64.items():
        assert got[name].grad is not None, name
        assert C.rel_err(got[name].grad, v.grad) < GTOL, name
        ref = g[f"{key}_d_{name}"]
        mine = got[name].grad if v.numel() <= 4000 else torch.from_numpy(C.grad_digest(got[name].grad.cpu())["sample"])
        assert C.rel_err(mine, ref) < GTOL, name


def test_mnf_layer_eval_branches(lb):
    """eval: posterior-mean branch is still stochastic in z (MNF:202-206), .kl == 0 unless calculate_log_probs."""
    case = C.mnf_layer_case(71, 9, 50, 13)
    layer = _make_mnf_layer(lb, case, 50, 13, (75,) * 4)
    layer.eval()
    nz = _cuda_noise(case["noise"])
    with torch.no_grad():
        act = layer(case["x"].cuda(), noise=nz)
        assert layer.kl == 0
        ref, _ = O.mnf_forward(case["x"], case["p"], case["noise"], sample=False, calc_kl=False)
        assert C.rel_err(act, ref) < TOL
        act_s = layer(case["x"].cuda(), sample=True, calculate_log_probs=True, noise=nz)
        ref_s, ref_kl = O.mnf_forward(case["x"], case["p"], case["noise"], sample=True, calc_kl=True)
        assert C.rel_err(act_s, ref_s) < TOL and abs(layer.kl.item() - ref_kl.item()) / abs(ref_kl.item()) < TOL
        # native noise: runs, finite, differs call to call
        a1, a2 = layer(case["x"].cuda(), sample=True), layer(case["x"].cuda(), sample=True)
        assert torch.isfinite(a1).all() and not torch.equal(a1, a2)


def test_mnf_mnist_net_matches_reference(lb):
    g = np.load(os.path.join(C.GOLDEN, "mnf_net_mnist.npz"))
    case = C.mnf_net_case(seed=90, batch=100)
    net = lb.mnf.BayesianNetwork()
    for l, p in zip(net.layers, case["layers"]):
        l.load_state_dict(C.flat_named(p))
    net = net.cuda().train()
    logp = net(case["x"].cuda(), sample=True, noises=[_cuda_noise(n) for n in case["noises"]])
    nll = torch.nn.functional.nll_loss(logp, case["y"].cuda(), reduction="sum")
    kl = net.kl()
    (nll + kl / C.NUM_BATCHES).backward()
    assert C.rel_err(logp.detach(), g["logp"]) < TOL
    assert torch.equal(logp.argmax(1).cpu(), torch.from_numpy(g["logp"]).argmax(1))
    assert abs(nll.item() - float(g["nll"])) / float(g["nll"]) < TOL and abs(kl.item() - float(g["kl"])) / float(g["kl"]) < TOL
    for li, l in enumerate(net.layers):
        for name, v in l.named_parameters():
            got = torch.from_numpy(C.grad_digest(v.grad.cpu())["sample"]) if v.numel() > 2000 else v.grad
            assert C.rel_err(got, g[f"l{li}_{name}"]) < GTOL, (li, name)


@pytest.mark.parametrize("kind", ["mnf", "mf", "lrt"])
def test_graphed_trainer_replays_the_eager_step_with_fresh_noise(lb, kind):
    """GraphedTrainer captures the eager modules' training step: with lr = 0 the parameters stay put and only the native
    noise changes from replay to replay (device step counter), with lr > 0 the loss goes down on a fixed batch."""
    torch.manual_seed(3)
    rng = np.random.default_rng(8)
    x = C.t(rng.uniform(0, 1, size=(64, 784)))
    y = torch.from_numpy(rng.integers(0, 10, size=(64,))).long()
    make = {"mnf": lambda: lb.mnf.BayesianNetwork(), "mf": lambda: lb.mf.BayesianNetwork(),
            "lrt": lambda: lb.BayesianNetwork()}[kind]
    objective = "elbo" if kind == "mf" else "kl"
    net = make().cuda()
    p0 = [p.detach().clone() for p in net.parameters()]
    tr = lb.GraphedTrainer(net, batch_size=64, num_batches=C.NUM_BATCHES, lr=0.0, objective=objective)
    outs = [tr.step(x, y) for _ in range(4)]
    assert all(np.isfinite(o["loss"]) for o in outs)
    assert len({o["nll"] for o in outs}) == 4                      # fresh noise on every replay
    assert all(torch.equal(a, b.detach()) for a, b in zip(p0, net.parameters()))
    assert tr.step_dev.item() == 3 + 4                             # warm-up steps + replays (the capture pass does not run)
    # pipelined host-buffer steps: statistics arrive one call late, flush() delivers the last
    outs = [tr.step_async(x, y) for _ in range(3)]
    last = tr.flush()
    assert outs[0] is None and all(np.isfinite(o["loss"]) and o["nll"] > 0 for o in outs[1:] + [last])
    assert len({o["nll"] for o in outs[1:] + [last]}) == 3 and tr.step_dev.item() == 3 + 4 + 3
    net2 = make().cuda()
    tr2 = lb.GraphedTrainer(net2, batch_size=64, num_batches=C.NUM_BATCHES, lr=1e-3, objective=objective)
    nll = [tr2.step(x, y)["nll"] for _ in range(60)]
    assert np.mean(nll[-10:]) < np.mean(nll[:10])                  # it trains


def test_fused_objective_matches_the_torch_formulation(lb):
    """lbbnn_nll_kl_objective_f32 (GraphedTrainer's loss head: nll_loss(log_softmax) + sum(kl) / NUM_BATCHES and
    d loss / d logits in one launch, backward started at the logits and kl terms) against F.log_softmax / F.nll_loss /
    net.kl() and loss.backward() on the SAME forward graph of the MNF network: loss, nll, and every parameter gradient
    (MNF:267-272)."""
    import ctypes
    import torch.nn.functional as F
    from lbbnn import _capi as K
    torch.manual_seed(4)
    rng = np.random.default_rng(9)
    net = lb.mnf.BayesianNetwork((72, 40, 24, 10)).cuda().train()
    x = C.t(rng.uniform(0, 1, size=(37, 72))).cuda()
    y = torch.from_numpy(rng.integers(0, 10, size=(37,))).long().cuda()
    logits = net._logits(x, sample=True)
    kls = [l.kl for l in net.layers]
    params = list(net.parameters())
    nll = F.nll_loss(F.log_softmax(logits, dim=1), y, reduction="sum")
    loss = nll + sum(kls) / C.NUM_BATCHES
    ref = torch.autograd.grad(loss, params, retain_graph=True, allow_unused=True)
    out, dlogits = torch.zeros(2, device="cuda"), torch.empty_like(logits)
    ptrs = (ctypes.c_void_p * len(kls))(*[k.data_ptr() for k in kls])
    K.check(K.lib.lbbnn_nll_kl_objective_f32(K.ptr(logits), K.ptr(y, torch.int64), 37, 10, ptrs, None, len(kls),
                                             1.0 / C.NUM_BATCHES, K.ptr(out), K.ptr(dlogits), K.current_stream()))
    assert abs(out[0].item() - loss.item()) <= 1e-6 * abs(loss.item())
    assert abs(out[1].item() - nll.item()) <= 1e-6 * abs(nll.item())
    for p in params:
        p.grad = None
    torch.autograd.backward([logits] + kls, [dlogits] + [torch.full_like(k, 1.0 / C.NUM_BATCHES) for k in kls])
    n_checked = 0
    for (name, p), r in zip(net.named_parameters(), ref):
        if r is None:
            assert p.grad is None, name
            continue
        assert C.rel_err(p.grad, r) < 2e-6, name
        n_checked += 1
    assert n_checked > 100                                          # weights, biases, q0 / r0 terms and both flows of 3 layers
    # the trainer takes this path for the MNF network, and the torch formulation when asked to
    tr = lb.GraphedTrainer(lb.mnf.BayesianNetwork((72, 40, 24, 10)).cuda(), batch_size=37, num_batches=C.NUM_BATCHES, lr=0.0,
                           objective="kl", in_features=72)
    assert tr._dlogits is not None
    tr_t = lb.GraphedTrainer(lb.mnf.BayesianNetwork((72, 40, 24, 10)).cuda(), batch_size=37, num_batches=C.NUM_BATCHES, lr=0.0,
                             objective="kl", in_features=72, fuse_objective=False)
    assert tr_t._dlogits is None
    a, b = tr.step(x.cpu(), y.cpu()), tr_t.step(x.cpu(), y.cpu())
    assert np.isfinite(a["loss"]) and np.isfinite(b["loss"]) and a["loss"] > a["nll"] > 0


@pytest.mark.parametrize("D,O", [(784, 400), (50, 10), (33, 7)])
def test_aux_kl_kernels_match_the_eager_formulation(lb, D, O):
    """log_q0 - log_rb (MNF:212-227) from the fused kernels (csrc/mnf_aux.cu) against the same terms written with torch
    ops in float64, values and every gradient."""
    import math
    from lbbnn.mnf import _AuxKL
    rng = np.random.default_rng(D + O)
    mk = lambda *s, scale=1.0, off=0.0: (C.t(rng.standard_normal(size=s)) * scale + off).cuda().requires_grad_(True)  # noqa: E731
    q0_mean, q0_lv, r0_c, b1, b2 = mk(D, scale=0.1), mk(D, scale=0.1, off=-3.0), mk(D, scale=0.1), mk(D, scale=0.1), mk(D, scale=0.1)
    z0, z2, z_b = mk(1, D, scale=0.3), mk(D, scale=0.5, off=1.0), mk(D)
    M0, V = mk(O, D, scale=0.05), (C.t(rng.random(size=(O, D))) * 1e-3 + 1e-5).cuda().requires_grad_(True)
    eps_r = C.t(rng.standard_normal(size=(O,))).cuda()
    ins = [q0_mean, q0_lv, z0, r0_c, b1, b2, z2, M0, V, eps_r, z_b]
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = _AuxKL.apply(*ins, ticket)
    out.backward()
    assert int(ticket) == 0
    got = [t.grad.clone() for t in ins if t.requires_grad]
    d = [t.detach().double().requires_grad_(t.requires_grad) for t in ins]
    q0_mean, q0_lv, z0, r0_c, b1, b2, z2, M0, V, eps_r, z_b = d
    log_q0 = (-0.5 * math.log(math.pi) - 0.5 * q0_lv - 0.5 * ((z0 - q0_mean) ** 2 / q0_lv.exp())).sum()
    a_r = torch.tanh((r0_c * z2) @ M0.T + ((r0_c ** 2) @ V.T).sqrt() * eps_r)
    mean_r, lvr = b1 * a_r.mean(), b2 * a_r.mean()
    log_rb = (-0.5 * math.log(math.pi) - 0.5 * lvr - 0.5 * ((z_b[-1] - mean_r) ** 2 / lvr.exp())).sum()
    ref = log_q0 - log_rb
    ref.backward()
    assert abs(out.item() - ref.item()) / abs(ref.item()) < 1e-5
    for g, t, name in zip(got, [t for t in d if t.requires_grad],
                          ["q0_mean", "q0_log_var", "z0", "r0_c", "r0_b1", "r0_b2", "z2", "M0", "V", "z_b"]):
        assert C.rel_err(g, t.grad.float()) < 2e-5, name

