"""GPU parity tests of the mean-field (full weight sampling) path vs the oracle and the reference's golden
outputs.  fp32 tolerance 1e-5 (max|a-b|/max|b|) on activations and weight-shaped gradients, 5e-5 on the
lgamma/digamma-derived scalars; hard inclusion masks bit-exact."""
import os

import numpy as np
import pytest
import torch

import cases as C
import lbbnn_oracle as O

pytestmark = pytest.mark.gpu
MF_NAMES = ["weight_mu", "weight_rho", "lambdal", "bias_mu", "bias_rho", "weight_a", "weight_b", "bias_a", "bias_b",
            "pa", "pb"]


@pytest.fixture(scope="module")
def lb():
    import lbbnn
    return lbbnn


def _cuda_noise(nz):
    return {k: v.cuda() for k, v in nz.items()}


def _make_layer(lb, case, i, o, sim):
    layer = lb.mf.BayesianLinear(i, o, 1, logprob_on_ws=sim).cuda()
    with torch.no_grad():
        for k, v in case["p"].items():
            getattr(layer, k).copy_(v)
    return layer


def test_exact_gamma_mask_bit_exact_and_relaxed_close(lb):
    case = C.mf_layer_case(42, 33, 130, 10)
    alpha = O.alpha_of(case["p"]["lambdal"])
    view = lb.mf.MFBernoulliView(alpha.cuda())
    view.exact = True
    g = view.rsample(case["u"].cuda())
    assert torch.equal(g.cpu(), O.exact_bernoulli_sample(alpha, case["u"]))
    view.exact = False
    a = alpha.cuda().requires_grad_(True)
    view.alpha = a
    g = view.rsample(case["u"].cuda())
    a_ref = alpha.clone().requires_grad_(True)
    g_ref = O.relaxed_bernoulli_rsample(a_ref, case["u"])
    assert (g.detach().cpu() - g_ref.detach()).abs().max().item() < 2e-3     # T=0.001 amplifies log rounding 1000x
    dg = torch.from_numpy(np.random.default_rng(0).standard_normal(g.shape).astype(np.float32))
    g.backward(dg.cuda())
    g_ref.backward(dg)
    assert (a.grad.cpu() - a_ref.grad).norm() / a_ref.grad.norm() < 5e-2
    # native draw: a Bernoulli(alpha) mask
    view.exact = True
    view.alpha = torch.full((400, 500), 0.3, device="cuda")
    assert abs(view.rsample().mean().item() - 0.3) < 5e-3


@pytest.mark.parametrize("key", ["ma_rel", "ma_ex", "mb_rel", "mb_ex", "sa_rel", "sa_ex"])
def test_mf_layer_matches_oracle_and_reference(lb, key):
    g = np.load(os.path.join(C.GOLDEN, "mf_layer.npz"))
    seed, b, i, o, sim = (int(v) for v in g[f"{key}_meta"])
    relaxed = key.endswith("_rel")
    case = C.mf_layer_case(seed, b, i, o, sim=bool(sim))
    cg0 = torch.from_numpy(g[f"{key}_gamma"])                 # the reference's own gamma (injected)
    # oracle in fp64 with gamma as a leaf
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in case["p"].items()}
    x64 = case["x"].double().clone().requires_grad_(True)
    cg64 = cg0.double().clone().requires_grad_(relaxed)
    nz64 = {k: v.double() for k, v in case["noise"].items()}
    act64, lp64, lq64 = O.mf_forward(x64, p64, cg64, nz64, exact=(not relaxed, False, False, False), logprob_on_ws=bool(sim))
    ((act64 * case["gout"].double()).sum() + (lq64 - lp64) / C.NUM_BATCHES).backward()

    layer = _make_layer(lb, case, i, o, bool(sim))
    layer.train()
    layer.gamma.exact = not relaxed
    x = case["x"].cuda().requires_grad_(True)
    cg = cg0.cuda().requires_grad_(relaxed)
    act = layer(x, cg, sample=True, noise=_cuda_noise(case["noise"]))
    ((act * case["gout"].cuda()).sum() + (layer.log_variational_posterior - layer.log_prior) / C.NUM_BATCHES).backward()

    assert C.rel_err(act, act64) < 1e-5
    assert C.rel_err(act, g[f"{key}_act"]) < 1e-5
    assert abs(layer.log_prior.item() - lp64.item()) / abs(lp64.item()) < 1e-5
    assert abs(layer.log_variational_posterior.item() - lq64.item()) / abs(lq64.item()) < 1e-5
    assert abs(layer.log_prior.item() - float(g[f"{key}_log_prior"])) / abs(float(g[f"{key}_log_prior"])) < 1e-5
    assert C.rel_err(x.grad, x64.grad) < 1e-5
    if relaxed:
        assert C.rel_err(cg.grad, cg64.grad) < 2e-5
    for k in MF_NAMES:
        assert C.rel_err(getattr(layer, k).grad, p64[k].grad) < 5e-5, k


def test_mf_means(lb):
    g = np.load(os.path.join(C.GOLDEN, "mf_layer.npz"))
    for key, (seed, b, i, o, sim) in {"ma": (41, 7, 37, 23, False), "mb": (42, 33, 130, 10, False)}.items():
        case = C.mf_layer_case(seed, b, i, o, sim=sim)
        layer = _make_layer(lb, case, i, o, sim)
        layer.eval()
        with torch.no_grad():
            layer.alpha = 1 / (1 + torch.exp(-layer.lambdal))
            med = layer(case["x"].cuda(), (layer.alpha > 0.5).float(), sample=False, medimean=True)
            jm = layer(case["x"].cuda(), None, sample=False, medimean=False)
        assert layer.log_prior == 0
        assert C.rel_err(med, g[f"{key}_medimean"]) < 1e-5 and C.rel_err(jm, g[f"{key}_jointmean"]) < 1e-5


def test_mf_mnist_sample_elbo_matches_reference(lb):
    g = np.load(os.path.join(C.GOLDEN, "mf_net_mnist.npz"))
    case = C.mf_net_case(seed=50, batch=100)
    # gammas: the oracle's relaxed draws (fp32, CPU) injected as constants; lambda still gets the alpha-path grads
    layers = [{k: v.clone().requires_grad_(True) for k, v in p.items()} for p in case["layers"]]
    loss_o, nll_o, lp_o, lq_o, _, gam = O.mf_net_elbo(case["x"], case["y"], layers, case["noises"], case["us"], C.NUM_BATCHES)
    net = lb.mf.BayesianNetwork().cuda()
    with torch.no_grad():
        for l, p in zip(net.layers, case["layers"]):
            for k, v in p.items():
                getattr(l, k).copy_(v)
    net.train()
    out = net.forward(case["x"].cuda(), *[t.detach().cuda() for t in gam], sample=True,
                      noises=[_cuda_noise(n) for n in case["noises"]])
    nll = torch.nn.functional.nll_loss(out, case["y"].cuda(), reduction="sum")
    lp, lq = net.log_prior(), net.log_variational_posterior()
    for name, val in (("nll", nll), ("log_prior", lp), ("log_q", lq)):
        assert abs(val.item() - float(g[name])) / abs(float(g[name])) < 1e-5, name
    # full sample_elbo through the native gamma kernel runs and yields a finite loss close to the reference's
    loss, _, _, _ = net.sample_elbo(case["x"].cuda(), case["y"].cuda(), noises=[[_cuda_noise(n) for n in case["noises"]]],
                                    us=[[u.cuda() for u in case["us"]]])
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])) < 1e-3
    for l in net.layers:
        for k in MF_NAMES:
            assert torch.isfinite(getattr(l, k).grad).all(), k


def _oracle_mc(net_params, x, seed, samples, lb, first=0, device="cpu", dtype=torch.float32):
    """Oracle predictive average fed with the kernels' exported Philox noise for sample indices first...  device / dtype:
    where and in which precision the oracle's torch expressions run (fp32 on the CPU = what the reference computes;
    fp64 = the truth both fp32 implementations are judged against).  The hard masks [u < alpha] are always formed from
    the fp32 alpha and the fp32 uniform, as the kernels (and torch.bernoulli) do."""
    L = len(net_params)
    stride = lb.mf.MCPredictor.NSTREAMS * L
    alphas = [O.alpha_of(p["lambdal"]) for p in net_params]          # fp32 on the CPU: the reference's own arithmetic
    net_params = [{k: v.to(device=device, dtype=dtype) for k, v in p.items()} for p in net_params]
    x = x.to(device=device, dtype=dtype)
    sum_logp = torch.zeros(x.shape[0], net_params[-1]["weight_mu"].shape[0], dtype=torch.float64, device=device)
    sum_prob = torch.zeros_like(sum_logp)
    for s in range(first, first + samples):
        h = x
        for i, p in enumerate(net_params):
            o, k = p["weight_mu"].shape
            base = i * lb.mf.MCPredictor.NSTREAMS + s * stride
            u = lb.philox_uniform((o, k), seed, base + 0).cpu()
            ew = lb.philox_normal((o, k), seed, base + 1).to(device=device, dtype=dtype)
            eb = lb.philox_normal((o,), seed, base + 2).to(device=device, dtype=dtype)
            g = O.exact_bernoulli_sample(alphas[i], u).to(device=device, dtype=dtype)
            h, _, _ = O.mf_forward(h, p, g, {"eps_w": ew, "eps_b": eb}, calc_log_probs=False)
            h = torch.relu(h) if i < L - 1 else torch.log_softmax(h, 1)
        sum_logp += h.double()
        pr = torch.sigmoid(h)
        sum_prob += (pr / pr.sum(1, keepdim=True)).double()
    return sum_logp, sum_prob


def test_torch_gamma_rsample_path_still_matches_reference(lb, monkeypatch):
    """lbbnn.mf.FUSED_TAU = False: the precisions through torch.distributions.Gamma(...).rsample() / _StdGammaReparam and
    autograd instead of BayesianLinear._tau_draw's explicit derivative factors (the default, which every other test here
    runs): same layer and network parity against the oracle and the reference goldens."""
    monkeypatch.setattr(lb.mf, "FUSED_TAU", False)
    test_mf_mnist_sample_elbo_matches_reference(lb)
    for key in ("ma_rel", "mb_ex", "sa_rel"):
        test_mf_layer_matches_oracle_and_reference(lb, key)


def test_fused_elbo_objective_matches_sample_elbo_formulation(lb):
    """GraphedTrainer's loss head for objective="elbo" (lbbnn_nll_kl_objective_f32 with the layers' log q at +1 / NUM_BATCHES
    and log prior at -1 / NUM_BATCHES, backward started at the logits and those terms) against
    loss = nll + (log q - log prior) / NUM_BATCHES of sample_elbo (MF:316-318) on the SAME forward graph: loss, nll and every
    parameter gradient."""
    import ctypes
    import torch.nn.functional as F
    from lbbnn import _capi as K
    torch.manual_seed(5)
    rng = np.random.default_rng(10)
    net = lb.mf.BayesianNetwork((72, 40, 24, 10)).cuda().train()
    x = C.t(rng.uniform(0, 1, size=(37, 72))).cuda()
    y = torch.from_numpy(rng.integers(0, 10, size=(37,))).long().cuda()
    logits, terms, signs = net._elbo_terms(x)
    assert len(terms) == 6 and signs == [1.0] * 3 + [-1.0] * 3
    params = list(net.parameters())
    nll = F.nll_loss(F.log_softmax(logits, dim=1), y, reduction="sum")
    loss = nll + (sum(terms[:3]) - sum(terms[3:])) / net.num_batches
    ref = torch.autograd.grad(loss, params, retain_graph=True, allow_unused=True)
    out, dlogits = torch.zeros(2, device="cuda"), torch.empty_like(logits)
    ptrs = (ctypes.c_void_p * 6)(*[k.data_ptr() for k in terms])
    scales = (ctypes.c_float * 6)(*signs)
    K.check(K.lib.lbbnn_nll_kl_objective_f32(K.ptr(logits), K.ptr(y, torch.int64), 37, 10, ptrs, scales, 6,
                                             1.0 / net.num_batches, K.ptr(out), K.ptr(dlogits), K.current_stream()))
    assert abs(out[0].item() - loss.item()) <= 2e-6 * abs(loss.item())
    assert abs(out[1].item() - nll.item()) <= 1e-6 * abs(nll.item())
    for p in params:
        p.grad = None
    torch.autograd.backward([logits] + terms, [dlogits] + [torch.full_like(k, sg / net.num_batches) for k, sg in zip(terms, signs)])
    n_checked = 0
    for (name, p), r in zip(net.named_parameters(), ref):
        if r is None:
            assert p.grad is None, name
            continue
        assert C.rel_err(p.grad, r) < 2e-6, name
        n_checked += 1
    assert n_checked >= 30
    # (a fresh network: the graph above keeps this one's AccumulateGrad nodes on the default stream, which a capture on
    # the trainer's stream must not touch)
    net2 = lb.mf.BayesianNetwork((72, 40, 24, 10)).cuda()
    tr = lb.GraphedTrainer(net2, batch_size=37, num_batches=net2.num_batches, lr=0.0, objective="elbo", in_features=72)
    assert tr._dlogits is not None
    o = tr.step(x.cpu(), y.cpu())
    assert np.isfinite(o["loss"]) and o["nll"] > 0


@pytest.mark.parametrize("use_graph,spl,gemm", [(False, 1, "simt"), (True, 1, "simt"), (False, 5, "simt"), (True, 8, "simt"),
                                                (True, 16, "auto"), (False, 5, "tc"), (True, 8, "tc")])
def test_mc_predictor_matches_oracle_and_is_split_invariant(lb, use_graph, spl, gemm):
    sizes = [(64, 48), (48, 40), (40, 10)]
    case = C.mf_net_case(seed=60, batch=50, sizes=sizes)
    rng = np.random.default_rng(1)
    for p in case["layers"]:
        p["lambdal"] = C.t(rng.normal(0.0, 2.0, size=tuple(p["lambdal"].shape)))
        p["weight_mu"] = p["weight_mu"] * 5          # decisive logits: argmax margins well above fp32 noise
    net = lb.mf.BayesianNetwork((64, 48, 40, 10)).cuda()
    with torch.no_grad():
        for l, p in zip(net.layers, case["layers"]):
            for k, v in p.items():
                getattr(l, k).copy_(v)
    S = 12
    # spl == 5 keeps the separate head GEMM + accumulation kernels covered; the others fuse the 10-class head
    mc = lb.mf.MCPredictor(net, batch=50, seed=77, use_graph=use_graph, samples_per_launch=spl, gemm=gemm,
                           fused_head=spl != 5)
    assert mc.n_tc == (3 if gemm == "tc" else 0) and mc.fused_head == (spl in (8, 16) and gemm != "tc")
    mc.run(case["x"].cuda(), S)
    res = mc.result(S)
    ref_logp, ref_prob = _oracle_mc(case["layers"], case["x"], 77, S, lb)
    assert C.rel_err(mc.sum_logp, ref_logp) < 1e-5 and C.rel_err(mc.sum_prob, ref_prob) < 1e-5
    assert torch.equal(res["pred"].cpu(), (ref_logp / S).argmax(1))          # ensemble predictions bit-exact
    # sharding: 3 "ranks" each take a slice of the sample indices; the fp64 partials sum to the same thing
    tot_l, tot_p = torch.zeros_like(mc.sum_logp), torch.zeros_like(mc.sum_prob)
    for r in range(3):
        first, cnt = lb.mf.shard_samples(S, 3, r)
        mc.run(case["x"].cuda(), cnt, first_sample=first)
        tot_l += mc.sum_logp
        tot_p += mc.sum_prob
    assert (tot_l - res["mean_logp"] * S).abs().max().item() < 1e-9
    assert torch.equal((tot_l / S).argmax(1), res["pred"])


@pytest.mark.parametrize("gemm", ["auto", "simt"])
def test_mc_predictor_at_the_real_shape_matches_oracle(lb, gemm):
    """BASELINE.json configs[3] at its real shape: the MF 784-400-600-10 net (reference init, UN-scaled weight_mu,
    inclusion probabilities spread over (0,1)), a 1000-input test batch, 64 MC weight samples, default launch
    configuration (3xTF32 tcgen05 GEMMs + fused head, two lanes) and the CUDA-core configuration -- against the oracle
    driven by the kernels' exported Philox draws, in fp64 (the truth).  Accumulators within 1e-5; ensemble argmax
    bit-exact on every row whose top-2 margin exceeds the fp32 tolerance (near-tie rows are counted and reported)."""
    case = C.mf_net_case(seed=63, batch=1000)
    rng = np.random.default_rng(3)
    for p in case["layers"]:
        p["lambdal"] = C.t(rng.normal(0.0, 2.0, size=tuple(p["lambdal"].shape)))
    net = lb.mf.BayesianNetwork().cuda()
    with torch.no_grad():
        for l, p in zip(net.layers, case["layers"]):
            for k, v in p.items():
                getattr(l, k).copy_(v)
    S = 64
    mc = lb.mf.MCPredictor(net, batch=1000, seed=4321, samples_per_launch=32, gemm=gemm)
    assert mc.n_tc == (2 if gemm == "auto" else 0) and mc.fused_head
    res = mc.predict(case["x"].cuda(), S)
    ref_logp, ref_prob = _oracle_mc(case["layers"], case["x"], 4321, S, lb, device="cuda", dtype=torch.float64)
    e_l, e_p = C.rel_err(mc.sum_logp, ref_logp), C.rel_err(mc.sum_prob, ref_prob)
    mean_ref = ref_logp / S
    top2 = mean_ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    near_tie = margin < 2e-5 * mean_ref.abs().max()
    agree = res["pred"] == mean_ref.argmax(1)
    print(f"[real-shape] MCPredictor {gemm}: sum_logp err {e_l:.2e}, sum_prob err {e_p:.2e}, near-tie rows {int(near_tie.sum())}, "
          f"smallest margin {margin.min().item():.3e}, argmax mismatches {int((~agree).sum())}", flush=True)
    assert e_l < 1e-5 and e_p < 1e-5
    assert bool((agree | near_tie).all()) and int(near_tie.sum()) <= 2
    # the reference's ensemble statistic: the first ten samples only (MF:416)
    first_logp, _ = _oracle_mc(case["layers"], case["x"], 4321, 10, lb, device="cuda", dtype=torch.float64)
    t2 = (first_logp / 10).topk(2, dim=1).values
    tie10 = (t2[:, 0] - t2[:, 1]) < 2e-5 * first_logp.abs().max() / 10
    assert bool(((res["pred_first"] == first_logp.argmax(1)) | tie10).all()) and int(tie10.sum()) <= 2
    # and the fp32 reference arithmetic on the CPU (what the reference itself computes) for the first samples
    cpu_logp, _ = _oracle_mc(case["layers"], case["x"], 4321, 4, lb)
    mc.run(case["x"].cuda(), 4)
    assert C.rel_err(mc.sum_logp, cpu_logp) < 1e-5


def test_mc_predictor_batched_equals_one_sample_kernels(lb):
    """Batching over samples changes neither the draws nor (beyond fp32 GEMM summation order) the statistics; the
    batched sampler is bit-identical to the one-sample sampler; MNIST-shape layers with ragged tiles (400, 600, 10)."""
    case = C.mf_net_case(seed=61, batch=37)
    rng = np.random.default_rng(2)
    for p in case["layers"]:
        p["lambdal"] = C.t(rng.normal(0.0, 2.0, size=tuple(p["lambdal"].shape)))
    net = lb.mf.BayesianNetwork().cuda()
    with torch.no_grad():
        for l, p in zip(net.layers, case["layers"]):
            for k, v in p.items():
                getattr(l, k).copy_(v)
    a = lb.mf.MCPredictor(net, batch=37, seed=5, samples_per_launch=1)
    b = lb.mf.MCPredictor(net, batch=37, seed=5, samples_per_launch=4, gemm="simt")
    a.run(case["x"].cuda(), 7, first_sample=3)
    b.run(case["x"].cuda(), 7, first_sample=3)
    assert C.rel_err(b.sum_logp, a.sum_logp) < 1e-6 and C.rel_err(b.sum_prob, a.sum_prob) < 1e-6
    assert torch.equal(a.result(7)["pred"], b.result(7)["pred"])
    # last sample of the last (partial, 3-sample) launch of b == the single sample a drew last: indices 3+6
    assert torch.equal(b.w[0][2], a.w[0][0]) and torch.equal(b.b[2][2], a.b[2][0])
    # tensor-core (3xTF32) GEMMs for the 400- and 600-wide layers: same draws (hi + lo == the fp32 weight), fp32 accuracy
    c = lb.mf.MCPredictor(net, batch=37, seed=5, samples_per_launch=4, lanes=1)
    assert c.n_tc == 2
    c.run(case["x"].cuda(), 7, first_sample=3)
    assert torch.equal(c.w[0] + c.w_lo[0], b.w[0]) and torch.equal(c.w[1] + c.w_lo[1], b.w[1])
    assert torch.equal(c.w[2], b.w[2]) and torch.equal(c.b[1], b.b[1])
    assert C.rel_err(c.sum_logp, a.sum_logp) < 1e-5 and C.rel_err(c.sum_prob, a.sum_prob) < 1e-5
    assert torch.equal(a.result(7)["pred"], c.result(7)["pred"])
    # two concurrent lanes (the default with tensor-core GEMMs): same samples, fp64 partials added in lane order
    d = lb.mf.MCPredictor(net, batch=37, seed=5, samples_per_launch=2)
    assert len(d.lanes) == 2
    d.run(case["x"].cuda(), 7, first_sample=3)
    assert (d.sum_logp - c.sum_logp).abs().max().item() < 1e-9 and (d.sum_prob - c.sum_prob).abs().max().item() < 1e-9
    assert int(d.lanes[0].counter) == 3 + 4 and int(d.lanes[1].counter) == 3 + 7
    # parameters changed between runs are picked up (sigma / alpha are recomputed by every run())
    with torch.no_grad():
        net.layers[0].lambdal.add_(0.5)
        net.layers[1].weight_rho.add_(0.3)
    a.run(case["x"].cuda(), 4)
    d.run(case["x"].cuda(), 4)
    assert C.rel_err(d.sum_logp, a.sum_logp) < 1e-5 and torch.equal(a.result(4)["pred"], d.result(4)["pred"])



def test_sim_study_training_trajectory_matches_oracle(lb):
    """BASELINE.json configs[0]: the 20 -> 1 simulation-study model (MFsim:250-300) trained for 5 SGD steps with the
    script's per-group learning rates (MFsim:359-374) on the synthetic stand-in for its CSVs; every step's objective
    terms and the parameter trajectory follow the oracle (relaxed gamma injected), and the inclusion
    probabilities alpha = sigmoid(lambda) the study reports agree."""
    case = C.sim_study_case(seed=70)
    B, nb = case["batch"], case["num_batches"]
    lrs = {"bias_mu": 1e-4, "bias_rho": 1e-4, "weight_mu": 1e-4, "weight_rho": 1e-4, "pa": 1e-3, "pb": 1e-3,
           "weight_a": 1e-3, "weight_b": 1e-3, "bias_a": 1e-3, "bias_b": 1e-3, "lambdal": 1e-3}
    ref = {k: v.clone().requires_grad_(True) for k, v in case["p"].items()}
    opt_ref = torch.optim.SGD([{"params": [ref[k]], "lr": lr} for k, lr in lrs.items()], lr=0.01)
    net = lb.mf.SimStudyNetwork(num_batches=nb).cuda()
    with torch.no_grad():
        for k, v in case["p"].items():
            getattr(net.l1, k).copy_(v)
    opt = torch.optim.SGD([{"params": [getattr(net.l1, k)], "lr": lr} for k, lr in lrs.items()], lr=0.01)
    net.train()
    for step in range(5):
        xb, yb = case["X"][step * B:(step + 1) * B], case["y"][step * B:(step + 1) * B]
        nz, u = case["noises"][step], case["us"][step]
        g = O.relaxed_bernoulli_rsample(O.alpha_of(ref["lambdal"].detach()), u)          # the injected gamma
        opt_ref.zero_grad()
        loss_r, nll_r, lp_r, lq_r, out_r = O.mfsim_elbo(xb, yb, ref, nz, u, nb, gamma=g)
        loss_r.backward()
        opt.zero_grad()
        net.l1.alpha = 1 / (1 + torch.exp(-net.l1.lambdal))
        out = net(xb.cuda(), g.cuda(), sample=True, noise=_cuda_noise(nz))
        nll = torch.nn.functional.binary_cross_entropy(out, yb.cuda().unsqueeze(1).float(), reduction="sum")
        loss = nll + (net.log_variational_posterior() - net.log_prior()) / nb
        loss.backward()
        assert C.rel_err(out.detach(), out_r.detach()) < 1e-5
        for a, b in ((nll, nll_r), (net.log_prior(), lp_r), (net.log_variational_posterior(), lq_r), (loss, loss_r)):
            assert abs(a.item() - b.item()) <= 2e-5 * abs(b.item()) + 1e-4, step
        for k in lrs:
            assert C.rel_err(getattr(net.l1, k).grad, ref[k].grad) < 5e-5, (step, k)
        opt_ref.step()
        opt.step()
        for k in lrs:
            assert C.rel_err(getattr(net.l1, k).data, ref[k].data) < 1e-6, (step, k)
    alpha = 1 / (1 + torch.exp(-net.l1.lambdal.detach().cpu()))
    assert C.rel_err(alpha, O.alpha_of(ref["lambdal"].detach())) < 1e-6
    assert torch.equal(alpha > 0.5, O.alpha_of(ref["lambdal"].detach()) > 0.5)      # median-probability model, bit-exact
    # the native path of the same model (sample_elbo draws its own gamma / noise) runs and yields finite statistics
    loss, lp, lq, nll, out = net.sample_elbo(case["X"][:B].cuda(), case["y"][:B].cuda())
    assert all(torch.isfinite(v).all() for v in (loss, lp, lq, nll, out)) and out.shape == (B, 1)
