"""Batched posterior-predictive loops for the LRT / MNF networks (lbbnn.EnsemblePredictor) and the MF `outofsample`
composition against tests/golden/ensemble.npz -- the reference's OWN networks run sample by sample under replayed noise with
the NumPy statistics of test_ensemble / outofsample (LRT:239-265, 305-341; MNF:287-318) -- and the MF ensemble / sparsity
helpers (f2) on the device against the NumPy restatement of the reference's counters (MF:376-433, 462-465, 612-637).
Tolerances: accumulators and probabilities 1e-5 (max-abs / max); argmax bit-exact except on rows whose top-2 margin is inside
that tolerance (counted; none expected)."""
import os

import numpy as np
import pytest
import torch

import cases as C
import lbbnn_oracle as O

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(C.GOLDEN, "ensemble.npz"))


@pytest.fixture(scope="module")
def lb():
    import lbbnn
    return lbbnn


def _same_argmax(pred, mean_ref, tol=2e-5):
    """pred == argmax(mean_ref) except where the reference's own top-2 margin is within the fp32 tolerance."""
    mean_ref = torch.as_tensor(mean_ref, dtype=torch.float64)
    top2 = mean_ref.topk(2, dim=1).values
    tie = (top2[:, 0] - top2[:, 1]) < tol * mean_ref.abs().max()
    agree = torch.as_tensor(pred).cpu() == mean_ref.argmax(1)
    assert int(tie.sum()) <= 1, f"{int(tie.sum())} near-tie rows"
    return bool((agree | tie).all())


def _load_lrt(lb, case):
    net = lb.BayesianNetwork().cuda()
    with torch.no_grad():
        for l, p in zip(net.layers, case["layers"]):
            for k, v in p.items():
                getattr(l, k).copy_(v)
    return net.eval()


@pytest.mark.parametrize("spl", [12, 5, 1])
def test_lrt_ensemble_predictor_matches_the_reference_loop(lb, spl):
    S, B = 12, 52
    case = C.ensemble_case(seed=120, batch=B, samples=S, kind="lrt")
    net = _load_lrt(lb, case)
    pred = lb.EnsemblePredictor(net, batch=B, samples_per_launch=spl, seed=11)
    eps = [e.cuda() for e in case["eps"]]
    out = pred.test_ensemble(case["x"].cuda(), S, ensemble_first=10, eps=eps)
    outputs = torch.from_numpy(G["lrt_outputs"]).double()
    assert C.rel_err(pred.sum_logp, outputs.sum(0)) < 1e-5                                # sum of the log-softmax outputs
    assert C.rel_err(out["mean_prob"], G["lrt_mean_prob"]) < 1e-5                          # `mydata_means`, LRT:249-260
    assert C.rel_err(out["entropy"], G["lrt_entropy"]) < 1e-5                              # LRT:325-330
    assert _same_argmax(out["ensemble"], outputs[0:10].mean(0)) and np.array_equal(
        outputs[0:10].mean(0).argmax(1).numpy(), G["lrt_ensemble"])                        # outputs[0:10].mean(0), LRT:262-263
    assert _same_argmax(out["posterior_mean"], G["lrt_posterior_mean_logp"])             # net(data, sample=False), LRT:264
    oos = pred.outofsample(case["x"].cuda(), S, eps=eps)
    assert C.rel_err(oos["entropy"], G["lrt_entropy"]) < 1e-5
    assert _same_argmax(oos["pred"], outputs[1:S].mean(0))                                 # outputs[1:S].mean(0), LRT:332-334
    # the eager module loop (lbbnn.predict_ensemble: one forward per sample) gives the same statistics
    calls = {"i": 0}

    def fwd(xx, sample):
        if not sample:
            return net(xx, sample=False)
        i = calls["i"]
        calls["i"] += 1
        return net(xx, sample=True, eps=[e[i] for e in eps])
    loop = lb.predict_ensemble(net, case["x"].cuda(), S, forward=fwd)
    assert C.rel_err(loop["mean_prob"], out["mean_prob"]) < 1e-6 and torch.equal(loop["ensemble"], out["ensemble"])


def test_lrt_ensemble_predictor_native_noise_is_launch_width_invariant(lb):
    """Sample s of layer l draws from its own Philox stream: the sums do not depend on how many samples share a launch
    (beyond the fp32 summation order of the GEMMs, whose contraction split follows the stacked batch: 1e-6), equal the run
    with those streams exported and injected, and layer 1's e_b / var_b are computed once per batch."""
    S, B = 9, 52
    case = C.ensemble_case(seed=122, batch=B, samples=S, kind="lrt")
    net = _load_lrt(lb, case)
    x = case["x"].cuda()
    runs = []
    for spl in (9, 4, 1):
        p = lb.EnsemblePredictor(net, batch=B, samples_per_launch=spl, seed=77)
        p.run(x, S)
        runs.append((p.sum_logp.clone(), p.sum_prob.clone(), p.first_logp.clone()))
        assert p.kernels_per_launch == 1 + 3 * 2 + 2         # expand + 2 fused layers + the two accumulations
    for r in runs[1:]:
        for a, b in zip(r, runs[0]):
            assert C.rel_err(a, b) < 1e-6
    p = lb.EnsemblePredictor(net, batch=B, samples_per_launch=9, seed=77)
    eps = [torch.stack([lb.philox_normal((B, o), 77, p._stream(li, s)) for s in range(S)]) for li, (_, o) in enumerate(p.sizes)]
    p.run(x, S, eps=eps)
    assert C.rel_err(p.sum_logp, runs[0][0]) < 1e-6
    # shards of the sample range add up (first_sample): what an MC-sample sharding over ranks relies on
    p.run(x, 4, first_sample=0)
    part = p.sum_logp.clone()
    p.run(x, 5, first_sample=4)
    assert C.rel_err(part + p.sum_logp, runs[0][0]) < 1e-6


def test_mnf_ensemble_predictor_matches_the_reference_loop(lb):
    S, B = 6, 52
    case = C.ensemble_case(seed=121, batch=B, samples=S, kind="mnf")
    net = lb.mnf.BayesianNetwork()
    for l, p in zip(net.layers, case["layers"]):
        l.load_state_dict(C.flat_named(p))
    net = net.cuda().eval()
    # the reference pushes all B rows of every draw through the flow and keeps the LAST one (MNF:187)
    z_noise = [{"eps_z": case["eps_z"][l][:, -1].cuda(), "z_masks": [m[:, -1].cuda() for m in case["z_masks"][l]]} for l in range(3)]
    eps = [e.cuda() for e in case["eps"]]
    outputs = torch.from_numpy(G["mnf_outputs"]).double()
    for spl in (6, 4):
        pred = lb.EnsemblePredictor(net, batch=B, samples_per_launch=spl, seed=5)
        assert pred.mnf
        pred.run(case["x"].cuda(), S, eps=eps, z_noise=z_noise)
        out = pred.result(S)
        assert C.rel_err(pred.sum_logp, outputs.sum(0)) < 1e-5
        assert C.rel_err(out["mean_prob"], G["mnf_mean_prob"]) < 1e-5 and C.rel_err(out["entropy"], G["mnf_entropy"]) < 1e-5
        assert _same_argmax(out["ensemble"], outputs[0:10].mean(0))
    oos = pred.outofsample(case["x"].cuda(), S, eps=eps, z_noise=z_noise)
    assert C.rel_err(oos["entropy"], G["mnf_entropy"]) < 1e-5 and _same_argmax(oos["pred"], outputs[1:S].mean(0))
    # native draws run, are finite and differ from call to call in z
    pred.run(case["x"].cuda(), 3)
    a = pred.sum_logp.clone()
    pred.run(case["x"].cuda(), 3)
    assert torch.isfinite(a).all() and not torch.equal(a, pred.sum_logp)


def _oracle_mf_outofsample(layers, x, seed, S, lb, masks):
    """MF outofsample body (MF:450-502) with the oracle: per sample a stochastic forward with the given masks (medimod) or
    the Bernoulli draw, the row-normalised expit average over ALL samples and the log-prob mean over samples 1..S-1."""
    L = len(layers)
    stride = lb.mf.MCPredictor.NSTREAMS * L
    logps, probs = [], torch.zeros(x.shape[0], layers[-1]["weight_mu"].shape[0], dtype=torch.float64)
    for s in range(S):
        h = x.double()
        for i, p in enumerate(layers):
            o, k = p["weight_mu"].shape
            base = i * lb.mf.MCPredictor.NSTREAMS + s * stride
            ew = lb.philox_normal((o, k), seed, base + 1).cpu().double()
            eb = lb.philox_normal((o,), seed, base + 2).cpu().double()
            if masks is None:
                g = O.exact_bernoulli_sample(O.alpha_of(p["lambdal"]), lb.philox_uniform((o, k), seed, base + 0).cpu()).double()
            else:
                g = masks[i].double()
            h, _, _ = O.mf_forward(h, {kk: v.double() for kk, v in p.items()}, g, {"eps_w": ew, "eps_b": eb}, calc_log_probs=False)
            h = torch.relu(h) if i < L - 1 else torch.log_softmax(h, 1)
        logps.append(h)
        pr = torch.sigmoid(h)
        probs += pr / pr.sum(1, keepdim=True)
    probs /= S
    return -(probs * torch.log(probs)).sum(1), torch.stack(logps[1:]).mean(0)


@pytest.mark.parametrize("medimod", [False, True])
def test_mf_outofsample_composition(lb, medimod):
    """lbbnn.mf_outofsample = MF:450-502 for one batch on the batched MC kernels: entropy of the averaged normalised expit and
    the prediction from outputs[1:S].mean(0); medimod fixes the masks to [alpha > 0.5] (MF:462-465) while weights are sampled."""
    sizes = [(64, 48), (48, 40), (40, 10)]
    case = C.mf_net_case(seed=64, batch=50, sizes=sizes)
    rng = np.random.default_rng(4)
    for p in case["layers"]:
        p["lambdal"] = C.t(rng.normal(0.0, 2.0, size=tuple(p["lambdal"].shape)))
        p["weight_mu"] = p["weight_mu"] * 3
    net = lb.mf.BayesianNetwork((64, 48, 40, 10)).cuda()
    with torch.no_grad():
        for l, p in zip(net.layers, case["layers"]):
            for k, v in p.items():
                getattr(l, k).copy_(v)
    lb.mf.refresh_inclusion(net)
    S = 7
    out = lb.mf_outofsample(net, case["x"].cuda(), S, medimod=medimod, seed=31, samples_per_launch=4)
    masks = [(O.alpha_of(p["lambdal"]) > 0.5).float() for p in case["layers"]] if medimod else None
    ent, rest = _oracle_mf_outofsample(case["layers"], case["x"], 31, S, lb, masks)
    assert C.rel_err(out["entropy"], ent) < 1e-5
    assert _same_argmax(out["pred"], rest)
    if medimod:      # the masks really are the median-probability model: bit-exact against alpha > 0.5
        for m, l in zip(lb.mf.median_probability_masks(net), net.layers):
            assert torch.equal(m.cpu(), (1 / (1 + torch.exp(-l.lambdal.detach().cpu())) > 0.5).float())


def test_mf_ensemble_statistics_on_the_device(lb):
    """f2: refresh_inclusion / median_probability_* / mask_statistics with the parameters on the GPU against the NumPy
    restatement of the reference's counters (MF:376-396, 427-433, 462-465, 612-637); injected masks and native draws."""
    torch.manual_seed(3)
    net = lb.mf.BayesianNetwork(sizes=(784, 400, 600, 10)).cuda()
    with torch.no_grad():
        for l in net.layers:
            l.lambdal.normal_(0, 2)
    mf = lb.mf
    mf.refresh_inclusion(net)
    tot = sum(l.lambdal.numel() for l in net.layers)
    over = 0
    for l, m in zip(net.layers, mf.median_probability_masks(net)):
        lam = l.lambdal.detach().cpu()
        alpha = 1 / (1 + torch.exp(-lam))                                                    # MF:612-616 (fp32, as the script)
        assert l.alpha.is_cuda and m.is_cuda and l.gamma.alpha is l.alpha and l.gamma.exact is True
        assert C.rel_err(l.alpha, alpha) < 1e-6
        assert torch.equal(m.cpu(), (l.alpha.cpu() > 0.5).float())                           # MF:462-465, bit-exact
        over += int((l.alpha.cpu() > 0.5).sum())
    assert abs(mf.median_probability_density(net).item() - over / tot) < 1e-12             # `os`, MF:634-637
    rng = np.random.default_rng(5)
    S = 4
    draws_np = [([(rng.random(tuple(l.alpha.shape)) < l.alpha.cpu().numpy()).astype(np.float32) for l in net.layers],
                 [(rng.random(tuple(l.alpha.shape)) < l.alpha.cpu().numpy()).astype(np.float32) for l in net.layers])
                for _ in range(S)]
    draws = [([torch.from_numpy(g).cuda() for g in ga], [torch.from_numpy(g).cuda() for g in gb]) for ga, gb in draws_np]
    got = mf.mask_statistics(net, S, draws=draws)
    spars, density = 0.0, []
    gt = [np.zeros(tuple(l.alpha.shape)) for l in net.layers]
    for ga, gb in draws_np:                                                                   # MF:376-396
        spars += sum(int((g > 0.5).sum()) for g in ga) / tot
        gt = [t + (g > 0.5) for t, g in zip(gt, ga)]
        density.append(np.concatenate([g.ravel() for g in gb]).mean())
    assert got["sparsity"].is_cuda
    assert abs(got["sparsity"].item() - spars / S) < 1e-12
    assert abs(got["ever_active"].item() - sum(int((t > 0).sum()) for t in gt) / tot) < 1e-12
    assert abs(got["density"].item() - float(np.mean(density))) < 1e-6
    nat = mf.mask_statistics(net, 20)
    mean_alpha = sum(l.alpha.sum().item() for l in net.layers) / tot
    assert abs(nat["sparsity"].item() - mean_alpha) < 5e-3 and abs(nat["density"].item() - mean_alpha) < 5e-3
