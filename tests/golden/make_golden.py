"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CLASSES on the seeded cases of
tests/cases.py.  Needs /root/reference (build container only); the resulting fixtures travel.

    python tests/golden/make_golden.py [lrt|mnf|mf|all]

Every stored value is an output of reference code (AST-sliced, see ref_harness.py) under replayed
noise; inputs are regenerated from the seeds by tests/cases.py, so the big MNIST-shape cases keep
only outputs plus gradient digests.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import cases as C  # noqa: E402
import ref_harness as H  # noqa: E402


def _load_params(layer, p):
    with torch.no_grad():
        for k, v in p.items():
            getattr(layer, k).copy_(v)


def _grads(layer, names):
    return {k: getattr(layer, k).grad.detach().clone() for k in names}


LRT_NAMES = ["weight_mu", "weight_rho", "lambdal", "bias_mu", "bias_rho"]


def golden_lrt():
    ns = H.load_reference_classes("LBBNN-GP-MF-LRT.py")
    Layer, Net = ns["BayesianLinear"], ns["BayesianNetwork"]

    # ---- single layer, odd sizes, full tensors ----------------------------------------------
    out = {}
    for tag, (seed, b, i, o, spread) in {
        "a": (11, 9, 37, 23, False),
        "b": (12, 5, 20, 1, True),      # sim-study shape (20 -> 1)
        "c": (13, 33, 130, 10, True),
    }.items():
        case = C.lrt_layer_case(seed, b, i, o, spread_lambda=spread)
        layer = Layer(i, o)
        _load_params(layer, case["p"])
        x = case["x"].clone().requires_grad_(True)
        layer.train()
        with H.replay(H.NoiseQueue([("normal", case["eps"])])):
            act = layer(x, sample=True)
        kl = layer.kl
        loss = (act * case["gout"]).sum() + kl / C.NUM_BATCHES
        loss.backward()
        out[f"{tag}_meta"] = np.array([seed, b, i, o, int(spread)])
        out[f"{tag}_act"] = act.detach().numpy()
        out[f"{tag}_kl"] = np.float64(kl.item())
        out[f"{tag}_dx"] = x.grad.numpy()
        for k, g in _grads(layer, LRT_NAMES).items():
            out[f"{tag}_d_{k}"] = g.numpy()
        layer.eval()
        with torch.no_grad():
            mean_act = layer(case["x"], sample=False)
            assert layer.kl == 0
            with H.replay(H.NoiseQueue([("normal", case["eps"])])):
                act_eval = layer(case["x"], sample=True, calculate_log_probs=True)
        out[f"{tag}_act_mean"] = mean_act.numpy()
        out[f"{tag}_act_eval_sample"] = act_eval.numpy()
        out[f"{tag}_kl_eval"] = np.float64(layer.kl.item())
    np.savez_compressed(os.path.join(HERE, "lrt_layer.npz"), **out)

    # ---- MNIST-shape network, one training objective + grads -----------------------------------
    out = {}
    case = C.lrt_net_case(seed=0, batch=100)
    net = Net()
    for lay, p in zip((net.l1, net.l2, net.l3), case["layers"]):
        _load_params(lay, p)
    net.train()
    with H.replay(H.NoiseQueue([("normal", e) for e in case["eps"]])):
        logp = net(case["x"].view(100, 1, 28, 28), sample=True)
    nll = torch.nn.functional.nll_loss(logp, case["y"], reduction="sum")
    kl = net.kl()
    loss = nll + kl / C.NUM_BATCHES
    loss.backward()
    out["logp"] = logp.detach().numpy()
    out["nll"] = np.float64(nll.item())
    out["kl"] = np.float64(kl.item())
    out["loss"] = np.float64(loss.item())
    for li, lay in enumerate((net.l1, net.l2, net.l3)):
        for k, g in _grads(lay, LRT_NAMES).items():
            d = C.grad_digest(g)
            for dk, dv in d.items():
                out[f"l{li}_{k}_{dk}"] = dv
            if g.numel() <= 6000:
                out[f"l{li}_{k}_full"] = g.numpy()
    net.eval()
    with torch.no_grad():
        mean_logp = net(case["x"], sample=False)
        with H.replay(H.NoiseQueue([("normal", e) for e in case["eps"]])):
            samp_logp = net(case["x"], sample=True)
    out["mean_logp"] = mean_logp.numpy()
    out["mean_argmax"] = mean_logp.argmax(1).numpy()
    out["eval_sample_logp"] = samp_logp.numpy()
    out["eval_sample_argmax"] = samp_logp.argmax(1).numpy()
    np.savez_compressed(os.path.join(HERE, "lrt_net_mnist.npz"), **out)
    print("lrt golden written; loss", loss.item(), "kl", kl.item())


MF_NAMES = ["weight_mu", "weight_rho", "lambdal", "bias_mu", "bias_rho", "weight_a", "weight_b", "bias_a", "bias_b",
            "pa", "pb"]


def _mf_queue(nz, u=None):
    """Draw order inside MF BayesianLinear.forward (MF:232-249): eps_w, eps_b, then the two Gamma.rsample."""
    q = []
    if u is not None:
        q.append(("uniform", u))
    q += [("normal", nz["eps_w"]), ("normal", nz["eps_b"]), ("gamma", nz["g0_w"]), ("gamma", nz["g0_b"])]
    return q


def golden_mf():
    out = {}
    for script, tagp, sim in (("LBBNN-GP-MF.py", "m", False), ("LBBNN-GP-MFsim_study.py", "s", True)):
        ns = H.load_reference_classes(script)
        Layer = ns["BayesianLinear"]
        shapes = {"a": (41, 7, 37, 23), "b": (42, 33, 130, 10)} if not sim else {"a": (43, 9, 20, 1)}
        for tag, (seed, b, i, o) in shapes.items():
            key = tagp + tag
            case = C.mf_layer_case(seed, b, i, o, sim=sim)
            for relaxed in (True, False):
                layer = Layer(i, o) if sim else Layer(i, o, 1)
                _load_params(layer, case["p"])
                layer.train()
                x = case["x"].clone().requires_grad_(True)
                layer.alpha = 1 / (1 + torch.exp(-layer.lambdal))
                layer.gamma.alpha = layer.alpha
                layer.gamma.exact = not relaxed
                with H.replay(H.NoiseQueue([("uniform", case["u"])])):
                    cg = layer.gamma.rsample()
                with H.replay(H.NoiseQueue(_mf_queue(case["noise"]))):
                    act = layer(x, cg, sample=True)
                loss = (act * case["gout"]).sum() + (layer.log_variational_posterior - layer.log_prior) / C.NUM_BATCHES
                loss.backward()
                k2 = key + ("_rel" if relaxed else "_ex")
                out[k2 + "_meta"] = np.array([seed, b, i, o, int(sim)])
                out[k2 + "_gamma"] = cg.detach().numpy()
                out[k2 + "_act"] = act.detach().numpy()
                out[k2 + "_log_prior"] = np.float64(layer.log_prior.item())
                out[k2 + "_log_q"] = np.float64(layer.log_variational_posterior.item())
                out[k2 + "_dx"] = x.grad.numpy()
                for name, g in _grads(layer, MF_NAMES).items():
                    out[k2 + "_d_" + name] = g.numpy()
            # eval-mode means (medimean / joint mean with the stale alpha)
            layer.eval()
            with torch.no_grad():
                layer.alpha = 1 / (1 + torch.exp(-layer.lambdal))
                med = layer(case["x"], (layer.alpha > 0.5).float(), sample=False, medimean=True)
                jm = layer(case["x"], cg, sample=False, medimean=False)
            out[key + "_medimean"] = med.numpy()
            out[key + "_jointmean"] = jm.numpy()
    np.savez_compressed(os.path.join(HERE, "mf_layer.npz"), **out)

    # MNIST-shape sample_elbo (SAMPLES=1) + grads
    ns = H.load_reference_classes("LBBNN-GP-MF.py")
    Net = ns["BayesianNetwork"]
    case = C.mf_net_case(seed=50, batch=100)
    net = Net()
    for lay, p in zip((net.l1, net.l2, net.l3), case["layers"]):
        _load_params(lay, p)
    net.train()
    q = [("uniform", u) for u in case["us"]]
    for nz in case["noises"]:
        q += _mf_queue(nz)
    with H.replay(H.NoiseQueue(q)):
        loss, log_prior, log_q, nll = net.sample_elbo(case["x"].view(100, 1, 28, 28), case["y"])
    loss.backward()
    out = {"loss": np.float64(loss.item()), "log_prior": np.float64(log_prior.item()), "log_q": np.float64(log_q.item()),
           "nll": np.float64(nll.item())}
    for li, lay in enumerate((net.l1, net.l2, net.l3)):
        for k, g in _grads(lay, MF_NAMES).items():
            for dk, dv in C.grad_digest(g).items():
                out[f"l{li}_{k}_{dk}"] = dv
            if g.numel() <= 6000:
                out[f"l{li}_{k}_full"] = g.numpy()
    np.savez_compressed(os.path.join(HERE, "mf_net_mnist.npz"), **out)
    print("mf golden written; loss", loss.item(), "log_prior", log_prior.item(), "log_q", log_q.item())


def golden_mfsim():
    """BASELINE.json configs[0]: the simulation-study network (LBBNN-GP-MFsim_study.py:250-300), one sample_elbo +
    backward on the first minibatch of tests/cases.py:sim_study_case, driven by the replayed noise."""
    case = C.sim_study_case(seed=70)
    B, nb = case["batch"], case["num_batches"]
    ns = H.load_reference_classes("LBBNN-GP-MFsim_study.py", BATCH_SIZE=B, NUM_BATCHES=nb, SAMPLES=1)
    net = ns["BayesianNetwork"]()
    _load_params(net.l1, case["p"])
    net.train()
    with H.replay(H.NoiseQueue([("uniform", case["us"][0])] + _mf_queue(case["noises"][0]))):
        loss, lp, lq, nll, outp = net.sample_elbo(case["X"][:B], case["y"][:B])
    loss.backward()
    out = {"loss": np.float64(loss.item()), "log_prior": np.float64(lp.item()), "log_q": np.float64(lq.item()),
           "nll": np.float64(nll.item()), "out": outp.detach().numpy(), "gamma": net.l1.gammas.detach().numpy()}
    for k, g in _grads(net.l1, MF_NAMES).items():
        out["d_" + k] = g.numpy()
    np.savez_compressed(os.path.join(HERE, "mfsim_net.npz"), **out)
    print("mfsim golden written; loss", loss.item(), "nll", nll.item())


def _mnf_queue(nz):
    """Draw order of MNF BayesianLinear.forward in training (MNF:182-235)."""
    q = [("normal", nz["eps_z"])] + [("uniform", 1.0 - m * 0.75) for m in nz["z_masks"]]      # u < .5 <=> mask = 1
    q += [("normal", nz["eps"]), ("normal", nz["eps_z2"])] + [("uniform", 1.0 - m * 0.75) for m in nz["z_masks2"]]
    q += [("normal", nz["eps_r"])] + [("uniform", 1.0 - m * 0.75) for m in nz["r_masks"]]
    return q


def _load_named(module, named):
    sd = dict(module.named_parameters())
    with torch.no_grad():
        for k, v in named.items():
            sd[k].copy_(v)


def golden_mnf():
    out = {}
    for script, flows, tagp, hs in (("LBBNN-GP-MF-MNF.py", "flows2", "m", (75, 75, 75, 75)),
                                     ("LBBNN-GP-MF-MNFsim_study.py", "flows_simstudy", "s", (50, 50, 50, 50, 50))):
        ns = H.load_reference_classes(script, flows_module=flows)
        Layer = ns["BayesianLinear"]
        shapes = {"a": (71, 6, 37, 23), "b": (72, 17, 130, 10)} if tagp == "m" else {"a": (73, 9, 20, 1)}
        for tag, (seed, b, i, o) in shapes.items():
            key = tagp + tag
            case = C.mnf_layer_case(seed, b, i, o, h_sizes=hs)
            layer = Layer(i, o, 2)
            _load_named(layer, C.flat_named(case["p"]))
            layer.train()
            x = case["x"].clone().requires_grad_(True)
            with H.replay(H.NoiseQueue(_mnf_queue(case["noise"]))):
                act = layer(x, sample=True)
            kl = layer.kl
            ((act * case["gout"]).sum() + kl / C.NUM_BATCHES).backward()
            out[key + "_meta"] = np.array([seed, b, i, o, len(hs), hs[0]])
            out[key + "_act"] = act.detach().numpy()
            out[key + "_kl"] = np.float64(kl.item())
            out[key + "_z"] = layer.z.detach().numpy()
            out[key + "_dx"] = x.grad.numpy()
            for name, prm in layer.named_parameters():
                out[key + "_d_" + name] = prm.grad.numpy() if prm.numel() <= 4000 else C.grad_digest(prm.grad)["sample"]
    np.savez_compressed(os.path.join(HERE, "mnf_layer.npz"), **out)

    # IAF-style 'MNF' flow (flows2:225-241) through PropagateFlow, batched and 1-D
    import flows2
    rng = np.random.default_rng(80)
    out = {}
    for kind in ("RNVP", "MNF"):
        tps = C.O.init_flow_params(rng, 24, 2, (75, 75, 75, 75), kind)
        flow = flows2.PropagateFlow(kind, 24, 2)
        named = C.flat_named({"z_flow": tps})
        _load_named(flow, {k.replace("z_flow.", ""): v for k, v in named.items()})
        for tag, shape in (("b", (5, 24)), ("v", (24,))):
            z = C.t(rng.standard_normal(size=shape)).requires_grad_(True)
            masks = C._masks(rng, 2, *shape)
            with H.replay(H.NoiseQueue([("uniform", 1.0 - m * 0.75) for m in masks])):
                zo, ld = flow(z)
            ((zo * zo).sum() + ld.sum()).backward()
            out[f"{kind}_{tag}_z"] = z.detach().numpy()
            out[f"{kind}_{tag}_masks"] = np.stack([m.numpy() for m in masks])
            out[f"{kind}_{tag}_out"] = zo.detach().numpy()
            out[f"{kind}_{tag}_logdet"] = ld.detach().numpy()
            out[f"{kind}_{tag}_dz"] = z.grad.numpy()
        for k, v in named.items():
            out[f"{kind}_p_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "flows.npz"), **out)

    # MNIST-shape MNF network objective
    ns = H.load_reference_classes("LBBNN-GP-MF-MNF.py", flows_module="flows2")
    case = C.mnf_net_case(seed=90, batch=100)
    net = ns["BayesianNetwork"]()
    for lay, p in zip((net.l1, net.l2, net.l3), case["layers"]):
        _load_named(lay, C.flat_named(p))
    net.train()
    q = []
    for nz in case["noises"]:
        q += _mnf_queue(nz)
    with H.replay(H.NoiseQueue(q)):
        logp = net(case["x"].view(100, 1, 28, 28), sample=True)
    nll = torch.nn.functional.nll_loss(logp, case["y"], reduction="sum")
    kl = net.kl()
    (nll + kl / C.NUM_BATCHES).backward()
    out = {"logp": logp.detach().numpy(), "nll": np.float64(nll.item()), "kl": np.float64(kl.item())}
    for li, lay in enumerate((net.l1, net.l2, net.l3)):
        for name, prm in lay.named_parameters():
            out[f"l{li}_{name}"] = C.grad_digest(prm.grad)["sample"] if prm.numel() > 2000 else prm.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "mnf_net_mnist.npz"), **out)
    print("mnf golden written; nll", nll.item(), "kl", kl.item())


def golden_vd():
    """variational_dropout.py: BayesianLayer (VD:55-68) on odd shapes and the 784-1200-1200-1200-10
    BNN (VD:71-85) with loss_fn (VD:88-106), all under replayed zeta.  The reference's alpha is `nn.Parameter(zeros)+0.2`,
    a non-leaf tensor: it is overwritten with the case's alpha (a leaf) so that its gradient can be stored too."""
    class _DS:
        def __init__(self, n):
            self.dataset = range(n)
    ns = H.load_reference_classes("variational_dropout.py", functions=("loss_fn",), device="cpu",
                                  config={"batch_size": 100}, train_loader=_DS(60000), val_loader=_DS(10000))
    Layer, BNN, loss_fn = ns["BayesianLayer"], ns["BNN"], ns["loss_fn"]
    out = {}
    for tag, (seed, b, n, m, spread) in {"a": (81, 9, 37, 23, False), "b": (82, 33, 130, 10, True), "c": (83, 5, 64, 1, True)}.items():
        case = C.vd_layer_case(seed, b, n, m, spread_alpha=spread)
        layer = Layer(n, m)
        with torch.no_grad():
            layer.theta.copy_(case["p"]["theta"])
        layer.alpha = case["p"]["alpha"].clone().requires_grad_(True)
        x = case["x"].clone().requires_grad_(True)
        with H.replay(H.NoiseQueue([("normal", case["zeta"])])):
            act = layer(x)
        (act * case["gout"]).sum().backward()
        out[f"{tag}_meta"] = np.array([seed, b, n, m, int(spread)])
        out[f"{tag}_act"] = act.detach().numpy()
        out[f"{tag}_dx"] = x.grad.numpy()
        out[f"{tag}_d_theta"] = layer.theta.grad.numpy()
        out[f"{tag}_d_alpha"] = layer.alpha.grad.numpy()
    case = C.vd_net_case(seed=80, batch=100)
    net = BNN()
    for lay, p in zip((net.l1, net.l2, net.l3, net.l4), case["layers"]):
        with torch.no_grad():
            lay.theta.copy_(p["theta"])
        lay.alpha = p["alpha"].clone().requires_grad_(True)
    net.train()
    with H.replay(H.NoiseQueue([("normal", z) for z in case["zetas"]])):
        pred = net(case["x"])
    loss = loss_fn(pred, case["y"], net)
    loss.backward()
    out["net_logp"] = pred.detach().numpy()
    out["net_loss"] = np.float64(loss.item())
    for li, lay in enumerate((net.l1, net.l2, net.l3, net.l4)):
        for k, v in C.grad_digest(lay.theta.grad).items():
            out[f"net_l{li}_theta_{k}"] = v
        out[f"net_l{li}_d_alpha"] = lay.alpha.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "vd.npz"), **out)


def golden_vd_train():
    """Two iterations of the training branch of run_epoch (VD:160-166) on the reference BNN with the script's optimizer,
    torch.optim.AdamW(model.parameters(), lr=1e-4) (VD:110): pins what the optimizer sees -- `alpha` is NOT among
    model.parameters() (VD:61 builds it as a non-leaf tensor), so only theta moves -- and the AdamW defaults."""
    class _DS:
        def __init__(self, n):
            self.dataset = range(n)
    ns = H.load_reference_classes("variational_dropout.py", functions=("loss_fn",), device="cpu",
                                  config={"batch_size": 100}, train_loader=_DS(60000), val_loader=_DS(10000))
    BNN, loss_fn = ns["BNN"], ns["loss_fn"]
    case = C.vd_net_case(seed=87, batch=100)
    net = BNN()
    for lay, p in zip((net.l1, net.l2, net.l3, net.l4), case["layers"]):
        with torch.no_grad():
            lay.theta.copy_(p["theta"])
    assert [n for n, _ in net.named_parameters()] == ["l1.theta", "l2.theta", "l3.theta", "l4.theta"]
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
    rng = np.random.default_rng(870)
    out = {"n_steps": np.int64(2)}
    net.train()
    for step in range(2):
        zetas = [C.t(rng.standard_normal(size=tuple(z.shape))) for z in case["zetas"]]
        net.zero_grad()
        with H.replay(H.NoiseQueue([("normal", z) for z in zetas])):
            pred = net(case["x"])
        loss = loss_fn(pred, case["y"], net)
        loss.backward()
        opt.step()
        out[f"loss_{step}"] = np.float64(loss.item())
    for li, lay in enumerate((net.l1, net.l2, net.l3, net.l4)):
        for k, v in C.grad_digest(lay.theta.detach()).items():
            out[f"theta_l{li}_{k}"] = v
        out[f"alpha_l{li}"] = lay.alpha.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "vd_train.npz"), **out)
    print("vd_train golden written; losses", out["loss_0"], out["loss_1"])


def golden_mnf_iaf():
    """MNF layer whose z / r flows are the IAF-style 'MNF' transform (flows2:225-241; Z_FLOW_TYPE = R_FLOW_TYPE = 'MNF'):
    the KL branch's log_det_q is the KL row's own log-determinant (sample_z() with batch 1, MNF:210)."""
    ns = H.load_reference_classes("LBBNN-GP-MF-MNF.py", flows_module="flows2", Z_FLOW_TYPE="MNF", R_FLOW_TYPE="MNF")
    Layer = ns["BayesianLinear"]
    out = {}
    for tag, (seed, b, i, o) in {"a": (171, 6, 37, 23), "b": (172, 5, 64, 10)}.items():
        case = C.mnf_layer_case(seed, b, i, o, kind="MNF")
        layer = Layer(i, o, 2)
        _load_named(layer, C.flat_named(case["p"]))
        layer.train()
        x = case["x"].clone().requires_grad_(True)
        with H.replay(H.NoiseQueue(_mnf_queue(case["noise"]))):
            act = layer(x, sample=True)
        kl = layer.kl
        ((act * case["gout"]).sum() + kl / C.NUM_BATCHES).backward()
        out[tag + "_meta"] = np.array([seed, b, i, o])
        out[tag + "_act"] = act.detach().numpy()
        out[tag + "_kl"] = np.float64(kl.item())
        out[tag + "_z"] = layer.z.detach().numpy()
        out[tag + "_dx"] = x.grad.numpy()
        for name, prm in layer.named_parameters():
            out[tag + "_d_" + name] = prm.grad.numpy() if prm.numel() <= 4000 else C.grad_digest(prm.grad)["sample"]
    np.savez_compressed(os.path.join(HERE, "mnf_layer_iaf.npz"), **out)
    print("mnf iaf golden written; kl", out["a_kl"], out["b_kl"])


def tensor_digest(v):
    """Position-sensitive fingerprint of a parameter tensor: shape, fp64 sum and abs-sum, head and strided sample."""
    f = v.detach().reshape(-1)
    return {"shape": np.array(v.shape, dtype=np.int64), "sum": np.float64(f.double().sum().item()),
            "abs": np.float64(f.double().abs().sum().item()), "head": f[:8].numpy().copy(),
            "sample": f[::max(1, f.numel() // 64)].numpy().copy()}


def golden_ctor():
    """state_dict keys (in order) and parameter fingerprints of the reference's own constructors under torch.manual_seed:
    the drop-in constructors must consume torch's RNG in the same order (a2, a8)."""
    out = {}

    def put(tag, module, extra=()):
        sd = module.state_dict()
        out[tag + "_keys"] = np.array(list(sd.keys()))
        for k, v in sd.items():
            for n, d in tensor_digest(v).items():
                out[f"{tag}|{k}|{n}"] = d
        for name in extra:     # non-parameter tensors drawn in the ctor (MF: gammas, alpha)
            for li, lay in enumerate(extra[name]):
                for n, d in tensor_digest(getattr(lay, name)).items():
                    out[f"{tag}|extra.l{li + 1}.{name}|{n}"] = d

    for seed in (0, 7):
        ns = H.load_reference_classes("LBBNN-GP-MF-LRT.py")
        torch.manual_seed(seed)
        put(f"lrt{seed}", ns["BayesianNetwork"]())
        ns = H.load_reference_classes("LBBNN-GP-MF-MNF.py", flows_module="flows2")
        torch.manual_seed(seed)
        put(f"mnf{seed}", ns["BayesianNetwork"]())
        ns = H.load_reference_classes("LBBNN-GP-MF.py")
        torch.manual_seed(seed)
        net = ns["BayesianNetwork"]()
        put(f"mf{seed}", net, {"gammas": (net.l1, net.l2, net.l3), "alpha": (net.l1, net.l2, net.l3)})
    ns = H.load_reference_classes("LBBNN-GP-MFsim_study.py")
    torch.manual_seed(3)
    net = ns["BayesianNetwork"]()
    put("mfsim3", net, {"gammas": (net.l1,), "alpha": (net.l1,)})
    ns = H.load_reference_classes("LBBNN-GP-MF-MNFsim_study.py", flows_module="flows_simstudy", p=21)
    torch.manual_seed(3)
    put("mnfsim3", ns["BayesianNetwork"]())
    ns = H.load_reference_classes("LBBNN-GP-MF-MNF.py", flows_module="flows2", Z_FLOW_TYPE="MNF", R_FLOW_TYPE="MNF")
    torch.manual_seed(5)
    put("mnfiaf5", ns["BayesianLinear"](40, 12, 2))
    np.savez_compressed(os.path.join(HERE, "ctor.npz"), **out)
    print("ctor golden written;", len(out), "entries")


def _reference_ensemble_statistics(outputs):
    """The NumPy side of the reference's loops on the stacked outputs (S, B, C): `mydata_means` (LRT:249-260), the ensemble
    prediction outputs[0:10].mean(0) (LRT:262-263), outofsample's entropies (LRT:325-330) and its prediction from
    outputs[1:S].mean(0) (LRT:332-334)."""
    from scipy.special import expit
    S, B, _ = outputs.shape
    means = None
    for i in range(S):
        tmp = expit(outputs[i].detach().cpu().numpy())
        for j in range(B):
            tmp[j] /= np.sum(tmp[j])
        means = tmp if means is None else means + tmp
    means /= S
    ent = np.array([-np.sum(means[j] * np.log(means[j])) for j in range(B)])
    return {"mean_prob": means, "ensemble": outputs[0:10].mean(0).max(1)[1].numpy(), "entropy": ent,
            "oos_pred": outputs[1:S].mean(0).max(1)[1].numpy()}


def golden_ensemble():
    """test_ensemble / outofsample bodies run with the reference's own LRT and MNF networks under replayed noise."""
    out = {}
    # ---- LRT ----
    S, B = 12, 52
    case = C.ensemble_case(seed=120, batch=B, samples=S, kind="lrt")
    ns = H.load_reference_classes("LBBNN-GP-MF-LRT.py")
    net = ns["BayesianNetwork"]()
    for lay, p in zip((net.l1, net.l2, net.l3), case["layers"]):
        _load_params(lay, p)
    net.eval()
    outputs = torch.zeros(S, B, 10)
    with torch.no_grad():
        for s in range(S):
            with H.replay(H.NoiseQueue([("normal", case["eps"][l][s]) for l in range(3)])):
                outputs[s] = net(case["x"].view(B, 1, 28, 28), sample=True)
        mean_out = net(case["x"].view(B, 1, 28, 28), sample=False)
    st = _reference_ensemble_statistics(outputs)
    out["lrt_outputs"] = outputs.numpy()
    out["lrt_posterior_mean_logp"] = mean_out.numpy()
    for k, v in st.items():
        out["lrt_" + k] = v
    # ---- MNF ----
    S, B = 6, 52
    case = C.ensemble_case(seed=121, batch=B, samples=S, kind="mnf")
    ns = H.load_reference_classes("LBBNN-GP-MF-MNF.py", flows_module="flows2")
    net = ns["BayesianNetwork"]()
    for lay, p in zip((net.l1, net.l2, net.l3), case["layers"]):
        _load_named(lay, C.flat_named(p))
    net.eval()
    outputs = torch.zeros(S, B, 10)
    with torch.no_grad():
        for s in range(S):
            q = []
            for l in range(3):      # eval + sample=True: sample_z(B) then eps (MNF:193-200); no KL branch
                q += [("normal", case["eps_z"][l][s])] + [("uniform", 1.0 - m[s] * 0.75) for m in case["z_masks"][l]]
                q += [("normal", case["eps"][l][s])]
            with H.replay(H.NoiseQueue(q)):
                outputs[s] = net(case["x"].view(B, 1, 28, 28), sample=True)
    st = _reference_ensemble_statistics(outputs)
    out["mnf_outputs"] = outputs.numpy()
    for k, v in st.items():
        out["mnf_" + k] = v
    np.savez_compressed(os.path.join(HERE, "ensemble.npz"), **out)
    print("ensemble golden written; lrt ensemble acc-agnostic preds", out["lrt_ensemble"][:8], "mnf", out["mnf_ensemble"][:8])


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("vd_train", "all"):
        golden_vd_train()
    if what in ("vd", "all"):
        golden_vd()
    if what in ("lrt", "all"):
        golden_lrt()
    if what in ("mnf", "all") and "golden_mnf" in globals():
        globals()["golden_mnf"]()
    if what in ("mf", "all") and "golden_mf" in globals():
        globals()["golden_mf"]()
    if what in ("mfsim", "all"):
        golden_mfsim()
    if what in ("mnf_iaf", "all"):
        golden_mnf_iaf()
    if what in ("ctor", "all"):
        golden_ctor()
    if what in ("ensemble", "all"):
        golden_ensemble()
