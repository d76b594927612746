"""GPU parity tests of the LRT path: CUDA kernels (through the C-ABI) vs the CPU oracle on the same
seeded inputs and the same injected noise, and vs the golden outputs of the reference itself.

Tolerance (BASELINE.json north_star): max|a-b|/max|b| <= 1e-5 per tensor in fp32;
masks / argmax predictions bit-exact."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases as C
import lbbnn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
NAMES = ["weight_mu", "weight_rho", "lambdal", "bias_mu", "bias_rho"]


@pytest.fixture(scope="module")
def lb():
    import lbbnn
    return lbbnn


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def _oracle_layer(case, var_mode="reference", sample=True, dtype=torch.float32):
    p = {k: v.to(dtype).clone().requires_grad_(True) for k, v in case["p"].items()}
    x = case["x"].to(dtype).clone().requires_grad_(True)
    act = O.lrt_forward(x, p, case["eps"].to(dtype), sample=sample, var_mode=var_mode)
    kl = O.lrt_kl(p)
    ((act * case["gout"].to(dtype)).sum() + kl / C.NUM_BATCHES).backward()
    return act.detach(), kl.detach(), x.grad, {k: v.grad for k, v in p.items()}


def _cuda_layer(lb, case, var_mode="reference", sample=True):
    p = {k: v.cuda().requires_grad_(True) for k, v in case["p"].items()}
    x = case["x"].cuda().requires_grad_(True)
    cfg = lb.LayerConfig(var_mode=var_mode)
    act, kl = lb.lrt_linear(x, p["weight_mu"], p["weight_rho"], p["lambdal"], p["bias_mu"], p["bias_rho"],
                            eps=case["eps"].cuda(), cfg=cfg, sample=sample, want_kl=True)
    ((act * case["gout"].cuda()).sum() + kl / C.NUM_BATCHES).backward()
    return act.detach(), kl.detach(), x.grad, {k: v.grad for k, v in p.items()}


SHAPES = [(11, 9, 37, 23, False), (12, 5, 20, 1, True), (13, 33, 130, 10, True),
          (21, 100, 784, 400, False), (22, 1000, 400, 600, True), (23, 257, 600, 10, True), (24, 1, 64, 64, False)]


@pytest.mark.parametrize("seed,b,i,o,spread", SHAPES)
@pytest.mark.parametrize("var_mode", ["reference", "exact"])
def test_layer_fwd_bwd_matches_oracle(lb, seed, b, i, o, spread, var_mode):
    case = C.lrt_layer_case(seed, b, i, o, spread_lambda=spread)
    ref = _oracle_layer(case, var_mode, dtype=torch.float64)      # fp64 truth
    ref32 = _oracle_layer(case, var_mode)
    got = _cuda_layer(lb, case, var_mode)
    assert C.rel_err(got[0], ref[0]) < TOL, "activations"
    assert abs(got[1].item() - ref[1].item()) / abs(ref[1].item()) < TOL, "kl"
    assert C.rel_err(got[2], ref[2]) < TOL, "dx"
    for k in NAMES:
        assert C.rel_err(got[3][k], ref[3][k]) < TOL, k
    # and no worse than ~the fp32 oracle itself is
    assert C.rel_err(got[0], ref32[0]) < TOL


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_layer_matches_reference_golden(lb, tag):
    g = np.load(os.path.join(C.GOLDEN, "lrt_layer.npz"))
    seed, b, i, o, spread = (int(v) for v in g[f"{tag}_meta"])
    case = C.lrt_layer_case(seed, b, i, o, spread_lambda=bool(spread))
    act, kl, dx, grads = _cuda_layer(lb, case)
    assert C.rel_err(act, g[f"{tag}_act"]) < TOL
    assert abs(kl.item() - float(g[f"{tag}_kl"])) / abs(float(g[f"{tag}_kl"])) < TOL
    assert C.rel_err(dx, g[f"{tag}_dx"]) < TOL
    for k in NAMES:
        assert C.rel_err(grads[k], g[f"{tag}_d_{k}"]) < TOL, k


def test_mean_branch_and_eval_kl(lb):
    case = C.lrt_layer_case(13, 33, 130, 10, spread_lambda=True)
    g = np.load(os.path.join(C.GOLDEN, "lrt_layer.npz"))
    layer = lb.BayesianLinear(130, 10).cuda()
    with torch.no_grad():
        for k, v in case["p"].items():
            getattr(layer, k).copy_(v)
    layer.eval()
    with torch.no_grad():
        mean = layer(case["x"].cuda(), sample=False)
        assert layer.kl == 0
        samp = layer(case["x"].cuda(), sample=True, calculate_log_probs=True, eps=case["eps"].cuda())
    assert C.rel_err(mean, g["c_act_mean"]) < TOL
    assert C.rel_err(samp, g["c_act_eval_sample"]) < TOL
    assert abs(layer.kl.item() - float(g["c_kl_eval"])) / float(g["c_kl_eval"]) < TOL


def _load_net(lb, case):
    net = lb.BayesianNetwork().cuda()
    with torch.no_grad():
        for l, p in zip(net.layers, case["layers"]):
            for k, v in p.items():
                getattr(l, k).copy_(v)
    return net


def test_mnist_net_training_objective_matches_reference(lb):
    g = np.load(os.path.join(C.GOLDEN, "lrt_net_mnist.npz"))
    case = C.lrt_net_case(seed=0, batch=100)
    net = _load_net(lb, case)
    net.train()
    logp = net(case["x"].cuda().view(100, 1, 28, 28), sample=True, eps=[e.cuda() for e in case["eps"]])
    nll = F.nll_loss(logp, case["y"].cuda(), reduction="sum")
    kl = net.kl()
    loss = nll + kl / C.NUM_BATCHES
    loss.backward()
    assert C.rel_err(logp, g["logp"]) < TOL
    for name, val in (("nll", nll), ("kl", kl), ("loss", loss)):
        assert abs(val.item() - float(g[name])) / abs(float(g[name])) < TOL, name
    for li, l in enumerate(net.layers):
        for k in NAMES:
            grad = getattr(l, k).grad
            d = C.grad_digest(grad.cpu())
            assert C.rel_err(d["sample"], g[f"l{li}_{k}_sample"]) < TOL, (li, k)
            assert abs(d["l2"] - float(g[f"l{li}_{k}_l2"])) / float(g[f"l{li}_{k}_l2"]) < TOL, (li, k)
            if f"l{li}_{k}_full" in g:
                assert C.rel_err(grad, g[f"l{li}_{k}_full"]) < TOL, (li, k)


def test_mnist_net_predictions_bit_exact(lb):
    g = np.load(os.path.join(C.GOLDEN, "lrt_net_mnist.npz"))
    case = C.lrt_net_case(seed=0, batch=100)
    net = _load_net(lb, case)
    net.eval()
    with torch.no_grad():
        mean_logp = net(case["x"].cuda(), sample=False)
        samp = net(case["x"].cuda(), sample=True, eps=[e.cuda() for e in case["eps"]])
    assert C.rel_err(mean_logp, g["mean_logp"]) < TOL
    assert np.array_equal(mean_logp.argmax(1).cpu().numpy(), g["mean_argmax"])
    assert np.array_equal(samp.argmax(1).cpu().numpy(), g["eval_sample_argmax"])


def test_native_philox_noise_is_reproducible_and_normal(lb):
    a = lb.philox_normal((1000, 400), seed=7, stream_id=3)
    b = lb.philox_normal((1000, 400), seed=7, stream_id=3)
    c = lb.philox_normal((1000, 400), seed=7, stream_id=4)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1) < 5e-3
    assert abs((a ** 4).mean().item() - 3) < 0.1          # kurtosis of N(0,1)
    u = lb.philox_uniform((100000,), seed=1, stream_id=0)
    assert 0 < u.min().item() and u.max().item() < 1 and abs(u.mean().item() - 0.5) < 5e-3
    # odd length: the tail elements come from the same quads
    assert torch.equal(lb.philox_normal((10,), 7, 3), a.flatten()[:10])
    # the noise-descriptor form (lbbnn_philox_normal_ex): same values; with a device step counter the stream moves with it
    from lbbnn import _capi as K
    out = torch.empty(4000, device="cuda")
    K.check(K.lib.lbbnn_philox_normal_ex(K.ptr(out), out.numel(), K.make_noise(None, 7, 3), K.current_stream()))
    assert torch.equal(out, a.flatten()[:4000])
    step = torch.full((1,), 2, dtype=torch.int64, device="cuda")
    K.check(K.lib.lbbnn_philox_normal_ex(K.ptr(out), out.numel(), K.make_noise(None, 7, 1, step, 1), K.current_stream()))
    assert torch.equal(out, a.flatten()[:4000])            # stream 1 + 2 * 1 = stream 3
    inj = torch.arange(4000, dtype=torch.float32, device="cuda")
    K.check(K.lib.lbbnn_philox_normal_ex(K.ptr(out), out.numel(), K.make_noise(inj), K.current_stream()))
    assert torch.equal(out, inj)                           # injected values pass through
    # tails: Box-Muller through the MUFU log / sincos still reaches |z| > 4 at this sample size and stays finite
    big = lb.philox_normal((4000, 1000), seed=11, stream_id=0)
    assert torch.isfinite(big).all() and big.abs().max().item() > 4.0
    frac = (big.abs() > 1.959964).float().mean().item()
    assert abs(frac - 0.05) < 1e-3


@pytest.mark.parametrize("b,i,o", [(100, 784, 400), (33, 130, 10), (7, 37, 23)])
def test_native_noise_matches_oracle_on_exported_eps(lb, b, i, o):
    """The layer draws eps inside the kernel; exporting the same Philox stream and feeding it to the
    oracle must give the same activations and gradients (fwd and bwd regenerate identical noise)."""
    case = C.lrt_layer_case(31, b, i, o)
    layer = lb.BayesianLinear(i, o).cuda()
    with torch.no_grad():
        for k, v in case["p"].items():
            getattr(layer, k).copy_(v)
    layer.train()
    x = case["x"].cuda().requires_grad_(True)
    act = layer(x, sample=True)
    ((act * case["gout"].cuda()).sum() + layer.kl / C.NUM_BATCHES).backward()
    seed, stream = layer.last_noise_key
    case["eps"] = lb.philox_normal((b, o), seed, stream).cpu()
    ref = _oracle_layer(case, dtype=torch.float64)
    assert C.rel_err(act, ref[0]) < TOL
    assert C.rel_err(x.grad, ref[2]) < TOL
    for k in NAMES:
        assert C.rel_err(getattr(layer, k).grad, ref[3][k]) < TOL, k


def test_cpu_tensors_are_rejected(lb):
    layer = lb.BayesianLinear(8, 4)
    with pytest.raises(lb.LbbnnError):
        layer(torch.zeros(2, 8), sample=True)


@pytest.mark.parametrize("use_graph,fused", [(False, False), (True, False), (False, True), (True, True)])
def test_trainer_step_matches_oracle_plus_torch_adam(lb, use_graph, fused):
    """Whole captured step (fwd, loss, bwd, Adam) vs oracle autograd + torch.optim.Adam, 3 steps; both the
    per-layer launch sequence and the single persistent step kernel (csrc/lrt_step.cu)."""
    case = C.lrt_net_case(seed=5, batch=100)
    net = _load_net(lb, case)
    tr = lb.LRTTrainer(net, batch_size=100, num_batches=C.NUM_BATCHES, lr=1e-3, use_graph=use_graph, inject_noise=True,
                       fused=fused)
    assert tr.fused == fused
    # the oracle runs in float64 (the fp32 CPU oracle carries its own ~1e-5 rounding through three layers; against fp64 the
    # fp32 kernels meet the stated 1e-5 on the loss, the KL and every gradient)
    layers = [{k: v.double().clone().requires_grad_(True) for k, v in p.items()} for p in case["layers"]]
    opt = torch.optim.Adam([v for p in layers for v in p.values()], lr=1e-3)
    rng = np.random.default_rng(99)
    for step in range(3):
        eps = [C.t(rng.standard_normal(size=tuple(e.shape))) for e in case["eps"]]
        for buf, e in zip(tr.eps_in, eps):
            buf.copy_(e)
        out = tr.step(case["x"], case["y"])
        opt.zero_grad()
        loss, nll, kl, _ = O.lrt_net_loss(case["x"].double(), case["y"], layers, [e.double() for e in eps], C.NUM_BATCHES)
        loss.backward()
        assert abs(out["nll"] - nll.item()) / abs(nll.item()) < TOL, step
        assert abs(out["kl"] - kl.item()) / abs(kl.item()) < TOL, step
        for li, (l, p) in enumerate(zip(net.layers, layers)):
            for k in NAMES:
                assert C.rel_err(getattr(l, k).grad, p[k].grad.float()) < TOL, (step, li, k, "grad")
                # Adam's first steps are ~lr*sign(g): drive torch's Adam with the trainer's own gradient so
                # the optimizer kernel is compared exactly instead of amplifying 1e-6 gradient differences
                p[k].grad = getattr(l, k).grad.detach().cpu().double().clone()
        opt.step()
        for li, (l, p) in enumerate(zip(net.layers, layers)):
            for k in NAMES:
                assert C.rel_err(getattr(l, k).data, p[k].data.float()) < 2e-6, (step, li, k, "param")
                with torch.no_grad():          # the next step starts from the trainer's fp32 parameters on both sides
                    p[k].copy_(getattr(l, k).data.detach().cpu().double())


@pytest.mark.parametrize("fused", [False, True])
def test_trainer_native_noise_changes_every_replay(lb, fused):
    case = C.lrt_net_case(seed=6, batch=100)
    net = _load_net(lb, case)
    tr = lb.LRTTrainer(net, batch_size=100, num_batches=C.NUM_BATCHES, lr=0.0, use_graph=True, fused=fused)
    a = tr.step(case["x"], case["y"])["nll"]
    b = tr.step(case["x"], case["y"])["nll"]
    assert a != b          # lr = 0: only the Philox stream (keyed by the device step counter) changed
    assert tr.step_dev.item() == 2


@pytest.mark.parametrize("sizes,batch", [((784, 400, 600, 10), 32), ((20, 1), 128), ((37, 23, 5), 7), ((130, 64, 64, 64, 3), 100),
                                         ((64, 48, 40), 9), ((100, 72), 50)])      # heads > 32 classes: the tiled last layer
def test_fused_step_equals_per_layer_step(lb, sizes, batch):
    """The persistent step kernel and the per-layer launch sequence draw the same Philox noise and follow the same
    formulas: after 3 steps of native-noise training their parameters, Adam state and stats agree (odd shapes:
    ragged tiles, K and N not multiples of 4, a single layer, five layers)."""
    pairs = list(zip(sizes[:-1], sizes[1:]))
    case = C.lrt_net_case(seed=11, batch=batch, sizes=pairs, classes=sizes[-1])
    outs = []
    for fused in (False, True):
        torch.manual_seed(0)
        net = lb.BayesianNetwork(sizes).cuda()
        with torch.no_grad():
            for l, p in zip(net.layers, case["layers"]):
                for k, v in p.items():
                    getattr(l, k).copy_(v)
        tr = lb.LRTTrainer(net, batch_size=batch, num_batches=C.NUM_BATCHES, lr=1e-3, seed=77, fused=fused,
                           materialize_grads=fused)
        hist = [tr.step(case["x"], case["y"]) for _ in range(3)]
        outs.append((hist, tr.flat.clone(), tr.exp_avg.clone(), tr.exp_avg_sq.clone(), tr.gflat.clone()))
    (h0, p0, m0, v0, g0), (h1, p1, m1, v1, g1) = outs
    for a, b in zip(h0, h1):
        assert abs(a["nll"] - b["nll"]) <= 2e-5 * abs(a["nll"]) and abs(a["kl"] - b["kl"]) <= 1e-5 * abs(a["kl"])
    assert C.rel_err(g1, g0) < 5e-5
    # Adam's first steps move every parameter by ~lr regardless of the gradient's size, so tiny gradient
    # differences can flip a step: compare the state the gradients drive (exp_avg) and bound the parameters by lr
    assert C.rel_err(m1, m0) < 5e-5 and C.rel_err(v1, v0) < 1e-4
    assert (p1 - p0).abs().max().item() < 2.5e-3
    assert (p1 - p0).abs().mean().item() < 2e-6


def test_fused_step_without_materialized_grads_matches(lb):
    case = C.lrt_net_case(seed=12, batch=100)
    res = []
    for mat in (True, False):
        net = _load_net(lb, case)
        tr = lb.LRTTrainer(net, batch_size=100, num_batches=C.NUM_BATCHES, lr=1e-3, seed=5, fused=True, materialize_grads=mat)
        for _ in range(2):
            out = tr.step(case["x"], case["y"])
        res.append((out, tr.flat.clone()))
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])      # same kernel, same order: bit-identical


def test_step_async_returns_every_steps_stats_one_call_late(lb):
    """The pipelined host API runs the same steps as the synchronous one: identical statistics, delivered one call late."""
    case = C.lrt_net_case(seed=13, batch=100)
    rng = np.random.default_rng(3)
    xs = [C.t(rng.uniform(0, 1, size=(100, 784))) for _ in range(5)]
    seqs = []
    for mode in ("sync", "async", "async_pinned"):
        net = _load_net(lb, case)
        tr = lb.LRTTrainer(net, batch_size=100, num_batches=C.NUM_BATCHES, lr=1e-3, seed=9, materialize_grads=False)
        if mode == "sync":
            out = [tr.step(x, case["y"]) for x in xs]
        else:
            pin = (lambda t_: t_.pin_memory()) if mode == "async_pinned" else (lambda t_: t_)
            ys = pin(case["y"])
            out = [tr.step_async(pin(x), ys) for x in xs]
            assert out[0] is None
            out = out[1:] + [tr.flush()]
        seqs.append((out, tr.flat.clone()))
    for other in seqs[1:]:
        assert other[0] == seqs[0][0] and torch.equal(other[1], seqs[0][1])


def test_multi_tensor_adam_matches_torch_adam(lb):
    """MultiTensorAdam (one launch over a table of tensors, lbbnn_adam_multi_f32) against torch.optim.Adam on a ragged
    parameter list -- sizes around the 1024-element block and the float4 boundaries, a parameter without gradient."""
    torch.manual_seed(5)
    shapes = [(400, 784), (75,), (1,), (1023,), (1024,), (1025,), (3, 341), (75, 75), (2049,)]
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    unused = torch.nn.Parameter(torch.randn(17, device="cuda"))
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ours = lb.MultiTensorAdam(ps + [unused], lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    ref = torch.optim.Adam(qs, lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    u0 = unused.detach().clone()
    for step in range(5):
        ours.zero_grad()
        ref.zero_grad()
        for p, q in zip(ps, qs):
            g = torch.randn_like(p) * (10.0 ** (step - 2))
            p.grad, q.grad = g.clone(), g.clone()
        ours.step()
        ref.step()
        for p, q in zip(ps, qs):
            assert C.rel_err(p.detach(), q.detach()) < 1e-6
    assert torch.equal(unused.detach(), u0) and int(ours.t_dev) == 5


def test_multi_tensor_adam_early_group_updates_during_the_backward(lb):
    """set_late_params: every parameter outside the late group is updated from inside loss.backward() (post-accumulate hooks
    -> one launch on a private stream once the last early gradient is final), the late group by step(); the trajectory is
    torch.optim.Adam's, eagerly and as a captured CUDA graph, and a backward that skips an early parameter falls back to the
    single launch."""
    torch.manual_seed(7)
    shapes = [(300, 70), (70,), (1025,), (64, 70), (3,)]
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    x = torch.randn(5, 300, device="cuda")

    def loss_of(t, scale):                       # t[2] and t[4] are used twice / last: several accumulation patterns
        h = torch.tanh(x @ t[0] + t[1])
        return scale * ((h @ t[3].T).sum() + (t[2] ** 2).sum() + t[2].sum() * t[4].sum() + (t[4] ** 3).sum())

    ours = lb.MultiTensorAdam(ps, lr=1e-2)
    ours.set_late_params([ps[2], ps[4]])
    ref = torch.optim.Adam(qs, lr=1e-2)
    for step in range(4):
        ours.zero_grad()
        ref.zero_grad()
        loss_of(ps, 10.0 ** (step - 1)).backward()
        assert ours._early_done                                      # launched from the hooks, before step()
        loss_of(qs, 10.0 ** (step - 1)).backward()
        ours.step()
        ref.step()
        torch.cuda.synchronize()
        for p, q in zip(ps, qs):
            assert C.rel_err(p.detach(), q.detach()) < 1e-6, step
    assert int(ours.t_dev) == 4
    # captured: the hooks' events and the private stream become graph branches
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    ours.zero_grad()
    with torch.cuda.graph(g, stream=s):
        ours.zero_grad()
        loss_of(ps, 1.0).backward()
        ours.step()
    ours.finish_capture()
    for _ in range(3):
        g.replay()
        ref.zero_grad()
        loss_of(qs, 1.0).backward()
        ref.step()
    torch.cuda.synchronize()
    for p, q in zip(ps, qs):
        assert C.rel_err(p.detach(), q.detach()) < 2e-6
    assert int(ours.t_dev) == 7
    # an early parameter without gradient: no early launch, step() updates what has one
    ours.zero_grad()
    ref.zero_grad()
    ((ps[0] ** 2).sum() + ps[2].sum()).backward()
    ((qs[0] ** 2).sum() + qs[2].sum()).backward()
    assert not ours._early_done
    ours.step()
    ref.step()
    for i in (0, 2):             # (the optimizer counts updates globally, torch per parameter: compare the updated ones)
        assert C.rel_err(ps[i].detach(), qs[i].detach()) < 2e-6
    assert int(ours.t_dev) == 8


def test_multi_tensor_adam_parameter_groups_match_torch_adam(lb):
    """f1: per-group learning rates.  The MF script's 33 parameter groups (MF:520-553: 1e-4 weights / biases, 1e-3 pa / pb,
    1e-5 Gamma hyper-parameters, 0.1 lambdal) through MultiTensorAdam (per-row lr in the device table) against
    torch.optim.Adam with the same groups, five steps of random gradients; plus a zero-lr group and a default-lr group."""
    torch.manual_seed(6)
    nets = [lb.mf.BayesianNetwork((72, 40, 24, 10)).cuda() for _ in range(2)]
    nets[1].load_state_dict(nets[0].state_dict())
    groups = [lb.mf.reference_param_groups(n) for n in nets]
    assert len(groups[0]) == 33 and sorted({g["lr"] for g in groups[0]}) == [1e-5, 1e-4, 1e-3, 0.1]
    assert {id(p) for g in groups[0] for p in [g["params"]]} == {id(p) for p in nets[0].parameters()}
    ours = lb.MultiTensorAdam(groups[0], lr=1e-4)
    ref = torch.optim.Adam(groups[1], lr=1e-4)
    for step in range(5):
        for p, q in zip(nets[0].parameters(), nets[1].parameters()):
            g = torch.randn_like(p) * (10.0 ** (step - 2))
            p.grad, q.grad = g.clone(), g.clone()
        ours.step()
        ref.step()
        for (name, p), q in zip(nets[0].named_parameters(), nets[1].parameters()):
            assert C.rel_err(p.detach(), q.detach()) < 1e-6, (step, name)
    # a frozen group (lr 0), a group that inherits the optimizer's lr, base lr 0 (GraphedTrainer's noise-only test uses it)
    ps = [torch.nn.Parameter(torch.randn(n, device="cuda")) for n in (5, 1030, 64)]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    mk = lambda t: [{"params": [t[0]], "lr": 0.0}, {"params": [t[1]]}, {"params": t[2], "lr": 0.05}]  # noqa: E731
    for base in (1e-2, 0.0):
        ours, ref = lb.MultiTensorAdam(mk(ps), lr=base), torch.optim.Adam(mk(qs), lr=base)
        for step in range(3):
            for p, q in zip(ps, qs):
                g = torch.randn_like(p)
                p.grad, q.grad = g.clone(), g.clone()
            ours.step()
            ref.step()
        for p, q in zip(ps, qs):
            assert C.rel_err(p.detach(), q.detach()) < 1e-6
