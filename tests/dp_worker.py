"""Worker of tests/test_dp_gpu.py, launched as  python -m torch.distributed.run --nproc-per-node N tests/dp_worker.py.

Data-parallel training on N GPUs must follow the single-GPU step on the concatenated minibatch: rows are
sharded, the loss is a SUM over rows (LRT:223) so gradients are sum-reduced, and the replicated KL gradient is added exactly
once.  Every rank runs (a) the data-parallel trainer on its shard with its shard of the injected noise and (b) a single-rank
trainer on the whole batch, and compares Adam's first moment after one step (exp_avg = (1 - beta1) * gradient: the gradient
itself, whether or not the trainer materialises `.grad`), the updated parameters, and the KL statistics.  Exits non-zero on a
mismatch; rank 0 prints one line per case."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "bayesian-neural-nets_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases as C  # noqa: E402


def _load(net, case):
    with torch.no_grad():
        for l, p in zip(net.layers, case["layers"]):
            for k, v in p.items():
                getattr(l, k).copy_(v)


def _set_eps(tr, eps):
    bufs = tr.eps_in if hasattr(tr, "eps_in") and tr.eps_in is not None else [d["eps"] for d in tr.tc]
    for b, e in zip(bufs, eps):
        b.copy_(e)


def run_case(name, make_trainer, dims, per_rank, rank, world, pg, tol):
    import lbbnn
    B = per_rank * world
    case = C.lrt_net_case(seed=31 + len(dims), batch=B, sizes=list(zip(dims[:-1], dims[1:])))
    lo, hi = rank * per_rank, (rank + 1) * per_rank
    # (a) data parallel on this rank's rows
    net_dp = lbbnn.BayesianNetwork(dims).cuda()
    _load(net_dp, case)
    tr_dp = make_trainer(net_dp, per_rank, pg)
    _set_eps(tr_dp, [e[lo:hi] for e in case["eps"]])
    out_dp = tr_dp.step(case["x"][lo:hi], case["y"][lo:hi])
    # (b) one GPU, whole batch
    net_1 = lbbnn.BayesianNetwork(dims).cuda()
    _load(net_1, case)
    tr_1 = make_trainer(net_1, B, None)
    _set_eps(tr_1, case["eps"])
    out_1 = tr_1.step(case["x"], case["y"])
    torch.cuda.synchronize()
    # Adam's first moment after step 1 is 0.1 x the gradient the update consumed.  With the sharded NVLS update a rank only
    # keeps the moments of the weight quads it owns (a contiguous 1 / world of every weight tensor; rank 0 owns the biases)
    worst = 0.0
    owned = 0
    for li, (l_dp, l_1) in enumerate(zip(net_dp.layers, net_1.layers)):
        for k in ("weight_mu", "weight_rho", "lambdal", "bias_mu", "bias_rho"):
            off = tr_dp._offsets[(id(l_dp), k)] if hasattr(tr_dp, "_offsets") else tr_dp.param_off[(id(l_dp), k)]
            off1 = tr_1._offsets[(id(l_1), k)] if hasattr(tr_1, "_offsets") else tr_1.param_off[(id(l_1), k)]
            n = getattr(l_dp, k).numel()
            lo_e, hi_e = tr_dp.owned_range(li, k)
            owned += hi_e - lo_e
            if hi_e > lo_e:
                g_dp, g_1 = tr_dp.exp_avg[off + lo_e:off + hi_e], tr_1.exp_avg[off1 + lo_e:off1 + hi_e]
                err = ((g_dp - g_1).double().norm() / g_1.double().norm()).item()
                worst = max(worst, err)
                assert err < tol, (name, k, "gradient (exp_avg) of the DP step differs from the single-GPU step", err)
                # ... and nothing outside the owned range was touched
                rest = torch.cat([tr_dp.exp_avg[off:off + lo_e], tr_dp.exp_avg[off + hi_e:off + n]])
                assert not rest.numel() or float(rest.abs().max()) == 0.0, (name, k, "moments outside the owned shard changed")
            # the updated parameters are replicated on every rank whatever the update scheme
            p_dp, p_1 = tr_dp.flat[off:off + n], tr_1.flat[off1:off1 + n]
            perr = C.rel_err(p_dp, p_1)
            assert perr < max(1e-5, tol), (name, k, "updated parameters of the DP step differ from the single-GPU step", perr)
    assert abs(out_dp["kl"] - out_1["kl"]) <= 1e-6 * abs(out_1["kl"]), (name, out_dp["kl"], out_1["kl"])
    # the nll each rank reports is its shard's; the sum over ranks is the whole batch's
    nll = torch.tensor([out_dp["nll"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(nll, group=pg)
    assert abs(nll.item() - out_1["nll"]) <= 2e-5 * abs(out_1["nll"]), (name, nll.item(), out_1["nll"])
    # the shards of all ranks cover every parameter exactly once
    cover = torch.tensor([owned], dtype=torch.int64, device="cuda")
    dist.all_reduce(cover, group=pg)
    n_all = sum(getattr(l, k).numel() for l in net_dp.layers for k in ("weight_mu", "weight_rho", "lambdal", "bias_mu", "bias_rho"))
    sharded = getattr(tr_dp, "dp_sharded", False) or getattr(tr_dp, "_step_dp", None) is not None
    assert int(cover) == (n_all if sharded else n_all * world), (name, int(cover), n_all)
    # every rank holds the same updated parameters
    flat = tr_dp.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0, group=pg)
    assert torch.equal(flat, ref), (name, "ranks diverged")
    if rank == 0:
        print(f"dp_worker {name}: world {world} OK, worst gradient rel err {worst:.2e}, allreduce {getattr(tr_dp, 'allreduce', 'nccl')}",
              flush=True)
    tr_dp.graph = None
    tr_1.graph = None


def main():
    import lbbnn
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    pg = dist.group.WORLD
    NB = C.NUM_BATCHES
    # the persistent step kernel (batch <= 128): raw (dM, dV) all-reduced between its two launches
    run_case("lrt_step_kernel", lambda net, B, g: lbbnn.LRTTrainer(net, batch_size=B, num_batches=NB, lr=1e-3, inject_noise=True,
                                                                   process_group=g, fused=True, materialize_grads=True),
             (136, 72, 40, 10), 32, rank, world, pg, 2e-5)
    # the same kernel with the exchange INSIDE the launch: flags over NVLink, in-switch reduction of each rank's shard of the
    # raw gradients, sharded chain rule + KL + Adam, multicast store of the parameters (needs NVSwitch multicast)
    for mode in ("sharded", "sharded_p2p"):
        os.environ["LBBNN_DP_ALLREDUCE"] = mode
        run_case(f"lrt_step_kernel_in_launch[{mode}]",
                 lambda net, B, g: lbbnn.LRTTrainer(net, batch_size=B, num_batches=NB, lr=1e-3, inject_noise=True, process_group=g,
                                                    fused=True, materialize_grads=False),
                 (136, 72, 40, 10), 32, rank, world, pg, 2e-5)
    os.environ.pop("LBBNN_DP_ALLREDUCE")
    # the per-layer launch sequence: .grad all-reduced, KL pre-scaled by 1/world
    run_case("lrt_per_layer", lambda net, B, g: lbbnn.LRTTrainer(net, batch_size=B, num_batches=NB, lr=1e-3, inject_noise=True,
                                                                 process_group=g, fused=False),
             (136, 72, 40, 10), 32, rank, world, pg, 2e-5)
    # the bf16 tensor-core trainer, fused per-layer update after the all-reduce of the raw gradients, and the unfused form
    for fused_update in (True, False):
        run_case(f"lrt_tensor_core(fused_update={fused_update})",
                 lambda net, B, g: lbbnn.LRTTensorCoreTrainer(net, batch_size=B, num_batches=NB, lr=1e-3, inject_noise=True,
                                                              process_group=g, fused_update=fused_update),
                 (136, 264, 72, 10), 64, rank, world, pg, 2e-4)
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
