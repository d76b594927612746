"""Host-side multi-rank logic on CPU (gloo, world_size 2): MC-sample sharding covers [0, S) exactly once and
the reduction of per-rank fp64 partial accumulators equals the single-rank sum; the DP rule for the replicated
KL gradient (scaled by 1/world on every rank, summed by the all-reduce) reproduces the single-rank gradient."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_samples(total, world, rank):      # same rule as lbbnn.mf.shard_samples (importable without CUDA)
    sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
    from lbbnn.mf import shard_samples
    return shard_samples(total, world, rank)


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = _shard_samples(total, world, rank)
    g = torch.Generator().manual_seed(123)
    per_sample = torch.randn(total, 7, 3, generator=g, dtype=torch.float64)      # what each MC sample contributes
    partial = per_sample[first:first + count].sum(0)
    covered = torch.zeros(total, dtype=torch.int64)
    covered[first:first + count] = 1
    dist.all_reduce(partial)
    dist.all_reduce(covered)
    # DP: data gradient differs per rank, KL gradient is replicated and pre-scaled by 1/world
    data_grad = torch.full((5,), float(rank + 1), dtype=torch.float64)
    kl_grad = torch.arange(5, dtype=torch.float64)
    g_rank = data_grad + kl_grad / world
    dist.all_reduce(g_rank)
    if rank == 0:
        torch.save({"partial": partial, "covered": covered, "full": per_sample.sum(0), "g": g_rank,
                    "g_expected": sum(torch.full((5,), float(r + 1), dtype=torch.float64) for r in range(world)) + kl_grad}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [1024, 1023, 3])
def test_mc_sharding_and_dp_kl_rule_world2(tmp_path, total):
    out = str(tmp_path / "r.pt")
    port = 29600 + (os.getpid() + total) % 300
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.equal(r["covered"], torch.ones(total, dtype=torch.int64))          # every sample exactly once
    assert (r["partial"] - r["full"]).abs().max().item() < 1e-12
    assert torch.equal(r["g"], r["g_expected"])


def test_shard_samples_partition():
    for total in (0, 1, 7, 1024):
        for world in (1, 2, 3, 8):
            spans = [_shard_samples(total, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for first, c in spans:
                assert first == pos
                pos += c


def _dp_worker(rank, world, port, out):
    """The data-parallel rule of the trainers on the real objective (oracle, CPU): rows sharded, per-rank objective
    nll(rank rows) + KL / (NUM_BATCHES * world), gradients SUM-reduced -> the single-process gradient of LRT:222-224."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cases as C
    import lbbnn_oracle as O
    sizes = [(24, 16), (16, 12), (12, 5)]
    case = C.lrt_net_case(seed=21, batch=12, sizes=sizes, classes=5)
    rows = slice(rank * 6, (rank + 1) * 6)
    layers = [{k: v.clone().requires_grad_(True) for k, v in p.items()} for p in case["layers"]]
    logp = O.lrt_net_forward(case["x"][rows], layers, [e[rows] for e in case["eps"]])
    nll = torch.nn.functional.nll_loss(logp, case["y"][rows], reduction="sum")
    kl = sum(O.lrt_kl(p) for p in layers)
    (nll + kl / (C.NUM_BATCHES * world)).backward()
    flat = torch.cat([v.grad.reshape(-1) for p in layers for v in p.values()])
    dist.all_reduce(flat)
    if rank == 0:
        ref = [{k: v.clone().requires_grad_(True) for k, v in p.items()} for p in case["layers"]]
        loss, _, _, _ = O.lrt_net_loss(case["x"], case["y"], ref, case["eps"], C.NUM_BATCHES)
        loss.backward()
        torch.save({"dp": flat, "single": torch.cat([v.grad.reshape(-1) for p in ref for v in p.values()])}, out)
    dist.destroy_process_group()


def test_dp_gradient_rule_on_the_lrt_objective_world2(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_dp_worker, args=(2, 29900 + os.getpid() % 90, out), nprocs=2, join=True)
    r = torch.load(out)
    assert (r["dp"] - r["single"]).abs().max().item() <= 2e-6 * r["single"].abs().max().item()
