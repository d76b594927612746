"""Ensemble / sparsity statistics of the MF driver loops against a NumPy restatement of the
reference's own expressions (LBBNN-GP-MF.py:376-396, 427-433, 462-465, 612-637).  Pure torch reductions, so they are
checked on CPU; on a GPU they run where the parameters live."""
import numpy as np
import torch

import cases as C  # noqa: F401  (sets sys.path)


def _net():
    import lbbnn
    torch.manual_seed(3)
    net = lbbnn.mf.BayesianNetwork(sizes=(20, 12, 9, 4))
    with torch.no_grad():
        for l in net.layers:
            l.lambdal.normal_(0, 2)
    return net, lbbnn.mf


def test_refresh_and_median_probability_model():
    net, mf = _net()
    mf.refresh_inclusion(net)
    tot = sum(l.lambdal.numel() for l in net.layers)
    over = 0
    for l, m in zip(net.layers, mf.median_probability_masks(net)):
        alpha = 1 / (1 + np.exp(-l.lambdal.detach().numpy().astype(np.float64)))        # MF:612-616
        assert np.allclose(l.alpha.numpy(), alpha, rtol=1e-6) and l.gamma.alpha is l.alpha and l.gamma.exact is True
        assert np.array_equal(m.numpy(), (l.alpha.numpy() > 0.5).astype(np.float32))     # MF:462-465
        over += int((l.alpha.numpy() > 0.5).sum())
    assert abs(mf.median_probability_density(net).item() - over / tot) < 1e-12          # `os`, MF:634-637


def test_mask_statistics_follow_the_reference_counters():
    net, mf = _net()
    mf.refresh_inclusion(net)
    rng = np.random.default_rng(5)
    S = 6
    draws = [([torch.from_numpy((rng.random(l.alpha.shape) < l.alpha.numpy()).astype(np.float32)) for l in net.layers],
              [torch.from_numpy((rng.random(l.alpha.shape) < l.alpha.numpy()).astype(np.float32)) for l in net.layers])
             for _ in range(S)]
    got = mf.mask_statistics(net, S, draws=draws)
    # the reference's loop (MF:376-396): spars += sum(g > 0.5) / total, gt += (g > 0.5), density[i] = cat(g').mean()
    tot = sum(l.alpha.numel() for l in net.layers)
    spars, density = 0.0, []
    gt = [np.zeros(l.alpha.shape) for l in net.layers]
    for ga, gb in draws:
        spars += sum(int((g.numpy() > 0.5).sum()) for g in ga) / tot
        gt = [t + (g.numpy() > 0.5) for t, g in zip(gt, ga)]
        density.append(np.concatenate([g.numpy().ravel() for g in gb]).mean())
    assert abs(got["sparsity"].item() - spars / S) < 1e-12                             # `spars / ctr`, MF:428
    assert abs(got["ever_active"].item() - sum(int((t > 0).sum()) for t in gt) / tot) < 1e-12      # `ps` x 10, MF:427
    assert abs(got["density"].item() - float(np.mean(density))) < 1e-7                 # `np.mean(density)`, MF:431
    # native draws: Bernoulli(alpha) on the parameters' device; the three statistics are consistent with alpha
    torch.manual_seed(0)
    nat = mf.mask_statistics(net, 200)
    mean_alpha = sum(l.alpha.sum().item() for l in net.layers) / tot
    assert abs(nat["sparsity"].item() - mean_alpha) < 0.02 and abs(nat["density"].item() - mean_alpha) < 0.02
    assert nat["ever_active"].item() >= nat["sparsity"].item()


def test_lrt_predict_ensemble_follows_the_reference_loop():
    """lbbnn.lrt.predict_ensemble against the reference's test_ensemble body (LRT:239-265) restated with NumPy, on a stub
    stochastic network (the helper is plain torch around `net(x, sample=...)`; the lbbnn layers themselves are CUDA-only)."""
    import lbbnn
    from scipy.special import expit

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(1)
            self.w = torch.randn(12, 5, generator=g)
            self.noise = [torch.randn(7, 5, generator=g) * 0.7 for _ in range(14)]
            self.calls = 0

        def forward(self, x, sample=False):
            z = x @ self.w
            if sample:
                z = z + self.noise[self.calls]
                self.calls += 1
            return torch.log_softmax(z, dim=1)

    x = torch.randn(7, 12, generator=torch.Generator().manual_seed(2))
    S = 14
    got = lbbnn.lrt.predict_ensemble(Stub(), x, S)
    ref = Stub()
    outputs = np.stack([ref(x, sample=True).numpy() for _ in range(S)])
    means = np.zeros((7, 5))
    for i in range(S):                                            # LRT:249-258
        tmp = expit(outputs[i].astype(np.float64))
        tmp /= tmp.sum(1, keepdims=True)
        means += tmp
    means /= S
    assert np.allclose(got["mean_prob"].numpy(), means, atol=1e-7)
    assert np.array_equal(got["ensemble"].numpy(), outputs[0:10].mean(0).argmax(1))          # LRT:262-263
    assert np.array_equal(got["posterior_mean"].numpy(), ref(x, sample=False).numpy().argmax(1))
    assert np.allclose(got["entropy"].numpy(), -(means * np.log(means)).sum(1), atol=1e-7)     # outofsample, MF:487-490
    assert got["density"] is None
