"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
torch.manual_seed(0)
dev = "cuda"
# persistent step kernel: MNIST shape (small classifier head) and an odd 4-layer stack with a wide head (generic path)
for sizes, B in (((784, 400, 600, 10), 100), ((37, 23, 50, 40), 7), ((20, 1), 33)):
    net = lbbnn.BayesianNetwork(sizes).to(dev)
    tr = lbbnn.LRTTrainer(net, batch_size=B, num_batches=600, use_graph=False)
    x = torch.rand(B, sizes[0]); y = torch.randint(0, sizes[-1], (B,))
    for _ in range(2):
        out = tr.step(x, y)
    print("step", sizes, out["nll"])
# per-layer path + eager modules
net = lbbnn.BayesianNetwork((50, 30, 10)).to(dev).train()
logp = net(torch.rand(9, 50, device=dev), sample=True)
(torch.nn.functional.nll_loss(logp, torch.randint(0, 10, (9,), device=dev), reduction="sum") + net.kl() / 600).backward()
# MNF (flows) and MF
mnf = lbbnn.mnf.BayesianNetwork((40, 24, 10)).to(dev).train()
lp = mnf(torch.rand(5, 40, device=dev), sample=True)
(torch.nn.functional.nll_loss(lp, torch.randint(0, 10, (5,), device=dev), reduction="sum") + mnf.kl() / 600).backward()
mf = lbbnn.mf.BayesianNetwork((40, 24, 10)).to(dev).train()
mf.sample_elbo(torch.rand(5, 40, device=dev), torch.randint(0, 10, (5,), device=dev))[0].backward()
# MC predictive, batched and one-sample kernels
for spl in (1, 4):
    mc = lbbnn.mf.MCPredictor(mf, batch=5, seed=3, use_graph=False, samples_per_launch=spl)
    mc.run(torch.rand(5, 40, device=dev), 6)
    print("mc", spl, mc.result(6)["pred"].tolist())
torch.cuda.synchronize()
print("ok")
