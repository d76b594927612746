"""Profiling target: a few eager MNF (or MF) MNIST-shape training steps through the drop-in modules, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
kind = sys.argv[1] if len(sys.argv) > 1 else "mnf"
torch.manual_seed(0)
net = (lbbnn.mnf.BayesianNetwork() if kind == "mnf" else lbbnn.mf.BayesianNetwork()).cuda().train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)
x = torch.rand(100, 784, device="cuda"); y = torch.randint(0, 10, (100,), device="cuda")
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    opt.zero_grad(set_to_none=True)
    if kind == "mnf":
        loss = torch.nn.functional.nll_loss(net(x, sample=True), y, reduction="sum") + net.kl() / 600
    else:
        loss = net.sample_elbo(x, y)[0]
    loss.backward(); opt.step()
torch.cuda.synchronize()
print("ok", float(loss))
