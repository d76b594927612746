"""The first section of __graft_entry__.smoke() (eager LRT net, batch 32: forward, loss, backward) on its own, for
compute-sanitizer (initcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("bayesian-neural-nets_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import cases as C
import lbbnn

case = C.lrt_net_case(seed=3, batch=32, sizes=[(784, 400), (400, 600), (600, 10)])
net = lbbnn.BayesianNetwork().cuda()
with torch.no_grad():
    for l, p in zip(net.layers, case["layers"]):
        for k, v in p.items():
            getattr(l, k).copy_(v)
net.train()
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    net.zero_grad()
    logp = net(case["x"].cuda(), sample=True, eps=[e.cuda() for e in case["eps"]])
    loss = torch.nn.functional.nll_loss(logp, case["y"].cuda(), reduction="sum") + net.kl() / 600
    loss.backward()
    torch.cuda.synchronize()
    print(rep, loss.item(), sum(float(p.grad.double().abs().sum()) for p in net.parameters()))
