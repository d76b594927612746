"""Profiling target: two batched launches of the MC predictive loop (MF 784-400-600-10, 1000 inputs, 21 samples per launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
torch.manual_seed(0)
net = lbbnn.mf.BayesianNetwork().cuda()
mc = lbbnn.mf.MCPredictor(net, batch=1000, seed=1, use_graph=False, samples_per_launch=21)
x = torch.rand(1000, 784, device="cuda")
mc.run(x, 42)
torch.cuda.synchronize()
print("ok", mc.result(42)["pred"][:5].tolist())
