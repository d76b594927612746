"""Profiling target: the captured MNF MNIST-shape training step (GraphedTrainer), a few replays, for ncu launch lists."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
torch.manual_seed(0)
lbbnn.manual_seed(5)
net = lbbnn.mnf.BayesianNetwork().cuda()
tr = lbbnn.GraphedTrainer(net, batch_size=100, num_batches=600, lr=1e-3)
tr.x.copy_(torch.rand(100, 784, device="cuda")); tr.y.copy_(torch.randint(0, 10, (100,), device="cuda"))
torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.cuda.profiler.start()
for _ in range(n):
    tr.step_device()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tr.stats.tolist())
