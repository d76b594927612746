"""Profiling target: two eager steps of the wide (4096-4096-4096-10, B=8192) bf16 trainer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
torch.manual_seed(0)
net = lbbnn.BayesianNetwork((4096, 4096, 4096, 10)).cuda()
tr = lbbnn.LRTTensorCoreTrainer(net, batch_size=8192, num_batches=600, use_graph=False)
tr.x.uniform_(0, 1); tr.y.random_(0, 10)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    tr.step_device()
torch.cuda.synchronize()
print("ok", tr.stats.tolist())
