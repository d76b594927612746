"""Small profiling target: a few eager (non-graph) LRT MNIST-shape training steps, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
torch.manual_seed(0)
net = lbbnn.BayesianNetwork().cuda()
tr = lbbnn.LRTTrainer(net, batch_size=100, num_batches=600, use_graph=False, materialize_grads=False)
tr.x.uniform_(0, 1); tr.y.random_(0, 10)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    tr.step_device()
torch.cuda.synchronize()
print("ok", tr.stats.tolist())
