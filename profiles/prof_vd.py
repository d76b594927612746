"""A few replays of the variational-dropout training step (vd_mnist shape) for the ncu launch list."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import lbbnn  # noqa: E402

torch.manual_seed(0)
net = lbbnn.vd.BNN().cuda()
tr = lbbnn.vd.VDTrainer(net, batch_size=100, num_batches=600.0, lr=1e-4, use_graph=False)
x = ((torch.rand(100, 784) - 0.1307) / 0.3081).cuda()
y = torch.randint(0, 10, (100,)).cuda()
tr.x.copy_(x)
tr.y.copy_(y)
for _ in range(3):
    tr.step_device()
torch.cuda.synchronize()
print("kernels per step", tr.kernels_per_step)
