"""Profiling target: the captured MNF (argv[1] = mf: MF) training step of GraphedTrainer, two replays after its warm-up --
for an ncu launch list of the step as bench.py runs it (one-launch objective, early / late Adam groups) and for
`ncu --set full -k regex:objective_kernel|adam_multi_kernel`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
kind = sys.argv[1] if len(sys.argv) > 1 else "mnf"
torch.manual_seed(0)
lbbnn.manual_seed(5)
if kind == "mf":
    net = lbbnn.mf.BayesianNetwork().cuda()
    tr = lbbnn.GraphedTrainer(net, batch_size=100, num_batches=600, lr=1e-4, objective="elbo",
                              param_groups=lbbnn.mf.reference_param_groups(net))
else:
    net = lbbnn.mnf.BayesianNetwork().cuda()
    tr = lbbnn.GraphedTrainer(net, batch_size=100, num_batches=600, lr=1e-3)
tr.x.copy_(torch.rand(100, 784, device="cuda")); tr.y.copy_(torch.randint(0, 10, (100,), device="cuda"))
torch.cuda.synchronize()
print("replays start", flush=True)
for _ in range(2):
    tr.step_device()
torch.cuda.synchronize()
print("ok", tr.stats.tolist())
