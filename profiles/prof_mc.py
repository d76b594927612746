"""Profiling target: launches of the MC predictive loop (MF 784-400-600-10, 1000 inputs, 32 samples per launch, 3xTF32
tensor-core GEMMs for the 400- and 600-wide layers, one lane, no graph)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
torch.manual_seed(0)
net = lbbnn.mf.BayesianNetwork().cuda()
with torch.no_grad():
    for l in net.layers:
        l.lambdal.normal_(0, 2)
mc = lbbnn.mf.MCPredictor(net, batch=1000, seed=1, use_graph=False, samples_per_launch=32, lanes=1,
                          gemm=sys.argv[1] if len(sys.argv) > 1 else "auto")
x = torch.rand(1000, 784, device="cuda")
mc.run(x, 96)
torch.cuda.synchronize()
print("ok", mc.n_tc, mc.result(96)["pred"][:5].tolist())
