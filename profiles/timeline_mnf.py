"""Kernel timeline (start, duration, stream) of one replay of the captured MNF (or, argv[2] = mf, MF) training step, from
torch.profiler (CUPTI)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
from torch.profiler import profile, ProfilerActivity
torch.manual_seed(0)
lbbnn.manual_seed(5)
kind = sys.argv[2] if len(sys.argv) > 2 else "mnf"
if kind == "mf":      # LBBNN-GP-MF.py: sample_elbo objective, the script's 33 Adam parameter groups
    net = lbbnn.mf.BayesianNetwork().cuda()
    tr = lbbnn.GraphedTrainer(net, batch_size=100, num_batches=600, lr=1e-4, objective="elbo",
                              param_groups=lbbnn.mf.reference_param_groups(net))
else:
    net = lbbnn.mnf.BayesianNetwork().cuda()
    tr = lbbnn.GraphedTrainer(net, batch_size=100, num_batches=600, lr=1e-3)
tr.x.copy_(torch.rand(100, 784, device="cuda")); tr.y.copy_(torch.randint(0, 10, (100,), device="cuda"))
for _ in range(20):
    tr.step_device()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.step_device()
    torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/mnf_trace.json"
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# last replay only
n = len(ev) // 3
ev = ev[2 * n:]
t0 = ev[0]["ts"]
with open(out.replace(".json", ".txt"), "w") as f:
    for e in ev:
        f.write(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} s{e['args'].get('stream')} {e['name'][:90]}\n")
    f.write(f"total {ev[-1]['ts'] + ev[-1]['dur'] - t0:.1f} us, {len(ev)} kernels\n")
    agg = {}
    for e in ev:
        k = e["name"][:70]
        agg[k] = (agg.get(k, (0, 0))[0] + e["dur"], agg.get(k, (0, 0))[1] + 1)
    for k, (d, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        f.write(f"SUM {d:9.1f} us x{c:4d} {k}\n")
os.remove(out)
print("ok")
