"""Kernel timeline of one 1024-sample MC posterior-predictive step (MCPredictor at BASELINE configs[3]), torch.profiler."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import numpy as np, torch, lbbnn
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
rng = np.random.default_rng(7)
params, _ = bench._mc_net_params(rng)
net = lbbnn.mf.BayesianNetwork(bench.MC_SIZES).to(dev)
with torch.no_grad():
    for l, p in zip(net.layers, params):
        for k, v in p.items():
            getattr(l, k).copy_(torch.as_tensor(v))
mc = lbbnn.mf.MCPredictor(net, batch=1000, seed=4321)
x = torch.rand(1000, 784, device=dev)
for _ in range(3):
    mc.run(x, 1024, first_sample=0); mc.result(1024)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    mc.run(x, 1024, first_sample=0); mc.result(1024)
    torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/mc_trace.json"
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
agg = {}
with open(out.replace(".json", ".txt"), "w") as f:
    for e in ev:
        f.write(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} s{e['args'].get('stream')} {e['name'][:100]}\n")
        k = e["name"][:60]
        agg[k] = (agg.get(k, (0, 0))[0] + e["dur"], agg.get(k, (0, 0))[1] + 1)
    f.write(f"total {ev[-1]['ts'] + ev[-1]['dur'] - t0:.1f} us, {len(ev)} kernels\n")
    for k, (d, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"SUM {d:9.1f} us x{c:4d} {k}\n")
os.remove(out)
print("ok")
