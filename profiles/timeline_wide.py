"""Kernel timeline (start, duration, stream) of one replay of the captured wide bf16 step, from torch.profiler (CUPTI)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
from torch.profiler import profile, ProfilerActivity
torch.manual_seed(0)
lbbnn.manual_seed(5)
B = 8192
net = lbbnn.BayesianNetwork((4096, 4096, 4096, 10)).cuda()
tr = lbbnn.LRTTensorCoreTrainer(net, batch_size=B, num_batches=600, lr=1e-3)
tr.x.copy_(torch.rand(B, 4096, device="cuda")); tr.y.copy_(torch.randint(0, 10, (B,), device="cuda"))
for _ in range(10):
    tr.step_device()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.step_device()
    torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/wide_trace.json"
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
n = len(ev) // 3
ev = ev[2 * n:]
t0 = ev[0]["ts"]
with open(out.replace(".json", ".txt"), "w") as f:
    for e in ev:
        f.write(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} s{e['args'].get('stream')} {e['name'][:110]}\n")
    f.write(f"total {ev[-1]['ts'] + ev[-1]['dur'] - t0:.1f} us, {len(ev)} kernels\n")
os.remove(out)
print("ok")
