"""Phase breakdown of the persistent LRT step kernel: clock64 stamps of CTA 0 at every phase boundary."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
from lbbnn import _capi as K
torch.manual_seed(0)
net = lbbnn.BayesianNetwork().cuda()
tr = lbbnn.LRTTrainer(net, batch_size=100, num_batches=600, use_graph=False, materialize_grads=False)
print(tr.schedule)
tr.x.uniform_(0, 1); tr.y.random_(0, 10)
stamps = torch.zeros(256, dtype=torch.int64, device="cuda")
K.check(K.lib.lbbnn_lrt_step_profile(stamps.data_ptr()))
acc = None
N = 20
for i in range(N + 3):
    stamps.zero_()
    if int(os.environ.get("LBBNN_PROBE_REPEAT", "0")):
        K.check(K.lib.lbbnn_lrt_step_f32(tr._step_desc, 7, tr.ws.data_ptr(), tr.ws.numel(), K.current_stream()))
    else:
        tr.step_device()
    torch.cuda.synchronize()
    s = stamps.cpu()
    if i == N + 2:
        for nm, base in (("fwd", 64), ("dW", 96), ("dX", 128)):
            for l in range(3):
                v = s[base + 8 * l: base + 8 * l + 6]
                if v[0] != 0:
                    print(f"{nm} item l{l}: [stage-B/A1, stage-A/B, sync, mac, reduce, emit] clk:", (v[1:] - v[:-1]).tolist())
    sub = s[40:46].clone()
    subu = s[48:55].clone()
    if i == N + 2:
        print("U of CTA 0: [l0 weights, l0 bias+sum, l1 weights, l1 bias+sum, l2 weights, l2 bias+sum] clk:", (subu[1:] - subu[:-1]).tolist())
    s = s[:40]
    n = int((s != 0).sum())

    d = (s[1:n] - s[:n - 1]).double()
    if i >= 3:
        acc = d if acc is None else acc + d
K.check(K.lib.lbbnn_lrt_step_profile(None))
acc /= N
names = []
REP = int(os.environ.get("LBBNN_PROBE_REPEAT", "0"))
for l in list(range(3)) * (2 if REP else 1):
    names += [f"F{l} items", f"F{l} barrier", f"Fe{l}", f"Fe{l} barrier"]
for l in (2, 1, 0):
    names += [f"B{l} items", f"B{l} barrier"]
    if l == 1:
        names += [f"Xe{l}", f"Xe{l} barrier"]
names += ["U"]
if len(names) != len(acc):     # small-classifier head: the last layer has its own forward+loss and backward phases
    names = []
    for l in range(2):
        names += [f"F{l} items", f"F{l} barrier", f"Fe{l}", f"Fe{l} barrier"]
    names += ["FL (last fwd+loss)", "FL barrier", "BL (last bwd)", "BL barrier", "B1 items", "B1 barrier", "Xe1", "Xe1 barrier",
              "B0 items", "B0 barrier", "U"]
if len(names) != len(acc):
    names = [f"interval {i}" for i in range(len(acc))]
mhz = 1965.0
tot = 0
for nm, c in zip(names, acc.tolist()):
    print(f"{nm:14s} {c:10.0f} clk  {c / mhz:8.2f} us")
    tot += c
print(f"total {tot / mhz:.2f} us ({len(acc)} intervals, {len(names)} names)")
