"""Phase breakdown of the persistent LRT step kernel with the data-parallel exchange inside the launch (clock64 stamps of CTA 0
and of the last-arriving CTA).   torchrun --nproc-per-node N profiles/prof_step_dp.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bayesian-neural-nets_b200"))
import torch, lbbnn
import torch.distributed as dist
from lbbnn import _capi as K
rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
torch.manual_seed(0)
net = lbbnn.BayesianNetwork().cuda()
tr = lbbnn.LRTTrainer(net, batch_size=100, num_batches=600, use_graph=False, materialize_grads=False, process_group=dist.group.WORLD)
tr.x.uniform_(0, 1); tr.y.random_(0, 10)
stamps = torch.zeros(256, dtype=torch.int64, device="cuda")
K.check(K.lib.lbbnn_lrt_step_profile(stamps.data_ptr()))
acc = None
tail = torch.zeros(2, dtype=torch.float64)
N = 50
for i in range(N + 5):
    stamps.zero_()
    for _ in range(4):
        tr.step_device()          # back to back, like the bench: the ranks stay in lockstep through the closing exchange
    torch.cuda.synchronize()
    s = stamps.cpu()
    n = int((s[:40] != 0).sum())
    d = (s[1:n] - s[:n - 1]).double()
    if i >= 5:
        acc = d if acc is None else acc + d
        tail += torch.tensor([float(s[201] - s[200]), float(s[202] - s[201])], dtype=torch.float64)
K.check(K.lib.lbbnn_lrt_step_profile(None))
if rank == 0:
    print("allreduce:", tr.allreduce, " world", dist.get_world_size())
    acc /= N
    names = []
    for l in range(2):
        names += [f"F{l} items", f"F{l} barrier", f"Fe{l}", f"Fe{l} barrier"]
    names += ["FL (last fwd+loss)", "FL barrier", "BL (last bwd)", "BL barrier", "B1 items", "B1 barrier", "Xe1", "Xe1 barrier",
              "B0 items", "B0 barrier", "DP flags (CTA 0)", "grid barrier", "U (sharded)", "finish (CTA 0)"]
    mhz, tot = 1965.0, 0.0
    for i, c in enumerate(acc.tolist()):
        nm = names[i] if i < len(names) else f"interval {i}"
        print(f"{nm:20s} {c:10.0f} clk  {c / mhz:8.2f} us")
        tot += c
    print(f"total {tot / mhz:.2f} us")
    print(f"last CTA: KL sums {tail[0].item() / N / mhz:.2f} us, closing flag exchange {tail[1].item() / N / mhz:.2f} us")
dist.barrier(); torch.cuda.synchronize(); dist.destroy_process_group()
