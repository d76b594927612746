"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel time, share, launch count.
    python profiles/summarize_launches.py gpurun_out/launches.csv [first_launch last_launch]"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hd = rows[h]
ki, vi, ui = hd.index("Kernel Name"), hd.index("Metric Value"), hd.index("Metric Unit")
out = []
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    name = r[ki].split("(")[0].replace("lbbnn::<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
    out.append((name, v * scale))
nums = [int(a) for a in sys.argv[2:] if a.lstrip("-").isdigit()]
lo = nums[0] if nums else 0
hi = nums[1] if len(nums) > 1 else len(out)
out = out[lo:hi]
tot = sum(v for _, v in out)
if "--list" in sys.argv:
    for i, (k, v) in enumerate(out):
        print(f"{lo + i:4d} {v:9.1f} us {100 * v / tot:5.1f}%  {k}")
agg = OrderedDict()
for k, v in out:
    a = agg.setdefault(k, [0.0, 0])
    a[0] += v
    a[1] += 1
for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v:10.1f} us {100 * v / tot:5.1f}%  x{n:<3d} {k}")
print(f"{tot:10.1f} us total over {len(out)} launches")
