"""Summarise `ncu --set full` reports (read here, no GPU needed) into a small JSON committed under profiles/.

    python profiles/ncu_summary.py out.json key=report.ncu-rep[:launch_index[:label]] ...

For every named launch: duration, DRAM bytes read / written, tensor-pipe and DRAM utilisation, L2 hit rate, issue activity,
registers.  bench.py reads `dram_bytes_read + dram_bytes_write` of its dominant kernel from profiles/r02_ncu_summary.json
(`roofline.traffic`): per launch, cold caches (ncu flushes between replays)."""
import csv
import io
import json
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed": "tc_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_lsu_wavefronts_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "smsp__inst_executed.sum": "warp_instructions",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3,
         "ns": 1e-3, "us": 1.0, "ms": 1e3, "second": 1e6}


def launches(report):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[head.index("Kernel Name")]}
        for i, name in enumerate(head):
            if name in METRICS and i < len(r) and r[i] != "":
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                key = METRICS[name]
                if key == "duration":
                    d["duration_us"] = v * SCALE.get(u, 1e-3)
                elif key.startswith("dram_bytes"):
                    d[key] = v * SCALE.get(u, 1.0)
                else:
                    d[key] = v
        out.append(d)
    return out


def main():
    dst, specs = sys.argv[1], sys.argv[2:]
    try:
        summary = json.load(open(dst))
    except (OSError, ValueError):
        summary = {}
    cache = {}
    for spec in specs:
        key, rest = spec.split("=", 1)
        parts = rest.split(":")
        rep, idx = parts[0], int(parts[1]) if len(parts) > 1 else 0
        if rep not in cache:
            cache[rep] = launches(rep)
        d = dict(cache[rep][idx])
        d["report"] = rep.split("/")[-1]
        d["launch_index"] = idx
        if len(parts) > 2:
            d["call"] = parts[2]
        if "duration_us" in d and "dram_bytes_read" in d:
            d["dram_GBps"] = (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["duration_us"] / 1e3
        summary[key] = d
    json.dump(summary, open(dst, "w"), indent=1, sort_keys=True)
    for k in [s.split("=")[0] for s in specs]:
        print(k, json.dumps(summary[k]))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--list":
        for i, d in enumerate(launches(sys.argv[2])):
            print(i, d["kernel"][:90], round(d.get("duration_us", 0), 1), "us")
    else:
        main()
