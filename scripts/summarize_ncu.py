#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call brought back (gpurun_out/) into the small text summaries committed under
profiles/:   launches.csv (gpu__time_duration pass)  -> profiles/<tag>_launches.txt
             prof*.ncu-rep (--set full)              -> profiles/<tag>_kernels.txt (+ per-kernel DRAM traffic json)"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rep = sys.argv[3] if len(sys.argv) > 3 else f"prof_{tag}.ncu-rep"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

lc = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(lc):
    lines = [l for l in open(lc) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python scripts/profile_step.py --steps {steps}\n")
        f.write(f"# cold-cache, serialised launch times of {steps} steady-state training steps (lego-shape, 4096 rays): compare SHARES, not absolutes\n")
        f.write(f"# total {tot / 1e3 / steps:.1f} us/step over {sum(a[0] for a in agg.values()) / steps:.0f} launches/step\n")
        f.write(f"{'us/step':>10} {'launches/step':>14} {'share':>7}  kernel\n")
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v / 1e3 / steps:10.1f} {c / steps:14.1f} {100 * v / tot:6.1f}%  {n[:150]}\n")
    print("wrote", f"{tag}_launches.txt")

rp = os.path.join(ROOT, "gpurun_out", rep)
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
            "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum"]
    traffic = {}
    with open(os.path.join(out_dir, f"{tag}_kernels.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ... python scripts/profile_step.py (report {rep})\n")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            f.write(f"\n== {name[:140]}\n")
            for w in want:
                if w in idx:
                    f.write(f"   {w:70s} {r[idx[w]]} {units[idx[w]]}\n")
            try:
                def tobytes(key):
                    v = float(r[idx[key]].replace(",", ""))
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[key]], 1)
                traffic.setdefault(name.split("(")[0].strip(), []).append(tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"))
            except Exception:
                pass
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(os.path.join(out_dir, f"{tag}_dram_bytes_per_launch.json"), "w"), indent=1)
    print("wrote", f"{tag}_kernels.txt")
