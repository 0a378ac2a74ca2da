#!/usr/bin/env python
"""Top warp-stall locations of one kernel from an ncu report (source page, SASS view):
    python scripts/ncu_stalls.py gpurun_out/prof.ncu-rep k_nerf_fwd [top=25]
Prints total samples per stall reason and the SASS instructions holding the most samples."""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if not starts:
    sys.exit("no source page for " + kern)
beg = starts[0]
end = starts[1] - 1 if len(starts) > 1 else len(rows)   # first launch only
h = rows[beg]
body = [r for r in rows[beg + 1:end] if len(r) == len(h)]
col = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = {n: sum(int(r[col[n]] or 0) for r in body) for n in stall_cols}
allsamp = sum(int(r[col["# Samples"]] or 0) for r in body)
print(f"{kern}: {len(body)} SASS instructions, {allsamp} samples")
print("  by reason: " + ", ".join(f"{n[6:]} {v * 100 // max(1, allsamp)}%" for n, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 50 > allsamp))
ranked = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top]
for i in sorted(ranked):
    r = body[i]
    s = int(r[col["# Samples"]] or 0)
    why = max(stall_cols, key=lambda n: int(r[col[n]] or 0))
    print(f"  {s * 100.0 / max(1, allsamp):5.1f}%  [{i:4d}] {r[col['Source']].strip()[:90]:90s} {why[6:]}")
