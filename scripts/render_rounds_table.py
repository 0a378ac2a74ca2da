#!/usr/bin/env python
"""ncu launch list of one device-driven render (gpurun_out/render_launches.csv) -> per-round table (profiles/<tag>_render_rounds.txt)."""
import csv, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "render_launches.csv")
dst = sys.argv[2] if len(sys.argv) > 2 else None
title = sys.argv[3] if len(sys.argv) > 3 else ""
lines = [l for l in open(src) if not l.startswith("==")]
rows = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
    rows.append((r["Kernel Name"], v))
cols = [("march", "k_march_infer"), ("encoder", "k_grid_fwd"), ("network", "k_nerf_fwd"), ("composite", "k_composite_infer"), ("compact", "k_compact_alive")]
rounds, cur, other = [], None, {}
for name, us in rows:
    key = next((c for c, pat in cols if pat in name), None)
    if key == "march":
        cur = dict.fromkeys([c for c, _ in cols], 0.0)
        rounds.append(cur)
    if key is None or cur is None:
        short = name.split("(")[0].split("::")[-1][:40]
        other[short] = other.get(short, 0.0) + us
        continue
    cur[key] += us
out = []
tot = sum(us for _, us in rows)
out.append(f"# {title}")
out.append(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): microseconds per launch; total {tot:.0f} us over {len(rows)} launches")
out.append("round " + " ".join(f"{c:>9}" for c, _ in cols))
for i, r in enumerate(rounds):
    out.append(f"{i:5d} " + " ".join(f"{r[c]:9.0f}" for c, _ in cols))
out.append("  sum " + " ".join(f"{sum(r[c] for r in rounds):9.0f}" for c, _ in cols))
for k, v in sorted(other.items(), key=lambda kv: -kv[1]):
    out.append(f"# outside the rounds: {k} {v:.0f} us")
text = "\n".join(out) + "\n"
if dst:
    open(dst, "w").write(text)
print(text)
