import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("step", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("blocking_read_value"))
print("vs gpu ref", d.get("vs_gpu_reference"))
if "eager" in d: print("eager", d["eager"].get("ms_per_step"))
print("with occ", d.get("value_with_occupancy_update", {}).get("ms_per_step"))
r = d.get("render", {})
print("render", r.get("ms_per_frame"), r.get("rounds"), r.get("resolved_schedule", "")[:40], "ref", r.get("reference_schedule", {}).get("ms_per_frame"))
for k, v in d.get("configs", {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("ms_per_step", "value", "ms_per_frame", "rounds", "distill_views_per_s")})
for k, v in d.get("kernels", {}).items():
    print("  ", k, round(v["mean_ms"] * 1e3, 1), v["bound"], round(v["frac"], 3))
