import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"])
r = d["roofline_large"]
if "kernels" not in r:
    print(r)
else:
    for n, v in r["kernels"].items():
        print("  ", n, round(v["ms"] * 1e3, 1), "us", round(v["frac"], 3), v.get("timing", "")[:170])
