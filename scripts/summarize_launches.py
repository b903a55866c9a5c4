"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: python scripts/summarize_launches.py in.csv out.json"""
import csv
import json
import re
import sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
rows = []
with open(src, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
    name = re.sub(r"^void\s+", "", name)
    agg[name][0] += 1
    agg[name][1] += ns
    total += ns
out = {"source": src, "total_ms": total / 1e6,
       "kernels": [{"kernel": k, "launches": n, "ms": t / 1e6, "share": t / total} for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, open(dst, "w"), indent=1)
for k in out["kernels"][:12]:
    print(f"{k['share'] * 100:6.2f} %  {k['ms']:9.3f} ms  {k['launches']:5d}  {k['kernel'][:110]}")
print(f"total {out['total_ms']:.2f} ms over {sum(k['launches'] for k in out['kernels'])} launches")
