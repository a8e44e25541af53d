"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name: count, total, share."""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float); cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name] += v * scale; cnt[name] += 1
T = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'us/launch':>10s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k[:60]:60s} {cnt[k]:8d} {tot[k]:10.3f} {100*tot[k]/T:6.1f}% {1e3*tot[k]/cnt[k]:10.1f}")
print(f"{'TOTAL':60s} {sum(cnt.values()):8d} {T:10.3f}")
