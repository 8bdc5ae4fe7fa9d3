"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import collections
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0][:70]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
total = sum(v for _, v in tot.values())
print(f"total {total/1e6:.3f} ms in {sum(c for c, _ in tot.values())} launches")
for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/1e6:10.3f} ms {100*v/total:5.1f}% {c:6d}  {k}")
