"""Cut ONE spatpcaCV job out of an ncu launch list of `bench.py` (a job starts at the first tps_fill_kernel
of its Omega build) and summarise it by kernel.  argv: in.csv out_job.csv out_summary.txt [job index]"""
import csv
import collections
import sys

src, out_csv, out_txt = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 1
with open(src, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
starts = []
for i, r in enumerate(rows):
    if "tps_fill_kernel" in r["Kernel Name"] and (not starts or i - starts[-1] > 1000):
        starts.append(i)
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
job = rows[lo:hi]
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["ID", "Kernel Name", "Stream", "Grid Size", "Block Size", "gpu__time_duration.sum [ns]"])
    for r in job:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
        w.writerow([r["ID"], r["Kernel Name"].split("(")[0][:90], r["Stream"], r["Grid Size"], r["Block Size"], int(v)])
tot = collections.OrderedDict()
for r in job:
    name = r["Kernel Name"].split("(")[0][:90]
    v = float(r["Metric Value"].replace(",", ""))
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(v for _, v in tot.values())
own = sum(v for k, (_, v) in tot.items() if "spb::" in k)
with open(out_txt, "w") as f:
    f.write(f"jobs found at launch indices {starts}; job {which}: launches {lo}..{hi}\n")
    f.write(f"one job: {len(job)} launches, {total/1e6:.1f} ms of serialised kernel time "
            f"(own kernels {own/1e6:.1f} ms = {100*own/total:.1f} %)\n")
    for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{v/1e6:10.3f} ms {100*v/total:5.1f}% {c:7d}  {k}\n")
print(open(out_txt).read()[:3000])
