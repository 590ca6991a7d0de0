"""Condenses an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X python bench.py ...)
into the per-kernel table committed under profiles/: launches, total and mean duration, share of GPU time.
  python tools/launch_summary.py gpurun_out/launches.csv profiles/name.csv"""
import csv
import re
import sys
from collections import OrderedDict

src, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"')) ]
hdr = rows[0]
ki, vi, ui, gi, bi = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
agg = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])[:110]
    ns = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}[r[ui]]
    a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_us", "mean_us", "share_of_gpu_time", "grid(last)", "block(last)"])
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([name, a[0], "%.1f" % (a[1] / 1e3), "%.2f" % (a[1] / 1e3 / a[0]), "%.4f" % (a[1] / total), a[2], a[3]])
    w.writerow(["TOTAL", sum(a[0] for a in agg.values()), "%.1f" % (total / 1e3), "", "1.0", "", ""])
print(open(out).read())
