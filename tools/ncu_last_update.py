"""Print the kernels of the last update in an ncu launch-list csv (gpu__time_duration.sum), in launch order
or aggregated.  usage: ncu_last_update.py file.csv [agg]"""
import csv
import sys
from collections import OrderedDict

rows = []
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v_us = v / 1e3 if u in ("ns", "nsecond") else v
        rows.append((r["Kernel Name"].split("(")[0][:70], r.get("Grid Size", ""), v_us))
idx = [i for i, r in enumerate(rows) if "ring_sample" in r[0]]
last = rows[idx[-1]:] if idx else rows
print("last update: launches", len(last), "total %.1f us" % sum(r[2] for r in last))
if len(sys.argv) > 2:
    agg = OrderedDict()
    for k, g, v in last:
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + v)
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.1f} us  x{c:<3d} {k}")
else:
    for r in last:
        print(f"{r[2]:8.1f}  {r[1]:>16s}  {r[0]}")
