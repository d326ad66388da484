"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count,
total and share of the captured time.  usage: summarize_launches.py file.csv [skip_first_n]"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"].split("(")[0], v_us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = OrderedDict()
for k, v in rows:
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
print(f"launches {len(rows)}  total {tot:.1f} us")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us  {100 * t / tot:5.1f}%  x{c:<4d} {k}")
