"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
python scripts/launch_agg.py <csv> [last_n_launches]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [(r["Kernel Name"], r.get("Grid Size", ""), float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
        for r in csv.DictReader(lines)]
if len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
agg = collections.OrderedDict()
for n, g, v, u in rows:
    a = agg.setdefault((re.sub(r"\(.*", "", n)[-56:], g), [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(r[2] for r in rows)
print(f"{len(rows)} launches, total {tot / 1e6:.3f} ms ({rows[0][3]})")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:56s} {k[1]:16s} n={c:5d} total={v / 1e6:9.3f} ms avg={v / c / 1e3:9.1f} us {100 * v / tot:5.1f}%")
