#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by (kernel, grid).
usage: python tools/launch_summary.py launches.csv [first_launch] [n_launches]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
body = [r for r in rows[hi + 1:] if len(r) > mv]
a = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else len(body)
body = body[a:a + n]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in body:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    k = (r[kn].split("(")[0][:48], r[gs])
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(body)} launches, {tot / 1e6:.3f} ms of kernel time (ncu: serialised, cold cache)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1] / 1e3:9.1f} us n={v[0]:4d} avg={v[1] / v[0] / 1e3:8.1f} us {100 * v[1] / tot:5.1f}%  {k[0]} grid={k[1]}")
