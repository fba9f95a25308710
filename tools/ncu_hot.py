#!/usr/bin/env python
"""Top stall lines from `ncu -i X.ncu-rep --page source --csv` (SASS view).
usage: python tools/ncu_hot.py file.csv [N] [kernel_index]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
# split into per-kernel sections
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "body": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["body"].append(r)
print(len(sections), "kernel sections:", [s["name"][:40] for s in sections])
s = sections[want]
hdr, body = s["hdr"], s["body"]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print(s["name"][:60], "total samples", tot, "instructions", len(body))
agg = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in stall_cols}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(range(len(body)), key=lambda k: -int(body[k][ix["# Samples"]] or 0))[:N]
for k in sorted(top):
    r = body[k]
    st = {hdr[i][6:]: int(r[i]) for i in stall_cols if int(r[i] or 0)}
    print(f"{k:5d} {r[ix['# Samples']]:>6} {r[ix['Source']][:64]:64s} {st}")
