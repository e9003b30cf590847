#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` dump by address window: executed instructions and stall samples.
usage: ncu_stalls.py src.csv [window_instrs]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
win = int(sys.argv[2]) if len(sys.argv) > 2 else 512
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = {}
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[0], 16)
    if base is None: base = a
    w = (a - base) // 16 // win
    d = agg.setdefault(w, {"n": 0, "exec": 0, "samples": 0, **{s: 0 for s in stalls}})
    d["n"] += 1
    d["exec"] += int(r[col["Instructions Executed"]] or 0)
    d["samples"] += int(r[col["# Samples"]] or 0)
    for s in stalls: d[s] += int(r[col[s]] or 0)
tot = sum(d["samples"] for d in agg.values())
tote = sum(d["exec"] for d in agg.values())
print("window(start instr)  exec%  samples%  top stalls")
for w in sorted(agg):
    d = agg[w]
    top = sorted(((d[s], s) for s in stalls), reverse=True)[:5]
    print(f"{w*win:6d} {100*d['exec']/tote:6.2f} {100*d['samples']/tot:6.2f}  " + " ".join(f"{s[6:]}={100*v/max(1,d['samples']):.0f}%" for v, s in top))
print("total samples", tot, "total warp-instr", tote)
allst = {s: sum(d[s] for d in agg.values()) for s in stalls}
print(" ".join(f"{s[6:]}={100*v/tot:.1f}%" for s, v in sorted(allst.items(), key=lambda x: -x[1])[:8]))
