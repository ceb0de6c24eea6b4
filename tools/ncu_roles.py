#!/usr/bin/env python
"""Per-region instruction and stall-sample totals from an ncu source page (csv), regions given as name:hexstart ...
   python tools/ncu_roles.py src.csv loader:3100 foldE:47a0 ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
regions = [(a.split(":")[0], int(a.split(":")[1], 16)) for a in sys.argv[2:]]
regions.sort(key=lambda x: x[1])
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
base = int(data[0][col["Address"]], 16) if data[0][col["Address"]].startswith("0x") else int(data[0][col["Address"]])
tot = {}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in data:
    a = r[col["Address"]]
    addr = (int(a, 16) if a.startswith("0x") else int(a)) - base
    name = "pre"
    for n, s in regions:
        if addr >= s: name = n
    t = tot.setdefault(name, {"inst": 0, "samples": 0, **{h: 0 for h in stall_cols}})
    t["inst"] += int(r[col["Instructions Executed"]] or 0)
    t["samples"] += int(r[col["# Samples"]] or 0)
    for h in stall_cols: t[h] += int(r[col[h]] or 0)
ti = sum(t["inst"] for t in tot.values()); ts = sum(t["samples"] for t in tot.values())
print(f"total warp-instr {ti}  samples {ts}")
for n, t in tot.items():
    top = sorted(((t[h], h[6:]) for h in stall_cols), reverse=True)[:5]
    print(f"{n:10s} inst {t['inst']:11d} ({100*t['inst']/ti:5.1f}%)  samples {t['samples']:7d} ({100*t['samples']/ts:5.1f}%)  " + " ".join(f"{h}={v}" for v, h in top if v))
