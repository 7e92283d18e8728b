#!/usr/bin/env python
"""Developer tool: hottest SASS instructions of an `ncu --page source --csv` dump (stall samples per instruction)."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hdr_i + 1:]:
    if r and r[0] == "Address":      # next kernel of a multi-launch dump: keep the first one only
        break
    if len(r) == len(hdr):
        data.append(r)
total = sum(int(r[col["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", total)
ranked = sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]
for r in ranked:
    n = int(r[col["# Samples"]] or 0)
    st = sorted(((int(r[col[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print(f"{n:7d} {100.0 * n / total:5.1f}%  {r[col['Source']][:90]:90s} {st}")
