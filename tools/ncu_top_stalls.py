#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: per kernel section, the top instructions by warp-stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for si, sec in enumerate(sections):
    if only is not None and si != only: continue
    hdr = sec["hdr"]; idx = {h: i for i, h in enumerate(hdr)}
    data = sec["data"]
    tot = sum(int(r[idx['# Samples']]) for r in data) or 1
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    print(f"== [{si}] {sec['name'][:110]}\n   total samples {tot}")
    for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:n]:
        s = int(r[idx['# Samples']])
        st = sorted(((int(r[idx[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{s:6d} {100*s/tot:5.1f}%  {r[idx['Source']].strip()[:76]:76s} {st}")
