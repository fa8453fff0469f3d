#!/usr/bin/env python
"""Per-opcode warp-stall samples and executed-instruction counts of one captured kernel, plus the samples of every
mbarrier wait (which warp role waits on which barrier).  Separates issue-bound from latency-bound kernels.

usage: ncu_opcode_hist.py <report.ncu-rep> <kernel-regex> [name-substring]"""
import collections, csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
sub = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and sub in r[1]]
if not starts:
    sys.exit(f"no captured kernel matches {kre!r} / {sub!r}")
s0 = starts[0]
hdr = rows[s0 + 1]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[s0 + 2:]:
    if r and r[0] == "Kernel Name": break
    if len(r) == len(hdr) and r[0].startswith("0x"): data.append(r)
tot_s = sum(int(r[idx["# Samples"]] or 0) for r in data)
tot_e = sum(int(r[idx["Instructions Executed"]] or 0) for r in data)
print(f"== {rows[s0][1][:100]}\n   {tot_s} samples, {tot_e} warp instructions executed")
agg = collections.defaultdict(lambda: [0, 0])
for r in data:
    src = r[idx["Source"]].strip()
    op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
    agg[op][0] += int(r[idx["# Samples"]] or 0); agg[op][1] += int(r[idx["Instructions Executed"]] or 0)
for op, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"  {op:12s} executed {e:11d} ({100 * e / max(tot_e, 1):5.1f}%)   samples {s:6d} ({100 * s / max(tot_s, 1):5.1f}%)")
print("  -- mbarrier waits (samples incl. the spin branch that follows)")
for i, r in enumerate(data):
    if "SYNCS.PHASECHK" in r[idx["Source"]]:
        s = int(r[idx["# Samples"]] or 0) + (int(data[i + 1][idx["# Samples"]] or 0) if i + 1 < len(data) else 0)
        if s:
            print(f"  {s:6d} {100 * s / max(tot_s, 1):5.1f}%  {r[idx['Source']].strip()[:90]}")
