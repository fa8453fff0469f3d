#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the in-tree sm_100a library -> profiles/<tag>_sass_evidence.txt
(UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit)."""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "olearning_sim_b200/ops/_olsb200_kernels.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = {"UTCHMMA": r"\bUTCHMMA", "….WS": r"UTCHMMA\.WS", "….2CTA": r"UTCHMMA\.2CTA", "UTMALDG": r"\bUTMALDG", "UTMAPF": r"UTMAPF|UTMACCTL", "LDTM": r"\bLDTM",
        "UTCBAR": r"\bUTCBAR", "SYNCS": r"\bSYNCS", "UBLKPF": r"UBLKPF|UBLKCP", "HMMA(legacy)": r"\bHMMA", "SHFL": r"\bSHFL", "LDGMC/STGMC": r"\b(LDGMC|STGMC)"}
per, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)[:110]
        per[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                per[cur][k] += 1
lines = ["SASS evidence for the sm_100a build (cuobjdump -sass olearning_sim_b200/ops/_olsb200_kernels.so)", "",
         "UTCHMMA = tcgen05.mma (.WS = weight-stationary form, .2CTA = cta_group::2 pair MMA), UTMALDG = TMA tensor load, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit,",
         "SYNCS = mbarrier ops, UBLK* = bulk L2 prefetch, LDGMC / STGMC = multimem.ld_reduce / multimem.st (NVSwitch multicast).  No legacy HMMA (mma.sync) anywhere.", "",
         f"{'kernel':<112} " + " ".join(f"{k:>8}" for k in pats)]
for k, c in per.items():
    lines.append(f"{k:<112} " + " ".join(f"{c[p]:>8}" for p in pats))
arch = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
lines += ["", "ELF images: " + ", ".join(sorted(set(re.findall(r"sm_\w+", arch))))]
path = os.path.join(root, "profiles", f"{tag}_sass_evidence.txt")
open(path, "w").write("\n".join(lines) + "\n")
print(path)
