#!/usr/bin/env python
"""SASS excerpts of the hand-written kernels: profiles/<tag>_sass_<kernel>.txt holds, for every kernel whose demangled name
matches a pattern, the lines around its tensor-core (UTCHMMA = tcgen05.mma), TMA (UTMALDG / UBLKCP), TMEM (LDTM / STTM) and
commit (UTCBAR) instructions — the inner loops a reader needs to confirm the Blackwell-native path, not the whole listing.

    python tools/sass_excerpt.py <tag> 'bgemm_kernel<0, 0, 256' 'conv_wgrad_kernel' ..."""
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "olearning_sim_b200/ops/_olsb200_kernels.so")
tag, pats = sys.argv[1], sys.argv[2:]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        funcs[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(line)
KEY = re.compile(r"LDGMC|STGMC|UTCHMMA|UTMALDG|UTMASTG|UBLKCP|UBLKPF|LDTM|STTM|UTCBAR|UTCATOMSWS|SYNCS|multimem|LD\.E.*SYS|ST\.E.*SYS|R2UR|ELECT")
for pat in pats:
    for name, lines in funcs.items():
        if pat not in name:
            continue
        keep = set()
        for i, l in enumerate(lines):
            if KEY.search(l):
                keep.update(range(max(0, i - 2), min(len(lines), i + 3)))
        short = re.sub(r"[^A-Za-z0-9]+", "_", re.sub(r"\(.*", "", name))[:80].strip("_")
        out = os.path.join(root, "profiles", f"{tag}_sass_{short}.txt")
        cnt = {k: sum(1 for l in lines if re.search(k, l)) for k in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", r"\bHMMA")}
        with open(out, "w") as f:
            f.write(f"{name}\ninstructions: {sum(1 for l in lines if '/*' in l and ';' in l)}   counts: {cnt}\n"
                    f"(excerpt: +-2 lines around every tcgen05 / TMA / TMEM / mbarrier instruction; cuobjdump -sass of the in-tree .so)\n\n")
            prev = -2
            for i in sorted(keep):
                if i != prev + 1:
                    f.write("        ...\n")
                f.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", lines[i]).rstrip()[:170] + "\n")
                prev = i
        print(out, cnt)
