#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines.

usage: ncu_source_lines.py <report.ncu-rep> <kernel-regex> [top_n] [name-substring]
(the regex matches the base name; the optional substring picks one template instance, e.g. "<32, 1, 64")

ncu's CSV source page is SASS-only, so the per-instruction samples are joined with `nvdisasm --print-line-info` of the
in-tree library by instruction offset inside the kernel (the library must be the build that was profiled)."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 20
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "olearning_sim_b200/ops/_olsb200_kernels.so")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
want = sys.argv[4] if len(sys.argv) > 4 else ""
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and want in r[1]]
if not starts:
    sys.exit(f"no captured kernel matches regex {kre!r} / substring {want!r}")
s0 = starts[0]
kname = rows[s0][1]
hdr = rows[s0 + 1]; idx = {h: i for i, h in enumerate(hdr)}
inst = []
for r in rows[s0 + 2:]:
    if len(r) != len(hdr) or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name": break        # next kernel instance
        continue
    inst.append((int(r[0], 16), r[idx["Source"]].strip(), int(r[idx["# Samples"]] or 0), r))
base = inst[0][0]
stall_cols = [h for h in hdr if h.startswith("stall_")]
# line info
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "--print-instruction-encoding", cubin], capture_output=True, text=True).stdout
short = re.sub(r"\(.*", "", kname).split("::")[-1].split("<")[0]
cur_fn, line, off2line = None, None, {}
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        cur_fn = m.group(1); continue
    if cur_fn is None or short not in cur_fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and line:
        off2line.setdefault((cur_fn, int(m.group(1), 16)), line)
# choose the function instance whose size matches best (template instances share the short name)
fns = collections.Counter(fn for fn, _ in off2line)
def _targs(name):          # template arguments as a normalised tuple of ints/bools
    return tuple(int(x) if x.lstrip("-").isdigit() else (1 if x == "true" else 0)
                 for x in re.findall(r"-?\d+|true|false", re.sub(r"\(int\)|\(bool\)", "", name.split("(CU")[0].split("<", 1)[-1])))
def _margs(fn):            # same from the mangled name: Li<digits>E / Lb<0|1>E
    return tuple(int(x) for x in re.findall(r"L[ib](\d+)E", fn))
cands = [fn for fn in fns if _margs(fn) == _targs(kname)] or list(fns)
best = max(cands, key=lambda fn: sum(1 for a, *_ in inst if (fn, a - base) in off2line))
agg, why = collections.Counter(), collections.defaultdict(collections.Counter)
tot = sum(s for _, _, s, _ in inst)
for a, src, s, r in inst:
    ln = off2line.get((best, a - base), ("?", 0))
    agg[ln] += s
    for h in stall_cols:
        v = int(r[idx[h]] or 0)
        if v: why[ln][h[6:]] += v
print(f"== {kname[:100]}\n   {tot} samples, function {best[:80]}")
srcs = {}
for (f, n), s in agg.most_common(topn):
    path = os.path.join(root, "olearning_sim_b200/ops/csrc", f)
    if f not in srcs and os.path.exists(path): srcs[f] = open(path).read().splitlines()
    text = srcs.get(f, [""] * (n + 1))[n - 1].strip() if f in srcs and n - 1 < len(srcs[f]) else ""
    top = ", ".join(f"{k} {v}" for k, v in why[(f, n)].most_common(2))
    print(f"{s:7d} {100 * s / tot:5.1f}%  {f}:{n:<5d} {text[:90]:<90}  [{top}]")
