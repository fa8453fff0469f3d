#!/usr/bin/env python
"""profiles/<tag>_ncu_*.csv (tools/ncu_export.py) -> profiles/<tag>_kernels.md: one row per profiled kernel launch.

    python tools/ncu_table.py r2
"""
import csv
import glob
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"


def num(v):
    try:
        return float(v.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(int\)", "", name)
    return name[:70]


recs = []
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", f"{tag}_ncu_*.csv"))):
    if f.endswith("_stalls.txt"):
        continue
    cur = None
    for row in csv.reader(open(f)):
        if len(row) < 4 or row[0] == "kernel":
            continue
        k, m, u, v = row[:4]
        if cur is None or cur["kernel"] != k or (m in cur and "issue_stalled" not in m):
            cur = {"kernel": k, "file": os.path.basename(f), "stalls": []}
            recs.append(cur)
        if "issue_stalled" in m:
            cur["stalls"].append((num(v) or 0.0, re.sub(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", r"\1", m)))
        else:
            cur[m] = (num(v), u)


def g(r, key, scale=1.0, nd=1):
    x = r.get(key)
    if not x or x[0] is None:
        return "—"
    return f"{x[0] * scale:,.{nd}f}"


def dur_us(r):
    x = r.get("gpu__time_duration.sum")
    if not x or x[0] is None:
        return "—"
    v, u = x
    f = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
    return f"{v * f:,.1f}"


def gb(r, key):
    x = r.get(key)
    if not x or x[0] is None:
        return "—"
    v, u = x
    f = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9)
    return f"{v * f:.3f}"


out = [f"# Kernel profiles ({tag}): `ncu --set full --clock-control none`, one row per captured launch", "",
       "Exported by `tools/ncu_export.py` from the `.ncu-rep` files of `tools/gpu_final_ncu.sh` (the per-report CSVs next to this "
       "file hold more metrics and, where the report was taken with `--import-source on`, `*_stalls.txt` lists the source lines "
       "with the most stall samples).  Durations are under the profiler's serialisation and cache control: read them as "
       "relative weights; the timed numbers are in `r2_results.md`.", "",
       "| report | kernel | grid x block | regs | dyn smem KB | µs | tensor pipe % | DRAM % of peak | DRAM GB r / w | L2 % | SM % | top issue-stall reasons |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in recs:
    st = ", ".join(f"{n} {v:.2f}" for v, n in sorted(r["stalls"], reverse=True)[:3])
    out.append("| `%s` | `%s` | %s x %s | %s | %s | %s | %s | %s | %s / %s | %s | %s | %s |" % (
        r["file"].replace(f"{tag}_ncu_", "").replace(".csv", ""), short(r["kernel"]), g(r, "launch__grid_size", 1, 0), g(r, "launch__block_size", 1, 0),
        g(r, "launch__registers_per_thread", 1, 0), g(r, "launch__shared_mem_per_block_dynamic", 1 / 1024 if (r.get("launch__shared_mem_per_block_dynamic") or (0, ""))[1] == "byte" else 1, 1),
        dur_us(r), g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        gb(r, "dram__bytes_read.sum"), gb(r, "dram__bytes_write.sum"), g(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), st))
open(os.path.join(ROOT, "profiles", f"{tag}_kernels.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:20]))
