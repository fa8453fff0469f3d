#!/usr/bin/env python
"""Export the judged numbers of ncu reports into small tracked files under profiles/.

    python tools/ncu_export.py <tag> gpurun_out/ncu_*.ncu-rep

For every report: profiles/<tag>_ncu_<name>.csv with the raw-page metrics that matter for a roofline reading (duration,
tensor-pipe activity, DRAM / L2 traffic and utilisation, registers, shared memory, occupancy, top warp-stall reasons) and
profiles/<tag>_ncu_<name>_stalls.txt with the 12 source lines that collect the most stall samples (needs -lineinfo)."""
import csv
import os
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_tensor.*|"
                  r"dram__bytes_(read|write)\.sum(\.per_second)?|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_bytes\.sum|"
                  r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|"
                  r"launch__grid_size|launch__block_size|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|lts__t_sector_hit_rate\.pct|"
                  r"l1tex__t_bytes_pipe_lsu_mem_global_op_st\.sum|l1tex__t_sectors_pipe_lsu_mem_global_op_st\.sum|l1tex__t_requests_pipe_lsu_mem_global_op_st\.sum|"
                  r"smsp__inst_executed\.sum|sm__cycles_elapsed\.max)$")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, reps = sys.argv[1], sys.argv[2:]
outdir = os.environ.get("OLSB_PROFILE_DIR") or os.path.join(root, "profiles")      # on the GPU box: export into gpurun_out/
os.makedirs(outdir, exist_ok=True)
for rep in reps:
    name = re.sub(r"^ncu_", "", os.path.splitext(os.path.basename(rep))[0])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    if len(rows) < 3:
        print("skip", rep)
        continue
    hdr, units = rows[0], rows[1]
    out = os.path.join(outdir, f"{tag}_ncu_{name}.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for r in rows[2:]:
            kn = r[hdr.index("Kernel Name")][:100]
            stalls = []
            for i, h in enumerate(hdr):
                if KEEP.match(h):
                    if "issue_stalled" in h:
                        try:
                            stalls.append((float(r[i].replace(",", "")), h, units[i], r[i]))
                        except ValueError:
                            pass
                        continue
                    w.writerow([kn, h, units[i], r[i]])
            for _, h, u, v in sorted(stalls, reverse=True)[:6]:
                w.writerow([kn, h, u, v])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader([l for l in src.splitlines() if not l.startswith("==")]))
    # source page: one section per kernel ("Kernel Name" row, header row, one row per SASS instruction)
    secs, cur, hdr_next = [], None, False
    for r in srows:
        if r and r[0] == "Kernel Name":
            cur = {"kernel": r[1] if len(r) > 1 else "?", "hdr": None, "rows": []}
            secs.append(cur)
            hdr_next = True
        elif cur is not None and hdr_next:
            cur["hdr"], hdr_next = r, False
        elif cur is not None:
            cur["rows"].append(r)
    lines = []
    for sec in secs:
        h = sec["hdr"] or []
        try:
            si = next(i for i, x in enumerate(h) if x.strip() in ("# Samples", "Samples"))
            ci = next(i for i, x in enumerate(h) if x.strip() == "Source")
        except StopIteration:
            continue
        ok = [r for r in sec["rows"] if len(r) > max(si, ci) and r[si].replace(",", "").isdigit()]
        tot = sum(int(r[si].replace(",", "")) for r in ok)
        if not tot:
            continue
        lines.append(f"== {sec['kernel'][:110]}  (total warp-stall samples {tot})")
        for r in sorted(ok, key=lambda r: -int(r[si].replace(",", "")))[:10]:
            n = int(r[si].replace(",", ""))
            lines.append(f"{n:8d} {100.0 * n / tot:5.1f}%  {r[ci].strip()[:140]}")
    if lines:
        with open(os.path.join(outdir, f"{tag}_ncu_{name}_stalls.txt"), "w") as f:
            f.write(f"SASS instructions with the most warp-stall samples, per kernel of report {name}\n" + "\n".join(lines) + "\n")
    print(out)
