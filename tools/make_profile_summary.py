#!/usr/bin/env python
"""Build profiles/<tag>_kernels.md from an ncu report (raw page), an ncu launch list and MEASURED_PEAKS.json.
usage: make_profile_summary.py <prof.ncu-rep> <launches.csv> <clients> <tag>"""
import csv, json, os, re, subprocess, sys, collections
rep, launches, C, tag = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "fallback": True}
B = 32
MB = 1 << 20
# algorithmic work per client per launch: (GFLOP, MB moved through HBM)
work = {
 "conv_taps_kernel<32, 1, 32":      ("conv1 fwd+pool",   32768 * 32 * 75 * 2 * 1e-9,            (B * 32 * 1024 + B * 16 * 16 * 32 * 3) / MB),
 "conv1_window_kernel":             ("conv1 fwd+pool",   32768 * 32 * 75 * 2 * 1e-9,            (B * 32 * 1024 + B * 16 * 16 * 32 * 3) / MB),
 "conv_fwd_fold_kernel":            ("conv2 fwd+pool",   8192 * 64 * 800 * 2 * 1e-9,            (B * 16 * 16 * 32 * 2 + B * 8 * 8 * 64 * 3 + 102400) / MB),
 "head_fc_xent_kernel":             ("fc+xent+bwd+unpool", 3 * B * 4096 * 10 * 2 * 1e-9,          (B * 4096 * 3 + B * 16 * 16 * 64 * 2 + 2 * 163840) / MB),
 "conv_dgrad_fold_kernel":          ("conv2 dgrad+gate", 8192 * 32 * 1600 * 2 * 1e-9,           (B * 16 * 16 * 64 * 2 + 2 * B * 16 * 16 * 32 * 2 + 102400) / MB),
 "conv_wgrad_kernel<64, 5, 128":    ("conv2 wgrad+SGD",  8192 * 64 * 800 * 2 * 1e-9,            (B * 16 * 16 * 32 * 2 + B * 16 * 16 * 64 * 2 + 2 * 204800 + 2 * 102400) / MB),
 "conv_wgrad_kernel<32, 1, 64":     ("conv1 wgrad+SGD",  32768 * 32 * 75 * 2 * 1e-9,            (B * 32 * 1024 + B * 16 * 16 * 32 * 3 + 3 * 10240) / MB),
}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try: return float(r[idx[k]].replace(",", ""))
    except Exception: return float("nan")
def gbytes(r, k):
    """DRAM byte counters come with a per-column unit (byte / Kbyte / Mbyte / Gbyte): normalise to GB"""
    u = rows[1][idx[k]].lower()
    return g(r, k) * {"byte": 1e-9, "kbyte": 1e-6, "mbyte": 1e-3, "gbyte": 1.0}.get(u, float("nan"))
seen = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    key = next((k for k in work if k in name), None)
    if key and key not in seen:
        seen[key] = r
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    agg[re.sub(r"\(.*", "", row["Kernel Name"])].append(v)
out = [f"# CNN-2 local-step kernels — ncu summary ({tag})", "",
       f"Source: `ncu --set full --clock-control none --import-source on`, one capture per kernel, {C} resident clients "
       f"(4 per SM), batch 32; launch times from `ncu --metrics gpu__time_duration.sum` over a whole benchmark process "
       f"(serialised, cold caches: compare shares).  Peaks (of measured, `MEASURED_PEAKS.json`): HBM copy "
       f"{peaks['hbm_gbs']:.0f} GB/s, cuBLAS bf16 {peaks['bf16_tflops']:.0f} TFLOP/s.", "",
       "| kernel | role | time (us) | tensor pipe active % | DRAM % of peak | DRAM GB r/w | regs | algorithmic GFLOP | achieved TFLOP/s (of measured) | algorithmic MB | achieved GB/s (of measured) | roofline-limited time (us) | fraction of roofline |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
tot_t = tot_floor = 0.0
for key, (role, gf, mb) in work.items():
    r = seen.get(key)
    if r is None: continue
    t_ms = g(r, "gpu__time_duration.sum"); unit = rows[1][idx["gpu__time_duration.sum"]]
    t_us = t_ms * (1e3 if unit == "ms" else (1e-3 if unit == "ns" else 1.0)) if unit != "us" else t_ms
    tf = gf * C / (t_us * 1e-6) / 1e3
    gbs = mb * MB * C / (t_us * 1e-6) / 1e9
    floor = max(gf * C / (peaks["bf16_tflops"] * 1e3), mb * MB * C / (peaks["hbm_gbs"] * 1e9)) * 1e6
    tot_t += t_us; tot_floor += floor
    out.append(f"| `{key}…` | {role} | {t_us:.0f} | {g(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
               f"{g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {gbytes(r,'dram__bytes_read.sum'):.2f} / {gbytes(r,'dram__bytes_write.sum'):.2f} | "
               f"{r[idx['launch__registers_per_thread']]} | {gf*C:.0f} | {tf:.0f} ({100*tf/peaks['bf16_tflops']:.0f}%) | {mb*C:.0f} | {gbs:.0f} ({100*gbs*1e9/(peaks['hbm_gbs']*1e9):.0f}%) | {floor:.0f} | {floor/t_us:.2f} |")
out += ["", f"Local step total: {tot_t:.0f} us measured vs {tot_floor:.0f} us roofline-limited "
        f"(slower of FLOPs at measured cuBLAS peak and algorithmic bytes at measured copy bandwidth, per kernel) → "
        f"**{tot_floor/tot_t:.2f} of roofline**.", "", "## Launch list (whole process)", "", "```"]
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:14]:
    out.append(f"{sum(v):10.1f} us {100*sum(v)/tot:5.1f}%  n={len(v):3d}  avg={sum(v)/len(v):9.1f} us  {k[:90]}")
out += ["```", ""]
os.makedirs(os.path.join(root, "profiles"), exist_ok=True)
path = os.path.join(root, "profiles", f"{tag}_kernels.md")
open(path, "w").write("\n".join(out))
print(path); print("\n".join(out[:16]))
