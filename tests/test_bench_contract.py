"""The driver-facing contract of bench.py / __graft_entry__.py that can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=300)


def test_reference_arm_reports_unavailable_and_exits_zero():
    r = _run("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "3")
    assert r.returncode == 0, r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and isinstance(d["unavailable"], str) and "\n" not in d["unavailable"]


def test_without_a_gpu_the_bench_fails_loudly_not_silently():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "3")
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert r.returncode != 0 and d["value"] is None and "CUDA" in d["error"]


def test_presets_cover_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    assert sorted(bench.PRESETS) == [2, 3, 4, 5]
    assert bench.PRESETS[2]["clients"] == 1024 and bench.PRESETS[2]["local_steps"] == 5 and bench.PRESETS[2]["model"] == "cnn2"
    assert bench.PRESETS[4]["algorithm"] == "fedprox" and bench.PRESETS[4]["hetero_steps"] and bench.PRESETS[4]["clients"] == 4096
    assert bench.METRIC == "fl_client_local_steps_per_sec"


def test_graft_entry_exposes_build_and_smoke():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    assert callable(g.build) and callable(g.smoke)
