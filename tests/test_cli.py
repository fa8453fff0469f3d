"""`python -m olearning_sim_b200`: in-process `run`, and `serve` + `submit` / `status` / `queue` over gRPC."""
import json
import os
import subprocess
import sys
import time

from tests.helpers import make_task_json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MLP = {"model": "mlp2", "local_steps": 1, "batch": 8, "shard": 8, "lr": 0.1, "device": "cpu", "backend": "torch"}


def _cli(*args, timeout=180):
    return subprocess.run([sys.executable, "-m", "olearning_sim_b200", *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=timeout)


def test_run_command_executes_a_task_to_completion(tmp_path):
    f = tmp_path / "t.json"
    f.write_text(json.dumps(make_task_json(task_id="cli_run", rounds=1, nums=(4,), operator_params=MLP)))
    r = _cli("run", str(f), "--timeout", "120")
    assert r.returncode == 0, r.stderr[-500:]
    assert json.loads(r.stdout.strip().splitlines()[-1]) == {"task_id": "cli_run", "status": "SUCCEEDED"}


def test_serve_submit_status_round_trip(tmp_path):
    srv = subprocess.Popen([sys.executable, "-m", "olearning_sim_b200", "serve", "--port", "0",
                            "--set", "taskMgr.scheduler_sleep_time=0.2", "--set", "taskMgr.release_sleep_time=0.2"],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, cwd=ROOT)
    try:
        info = json.loads(srv.stdout.readline())
        assert {"TaskMgr", "ResourceMgr", "TaskOperatorOrientedDeviceFlow"} <= set(info["services"])
        f = tmp_path / "t.json"
        f.write_text(json.dumps(make_task_json(task_id="cli_srv", rounds=1, nums=(4,), operator_params=MLP)))
        r = _cli("submit", "--address", info["address"], str(f))
        assert r.returncode == 0 and json.loads(r.stdout)["is_success"], r.stderr[-500:]
        status = None
        for _ in range(150):
            status = json.loads(_cli("status", "--address", info["address"], "cli_srv").stdout)["taskStatus"]
            if status in ("SUCCEEDED", "FAILED", "STOPPED"):
                break
            time.sleep(0.4)
        assert status == "SUCCEEDED"
        assert "tasks" in json.loads(_cli("queue", "--address", info["address"]).stdout)
    finally:
        srv.terminate()
        srv.wait(10)
