"""Command line of the framework (``python -m olearning_sim_b200 ...``).

The reference is started with ``simu_session.py`` (one gRPC server mounting the task / resource / cluster services,
ols_core/simu_session.py:34-71) and driven from other processes through gRPC stubs.  The same here:

    serve   start the control plane (TaskMgr, ResourceMgr, DeviceFlow, cluster manager) behind one gRPC server
    submit  send a task JSON to a running server        status / stop / queue   the other TaskMgr calls
    run     run one task JSON to completion in this process (no server) and print its final status
"""
from __future__ import annotations

import argparse
import json
import sys
import time


def _serve(a) -> int:
    from . import build_control_plane
    from .api.grpc_service import SimulatorSession
    from .cluster.manager import GpuClusterManager
    from .utils.config import load_config
    from .utils.misc_repos import PerformanceManager
    cfg = load_config([a.config] if a.config else None, overrides=a.set)
    if a.svc == 1 and a.remote:
        # the reference's split deployment: this process is the TaskMgr only; ResourceMgr / phone manager / DeviceFlow are
        # other `serve --svc 2|6|5` processes reached over gRPC ([grpc] resmgr / phonemgr, deviceflow.grpc_service_*)
        from . import build_task_service
        tm, rm, df = build_task_service(cfg, start_threads=True)
    else:
        tm, rm, df = build_control_plane(cfg, start_threads=True)
    from .utils.metrics import Metrics
    metrics = Metrics()
    perf = PerformanceManager(metrics=metrics)
    tm._job_context["performance_mgr"] = perf              # round records of in-process jobs feed the metrics
    if a.metrics_port:
        import collections
        import threading
        metrics.start_http_server(a.metrics_port, a.host)

        def poll_tasks():                                   # tasks-per-status gauge from the task table
            while True:
                rows = tm._task_repo.all_rows()
                counts = collections.Counter(str(r.get("task_status", "")) for r in rows)
                metrics.observe_tasks({st: counts.get(st, 0) for st in ("UNDONE", "QUEUED", "RUNNING", "SUCCEEDED", "FAILED", "STOPPED")})
                time.sleep(2.0)
        threading.Thread(target=poll_tasks, name="metrics-tasks", daemon=True).start()
    sess = SimulatorSession(port=a.port or cfg.grpc.taskmgr_port, task_manager=tm, resource_manager=rm, deviceflow=df,
                            cluster_manager=GpuClusterManager(), performance_manager=perf, svc=a.svc,
                            max_workers=cfg.grpc.max_workers, host=a.host, session_sleep_time=cfg.runtime.session_sleep_time,
                            phone_manager=rm.phone_mgr)
    print(json.dumps({"address": sess.address, "services": sess.mounted}), flush=True)
    sess.start(block=True)
    return 0


def _stub(a):
    from .api.grpc_service import JsonStub, connect
    return JsonStub(connect(a.address), "TaskMgr")


def _submit(a) -> int:
    task = json.load(open(a.task))
    ok = _stub(a).submitTask({"task": task})["is_success"]
    print(json.dumps({"task_id": task.get("task_id"), "is_success": ok}))
    return 0 if ok else 1


def _status(a) -> int:
    print(json.dumps(_stub(a).getTaskStatus({"taskID": a.task_id})))
    return 0


def _stop(a) -> int:
    print(json.dumps(_stub(a).stopTask({"taskID": a.task_id})))
    return 0


def _queue(a) -> int:
    print(json.dumps(_stub(a).getTaskQueue({})))
    return 0


def _run(a) -> int:
    from . import build_control_plane
    from .api.schema import TaskStatus
    from .utils.config import load_config
    cfg = load_config([a.config] if a.config else None, overrides=a.set)
    tm, _, _ = build_control_plane(cfg, start_threads=False)
    task = json.load(open(a.task))
    if not tm.submitTask(task):
        print(json.dumps({"task_id": task.get("task_id"), "error": "rejected by validation"}))
        return 1
    tm.schedule_once(inline=True)
    st = tm.wait_for(task["task_id"], a.timeout)
    deadline = time.time() + 5
    while not tm.reap_once() and time.time() < deadline:      # release resources / unregister flows
        time.sleep(0.1)
    print(json.dumps({"task_id": task["task_id"], "status": st}))
    return 0 if st == TaskStatus.SUCCEEDED else 1


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="olearning_sim_b200", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("serve", help="start the gRPC services")
    s.add_argument("--config", default=None, help="INI / YAML / JSON config (reference names accepted)")
    s.add_argument("--set", action="append", default=[], metavar="section.key=value")
    s.add_argument("--svc", type=int, default=0, help="0 all, 1 task, 2 resource, 3 cluster, 4 performance, 5 deviceflow, 6 phone manager")
    s.add_argument("--remote", action="store_true", help="with --svc 1: use the ResourceMgr / DeviceFlow services of other processes (addresses from the config)")
    s.add_argument("--host", default="127.0.0.1")
    s.add_argument("--port", type=int, default=0)
    s.add_argument("--metrics-port", type=int, default=0, help="expose Prometheus metrics on this port (0 = off)")
    s.set_defaults(fn=_serve)
    for name, fn in (("submit", _submit), ("status", _status), ("stop", _stop), ("queue", _queue)):
        c = sub.add_parser(name)
        c.add_argument("--address", default="127.0.0.1:60061")
        if name == "submit":
            c.add_argument("task", help="task JSON file")
        elif name != "queue":
            c.add_argument("task_id")
        c.set_defaults(fn=fn)
    r = sub.add_parser("run", help="run one task JSON in-process")
    r.add_argument("task")
    r.add_argument("--config", default=None)
    r.add_argument("--set", action="append", default=[], metavar="section.key=value")
    r.add_argument("--timeout", type=float, default=3600.0)
    r.set_defaults(fn=_run)
    a = ap.parse_args(argv)
    return a.fn(a)


if __name__ == "__main__":
    sys.exit(main())
