"""Job executors — what the reference delegates to the Ray Jobs API.

The reference submits ``python3 run_task.py --task '<json>'`` to a Ray head through ``JobSubmissionClient``
(ols_core/taskMgr/task_runner.py:41-87) and polls ``get_job_status`` / ``stop_job``
(ols_core/taskMgr/task_manager.py:429,522).  Here a job is the same round-loop driver (``control/run_task.py``)
launched either

* in-process on a worker thread (``InProcessJobExecutor`` — tests, single-GPU service), or
* as a child process group, one rank per GPU via ``torch.distributed.run`` (``SubprocessJobExecutor``),

behind the same three calls: ``submit_job`` / ``get_job_status`` / ``stop_job``.
"""
from __future__ import annotations

import json
import os
import signal
import subprocess
import sys
import threading
import uuid
from typing import Any, Callable, Dict, Optional

from ..api.schema import TaskStatus


class JobStatus:
    PENDING = TaskStatus.PENDING
    RUNNING = TaskStatus.RUNNING
    SUCCEEDED = TaskStatus.SUCCEEDED
    FAILED = TaskStatus.FAILED
    STOPPED = TaskStatus.STOPPED


class JobExecutor:
    def submit_job(self, task_json: Dict[str, Any], context: Optional[Dict[str, Any]] = None) -> str:
        raise NotImplementedError

    def get_job_status(self, job_id: str) -> str:
        raise NotImplementedError

    def stop_job(self, job_id: str) -> bool:
        raise NotImplementedError

    def get_job_error(self, job_id: str) -> str:
        return ""


class InProcessJobExecutor(JobExecutor):
    """Runs the round-loop driver on a daemon thread of this process."""

    def __init__(self, runner_factory: Optional[Callable[..., Any]] = None):
        self._jobs: Dict[str, Dict[str, Any]] = {}
        self._lock = threading.Lock()
        self._factory = runner_factory

    def submit_job(self, task_json, context=None) -> str:
        job_id = f"job_{uuid.uuid4().hex[:12]}"
        stop_evt = threading.Event()
        rec: Dict[str, Any] = {"status": JobStatus.PENDING, "stop": stop_evt, "error": ""}

        def target():
            from .run_task import RoundRunner
            rec["status"] = JobStatus.RUNNING
            try:
                factory = self._factory or RoundRunner
                runner = factory(task_json, stop_event=stop_evt, **(context or {}))
                ok = runner.run()
                if stop_evt.is_set():
                    rec["status"] = JobStatus.STOPPED
                else:
                    rec["status"] = JobStatus.SUCCEEDED if ok else JobStatus.FAILED
            except Exception as e:  # noqa: BLE001
                import traceback
                rec["error"] = f"{type(e).__name__}: {e}\n{traceback.format_exc(limit=6)}"
                rec["status"] = JobStatus.STOPPED if stop_evt.is_set() else JobStatus.FAILED

        th = threading.Thread(target=target, name=job_id, daemon=True)
        rec["thread"] = th
        with self._lock:
            self._jobs[job_id] = rec
        th.start()
        return job_id

    def get_job_status(self, job_id: str) -> str:
        with self._lock:
            rec = self._jobs.get(job_id)
        if rec is None:
            raise KeyError(job_id)
        return rec["status"]

    def get_job_error(self, job_id: str) -> str:
        with self._lock:
            return self._jobs.get(job_id, {}).get("error", "")

    def stop_job(self, job_id: str) -> bool:
        with self._lock:
            rec = self._jobs.get(job_id)
        if rec is None:
            return False
        rec["stop"].set()
        return True

    def wait(self, job_id: str, timeout: Optional[float] = None) -> str:
        with self._lock:
            rec = self._jobs[job_id]
        rec["thread"].join(timeout)
        return rec["status"]


class SubprocessJobExecutor(JobExecutor):
    """One child process group per job: ``python -m torch.distributed.run --nproc-per-node N -m
    olearning_sim_b200.control.run_task --task-file f.json`` (N = GPUs granted by the ResourceManager)."""

    def __init__(self, nproc_per_node: int = 1, workdir: Optional[str] = None, repo_path: Optional[str] = None,
                 master_port_base: int = 29600, max_restarts: int = 0, shrink_on_restart: bool = False,
                 env: Optional[Dict[str, str]] = None):
        """``max_restarts`` > 0 turns the executor into an elastic supervisor: a job whose process group dies (rank
        loss) is relaunched from its last round checkpoint, on one rank fewer when ``shrink_on_restart`` — clients are
        re-sharded by world size, so the survivors simply take over the lost rank's clients."""
        self.nproc, self.workdir, self.repo_path = nproc_per_node, workdir or "/tmp/olsb_jobs", repo_path
        self.max_restarts, self.shrink, self.env = max_restarts, shrink_on_restart, env
        self._port = master_port_base
        self._jobs: Dict[str, Dict[str, Any]] = {}
        self._lock = threading.Lock()
        os.makedirs(self.workdir, exist_ok=True)

    def _launch(self, job_id: str, path: str, nproc: int, attempt: int, port: Optional[int] = None):
        if port is None:
            with self._lock:
                port = self._port
                self._port += 1
        if nproc > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                   "--master-addr", "127.0.0.1", "--master-port", str(port), "-m",
                   "olearning_sim_b200.control.run_task", "--task-file", path]
        else:
            cmd = [sys.executable, "-m", "olearning_sim_b200.control.run_task", "--task-file", path]
        log = open(os.path.join(self.workdir, f"{job_id}.log"), "a")
        log.write(f"--- attempt {attempt}, {nproc} rank(s)\n")
        log.flush()
        env = dict(os.environ, **(self.env or {}))
        return subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT, start_new_session=True, env=env), log.name

    @staticmethod
    def _json_context(context: Optional[Dict[str, Any]]) -> Dict[str, Any]:
        """The context crosses a process boundary as JSON: live objects the in-process executor hands to the runner
        (the task repo, the DeviceFlow service) are dropped — the child re-opens the repo from ``repo_path`` and talks
        to DeviceFlow through its own client."""
        out: Dict[str, Any] = {}
        for k, v in (context or {}).items():
            try:
                json.dumps(v)
            except (TypeError, ValueError):
                continue
            out[k] = v
        return out

    def submit_job(self, task_json, context=None) -> str:
        job_id = f"job_{uuid.uuid4().hex[:12]}"
        path = os.path.join(self.workdir, f"{job_id}.json")
        ctx = self._json_context(context)
        if self.repo_path is not None or "repo_path" not in ctx:
            ctx["repo_path"] = self.repo_path
        # ranks = GPUs granted by the ResourceManager (context["gpus"]) capped by the executor's configured width
        granted = ctx.get("gpus")
        nproc = max(1, min(int(granted), self.nproc)) if isinstance(granted, (int, float)) and granted else self.nproc
        with open(path, "w") as f:
            json.dump({"task": task_json, "context": ctx}, f)
        proc, logname = self._launch(job_id, path, nproc, 0)
        with self._lock:
            self._jobs[job_id] = {"proc": proc, "stopped": False, "log": logname, "path": path, "nproc": nproc,
                                  "restarts": 0}
        return job_id

    def get_job_status(self, job_id: str) -> str:
        # poll + relaunch decision + process swap are one critical section: the reaper thread and a getTaskStatus RPC
        # may observe the same dead process group, and only one of them may start its replacement
        with self._lock:
            rec = self._jobs.get(job_id)
            if rec is None:
                raise KeyError(job_id)
            rc = rec["proc"].poll()
            if rc is None:
                return JobStatus.RUNNING
            if rec["stopped"]:
                return JobStatus.STOPPED
            if rc == 0:
                return JobStatus.SUCCEEDED
            if rec["restarts"] >= self.max_restarts:
                return JobStatus.FAILED
            # elastic recovery: relaunch from the last checkpoint, optionally without the lost rank
            rec["restarts"] += 1
            if self.shrink and rec["nproc"] > 1:
                rec["nproc"] -= 1
            attempt, nproc, path = rec["restarts"], rec["nproc"], rec["path"]
            port = self._port
            self._port += 1
            rec["proc"], _ = self._launch(job_id, path, nproc, attempt, port=port)
            return JobStatus.RUNNING

    def get_job_error(self, job_id: str) -> str:
        with self._lock:
            rec = self._jobs.get(job_id)
        if not rec:
            return ""
        try:
            with open(rec["log"]) as f:
                return f.read()[-4000:]
        except OSError:
            return ""

    def stop_job(self, job_id: str) -> bool:
        with self._lock:
            rec = self._jobs.get(job_id)
        if rec is None:
            return False
        rec["stopped"] = True
        proc = rec["proc"]
        if proc.poll() is None:
            try:
                os.killpg(proc.pid, signal.SIGTERM)      # exact process group we started
            except ProcessLookupError:
                pass
        return True
