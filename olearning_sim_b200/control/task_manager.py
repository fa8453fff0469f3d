"""TaskManager — owner of the task lifecycle (queue, scheduling, launch, status fusion, stop, reap, watchdog).

Counterpart of ols_core/taskMgr/task_manager.py (1200 LoC):

    submitTask :186-253        stopTask/stop_task/stop_operation :348-455      getTaskStatus/get_task_status :457-608
    combine_task_status :610-697 (truth table in its docstring)                calculate_conditions :699-889
    getTaskQueue :891-893      changeScheduler :914-915 (`pass` there)         threading_submit_task :917-1051
    run :1053-1069 (5 s)       releaseResource :1071-1148 (10 s)               interruptTask :1150-1200 (300 s)
    get_taskqueue_from_repo :89-155 (restart recovery)

Differences: every loop body is also callable synchronously (``schedule_once`` / ``reap_once`` / ``watchdog_once``)
so tests drive the FSM deterministically; shared state is lock-protected; Ray Jobs are replaced by a ``JobExecutor``;
enum names are validated by table lookup (the reference ``eval``s them).
"""
from __future__ import annotations

import concurrent.futures
import json
import threading
import time
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

from ..api.schema import TaskConfig, TaskStatus, json2taskconfig, taskconfig2json
from ..api.validate import ValidateParameters
from ..utils.logger import Logger
from .jobs import InProcessJobExecutor, JobExecutor
from .repo import TaskTableRepo
from .resource_manager import ResourceManager
from .scheduling import DefaultStrategy, TaskQueue, TaskScheduler
from .task_runner import TaskRunner

logger = Logger()
TIME_FORMAT = "%Y-%m-%d %H:%M:%S"


def get_current_time() -> str:
    return time.strftime(TIME_FORMAT)


def get_time_difference(t0: str, t1: Optional[str] = None) -> float:
    a = time.mktime(time.strptime(t0, TIME_FORMAT))
    b = time.mktime(time.strptime(t1, TIME_FORMAT)) if t1 else time.time()
    return b - a


@dataclass
class TaskMgrConfig:
    """[taskMgr] section of config/config.conf:38-43"""
    scheduler_sleep_time: float = 5.0
    release_sleep_time: float = 10.0
    interrupt_sleep_time: float = 300.0
    interrupt_queue_time: float = 3600.0
    interrupt_running_time: float = 172800.0
    auto_create_rows: bool = True      # the reference requires a GUI-created UNDONE row; standalone use creates it
    submit_workers: int = 4


class TaskManager:
    def __init__(self, repo: Optional[TaskTableRepo] = None, resource_manager: Optional[ResourceManager] = None,
                 deviceflow: Optional[Any] = None, executor: Optional[JobExecutor] = None,
                 config: Optional[TaskMgrConfig] = None, start_threads: bool = False,
                 job_context: Optional[Dict[str, Any]] = None):
        self.config = config or TaskMgrConfig()
        self._task_repo = repo or TaskTableRepo()
        self._resmgr = resource_manager or ResourceManager()
        self._deviceflow = deviceflow
        self._executor = executor or InProcessJobExecutor()
        self._job_context = dict(job_context or {})
        self._job_context.setdefault("repo", self._task_repo)
        self._job_context.setdefault("deviceflow", deviceflow)
        self._task_queue = TaskQueue()
        self._scheduler = TaskScheduler(self._resmgr)
        self._task_running: Dict[str, Dict[str, Any]] = {}     # task_id -> {job_id, task, task_request}
        self._lock = threading.RLock()
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=self.config.submit_workers)
        self._stop_evt = threading.Event()
        self._threads: List[threading.Thread] = []
        self.get_taskqueue_from_repo()
        if start_threads:
            self.start()

    # ================================================================================================
    # restart recovery
    # ================================================================================================
    def get_taskqueue_from_repo(self) -> None:
        """Rebuild the queue (QUEUED rows ordered by in_queue_time, releasing anything they still hold) and the
        running set (resource_occupied == 1) from taskmgr_table."""
        queued = []
        for row in self._task_repo.all_rows():
            tid = row["task_id"]
            try:
                task = json2taskconfig(row["task_params"]) if row.get("task_params") else None
            except ValueError:
                task = None
            if row.get("task_status") == TaskStatus.QUEUED and task is not None:
                if row.get("resource_occupied"):
                    if self._scheduler.release(task):
                        self._task_repo.set_task_value(tid, "resource_occupied", 0)
                queued.append((row.get("in_queue_time") or "", task))
            elif row.get("resource_occupied") == 1 and task is not None:
                self._task_running[tid] = {"job_id": row.get("job_id"), "task": task,
                                           "task_request": DefaultStrategy().get_task_request_resource(task)}
        queued.sort(key=lambda x: x[0])
        self._task_queue.setTaskQueue([t for _, t in queued])

    # ================================================================================================
    # RPC surface
    # ================================================================================================
    def set_job_context(self, key: str, value: Any) -> None:
        """JSON-safe settings handed to every job (e.g. ``deviceflow_address`` once a gRPC session is listening)."""
        self._job_context[key] = value

    def submitTask(self, request: Any) -> bool:
        task = request if isinstance(request, TaskConfig) else json2taskconfig(request)
        ok, msg = ValidateParameters().validate_task_parameters(task)
        if not ok:
            logger.error(task.task_id, "TaskMgr", "task_manager", f"[submitTask]: validation failed: {msg}")
            return False
        tid = task.task_id
        with self._lock:
            if not self._task_repo.check_task_in_database(tid):
                if not self.config.auto_create_rows:
                    logger.error(tid, "TaskMgr", "task_manager", "[submitTask]: the submitted task is not in the repo")
                    return False
                self._task_repo.create_task(tid, task.user_id)
            status = self._task_repo.get_task_value(tid, "task_status")
            if status != TaskStatus.UNDONE:
                logger.error(tid, "TaskMgr", "task_manager", f"[submitTask]: task exists with status {status}")
                return False
            tj = taskconfig2json(task)
            total_simulation = {
                "max_round": task.operatorflow.flow_setting.round,
                "operator_name_list": task.operator_names(),
                "data_name_list": task.data_names(),
                "total_simulation": [{"name": d["name"], "simulation_target": d["total_simulation"]}
                                     for d in tj["target"]["data"]],
            }
            if not self._task_repo.set_item_values("task_id", tid, {
                    "task_status": TaskStatus.QUEUED, "task_params": json.dumps(tj),
                    "total_simulation": json.dumps(total_simulation), "in_queue_time": get_current_time(),
                    "user_id": task.user_id}):
                return False
            self._task_queue.add(task)
        return True

    def stopTask(self, task_id: str) -> bool:
        return self.stop_task(task_id)

    def getTaskStatus(self, task_id: str) -> str:
        return self.get_task_status(task_id)

    def getTaskQueue(self) -> List[str]:
        return self._task_queue.getTaskIDsQueue()

    def changeScheduler(self, strategy: Optional[str]) -> bool:
        """`pass` in the reference; here it swaps the registered scheduling strategy."""
        self._scheduler.set_strategy(strategy)
        return True

    # ================================================================================================
    # stop
    # ================================================================================================
    def stop_task(self, task_id: str, job_id: Optional[str] = None) -> bool:
        occupied = self._task_repo.get_task_value(task_id, "resource_occupied")
        status = self._task_repo.get_task_value(task_id, "task_status")
        if occupied == 1:
            params = self._task_repo.get_task_value(task_id, "task_params")
            if params is None:
                return False
            try:
                task = json2taskconfig(params)
            except ValueError:
                return False
            return self.stop_operation(task_id, status, task, job_id)
        if task_id in self._task_queue:
            self._task_queue.delete(task_id)
            return bool(self._task_repo.set_task_value(task_id, "task_status", TaskStatus.STOPPED))
        return status in TaskStatus.ALL

    def stop_operation(self, task_id: str, task_status: str, task: TaskConfig, job_id: Optional[str] = None) -> bool:
        logical_ok = phone_ok = True
        if task.logical_simulation.resource_request:
            job_id = job_id or self._task_repo.get_task_value(task_id, "job_id")
            if job_id is not None:
                try:
                    self._executor.stop_job(job_id)
                except Exception as e:  # noqa: BLE001
                    logger.error(task_id, "TaskMgr", "task_manager", f"[stop_operation]: stop_job failed: {e}")
                    logical_ok = False
        if any(sum(r.num_request) > 0 for r in task.device_simulation.resource_request):
            try:
                phone_ok = bool(self._resmgr.phone_mgr.stopDevice(task_id))
            except Exception as e:  # noqa: BLE001
                logger.error(task_id, "TaskMgr", "task_manager", f"[stop_operation]: stopDevice failed: {e}")
                phone_ok = False
        return logical_ok and phone_ok

    # ================================================================================================
    # status
    # ================================================================================================
    def _fail(self, task_id: str) -> str:
        self._task_repo.set_task_value(task_id, "task_status", TaskStatus.FAILED)
        return TaskStatus.FAILED

    def get_task_status(self, task_id: str) -> str:
        repo = self._task_repo
        if not repo.check_task_in_database(task_id):
            return TaskStatus.MISSING
        if task_id in self._task_queue:
            return TaskStatus.QUEUED
        if repo.get_task_value(task_id, "resource_occupied") != 1:
            status = repo.get_task_value(task_id, "task_status")
            return status if status in TaskStatus.ALL else self._fail(task_id)
        params = repo.get_task_value(task_id, "task_params")
        if params is None:
            return self._fail(task_id)
        try:
            task = json2taskconfig(params)
        except ValueError:
            return self._fail(task_id)
        # (1) logical side: live job status
        if task.logical_simulation.resource_request:
            try:
                logical_status = str(self._executor.get_job_status(repo.get_task_value(task_id, "job_id")))
            except Exception:  # noqa: BLE001
                if repo.get_task_value(task_id, "logical_target") is None:
                    logical_status = TaskStatus.SUCCEEDED
                else:
                    return self._fail(task_id)
        else:
            logical_status = TaskStatus.SUCCEEDED
        # (2) device side: ask the PhoneMgr, mirror its answer into the repo
        device_result: Dict[str, Any] = {}
        if any(sum(r.num_request) > 0 for r in task.device_simulation.resource_request):
            try:
                st = self._resmgr.phone_mgr.getDeviceTaskStatus(task_id)
                device_result = {"is_finished": bool(st.get("is_finished", True)),
                                 "device_result": list(st.get("device_result", []))}
                repo.set_item_values("task_id", task_id, {
                    "device_round": int(st.get("device_round", 0)), "device_operator": str(st.get("device_operator", "")),
                    "device_result": json.dumps(device_result)})
            except Exception as e:  # noqa: BLE001
                logger.error(task_id, "TaskMgr", "task_manager", f"[get_task_status]: phone status error: {e}")
                return TaskStatus.FAILED
        final = self.combine_task_status(task_id, logical_status, device_result)
        repo.set_task_value(task_id, "task_status", final)
        return final

    def combine_task_status(self, task_id: str, logical_task_status: str, device_task_result: Dict[str, Any]) -> str:
        device_task_status = device_task_result.get("is_finished", True)
        ts = self._task_repo.get_task_value(task_id, "total_simulation")
        if ts is None:
            return TaskStatus.FAILED
        ls, lrf, ds, drf = self.calculate_conditions(task_id, json.loads(ts), device_task_result)
        return combine_status(ls, lrf, logical_task_status, ds, drf, device_task_status)

    def calculate_conditions(self, task_id: str, task_params: Dict[str, Any],
                             device_task_result: Dict[str, Any]) -> Tuple[bool, bool, bool, bool]:
        repo = self._task_repo
        lt = repo.get_task_value(task_id, "logical_target")
        lr = repo.get_task_value(task_id, "logical_result")
        dt = repo.get_task_value(task_id, "device_target")
        return calculate_conditions(
            task_params,
            logical_target=json.loads(lt).get("logical_target", []) if lt else None,
            logical_result=json.loads(lr).get("logical_result", []) if lr else None,
            logical_round=repo.get_task_value(task_id, "logical_round"),
            logical_operator=repo.get_task_value(task_id, "logical_operator"),
            device_target=json.loads(dt).get("device_target", []) if dt else None,
            device_result=device_task_result.get("device_result", []),
            device_round=repo.get_task_value(task_id, "device_round"),
            device_operator=repo.get_task_value(task_id, "device_operator"))

    # ================================================================================================
    # scheduler loop / launch
    # ================================================================================================
    def schedule_once(self, inline: bool = False) -> Optional[str]:
        """One iteration of ``run``: pick the next feasible task, remove it from the queue, launch it."""
        with self._lock:
            res = self._scheduler.run(self._task_queue.getTaskQueue())
            if res is None:
                return None
            self._task_queue.delete(res.task.task_id)
        if inline:
            self.threading_submit_task(res.task, res.task_request)
        else:
            self._pool.submit(self.threading_submit_task, res.task, res.task_request)
        return res.task.task_id

    def threading_submit_task(self, task: TaskConfig, task_request: Dict[str, Any]) -> bool:
        tid, repo = task.task_id, self._task_repo
        if not self._scheduler.freeze(task, task_request):
            # could not get the resources after all: back to the queue (it stays QUEUED)
            self._task_queue.add(task)
            return False
        repo.set_task_value(tid, "resource_occupied", 1)
        if repo.get_task_value(tid, "resource_occupied") != 1:       # write-back verification (:936-948)
            self._scheduler.release(task, task_request)
            return bool(self._fail(tid)) and False
        runner = TaskRunner(task, repo=repo, executor=self._executor, deviceflow=self._deviceflow,
                            phone_mgr=self._resmgr.phone_mgr, job_context=self._job_context)
        try:
            ok, job_id = runner.submit()
        except Exception as e:  # noqa: BLE001
            logger.error(tid, "TaskMgr", "task_manager", f"[threading_submit_task]: submit raised {e}")
            ok, job_id = False, None
        if ok:
            with self._lock:
                self._task_running[tid] = {"job_id": job_id, "task": task, "task_request": task_request}
            repo.set_item_values("task_id", tid, {"job_id": job_id, "task_status": TaskStatus.RUNNING,
                                                  "submit_task_time": get_current_time()})
            if repo.get_task_value(tid, "job_id") != job_id:          # lost write: do not leave an orphan job
                self._executor.stop_job(job_id) if job_id else None
                self._fail(tid)
                return False
            return True
        # failure path (:1001-1051): stop whatever started, FAILED, release now or leave it to the reaper
        self.stop_operation(tid, TaskStatus.FAILED, task, job_id)
        self._fail(tid)
        if self._scheduler.release(task, task_request):
            repo.set_item_values("task_id", tid, {"resource_occupied": 0, "task_finished_time": get_current_time()})
        else:
            with self._lock:
                self._task_running[tid] = {"job_id": job_id, "task": task, "task_request": task_request}
        return False

    # ================================================================================================
    # reaper / watchdog
    # ================================================================================================
    def reap_once(self) -> List[str]:
        """Terminal tasks: stop, release resources, clear resource_occupied, unregister from DeviceFlow.
        A failed step leaves the task in the running set so the next pass retries (reference: re-queued forever)."""
        released = []
        with self._lock:
            running = dict(self._task_running)
        for tid, rec in running.items():
            status = self.get_task_status(tid)
            if status not in TaskStatus.TERMINAL:
                continue
            task = rec["task"]
            if not self.stop_task(tid, rec.get("job_id")):
                continue
            if not self._scheduler.release(task, rec.get("task_request")):
                continue
            self._task_repo.set_item_values("task_id", tid, {"resource_occupied": 0,
                                                             "task_finished_time": get_current_time(),
                                                             "task_status": status})
            if self._deviceflow is not None and task.uses_deviceflow():
                try:
                    if not self._deviceflow.CheckDeviceflowDispatchFinished(tid):
                        self._task_repo.set_task_value(tid, "resource_occupied", 0)
                    self._deviceflow.UnRegisterTask(tid)
                except Exception as e:  # noqa: BLE001
                    logger.error(tid, "TaskMgr", "task_manager", f"[releaseResource]: deviceflow unregister failed: {e}")
            with self._lock:
                self._task_running.pop(tid, None)
            released.append(tid)
        return released

    def watchdog_once(self, now: Optional[float] = None) -> List[str]:
        """Stop tasks queued longer than interrupt_queue_time or running longer than interrupt_running_time."""
        stopped = []
        for row in self._task_repo.all_rows():
            tid, status = row["task_id"], row.get("task_status")
            try:
                if status == TaskStatus.QUEUED and row.get("in_queue_time"):
                    if get_time_difference(row["in_queue_time"]) + (0 if now is None else now - time.time()) > self.config.interrupt_queue_time:
                        if self.stop_task(tid):
                            stopped.append(tid)
                elif status == TaskStatus.RUNNING and row.get("submit_task_time"):
                    if get_time_difference(row["submit_task_time"]) + (0 if now is None else now - time.time()) > self.config.interrupt_running_time:
                        if self.stop_task(tid):
                            stopped.append(tid)
            except (ValueError, TypeError):
                continue
        return stopped

    def _loop(self, fn, period: float) -> None:
        while not self._stop_evt.is_set():
            try:
                fn()
            except Exception as e:  # noqa: BLE001 — a loop must survive any single failure
                logger.error("", "TaskMgr", "task_manager", f"[{fn.__name__}]: {e}")
            self._stop_evt.wait(period)

    def start(self) -> None:
        self._threads = [
            threading.Thread(target=self._loop, args=(self.schedule_once, self.config.scheduler_sleep_time), name="tm-run", daemon=True),
            threading.Thread(target=self._loop, args=(self.reap_once, self.config.release_sleep_time), name="tm-reap", daemon=True),
            threading.Thread(target=self._loop, args=(self.watchdog_once, self.config.interrupt_sleep_time), name="tm-watchdog", daemon=True),
        ]
        for t in self._threads:
            t.start()

    def shutdown(self) -> None:
        self._stop_evt.set()
        self._pool.shutdown(wait=False, cancel_futures=True)

    # convenience for tests / embedded use
    def wait_for(self, task_id: str, timeout: float = 60.0, poll: float = 0.05) -> str:
        t0 = time.time()
        while time.time() - t0 < timeout:
            st = self.get_task_status(task_id)
            if st in TaskStatus.TERMINAL:
                return st
            time.sleep(poll)
        return self.get_task_status(task_id)


# ====================================================================================================
# pure functions (unit-testable without a repo)
# ====================================================================================================
def combine_status(logical_success: bool, logical_round_failed: bool, logical_task_status: str, device_success: bool,
                   device_round_failed: bool, device_task_status: bool) -> str:
    """Truth table of ols_core/taskMgr/task_manager.py:626-697 (2*2*5*2*2*2 = 160 inputs, 90 reachable)."""
    if logical_success and logical_round_failed:
        return TaskStatus.FAILED
    if device_success and device_round_failed:
        return TaskStatus.FAILED
    if logical_success and device_success:
        return TaskStatus.SUCCEEDED
    if (not logical_success and not logical_round_failed and logical_task_status == TaskStatus.STOPPED
            and not device_round_failed and device_task_status is True):
        return TaskStatus.STOPPED
    if not logical_success and logical_task_status in (TaskStatus.SUCCEEDED, TaskStatus.FAILED, TaskStatus.STOPPED):
        return TaskStatus.FAILED
    if not logical_success and logical_round_failed:
        return TaskStatus.FAILED
    if not device_success and device_task_status is True:
        return TaskStatus.FAILED
    if not device_success and device_round_failed:
        return TaskStatus.FAILED
    return TaskStatus.RUNNING


def _sim(d: Dict[str, Any], key: str) -> List[int]:
    return list(d.get("simulation_target", {}).get(key, []))


def calculate_conditions(task_params: Dict[str, Any], logical_target: Optional[List[dict]],
                         logical_result: Optional[List[dict]], logical_round: Optional[int],
                         logical_operator: Optional[str], device_target: Optional[List[dict]],
                         device_result: List[dict], device_round: Optional[int],
                         device_operator: Optional[str]) -> Tuple[bool, bool, bool, bool]:
    """(logical_success, logical_round_failed, device_success, device_round_failed) —
    ols_core/taskMgr/task_manager.py:699-889.  ``*_target`` is None when that side is not used by the task;
    ``logical_result`` is None until the first operator of the first round has reported."""
    max_round = task_params.get("max_round", 0)
    op_names = task_params.get("operator_name_list", [])
    data_names = task_params.get("data_name_list", [])
    total = task_params.get("total_simulation", [])
    last_op = op_names[-1] if op_names else None

    def side_success(target, result, cur_round, cur_op):
        names = [r.get("name", "") for r in result]
        if cur_round is None or cur_round < max_round or cur_op != last_op:
            return False, names
        checks = []
        for i, dt in enumerate(total):
            name = data_names[i]
            if name in names:
                nums = _sim(target[i], "nums") if i < len(target) else []
                dyn = _sim(dt, "dynamic_nums")
                succ = _sim(result[names.index(name)], "success_num")
                checks.append(all(x >= a - z for x, a, z in zip(succ, nums, dyn)))
        return bool(checks) and all(checks), names

    # logical side
    l_res: List[dict] = []
    l_names: List[str] = []
    l_round, l_op = None, None
    if logical_target is not None:
        logical_success = logical_round_failed = False
        if logical_result is not None:
            l_res, l_round, l_op = logical_result, logical_round, logical_operator
            logical_success, l_names = side_success(logical_target, l_res, l_round, l_op)
    else:
        logical_success, logical_round_failed = True, False
    # device side
    d_res: List[dict] = []
    d_names: List[str] = []
    d_round, d_op = None, None
    if device_target is not None:
        device_success = device_round_failed = False
        if device_result:
            d_res, d_round, d_op = device_result, device_round, device_operator
            device_success, d_names = side_success(device_target, d_res, d_round, d_op)
    else:
        device_success, device_round_failed = True, False

    combine_data_status: List[bool] = []
    for i, dt in enumerate(total):
        name = data_names[i]
        nums = _sim(dt, "nums")
        dyn = _sim(dt, "dynamic_nums") or [0] * len(nums)
        if name in l_names:
            r = l_res[l_names.index(name)]
            lf, lsucc = _sim(r, "failed_num"), _sim(r, "success_num")
        else:
            lf, lsucc = [0] * len(dyn), [0] * len(nums)
        if name in d_names:
            r = d_res[d_names.index(name)]
            df, dsucc = _sim(r, "failed_num"), _sim(r, "success_num")
        else:
            df, dsucc = [0] * len(dyn), [0] * len(nums)
        failed_cmp: List[bool] = []
        if not l_res or not d_res:
            failed_cmp = [x < y + z for x, y, z in zip(dyn, lf, df)]
        if l_round is not None and d_round is not None and l_round == d_round and l_op == d_op:
            failed_cmp = [x < y + z for x, y, z in zip(dyn, lf, df)]
        if failed_cmp and True in failed_cmp:
            if not l_res and d_res:
                logical_round_failed, device_round_failed = False, True
            elif l_res and not d_res:
                logical_round_failed, device_round_failed = True, False
            else:
                logical_round_failed, device_round_failed = True, True
            break
        succ_cmp: List[bool] = []
        if not l_res or not d_res:
            succ_cmp = [x + y >= a - z for x, y, a, z in zip(lsucc, dsucc, nums, dyn)]
        if l_round is not None and d_round is not None and l_round == d_round:
            succ_cmp = [x + y >= a - z for x, y, a, z in zip(lsucc, dsucc, nums, dyn)]
        if succ_cmp:
            combine_data_status.append(False not in succ_cmp)
    # a failed round can never also be a success (the reference's own "impossible state" guard)
    if l_res and l_round is not None and l_round >= max_round and l_op == last_op:
        if combine_data_status and False not in combine_data_status and not logical_round_failed:
            logical_success = True
    if d_res and d_round is not None and d_round >= max_round and d_op == last_op:
        if combine_data_status and False not in combine_data_status and not device_round_failed:
            device_success = True
    if logical_round_failed:
        logical_success = False
    if device_round_failed:
        device_success = False
    return logical_success, logical_round_failed, device_success, device_round_failed
