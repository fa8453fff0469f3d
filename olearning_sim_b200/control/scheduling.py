"""Task queue, scheduler strategy and task scheduler.

* ``TaskQueue``           — ols_core/taskMgr/task_queue.py:16-49 (a bare list mutated from gRPC threads *and* the
                            scheduler thread; here every operation takes a lock).
* ``DefaultStrategy``     — ols_core/taskMgr/utils/scheduler_strategy.py:37-188: demand per task
                            (cpu = sum num_cpus[class] * num_request[class], mem = 1.0 * num_request "for display"),
                            feasibility filter, score = (n-i)/n + priority/10, argmax (first wins on ties).
* ``TaskScheduler``       — ols_core/taskMgr/task_scheduler.py:22-261: available resources, freeze / release with
                            type 0 (both) / 1 (cluster only) / 2 (phones only).

B200 extension: a compute unit may also declare ``gpu_fraction``; the ledger then tracks GPU share next to the
reference's CPU/mem numbers.
"""
from __future__ import annotations

import threading
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional

from ..api.schema import TaskConfig
from ..utils.logger import Logger

logger = Logger()


class TaskQueue:
    def __init__(self) -> None:
        self._q: List[TaskConfig] = []
        self._lock = threading.RLock()

    def add(self, task: TaskConfig) -> bool:
        with self._lock:
            if any(t.task_id == task.task_id for t in self._q):
                return False
            self._q.append(task)
            return True

    def getTaskQueue(self) -> List[TaskConfig]:
        with self._lock:
            return list(self._q)

    def setTaskQueue(self, tasks: List[TaskConfig]) -> None:
        with self._lock:
            self._q = list(tasks)

    def getTaskIDsQueue(self) -> List[str]:
        with self._lock:
            return [t.task_id for t in self._q]

    def delete(self, task_id: str) -> bool:
        with self._lock:
            n = len(self._q)
            self._q = [t for t in self._q if t.task_id != task_id]
            return len(self._q) != n

    def __contains__(self, task_id: str) -> bool:
        with self._lock:
            return any(t.task_id == task_id for t in self._q)

    def __len__(self) -> int:
        with self._lock:
            return len(self._q)


class TaskSchedulerRes:
    def __init__(self, task: TaskConfig, task_request: Dict[str, Any]):
        self.task = task
        self.task_request = task_request


class SchedulerStrategy(ABC):
    @abstractmethod
    def schedule_next_task(self, task_queue: List[TaskConfig], available_resources: Dict[str, Any]) -> Optional[TaskSchedulerRes]:
        raise NotImplementedError


class DefaultStrategy(SchedulerStrategy):
    def get_task_request_resource(self, task: TaskConfig) -> Dict[str, Any]:
        units: Dict[str, int] = {}
        for rr in task.logical_simulation.resource_request:
            for dev, n in zip(rr.devices, rr.num_request):
                units[dev] = int(units.get(dev, 0)) + int(n)
        cu = task.logical_simulation.computation_unit
        setting = {dev: s for dev, s in zip(cu.devices, cu.setting)}
        cpu = mem = gpu = 0.0
        for dev, n in units.items():
            s = setting.get(dev)
            cpu += (s.num_cpus if s else 0) * n
            mem += 1.0 * n                      # placeholder "for display" in the reference (:76-79)
            gpu += (s.gpu_fraction if s else 0.0) * n
        phones: Dict[str, int] = {}
        for rr in task.device_simulation.resource_request:
            for dev, n in zip(rr.devices, rr.num_request):
                phones[dev] = int(phones.get(dev, 0)) + int(n)
        req: Dict[str, Any] = {"logical_simulation": {"cpu": cpu, "mem": mem},
                               "device_simulation": {task.user_id: phones}}
        if gpu:
            req["logical_simulation"]["gpu"] = gpu
        return req

    def check_resource_availability(self, task_request: Dict[str, Any], available: Dict[str, Any]) -> bool:
        rl = task_request.get("logical_simulation", {})
        al = available.get("logical_simulation", {})
        if not (rl.get("cpu", 0.0) <= al.get("cpu", 0.0) and rl.get("mem", 0.0) <= al.get("mem", 0.0)):
            return False
        if rl.get("gpu", 0.0) > al.get("gpu", 0.0) + 1e-9:
            return False
        dev_req = task_request.get("device_simulation", {})
        if not dev_req:
            return True
        user = next(iter(dev_req))
        for phone_type, n in dev_req[user].items():
            if n > available.get("device_simulation", {}).get(user, {}).get(phone_type, 0):
                return False
        return True

    def schedule_task(self, waiting: List[Dict[str, Any]]) -> int:
        n = len(waiting)
        scores = [(n - i) / n + waiting[i]["task_priority"] / 10 for i in range(n)]
        return scores.index(max(scores))

    def schedule_next_task(self, task_queue, available_resources):
        waiting = []
        for task in task_queue:
            req = self.get_task_request_resource(task)
            if self.check_resource_availability(req, available_resources):
                waiting.append({"task": task, "task_priority": task.target.priority, "task_request": req})
        if not waiting:
            if task_queue:
                logger.info(task_id="", system_name="TaskMgr", module_name="task_scheduler",
                            message="[schedule_next_task]: There are not enough resources for all submitted tasks!")
            return None
        i = self.schedule_task(waiting)
        return TaskSchedulerRes(waiting[i]["task"], waiting[i]["task_request"])


class StrategyFactory:
    _registry: Dict[str, type] = {"default": DefaultStrategy}

    @classmethod
    def register(cls, name: str, strategy_cls: type) -> None:
        cls._registry[name] = strategy_cls

    @classmethod
    def create_strategy(cls, strategy: Optional[str] = None) -> SchedulerStrategy:
        return cls._registry.get(strategy or "default", DefaultStrategy)()


class TaskScheduler:
    """Talks to a ResourceManager (in-process object or gRPC client exposing the same methods)."""

    def __init__(self, resource_manager: Any, strategy: Optional[str] = None):
        self._resmgr = resource_manager
        self._strategy = StrategyFactory.create_strategy(strategy)

    def set_strategy(self, strategy: Optional[str]) -> None:
        self._strategy = StrategyFactory.create_strategy(strategy)

    def get_available_resources(self) -> Dict[str, Any]:
        res = self._resmgr.getResource()
        return {"logical_simulation": dict(res.get("cluster", {})), "device_simulation": dict(res.get("phone", {}))}

    @staticmethod
    def _request_type(task_request: Dict[str, Any]) -> int:
        """0 = both sides, 1 = cluster only, 2 = phones only, -1 = nothing to request (task_scheduler.py:71-174)."""
        rl = task_request.get("logical_simulation", {})
        need_cluster = rl.get("cpu", 0) > 0 or rl.get("mem", 0) > 0 or rl.get("gpu", 0) > 0
        phones = next(iter(task_request.get("device_simulation", {}).values()), {})
        need_phone = any(v > 0 for v in phones.values())
        if need_cluster and need_phone:
            return 0
        if need_cluster:
            return 1
        if need_phone:
            return 2
        return -1

    def freeze(self, task: TaskConfig, task_request: Dict[str, Any]) -> bool:
        typ = self._request_type(task_request)
        if typ < 0:
            return True
        rl = task_request.get("logical_simulation", {})
        phones = next(iter(task_request.get("device_simulation", {}).values()), {})
        ok, msg = self._resmgr.requestResource(task_id=task.task_id, user_id=task.user_id, cpu=rl.get("cpu", 0.0),
                                               mem=rl.get("mem", 0.0), gpu=rl.get("gpu", 0.0), phones=phones, type=typ)
        if not ok:
            logger.error(task.task_id, "TaskMgr", "task_scheduler", f"[freeze]: request resource failed: {msg}")
        return ok

    def release(self, task: TaskConfig, task_request: Optional[Dict[str, Any]] = None) -> bool:
        if task_request is None:
            task_request = DefaultStrategy().get_task_request_resource(task)
        typ = self._request_type(task_request)
        if typ < 0:
            return True
        phones = next(iter(task_request.get("device_simulation", {}).values()), {})
        ok, msg = self._resmgr.releaseResource(task_id=task.task_id, user_id=task.user_id, phones=phones, type=typ)
        if not ok:
            logger.error(task.task_id, "TaskMgr", "task_scheduler", f"[release]: release resource failed: {msg}")
        return ok

    def run(self, task_queue: List[TaskConfig]) -> Optional[TaskSchedulerRes]:
        if not task_queue:
            return None
        return self._strategy.schedule_next_task(task_queue, self.get_available_resources())
