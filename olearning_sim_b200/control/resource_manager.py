"""ResourceManager — the cluster resource ledger (CPU / mem as in the reference, plus GPU share and HBM).

Counterpart of ols_core/resourceMgr/resource_manager.py:18-332.  The reference reads totals from
``ray.cluster_resources()`` (:50-54); here they come from the local box (CPU cores, host RAM, visible GPUs and
their HBM).  Ledger rows live in ``resmgr_table`` (status 0 = held) exactly as in the reference (:73-95,118-123),
but check-then-insert is atomic (the reference is racy: :151-177).  Phone resources are delegated to a PhoneMgr
(``FakePhoneMgr`` in-process, or any object with the same six methods — ols_core/proto/phoneMgr.proto:53-66).
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass
from enum import IntEnum
from typing import Any, Dict, List, Optional, Tuple

from ..utils.logger import Logger
from .repo import ResTableRepo

logger = Logger()


class ServerClusterResStatus(IntEnum):
    """ols_core/proto/resourceService.proto ServerClusterResStatus"""
    SUCCESS = 0
    FAIL = 1
    NOT_ENOUGH_RES = 2
    TASK_REPEAT = 3
    TASKID_NULL = 4
    TASKID_NOT_FOUND = 5


@dataclass
class ResourceDetail:
    """ols_core/resourceMgr/res_param.py:5-14"""
    task_id: str
    cpu: float
    mem: float
    gpu: float = 0.0


@dataclass
class Resource:
    """ols_core/resourceMgr/res_param.py:12-14 (+ gpu: the unit the B200 executor schedules)"""
    cpu: float = 0.0
    mem: float = 0.0
    gpu: float = 0.0


def detect_cluster_totals() -> Dict[str, float]:
    cpu = float(os.cpu_count() or 1)
    mem_gb = 0.0
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemTotal:"):
                    mem_gb = float(line.split()[1]) / 1e6
                    break
    except OSError:
        pass
    gpu = 0.0
    hbm_gb = 0.0
    try:
        import torch
        if torch.cuda.is_available():
            gpu = float(torch.cuda.device_count())
            hbm_gb = sum(torch.cuda.get_device_properties(i).total_memory for i in range(int(gpu))) / 1e9
    except Exception:  # noqa: BLE001
        pass
    return {"cpu": cpu, "mem": mem_gb, "gpu": gpu, "hbm": hbm_gb}


class FakePhoneMgr:
    """In-process stand-in for the external PhoneMgr service (client stubs only in the reference:
    ols_core/proto/phoneMgr.proto:53-66).  Phones are counted per (user, device class)."""

    def __init__(self, inventory: Optional[Dict[str, Dict[str, int]]] = None):
        self._total: Dict[str, Dict[str, int]] = {u: dict(v) for u, v in (inventory or {}).items()}
        self._held: Dict[str, Tuple[str, Dict[str, int]]] = {}
        self._tasks: Dict[str, Dict[str, Any]] = {}
        self._lock = threading.RLock()

    def getDeviceAvailableResource(self) -> Dict[str, Dict[str, int]]:
        with self._lock:
            avail = {u: dict(v) for u, v in self._total.items()}
            for user, phones in self._held.values():
                for k, n in phones.items():
                    avail.setdefault(user, {})[k] = avail.get(user, {}).get(k, 0) - n
            return avail

    def requestDeviceResource(self, task_id: str, user_id: str, phones: Dict[str, int]) -> bool:
        with self._lock:
            avail = self.getDeviceAvailableResource().get(user_id, {})
            if any(n > avail.get(k, 0) for k, n in phones.items()):
                return False
            self._held[task_id] = (user_id, dict(phones))
            return True

    def releaseDeviceResource(self, task_id: str) -> bool:
        with self._lock:
            return self._held.pop(task_id, None) is not None

    def submitTask(self, device_task: Dict[str, Any]) -> bool:
        with self._lock:
            self._tasks[device_task["task_id"]] = {"config": device_task, "is_finished": False, "device_result": [],
                                                   "device_round": 0, "device_operator": ""}
            return True

    def stopDevice(self, task_id: str) -> bool:
        with self._lock:
            t = self._tasks.get(task_id)
            if t:
                t["is_finished"] = True
            return True

    def getDeviceTaskStatus(self, task_id: str) -> Dict[str, Any]:
        with self._lock:
            t = self._tasks.get(task_id)
            if t is None:
                return {"is_finished": True, "device_result": []}
            return {k: t[k] for k in ("is_finished", "device_result", "device_round", "device_operator")}

    # PhoneManager service (proto only in the reference: phoneMgr.proto connectMSPDevices / getLocalDeviceResource)
    def connectMSPDevices(self, user_id: str, devices: Dict[str, int]) -> bool:
        """Attach ``devices`` ({device class: count}) to ``user_id``'s inventory."""
        with self._lock:
            inv = self._total.setdefault(user_id, {})
            for k, n in (devices or {}).items():
                inv[k] = inv.get(k, 0) + int(n)
            return True

    def getLocalDeviceResource(self) -> Dict[str, Dict[str, int]]:
        with self._lock:
            return {u: dict(v) for u, v in self._total.items()}

    # test hook: let a test play the phones' role
    def report(self, task_id: str, device_round: int, device_operator: str, device_result: List[dict],
               is_finished: bool = False) -> None:
        with self._lock:
            t = self._tasks[task_id]
            t.update(device_round=device_round, device_operator=device_operator, device_result=device_result,
                     is_finished=is_finished)


class ResourceManager:
    def __init__(self, repo: Optional[ResTableRepo] = None, totals: Optional[Dict[str, float]] = None,
                 phone_mgr: Optional[Any] = None):
        self._repo = repo or ResTableRepo()
        self._totals = dict(totals) if totals is not None else detect_cluster_totals()
        self._totals.setdefault("gpu", 0.0)
        self._phone = phone_mgr if phone_mgr is not None else FakePhoneMgr()
        self._lock = self._repo.lock()

    # ---- cluster side ------------------------------------------------------------------------------
    def _held(self) -> List[Dict[str, Any]]:
        return self._repo.held()

    def getClusterTotalResource(self) -> Dict[str, float]:
        return {"cpu": self._totals["cpu"], "mem": self._totals["mem"], "gpu": self._totals["gpu"]}

    def getClusterAvailableResource(self) -> Dict[str, float]:
        with self._lock:
            held = self._held()
            return {"cpu": self._totals["cpu"] - sum(r["cpu"] or 0 for r in held),
                    "mem": self._totals["mem"] - sum(r["mem"] or 0 for r in held),
                    "gpu": self._totals["gpu"] - sum(r["gpu"] or 0 for r in held)}

    def getClusterResourceDetail(self) -> List[ResourceDetail]:
        return [ResourceDetail(r["task_id"], r["cpu"] or 0.0, r["mem"] or 0.0, r["gpu"] or 0.0) for r in self._held()]

    def getDetailResbyTaskId(self, task_id: str) -> Optional[ResourceDetail]:
        for r in self._held():
            if r["task_id"] == task_id:
                return ResourceDetail(task_id, r["cpu"] or 0.0, r["mem"] or 0.0, r["gpu"] or 0.0)
        return None

    def requestClusterResource(self, task_id: str, cpu: float, mem: float, gpu: float = 0.0) -> ServerClusterResStatus:
        if not task_id:
            return ServerClusterResStatus.TASKID_NULL
        if cpu < 0 or mem < 0 or gpu < 0:
            return ServerClusterResStatus.FAIL
        with self._lock:                                   # atomic check-then-insert
            avail = self.getClusterAvailableResource()
            if cpu > avail["cpu"] + 1e-9 or mem > avail["mem"] + 1e-9 or gpu > avail["gpu"] + 1e-9:
                return ServerClusterResStatus.NOT_ENOUGH_RES
            if self.getDetailResbyTaskId(task_id) is not None:
                return ServerClusterResStatus.TASK_REPEAT
            ok = self._repo.insert_data(task_id, cpu, mem, gpu)
        return ServerClusterResStatus.SUCCESS if ok else ServerClusterResStatus.FAIL

    def releaseClusterResource(self, task_id: str) -> ServerClusterResStatus:
        with self._lock:
            if self.getDetailResbyTaskId(task_id) is None:
                return ServerClusterResStatus.TASKID_NOT_FOUND
            ok = self._repo.delete_by_taskid(task_id)
        return ServerClusterResStatus.SUCCESS if ok else ServerClusterResStatus.FAIL

    # ---- combined (cluster + phones) — resource_manager.py:262-332 ------------------------------------
    def getResource(self) -> Dict[str, Any]:
        return {"cluster": self.getClusterAvailableResource(), "phone": self._phone.getDeviceAvailableResource()}

    def requestResource(self, task_id: str, user_id: str = "", cpu: float = 0.0, mem: float = 0.0, gpu: float = 0.0,
                        phones: Optional[Dict[str, int]] = None, type: int = 0) -> Tuple[bool, str]:  # noqa: A002
        phones = phones or {}
        if type in (0, 1):
            st = self.requestClusterResource(task_id, cpu, mem, gpu)
            if st != ServerClusterResStatus.SUCCESS:
                return False, st.name
        if type in (0, 2):
            if not self._phone.requestDeviceResource(task_id, user_id, phones):
                if type == 0:
                    self.releaseClusterResource(task_id)   # roll back the half-acquired request
                return False, "PHONE_NOT_ENOUGH_RES"
        return True, "SUCCESS"

    def releaseResource(self, task_id: str, user_id: str = "", phones: Optional[Dict[str, int]] = None,
                        type: int = 0) -> Tuple[bool, str]:  # noqa: A002
        ok = True
        msg = "SUCCESS"
        if type in (0, 1):
            st = self.releaseClusterResource(task_id)
            if st not in (ServerClusterResStatus.SUCCESS, ServerClusterResStatus.TASKID_NOT_FOUND):
                ok, msg = False, st.name
        if type in (0, 2):
            self._phone.releaseDeviceResource(task_id)
        return ok, msg

    # declared-but-unimplemented in the reference (``pass`` at :253-260); here they answer from the ledger
    def getPhoneResource(self) -> Dict[str, Dict[str, int]]:
        return self._phone.getDeviceAvailableResource()

    def getTotalResource(self) -> Dict[str, Any]:
        return {"cluster": self.getClusterTotalResource(), "phone": getattr(self._phone, "_total", {})}

    def getVMResource(self) -> Dict[str, float]:
        return dict(self._totals)

    @property
    def phone_mgr(self) -> Any:
        return self._phone
