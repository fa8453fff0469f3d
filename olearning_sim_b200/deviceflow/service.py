"""DeviceFlowService — the device-behaviour simulator service (servicer + background loops + restart recovery).

RPC-shaped methods mirror ``TaskOperatorOrientedDeviceFlow`` (ols_core/proto/deviceflow.proto:63-72) as
implemented in ols_core/deviceflow/grpc_service/deviceflow_server.py:

    GetDeviceflowPulsarClient (:167-172)  GetDeviceflowWebsocket (:175-179)  NotifyStart (:182-289)
    NotifyComplete (:291-350)             RegisterTask (:353-385)            UnRegisterTask (:391-401)
    CheckDeviceflowDispatchFinished (:403-427)                               GetTotalComputeResources (`pass` :388-389)

Three loops — ``sort`` (:429-430), ``dispatch`` (:432-451, 2 s poll) and ``flow_release`` (:453-473, 2 s poll) — run as
daemon threads (``start()``) or are stepped deterministically with ``pump()`` in tests.  All shared dicts are guarded
by one lock (the reference shares them unlocked between gRPC handlers and three threads).
"""
from __future__ import annotations

import concurrent.futures
import json
import random
import threading
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

from ..control.repo import FlowTableRepo
from ..utils.logger import Logger
from .bus import Clock, InMemoryBus
from .core import DeviceFlow, Dispatcher, Sorter, TaskOrientedDeviceFlowRegistry
from .strategy import Strategy
from .validate import ComputeResource, ValidateParameters

logger = Logger()


@dataclass
class DeviceFlowConfig:
    """config/deviceflow_config.yaml:3-19 of the reference"""
    grpc_service_address: str = "127.0.0.1"
    grpc_service_port: int = 50051
    max_workers: int = 10
    shelfroom_url: str = "mem://local"
    shelfroom_namespace: str = "public/shelf_room"
    source_url: str = "mem://local"
    source_topic: str = "persistent://public/default/deviceflow_inbound"
    source_websocket_url: str = ""
    poll_seconds: float = 2.0


@dataclass
class NotifyRequest:
    task_id: str
    routing_key: str
    compute_resource: str
    strategy: str = ""
    outbound_service: Dict[str, Dict[str, Any]] = field(default_factory=dict)


class DeviceFlowService:
    def __init__(self, bus: Optional[InMemoryBus] = None, repo: Optional[FlowTableRepo] = None,
                 config: Optional[DeviceFlowConfig] = None, clock: Optional[Clock] = None, seed: Optional[int] = None,
                 now=None, dispatcher_timeout_millis: int = 1000, dashboard: Optional[Any] = None):
        self.config = config or DeviceFlowConfig()
        self.dashboard = dashboard
        self._device_reports: Dict[str, Dict[str, Any]] = {}      # flow_id -> on-device dispatch summary
        self.finished_flows: Dict[str, Dict[str, Any]] = {}       # flow_id -> final counts of device-dispatched flows
        self.bus = bus or InMemoryBus()
        self._repo = repo or FlowTableRepo()
        self._clock = clock or Clock()
        self._seed = seed
        self._now = now
        self._dtimeout = dispatcher_timeout_millis
        self._lock = threading.RLock()
        self._registry: Dict[str, Dict[str, Any]] = {}
        self._flow: Dict[str, Dict[str, Any]] = {}
        self._sorter_params: Dict[str, Dict[str, Any]] = {}
        self._dispatcher: Dict[str, Dispatcher] = {}
        self._futures: Dict[str, concurrent.futures.Future] = {}
        self._deviceflow = DeviceFlow(self.bus, self._repo, self.config.source_url, self.config.source_topic,
                                      self.config.shelfroom_url, self.config.shelfroom_namespace)
        self._registrar = TaskOrientedDeviceFlowRegistry(self._repo)
        self._sorter = Sorter(self._deviceflow._inbound_room, self._deviceflow._shelf_room)
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=self.config.max_workers)
        self._threads: List[threading.Thread] = []
        self._stop_evt = threading.Event()
        self.initiate_from_repo()

    # ---- helpers -----------------------------------------------------------------------------------
    def _rng(self, flow_id: str) -> random.Random:
        return random.Random(f"{self._seed}:{flow_id}") if self._seed is not None else random.Random()

    def _new_dispatcher(self, task_id: str, flow_id: str, is_finished: bool = False) -> Dispatcher:
        f = self._flow[flow_id]
        return Dispatcher(task_id=task_id, flow_id=f["flow_id"], shelf_room=self._deviceflow._shelf_room,
                          strategy=f["strategy"], outbound_service=f["outbound_service"], is_finished=is_finished,
                          clock=self._clock, rng=self._rng(flow_id), default_bus=self.bus, now=self._now,
                          timeout_millis=self._dtimeout, dashboard=self.dashboard)

    # ---- restart recovery (deviceflow_server.py:83-164) ---------------------------------------------------
    def initiate_from_repo(self) -> None:
        with self._lock:
            for row in self._repo.all_rows():
                tid = row["task_id"]
                if row.get("task_registry"):
                    try:
                        self._registry[tid] = json.loads(row["task_registry"])
                        self._registrar._task_registry[tid] = self._registry[tid]
                    except ValueError:
                        continue
            for row in self._repo.all_rows():
                fid = row.get("flow_id")
                if not fid or not row.get("flow"):
                    continue
                try:
                    params = json.loads(row["flow"])
                except ValueError:
                    continue
                self._flow[fid] = params
                tid = params["task_id"]
                if params.get("to_sort"):
                    self._deviceflow._shelf_room.add_shelf(fid)
                    self._sorter_params[fid] = {"task_id": tid, "notify_status": params["notify_start_called"]}
                real_time = Strategy.check_real_time_dispatch(params["strategy"])
                all_complete = tid in self._registry and DeviceFlow.check_all_notify_complete(self._registry[tid], params)
                if params.get("to_sort") and (real_time or all_complete):
                    d = self._new_dispatcher(tid, fid, is_finished=bool(params.get("isFinished")))
                    if all_complete:
                        d.release_dispatch()
                    self._dispatcher[fid] = d
                    params["to_dispatch"] = False          # re-arm: the dispatch loop restarts it
            self._sorter.update(self._sorter_params)

    # ---- RPCs --------------------------------------------------------------------------------------
    def GetDeviceflowPulsarClient(self) -> Dict[str, str]:
        return {"url": self.config.source_url, "topic": self.config.source_topic}

    def GetDeviceflowWebsocket(self) -> Dict[str, str]:
        return {"url": self.config.source_websocket_url}

    def GetTotalComputeResources(self, task_id: str) -> List[str]:
        """`pass` in the reference; answers from the registry here."""
        with self._lock:
            return list(self._registry.get(task_id, {}).get("total_compute_resources", []))

    def RegisterTask(self, task_id: str, total_compute_resources: List[str]) -> bool:
        if any(cr not in ComputeResource.__members__ for cr in total_compute_resources):
            logger.error(task_id, "Deviceflow", "RegisterTask", f"invalid compute resources {total_compute_resources}")
            return False
        with self._lock:
            ok = self._registrar.register_task(task_id, list(total_compute_resources))
            if ok:
                self._registry[task_id] = self._registrar.get(task_id)
            return ok

    def UnRegisterTask(self, task_id: str) -> bool:
        with self._lock:
            self._registrar.unregister_task(task_id)
            for fid in [f for f, p in self._flow.items() if p["task_id"] == task_id]:
                self.stop(task_id=task_id, flow_id=fid, delete_flow_id=True)
            return self.stop(task_id=task_id, delete_task_id=True)

    def NotifyStart(self, req: NotifyRequest) -> bool:
        ok, msg = ValidateParameters().check_params_of_notify_start(req)
        if not ok:
            logger.error(req.task_id, "Deviceflow", "NotifyStart", f"parameter validation failed: {msg}")
            return False
        task_id, flow_id = req.task_id, req.routing_key
        with self._lock:
            if not self._registry.get(task_id):
                logger.error(task_id, "Deviceflow", "NotifyStart", f"task_id={task_id} is not registered")
                return False
            ok, params = self._deviceflow.notify_start(self._flow, task_id, flow_id, req.strategy,
                                                       req.compute_resource, req.outbound_service)
            if not ok:
                self.stop(task_id=task_id, flow_id=flow_id, delete_flow_id=True)
                return False
            self._flow[flow_id] = params
            shelf = self._deviceflow._shelf_room
            first = flow_id not in self._sorter_params
            if not shelf.has_shelf(flow_id):
                shelf.add_shelf(flow_id)
            self._sorter_params[flow_id] = {"task_id": task_id, "notify_status": params["notify_start_called"]}
            params["to_sort"] = True
            self._deviceflow.update_flow(flow_id, task_id, params)
            self._sorter.update(self._sorter_params)
            if first and Strategy.check_real_time_dispatch(params["strategy"]):
                # real-time flows are armed at NotifyStart and released at NotifyComplete
                self._dispatcher[flow_id] = self._new_dispatcher(task_id, flow_id)
            return True

    def ReportDeviceDispatch(self, task_id: str, flow_id: str, compute_resource: str, summary: Dict[str, Any]) -> bool:
        """GPU-native operators apply the flow's strategy on the device (``deviceflow_mask_kernel``): the client updates
        never travel over the bus, so instead of messages the engine reports what the dispatcher would have counted —
        ``{"participants", "forwarded", "dropped", "late", ...}`` for this flow.  When every compute resource has called
        NotifyComplete the flow is closed with these numbers (no shelf to drain) and
        ``CheckDeviceflowDispatchFinished`` sees it as finished, exactly as after a bus-driven dispatch."""
        with self._lock:
            if flow_id not in self._flow or self._flow[flow_id]["task_id"] != task_id:
                return False
            rep = self._device_reports.setdefault(flow_id, {"task_id": task_id, "by_resource": {}})
            rep["by_resource"][compute_resource] = dict(summary)
            for k in ("participants", "forwarded", "dropped", "late"):
                rep[k] = sum(int(v.get(k, 0)) for v in rep["by_resource"].values())
            if self.dashboard is not None and summary.get("amount"):
                try:
                    self.dashboard.write_dispatch_curves(task_id, summary.get("timing", []), summary["amount"],
                                                         self._flow[flow_id]["strategy"])
                except Exception as e:  # noqa: BLE001
                    logger.warning(task_id, "Deviceflow", "ReportDeviceDispatch", f"dashboard write failed: {e}")
            return True

    def device_report(self, flow_id: str) -> Optional[Dict[str, Any]]:
        with self._lock:
            r = self._device_reports.get(flow_id)
            return None if r is None else json.loads(json.dumps(r))

    def NotifyComplete(self, req: NotifyRequest) -> bool:
        task_id, flow_id = req.task_id, req.routing_key
        if req.compute_resource not in ComputeResource.__members__:
            return False
        with self._lock:
            ok, params = self._deviceflow.notify_complete(self._flow, task_id, flow_id, req.compute_resource)
            if not ok:
                self.stop(task_id=task_id, flow_id=flow_id, delete_flow_id=True)
                return False
            all_done = task_id in self._registry and DeviceFlow.check_all_notify_complete(self._registry[task_id], params)
            if all_done and flow_id in self._device_reports and \
                    set(self._device_reports[flow_id]["by_resource"]) >= set(params["notify_complete_called"]):
                # dispatched on the device: nothing on the shelf, close the flow with the reported numbers
                params["isFinished"] = True
                params["device_dispatch"] = {k: self._device_reports[flow_id].get(k, 0)
                                             for k in ("participants", "forwarded", "dropped", "late")}
                self._deviceflow.update_flow(flow_id, task_id, params)
                self.finished_flows[flow_id] = dict(params["device_dispatch"], task_id=task_id)
                d = self._dispatcher.pop(flow_id, None)
                if d is not None:
                    d.stop_dispatch()
                self.stop(task_id=task_id, flow_id=flow_id, delete_flow_id=True)
                return True
            if task_id in self._registry and DeviceFlow.check_all_notify_complete(self._registry[task_id], params):
                if Strategy.check_real_time_dispatch(params["strategy"]):
                    self._dispatcher[flow_id].release_dispatch()
                else:
                    if not self._deviceflow._shelf_room.has_shelf(flow_id):
                        self.stop(task_id=task_id, flow_id=flow_id, delete_flow_id=True)
                        return False
                    d = self._new_dispatcher(task_id, flow_id)
                    d.release_dispatch()
                    self._dispatcher[flow_id] = d
            return True

    def CheckDeviceflowDispatchFinished(self, task_id: str) -> bool:
        with self._lock:
            flow_ids = [f for f in self._repo.get_values_by_conditions("flow_id", task_id=task_id) if f is not None]
            return all(self._flow.get(f, {}).get("isFinished", True) for f in flow_ids)

    # ---- teardown (deviceflow_server.py:475-541) ---------------------------------------------------------
    def stop(self, flow_id: Optional[str] = None, task_id: Optional[str] = None, delete_flow_id: bool = False,
             delete_task_id: bool = False) -> bool:
        if not delete_flow_id and not delete_task_id:
            return False
        if flow_id is None and task_id is None:
            return False
        ok = True
        with self._lock:
            if delete_task_id:
                if task_id is None:
                    return False
                self._registry.pop(task_id, None)
                self._repo.delete_item("task_id", task_id)
            if delete_flow_id:
                if flow_id is None or task_id is None:
                    return False
                self._flow.pop(flow_id, None)
                self._sorter_params.pop(flow_id, None)
                d = self._dispatcher.pop(flow_id, None)
                if d is not None:
                    d.stop_dispatch()
                self._futures.pop(flow_id, None)
                self._repo.delete_item("flow_id", flow_id)
                self._sorter.update(self._sorter_params)
                self._deviceflow._shelf_room.remove_shelf(flow_id)
        return ok

    # ---- loops -------------------------------------------------------------------------------------
    def _dispatch_iteration(self, inline: bool = False) -> None:
        with self._lock:
            todo = [(fid, d) for fid, d in self._dispatcher.items() if not self._flow.get(fid, {}).get("to_dispatch")]
            for fid, d in todo:
                self._flow[fid]["to_dispatch"] = True
                if not self._deviceflow.update_flow(fid, self._flow[fid]["task_id"], self._flow[fid]):
                    self._flow[fid]["to_dispatch"] = False
                    continue
                if not inline:
                    self._futures[fid] = self._pool.submit(d.dispatch)
        if inline:
            for _, d in todo:
                d.dispatch()

    def _release_iteration(self) -> None:
        with self._lock:
            for fid, d in list(self._dispatcher.items()):
                if d.is_finished:
                    self._flow[fid]["isFinished"] = True
                    if not self._deviceflow.update_flow(fid, self._flow[fid]["task_id"], self._flow[fid]):
                        continue
                    self.stop(task_id=self._flow[fid]["task_id"], flow_id=fid, delete_flow_id=True)

    def pump(self, sort: bool = True) -> None:
        """One deterministic iteration of all three loops (tests / in-process use with a FakeClock):
        drain the inbound topic, run newly armed dispatchers inline, release finished flows."""
        if sort:
            self._sorter.drain()
        self._dispatch_iteration(inline=True)
        self._release_iteration()

    def start(self) -> None:
        def loop(fn, period):
            while not self._stop_evt.is_set():
                fn()
                if period:
                    self._stop_evt.wait(period)
        self._threads = [
            threading.Thread(target=lambda: loop(lambda: self._sorter.sort_once(200), 0), name="df-sort", daemon=True),
            threading.Thread(target=lambda: loop(self._dispatch_iteration, self.config.poll_seconds), name="df-dispatch", daemon=True),
            threading.Thread(target=lambda: loop(self._release_iteration, self.config.poll_seconds), name="df-release", daemon=True),
        ]
        for t in self._threads:
            t.start()

    def shutdown(self) -> None:
        self._stop_evt.set()
        self._sorter.stop()
        with self._lock:
            for d in self._dispatcher.values():
                d.stop_dispatch()
        self._pool.shutdown(wait=False, cancel_futures=True)

    # ---- introspection -----------------------------------------------------------------------------
    def stats(self) -> Dict[str, Any]:
        with self._lock:
            return {"registered": sorted(self._registry), "flows": sorted(self._flow),
                    "dispatchers": {k: {"forwarded": d.forwarded, "dropped": d.dropped, "discarded": d.discarded,
                                        "finished": d.is_finished} for k, d in self._dispatcher.items()},
                    "sorter": {"shelved": self._sorter.shelved, "discarded": self._sorter.discarded},
                    "device_dispatched": dict(self.finished_flows)}
