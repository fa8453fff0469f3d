"""Sorter, Dispatcher, task registry and flow state of DeviceFlow.

* ``Sorter``      ols_core/deviceflow/non_grpc/sorter.py:20-91 — single consumer on the inbound topic; a message is
                  shelved only while its (flow_id, compute_resource) is between NotifyStart and teardown, otherwise it
                  is acked and discarded.
* ``Dispatcher``  ols_core/deviceflow/non_grpc/dispatcher.py:47-252 — real-time mode (batch release with i.i.d.
                  drops, flush on release) and flow mode (timed slots, drop lists, leftovers discarded).
* ``TaskOrientedDeviceFlowRegistry`` ols_core/deviceflow/non_grpc/registry.py:14-112.
* ``DeviceFlow``  ols_core/deviceflow/non_grpc/deviceflow.py:29-197 — flow parameter bookkeeping + persistence.

All timing goes through an injectable ``Clock`` and all randomness through an injectable ``random.Random`` so the
behaviour is reproducible and testable with a ``FakeClock``.
"""
from __future__ import annotations

import concurrent.futures
import json
import random
import threading
from typing import Any, Callable, Dict, List, Optional

from ..control.repo import FlowTableRepo
from ..utils.logger import Logger
from .bus import BusClosed, BusTimeout, Clock, Consumer, InMemoryBus
from .rooms import InboundRoom, ShelfRoom, make_outbound_producer, parse_envelope
from .strategy import Strategy

TIMEOUT_MILLIS = 1000
logger = Logger()


class Sorter:
    def __init__(self, inbound_room: InboundRoom, shelf_room: ShelfRoom, subscription: str = "sorter"):
        self._consumer = inbound_room.get_consumer(subscription)
        self._shelf_room = shelf_room
        self._params: Dict[str, Dict[str, Any]] = {}
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self.shelved = 0
        self.discarded = 0

    def update(self, params: Dict[str, Dict[str, Any]]) -> None:
        with self._lock:
            self._params = {k: {"task_id": v["task_id"], "notify_status": dict(v["notify_status"])} for k, v in params.items()}

    def stop(self) -> None:
        self._stop.set()

    def sort_once(self, timeout_millis: int = TIMEOUT_MILLIS) -> bool:
        """Process at most one inbound message.  Returns False on timeout."""
        try:
            msg = self._consumer.receive(timeout_millis=timeout_millis)
        except (BusTimeout, BusClosed):
            return False
        raw = msg.value()
        try:
            env = parse_envelope(raw)
            flow_id, cr = env["routing_key"], env["compute_resource"]
            with self._lock:
                ok = bool(self._params.get(flow_id, {}).get("notify_status", {}).get(cr, False))
            if ok:
                self._shelf_room.put_on_shelf(flow_id, str(env["message"]).encode("utf-8"))
                self.shelved += 1
            else:
                self.discarded += 1
                logger.error(task_id="", system_name="Deviceflow", module_name="Sorter",
                             message=f"[sort]: message of flow_id={flow_id} compute_resource={cr} discarded")
        except Exception as e:  # noqa: BLE001 — malformed envelope / missing shelf: drop, never wedge the sorter
            self.discarded += 1
            logger.error(task_id="", system_name="Deviceflow", module_name="Sorter", message=f"[sort]: bad message: {e}")
        self._consumer.acknowledge(msg)
        return True

    def sort(self) -> None:
        while not self._stop.is_set():
            self.sort_once()

    def drain(self, max_messages: int = 1 << 30) -> int:
        n = 0
        while n < max_messages and self.sort_once(timeout_millis=0):
            n += 1
        return n


class Dispatcher:
    def __init__(self, task_id: str, flow_id: str, shelf_room: ShelfRoom, strategy: str,
                 outbound_service: Dict[str, Dict[str, Any]], is_finished: bool = False,
                 clock: Optional[Clock] = None, rng: Optional[random.Random] = None,
                 default_bus: Optional[InMemoryBus] = None, now: Optional[Callable] = None,
                 timeout_millis: int = TIMEOUT_MILLIS, dashboard: Optional[Any] = None):
        self._task_id, self._flow_id, self._shelf_room = task_id, flow_id, shelf_room
        self._dashboard = dashboard
        self._strategy, self._outbound_service = strategy, outbound_service
        self._is_finished = is_finished
        self._consumer: Consumer = shelf_room.get_consumer(flow_id)
        self._producer = None
        self._release_event, self._stop_event = threading.Event(), threading.Event()
        self._clock = clock or Clock()
        self._rng = rng or random.Random()
        self._default_bus, self._now, self._timeout = default_bus, now, timeout_millis
        self.forwarded = 0
        self.dropped = 0
        self.discarded = 0
        self.schedule = None

    # ---- control -----------------------------------------------------------------------------------
    def stop_dispatch(self) -> None:
        self._stop_event.set()

    def release_dispatch(self) -> None:
        self._release_event.set()

    @property
    def is_finished(self) -> bool:
        return self._is_finished

    def dispatch(self) -> None:
        self._producer = make_outbound_producer(self._outbound_service, self._default_bus)
        if Strategy.check_real_time_dispatch(self._strategy):
            self.dispatch_real_time_with_strategy()
        else:
            self.dispatch_flow_with_strategy()

    # ---- helpers -----------------------------------------------------------------------------------
    def _send(self, message: bytes) -> None:
        if self._producer is not None:
            self._producer.send(message)
        self.forwarded += 1

    def _forward_batch(self, batch: List[bytes], p: float) -> None:
        for m in batch:
            if p == 0:
                self._send(m)
            elif 0 < p < 1:
                if self._rng.random() > p:
                    self._send(m)
                else:
                    self.dropped += 1
            else:
                self.dropped += 1

    # ---- real-time mode (dispatcher.py:84-171) ---------------------------------------------------------
    def dispatch_real_time_with_strategy(self) -> None:
        sizes, p = Strategy.real_time_strategy_analysis(self._strategy)
        logger.info(self._task_id, "Deviceflow", "Dispatcher", f"[dispatch_real_time]: batch sizes {sizes} drop probability {p}")
        if not sizes:
            return
        k = 0
        batch: List[bytes] = []
        while not self._stop_event.is_set():
            try:
                msg = self._consumer.receive(timeout_millis=self._timeout)
                batch.append(msg.value())
                self._consumer.acknowledge(msg)
                if len(batch) == sizes[k % len(sizes)]:
                    self._forward_batch(batch, p)
                    batch = []
                    k += 1
            except BusTimeout:
                if self._release_event.is_set():
                    self._forward_batch(batch, p)      # flush the partial batch
                    self._is_finished = True
                    self.stop_dispatch()
                    break
                if getattr(self._consumer, "_bus", None) is not None and self._consumer._bus.virtual_time:
                    break                              # virtual time: nothing more will arrive during this call
            except Exception:  # noqa: BLE001
                break

    # ---- flow mode (dispatcher.py:174-252) ---------------------------------------------------------------
    def dispatch_flow_with_strategy(self) -> None:
        timing, amount, drops = Strategy.flow_strategy_analysis(self._strategy, self._flow_id, rng=self._rng, now=self._now)
        self.schedule = (timing, amount, drops)
        logger.info(self._task_id, "Deviceflow", "Dispatcher",
                    f"[dispatch_flow]: timing {timing} amount {amount}")
        if len(timing) != len(amount) or not timing:
            self.clean_remain_message()
            return
        if self._dashboard is not None:
            # display tables of the dispatch curve (dispatcher.py:254-395); failures never stop the dispatch
            try:
                self._dashboard.write_dispatch_curves(self._task_id, timing, amount, self._strategy)
            except Exception as e:  # noqa: BLE001
                logger.warning(self._task_id, "Deviceflow", "Dispatcher", f"[dashboard]: writing the curve tables failed: {e}")
        if timing[0] > 0:
            self._clock.sleep(timing[0])
        with concurrent.futures.ThreadPoolExecutor(max_workers=1) as ex:   # slots stay ordered
            for i in range(len(timing)):
                drop_list = drops[i] if i < len(drops) else []
                ex.submit(self.send_with_time_and_amount, amount[i], drop_list)
                if i < len(timing) - 1:
                    self._clock.sleep(timing[i + 1])
        self.clean_remain_message()

    def send_with_time_and_amount(self, send_amount: int, drop_list: List[int]) -> None:
        drop = set(drop_list)
        for i in range(send_amount):
            if self._stop_event.is_set():
                break
            try:
                msg = self._consumer.receive(timeout_millis=self._timeout)
            except (BusTimeout, BusClosed):
                break
            if i in drop:
                self._consumer.acknowledge(msg)
                self.dropped += 1
                continue
            self._send(msg.value())
            self._consumer.acknowledge(msg)

    def clean_remain_message(self) -> None:
        """Everything still on the shelf after the schedule is consumed and discarded — late clients are lost."""
        while not self._stop_event.is_set():
            try:
                msg = self._consumer.receive(timeout_millis=self._timeout)
                self._consumer.acknowledge(msg)
                self.discarded += 1
            except (BusTimeout, BusClosed):
                if self._release_event.is_set():
                    self._is_finished = True
                break


class TaskOrientedDeviceFlowRegistry:
    """task_id -> {"total_compute_resources": [...]}; persisted as one row per task (registry.py:27-89)."""

    def __init__(self, repo: FlowTableRepo):
        self._repo = repo
        self._task_registry: Dict[str, Dict[str, Any]] = {}

    def register_task(self, task_id: str, total_compute_resources: List[str]) -> bool:
        entry = {"total_compute_resources": list(total_compute_resources)}
        rows = self._repo.get_values_by_conditions("task_id", task_id=task_id)
        payload = json.dumps(entry)
        if not rows:
            ok = self._repo.add_item({"task_id": task_id, "task_registry": payload})
        else:
            ok = self._repo.set_item_value("task_id", task_id, "task_registry", payload)
        if ok:
            self._task_registry[task_id] = entry
        return ok

    def unregister_task(self, task_id: str) -> bool:
        self._task_registry.pop(task_id, None)
        return True

    def get(self, task_id: str) -> Optional[Dict[str, Any]]:
        return self._task_registry.get(task_id)

    def save_state(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump(self._task_registry, f)

    def load_state(self, path: str) -> None:
        with open(path) as f:
            self._task_registry = json.load(f)


class DeviceFlow:
    """Flow bookkeeping: merges the NotifyStart calls of the logical and the device side, checks consistency,
    persists ``flow_params`` (deviceflow.py:29-197)."""

    def __init__(self, bus: InMemoryBus, repo: FlowTableRepo, inbound_url: str = "mem://local",
                 inbound_topic: str = "deviceflow_inbound", shelfroom_url: str = "mem://local",
                 shelfroom_namespace: str = "shelf_room"):
        self.sql_table = repo
        self._inbound_room = InboundRoom(bus, inbound_url, inbound_topic)
        self._shelf_room = ShelfRoom(bus, shelfroom_url, shelfroom_namespace)

    def notify_start(self, flow: Dict[str, Dict[str, Any]], task_id: str, flow_id: str, strategy: str,
                     compute_resource: str, outbound_service: Dict[str, Dict[str, Any]]):
        outbound = {k: dict(v) for k, v in (outbound_service or {}).items() if v}
        if flow_id not in flow:
            params = {"isFinished": False, "to_sort": False, "to_dispatch": False, "task_id": task_id,
                      "flow_id": flow_id, "outbound_service": outbound, "strategy": strategy,
                      "notify_start_called": {}, "notify_complete_called": {}}
            if not self.update_flow_id(flow_id, task_id):
                return False, {}
        else:
            params = flow[flow_id]
            if task_id != params["task_id"]:
                logger.error(task_id, "Deviceflow", "deviceflow", f"[notify_start]: task_id mismatch for flow_id={flow_id}")
                return False, {}
            if strategy != params["strategy"]:
                logger.error(task_id, "Deviceflow", "deviceflow", f"[notify_start]: strategy mismatch for flow_id={flow_id}")
                return False, {}
            for kind, val in outbound.items():
                have = params["outbound_service"].get(kind)
                if have:
                    if have != val:
                        logger.error(task_id, "Deviceflow", "deviceflow", f"[notify_start]: outbound {kind} mismatch for flow_id={flow_id}")
                        return False, {}
                else:
                    params["outbound_service"][kind] = val
        params["notify_start_called"][compute_resource] = True
        if not self.update_flow(flow_id, task_id, params):
            return False, {}
        return True, params

    def notify_complete(self, flow: Dict[str, Dict[str, Any]], task_id: str, flow_id: str, compute_resource: str):
        if flow_id not in flow:
            return False, {}
        params = flow[flow_id]
        if task_id != params["task_id"]:
            return False, {}
        params["notify_complete_called"][compute_resource] = True
        if not self.update_flow(flow_id, task_id, params):
            return False, {}
        return True, params

    @staticmethod
    def check_all_notify_start(task_registry: Dict[str, Any], params: Dict[str, Any]) -> bool:
        return (len(task_registry["total_compute_resources"]) == len(params["notify_start_called"])
                and all(params["notify_start_called"].values()))

    @staticmethod
    def check_all_notify_complete(task_registry: Dict[str, Any], params: Dict[str, Any]) -> bool:
        return (len(task_registry["total_compute_resources"]) == len(params["notify_complete_called"])
                and all(params["notify_complete_called"].values()))

    def update_flow(self, flow_id: str, task_id: str, params: Dict[str, Any]) -> bool:
        return self.sql_table.set_item_value("flow_id", flow_id, "flow", json.dumps(params))

    def update_flow_id(self, flow_id: str, task_id: str) -> bool:
        rows = self.sql_table.get_values_by_conditions("task_id", flow_id=flow_id, task_id=task_id)
        if len(rows) == 0:
            return self.sql_table.add_item({"task_id": task_id, "flow_id": flow_id})
        return len(rows) == 1
