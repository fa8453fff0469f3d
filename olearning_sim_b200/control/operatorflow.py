"""OperatorFlow — per-round start / stop gates of a task.

Counterpart of ols_core/taskMgr/utils/operatorflow.py:26-351.  Strategies:

* ``""``                               no-op
* ``waiting_for_global_aggregation``   (:135-237) poll an external *selection service* for the current round index;
                                       ``stop`` requires ``current_round - previous_round == 1``
* ``sample_and_aggregation``           (:240-291) per round sample N client files (default 7000) into a round folder
                                       by symlink, then rendezvous with an external aggregator through marker files
                                       ``simulation_finished.txt`` / ``aggregation_finished.txt``
* ``sample_dc_and_aggregation``        (:294-352) same with bundles (default 500 bundles x 10 files)

The reference hard-codes site paths, sample sizes and a websocket URL; here they are parameters (``gate_params``)
with the reference's values as defaults, the selection service is a pluggable callable, the clock is injectable.
"""
from __future__ import annotations

import os
import random
import time
from typing import Any, Callable, Dict, Optional

from ..utils.logger import Logger

logger = Logger()


class GateTimeout(RuntimeError):
    pass


class OperatorFlowStrategy:
    name = ""

    def __init__(self, task_id: str, params: Dict[str, Any], gate_params: Dict[str, Any], sleep=time.sleep,
                 rng: Optional[random.Random] = None):
        self.task_id = task_id
        self.wait_interval = int(params.get("wait_interval", 0) or 0)
        self.total_timeout = int(params.get("total_timeout", 0) or 0)
        self.gp = gate_params
        self._sleep = sleep
        self._rng = rng or random.Random()

    def _poll(self, pred: Callable[[], bool]) -> bool:
        """wait_interval / total_timeout polling (operatorflow.py:139-156, 284-291)"""
        waited = 0
        while True:
            if pred():
                return True
            if waited >= self.total_timeout:
                return False
            step = max(self.wait_interval, 1)
            self._sleep(step)
            waited += step

    def start(self, round_idx: int) -> bool:
        return True

    def stop(self, round_idx: int) -> bool:
        return True


class NoGate(OperatorFlowStrategy):
    name = ""


class WaitingForGlobalAggregation(OperatorFlowStrategy):
    name = "waiting_for_global_aggregation"
    # task_id -> callable returning the selection service's current round (tests / in-process aggregator)
    services: Dict[str, Callable[[], int]] = {}

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._previous_round: Optional[int] = None

    def _ask_websocket(self, url: str) -> Optional[int]:
        """The reference's selection-service exchange (operatorflow.py:186-237): connect to ``service_url``, send one
        REGISTER_MESSAGE for the task, read the first answer and take ``data.round_idx`` from it."""
        import json
        import uuid
        from websockets.sync.client import connect
        cid, mid = str(uuid.uuid4()), str(uuid.uuid4())
        msg = {"header": {"messageId": mid, "messageType": "REGISTER_MESSAGE", "clientId": cid},
               "data": {"message_id": mid, "client_id": cid, "task_id": self.task_id}}
        timeout = float(self.gp.get("service_timeout", 5.0))
        with connect(url, open_timeout=timeout) as ws:
            ws.send(json.dumps(msg))
            answer = json.loads(ws.recv(timeout=timeout))
        r = (answer.get("data") or {}).get("round_idx")
        return None if r is None else int(r)

    def _current_round(self) -> Optional[int]:
        """``selection_service``: a callable (in-process aggregator / tests) or the ``ws://`` URL of the selection service
        (``service_url`` of the reference's selection_config.yaml); also looked up per task in ``services``."""
        svc = self.gp.get("selection_service") or self.gp.get("service_url") or self.services.get(self.task_id)
        if svc is None:
            return None
        try:
            if isinstance(svc, str):
                return self._ask_websocket(svc)
            return int(svc())
        except Exception as e:  # noqa: BLE001
            logger.error(self.task_id, "TaskMgr", "operatorflow", f"[selection service]: {e}")
            return None

    def start(self, round_idx: int) -> bool:
        ok = self._poll(lambda: self._current_round() is not None)
        if ok:
            self._previous_round = self._current_round()
        return ok

    def stop(self, round_idx: int) -> bool:
        def advanced():
            cur = self._current_round()
            return cur is not None and self._previous_round is not None and cur - self._previous_round == 1
        return self._poll(advanced)


class SampleAndAggregation(OperatorFlowStrategy):
    name = "sample_and_aggregation"
    default_sample = 7000

    def _dirs(self, round_idx: int):
        root = self.gp.get("root", f"/tmp/olsb_gate/{self.task_id}")
        return (self.gp.get("source_dir", os.path.join(root, "all_clients")),
                os.path.join(root, f"round_{round_idx}"))

    def _sample(self, src: str, dst: str, n: int) -> int:
        os.makedirs(dst, exist_ok=True)
        files = sorted(os.listdir(src)) if os.path.isdir(src) else []
        picked = files if len(files) <= n else self._rng.sample(files, n)
        for f in picked:
            link = os.path.join(dst, f)
            if not os.path.lexists(link):
                os.symlink(os.path.join(src, f), link)
        return len(picked)

    def start(self, round_idx: int) -> bool:
        src, dst = self._dirs(round_idx)
        n = self._sample(src, dst, int(self.gp.get("sample_num", self.default_sample)))
        logger.info(self.task_id, "TaskMgr", "operatorflow", f"[{self.name}.start]: round {round_idx} sampled {n} client files")
        for marker in ("simulation_finished.txt", "aggregation_finished.txt"):
            p = os.path.join(dst, marker)
            if os.path.exists(p):
                os.remove(p)
        return True

    def stop(self, round_idx: int) -> bool:
        _, dst = self._dirs(round_idx)
        os.makedirs(dst, exist_ok=True)
        with open(os.path.join(dst, "simulation_finished.txt"), "w") as f:
            f.write(str(round_idx))
        if self.gp.get("aggregator") is not None:        # in-process aggregator hook instead of an external service
            self.gp["aggregator"](round_idx, dst)
        return self._poll(lambda: os.path.exists(os.path.join(dst, "aggregation_finished.txt")))


class SampleDcAndAggregation(SampleAndAggregation):
    name = "sample_dc_and_aggregation"

    def _sample(self, src: str, dst: str, n: int) -> int:
        bundles = int(self.gp.get("bundle_num", 500))
        per = int(self.gp.get("files_per_bundle", 10))
        os.makedirs(dst, exist_ok=True)
        files = sorted(os.listdir(src)) if os.path.isdir(src) else []
        total = 0
        for b in range(bundles):
            if not files:
                break
            picked = files if len(files) <= per else self._rng.sample(files, per)
            bdir = os.path.join(dst, f"bundle_{b}")
            os.makedirs(bdir, exist_ok=True)
            for f in picked:
                link = os.path.join(bdir, f)
                if not os.path.lexists(link):
                    os.symlink(os.path.join(src, f), link)
            total += len(picked)
        return total


class OperatorFlowStrategyFactory:
    _registry = {c.name: c for c in (NoGate, WaitingForGlobalAggregation, SampleAndAggregation, SampleDcAndAggregation)}

    @classmethod
    def register(cls, strategy_cls) -> None:
        cls._registry[strategy_cls.name] = strategy_cls

    @classmethod
    def create(cls, name: str, *a, **k) -> OperatorFlowStrategy:
        if name not in cls._registry:
            raise ValueError(f"unknown operatorflow strategy {name!r}")
        return cls._registry[name](*a, **k)


class OperatorFlow:
    """``start()`` / ``stop()`` are called once per round by the round loop (run_task.py:221,315-322)."""

    def __init__(self, task_id: str, rounds: int, start_params: Dict[str, Any], stop_params: Dict[str, Any],
                 gate_params: Optional[Dict[str, Any]] = None, sleep=time.sleep, rng: Optional[random.Random] = None):
        gp = gate_params or {}
        self.task_id, self.rounds = task_id, rounds
        self._start = OperatorFlowStrategyFactory.create(start_params.get("strategy", "") or "", task_id, start_params, gp, sleep, rng)
        same = (start_params.get("strategy", "") or "") == (stop_params.get("strategy", "") or "")
        self._stop = self._start if same else OperatorFlowStrategyFactory.create(
            stop_params.get("strategy", "") or "", task_id, stop_params, gp, sleep, rng)
        if same:
            self._stop.wait_interval = int(stop_params.get("wait_interval", self._stop.wait_interval) or 0)
            self._stop_timeout = int(stop_params.get("total_timeout", self._stop.total_timeout) or 0)
        else:
            self._stop_timeout = self._stop.total_timeout

    def start(self, round_idx: int) -> bool:
        return self._start.start(round_idx)

    def stop(self, round_idx: int) -> bool:
        self._stop.total_timeout = self._stop_timeout
        return self._stop.stop(round_idx)
