"""Prometheus metrics of the control plane and the round loop.

The reference has no service metrics of its own (only the Pulsar helm chart brings Prometheus / Grafana,
for the broker; progress lives in DB cells).  Here the PerformanceManager's per-round records are also exported as
Prometheus series, so the usual scrape -> dashboard path works for the simulator itself:

    olsb_tasks{status=...}                gauge   tasks per status as seen by the TaskManager
    olsb_rounds_total{task_id}            counter completed rounds
    olsb_client_steps_total{task_id}      counter simulated client local steps
    olsb_round_seconds{task_id}           gauge   wall time of the last round
    olsb_phase_ms{task_id, phase}         gauge   broadcast / local_steps / reduce_update of the last profiled round

``prometheus_client`` is optional: without it every call is a no-op and ``start_http_server`` returns ``None``.
"""
from __future__ import annotations

import threading
from typing import Any, Dict, Optional

try:  # optional dependency
    import prometheus_client as _pc
except ImportError:  # pragma: no cover - the image ships it, deployments may not
    _pc = None


class Metrics:
    def __init__(self, registry: Optional[Any] = None):
        self.enabled = _pc is not None
        self._lock = threading.Lock()
        if not self.enabled:
            return
        self.registry = registry or _pc.CollectorRegistry()
        self.tasks = _pc.Gauge("olsb_tasks", "tasks per status", ["status"], registry=self.registry)
        self.rounds = _pc.Counter("olsb_rounds_total", "completed rounds", ["task_id"], registry=self.registry)
        self.client_steps = _pc.Counter("olsb_client_steps_total", "simulated client local steps", ["task_id"],
                                        registry=self.registry)
        self.round_seconds = _pc.Gauge("olsb_round_seconds", "wall time of the last round", ["task_id"], registry=self.registry)
        self.phase_ms = _pc.Gauge("olsb_phase_ms", "phase time of the last profiled round", ["task_id", "phase"],
                                  registry=self.registry)

    # ---- feeders -------------------------------------------------------------------------------------------------
    def observe_round(self, rec: Dict[str, Any]) -> None:
        """``rec`` = a PerformanceManager.commitJob record ({task_id, t_round_ms, client_steps, phases_ms})."""
        if not self.enabled:
            return
        tid = str(rec.get("task_id", ""))
        self.rounds.labels(tid).inc()
        self.client_steps.labels(tid).inc(float(rec.get("client_steps", 0)))
        self.round_seconds.labels(tid).set(float(rec.get("t_round_ms", 0.0)) / 1e3)
        for phase, ms in (rec.get("phases_ms") or {}).items():
            self.phase_ms.labels(tid, phase).set(float(ms))

    def observe_tasks(self, status_counts: Dict[str, int]) -> None:
        if not self.enabled:
            return
        for status, n in status_counts.items():
            self.tasks.labels(status).set(n)

    # ---- exposition -----------------------------------------------------------------------------------------------
    def render(self) -> str:
        return _pc.generate_latest(self.registry).decode() if self.enabled else ""

    def start_http_server(self, port: int, addr: str = "127.0.0.1"):
        if not self.enabled:
            return None
        return _pc.start_http_server(port, addr=addr, registry=self.registry)
