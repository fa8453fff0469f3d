"""Dashboard tables of the DeviceFlow dispatcher and the task runner.

The reference writes three display-only MySQL tables from "temporary" code with a hard-coded DSN
(ols_core/deviceflow/non_grpc/dispatcher.py:254-395 ``operation_amount_table`` / ``accumulated_amount_table``,
ols_core/taskMgr/task_runner.py:226-256 ``task_target``).  Here the same rows are produced by pure functions (testable
without a database) and stored through ``DashboardRepo`` (SQLite by default, any ``SqlDataBase``-like object works):

* ``operation_rows``   one row per dispatch slot: (cumulative time, amount).  Relative schedules get a leading (0, 0) row
                       when the first slot is not at t=0; absolute schedules are re-based so the first slot is t=0.
* ``accumulated_rows`` the cumulative-amount curve sampled every second, each slot's step limited to ``rate_cap``
                       messages per second (reference: 750, dispatcher.py:331) so a 5000-message burst shows as a ramp.
* ``task_target_row``  total simulated devices per class (high / middle / low).

Only flow strategies are shown; real-time (batch) dispatch has no curve, as in the reference.
"""
from __future__ import annotations

import json
from typing import Any, Dict, List, Optional, Sequence, Tuple

from ..control.repo import SqlDataBase

RATE_CAP = 750


def flow_time_type(strategy: Any) -> Optional[str]:
    """'relative' / 'absolute' for a flow strategy, None when the strategy has no curve to show."""
    st = json.loads(strategy) if isinstance(strategy, (str, bytes)) else dict(strategy)
    fd = st.get("flow_dispatch", {})
    if not fd.get("use_strategy", False):
        return None
    for key in ("specific_timing", "specific_interval"):
        if fd.get(key, {}).get("use", False):
            tt = fd[key].get("time_type", "")
            return tt if tt in ("relative", "absolute") else None
    return None


def cumulative_times(timing: Sequence[float], time_type: str) -> List[float]:
    out, acc = [], 0.0
    for i, dt in enumerate(timing):
        acc = 0.0 if (time_type == "absolute" and i == 0) else acc + dt
        out.append(acc)
    return out


def operation_rows(timing: Sequence[float], amount: Sequence[int], time_type: str) -> List[Tuple[float, int]]:
    rows: List[Tuple[float, int]] = []
    if not timing:
        return rows
    if time_type == "relative" and timing[0] > 0:
        rows.append((0, 0))
    rows += list(zip(cumulative_times(timing, time_type), [int(a) for a in amount]))
    return rows


def accumulated_rows(timing: Sequence[float], amount: Sequence[int], time_type: str, rate_cap: int = RATE_CAP) -> List[Tuple[int, int]]:
    ts = [int(t) for t in cumulative_times(timing, time_type)]
    if not ts:
        return []
    pts: List[Tuple[int, int]] = [(0, 0)]
    pts += [(k, 0) for k in range(1, ts[0])]                       # silence before the first slot
    done = 0
    last = len(ts) - 1
    for i, t in enumerate(ts):
        if i < last:
            seg = range(1 if (i == 0 and t == 0) else t, ts[i + 1])
        else:
            prev = ts[i - 1] if i > 0 else t
            seg = range(t, max(t + 2, t + (t - prev) // 2))      # short tail after the last slot
        for j, k in enumerate(seg):
            pts.append((k, min(done + int(amount[i]), done + rate_cap * (j + 1))))
        done += int(amount[i])
    return pts


def task_target_row(task_json: Dict[str, Any]) -> Dict[str, int]:
    tot = {"High": 0, "Middle": 0, "Low": 0}
    for d in task_json.get("target", {}).get("data", []):
        ts = d.get("total_simulation", {})
        for dev, n in zip(ts.get("devices", []), ts.get("nums", [])):
            if dev in tot:
                tot[dev] += int(n)
    return {k.lower(): v for k, v in tot.items() if v > 0}


class DashboardRepo:
    """operation_amount_table / accumulated_amount_table / task_target with replace-by-task_id writes."""

    def __init__(self, path: str = ":memory:", shared_key: Optional[str] = None):
        key = shared_key or (None if path != ":memory:" else f"dashboard:{id(self)}")
        self.operation = SqlDataBase(path, "operation_amount_table", ["task_id", "task_time", "amount"], key)
        self.accumulated = SqlDataBase(path, "accumulated_amount_table", ["task_id", "task_time", "amount"], key)
        self.target = SqlDataBase(path, "task_target", ["task_id", "high", "middle", "low"], key)

    def write_dispatch_curves(self, task_id: str, timing: Sequence[float], amount: Sequence[int], strategy: Any,
                              rate_cap: int = RATE_CAP) -> bool:
        tt = flow_time_type(strategy)
        if tt is None or not timing:
            return False
        for table, rows in ((self.operation, operation_rows(timing, amount, tt)),
                            (self.accumulated, accumulated_rows(timing, amount, tt, rate_cap))):
            table.delete_item("task_id", task_id)
            for t, a in rows:
                table.add_item({"task_id": task_id, "task_time": t, "amount": a})
        return True

    def write_task_target(self, task_json: Dict[str, Any]) -> bool:
        tid = task_json.get("task_id", "")
        self.target.delete_item("task_id", tid)
        return self.target.add_item({"task_id": tid, **task_target_row(task_json)})

    def curve(self, task_id: str, which: str = "accumulated") -> List[Tuple[float, int]]:
        t = self.accumulated if which == "accumulated" else self.operation
        return [(float(r["task_time"]), int(r["amount"])) for r in t.get_data_by_item_values("task_id", [task_id])]
