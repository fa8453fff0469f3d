"""Persistent tables (SQLite) behind the control plane.

Same table and column names as the reference's MySQL tables (used from
ols_core/taskMgr/utils/utils.py:29-267 ``TaskTableRepo``, ols_core/resourceMgr/utils/utils.py ``ResTableRepo``,
ols_core/utils/repo_utils.py:19-345 ``SqlDataBase``), same method names, but: one process-wide connection guarded
by a lock (the reference opens a fresh ``engine.execute`` per statement and has no locking), parametrised SQL only,
``:memory:`` for tests and a file path for restart recovery.
"""
from __future__ import annotations

import itertools
import json
import sqlite3
import threading
import time
from typing import Any, Dict, Iterable, List, Optional, Sequence

TASK_COLUMNS = ["task_id", "task_status", "task_params", "total_simulation", "in_queue_time", "resource_occupied",
                "job_id", "submit_task_time", "task_finished_time", "logical_target", "logical_result",
                "logical_round", "logical_operator", "device_target", "device_result", "device_round",
                "device_operator", "user_id"]
RES_COLUMNS = ["task_id", "cpu", "mem", "gpu", "status", "freeze_time", "release_time", "detail"]
FLOW_COLUMNS = ["task_id", "flow_id", "task_registry", "flow"]
LOG_COLUMNS = ["task_id", "other_number", "system_name", "module_name", "message", "log_type", "ts"]
_mem_counter = itertools.count()
_INT_COLS = {"resource_occupied", "logical_round", "device_round", "status"}
_REAL_COLS = {"cpu", "mem", "gpu", "ts"}


class SqlDataBase:
    """Generic single-table repo addressed by an arbitrary ``identify_name`` column."""

    _shared: Dict[str, "tuple[sqlite3.Connection, threading.RLock]"] = {}
    _shared_guard = threading.Lock()

    def __init__(self, path: str = ":memory:", table: str = "table", columns: Sequence[str] = (),
                 shared_key: Optional[str] = None):
        self.path, self.table, self.columns = path, table, list(columns)
        key = shared_key or (path if path != ":memory:" else f":memory:{next(_mem_counter)}")
        with SqlDataBase._shared_guard:
            if key not in SqlDataBase._shared:
                conn = sqlite3.connect(path, check_same_thread=False, isolation_level=None)
                conn.execute("PRAGMA journal_mode=WAL") if path != ":memory:" else None
                SqlDataBase._shared[key] = (conn, threading.RLock())
            self._conn, self._lock = SqlDataBase._shared[key]
        self._key = key
        self._create()

    def _coltype(self, c: str) -> str:
        return "INTEGER" if c in _INT_COLS else ("REAL" if c in _REAL_COLS else "TEXT")

    def _create(self) -> None:
        cols = ", ".join(f'"{c}" {self._coltype(c)}' for c in self.columns)
        with self._lock:
            self._conn.execute(f'CREATE TABLE IF NOT EXISTS "{self.table}" (_rowid INTEGER PRIMARY KEY AUTOINCREMENT, {cols})')

    def _check(self, *cols: str) -> None:
        for c in cols:
            if c not in self.columns:
                raise KeyError(f"unknown column {c!r} for table {self.table}")

    # ---- reference-compatible API ------------------------------------------------------------------
    def add_item(self, item: Dict[str, Any], log_params: Optional[dict] = None) -> bool:
        """item maps column -> value (or single-element list, as the reference passes pandas-style columns)."""
        flat = {k: (v[0] if isinstance(v, (list, tuple)) and len(v) == 1 else v) for k, v in item.items()}
        self._check(*flat.keys())
        cols = ", ".join(f'"{c}"' for c in flat)
        ph = ", ".join("?" for _ in flat)
        with self._lock:
            self._conn.execute(f'INSERT INTO "{self.table}" ({cols}) VALUES ({ph})', list(flat.values()))
        return True

    def set_item_value(self, identify_name: str, identify_value: Any, item: str, value: Any,
                       log_params: Optional[dict] = None) -> bool:
        self._check(identify_name, item)
        with self._lock:
            cur = self._conn.execute(f'UPDATE "{self.table}" SET "{item}"=? WHERE "{identify_name}"=?',
                                     (value, identify_value))
            return cur.rowcount > 0

    def set_item_values(self, identify_name: str, identify_value: Any, values: Dict[str, Any]) -> bool:
        self._check(identify_name, *values.keys())
        sets = ", ".join(f'"{c}"=?' for c in values)
        with self._lock:
            cur = self._conn.execute(f'UPDATE "{self.table}" SET {sets} WHERE "{identify_name}"=?',
                                     [*values.values(), identify_value])
            return cur.rowcount > 0

    def get_item_value(self, identify_name: str, identify_value: Any, item: str) -> Any:
        self._check(identify_name, item)
        with self._lock:
            row = self._conn.execute(f'SELECT "{item}" FROM "{self.table}" WHERE "{identify_name}"=? LIMIT 1',
                                     (identify_value,)).fetchone()
        return None if row is None else row[0]

    def check_item_in_database(self, identify_name: str, identify_value: Any) -> bool:
        self._check(identify_name)
        with self._lock:
            row = self._conn.execute(f'SELECT 1 FROM "{self.table}" WHERE "{identify_name}"=? LIMIT 1',
                                     (identify_value,)).fetchone()
        return row is not None

    def get_values_by_conditions(self, item: str, **conditions: Any) -> List[Any]:
        self._check(item, *conditions.keys())
        where = " AND ".join(f'"{k}"=?' for k in conditions) or "1=1"
        with self._lock:
            rows = self._conn.execute(f'SELECT "{item}" FROM "{self.table}" WHERE {where}',
                                      list(conditions.values())).fetchall()
        return [r[0] for r in rows]

    def get_data_by_item_values(self, item: str, values: Iterable[Any]) -> List[Dict[str, Any]]:
        self._check(item)
        vals = list(values)
        if not vals:
            return []
        ph = ", ".join("?" for _ in vals)
        with self._lock:
            cur = self._conn.execute(f'SELECT * FROM "{self.table}" WHERE "{item}" IN ({ph}) ORDER BY _rowid', vals)
            names = [d[0] for d in cur.description]
            return [dict(zip(names, r)) for r in cur.fetchall()]

    def get_all_items_not_None(self, item: str) -> List[Dict[str, Any]]:
        self._check(item)
        with self._lock:
            cur = self._conn.execute(f'SELECT * FROM "{self.table}" WHERE "{item}" IS NOT NULL ORDER BY _rowid')
            names = [d[0] for d in cur.description]
            return [dict(zip(names, r)) for r in cur.fetchall()]

    def get_unique_item(self, item: str) -> List[Any]:
        self._check(item)
        with self._lock:
            return [r[0] for r in self._conn.execute(f'SELECT DISTINCT "{item}" FROM "{self.table}"').fetchall()]

    def all_rows(self) -> List[Dict[str, Any]]:
        with self._lock:
            cur = self._conn.execute(f'SELECT * FROM "{self.table}" ORDER BY _rowid')
            names = [d[0] for d in cur.description]
            return [dict(zip(names, r)) for r in cur.fetchall()]

    def delete_item(self, identify_name: str, identify_value: Any, log_params: Optional[dict] = None) -> bool:
        self._check(identify_name)
        with self._lock:
            cur = self._conn.execute(f'DELETE FROM "{self.table}" WHERE "{identify_name}"=?', (identify_value,))
            return cur.rowcount > 0

    def lock(self) -> threading.RLock:
        """Expose the table lock so callers can make check-then-insert sequences atomic
        (the reference's ResourceManager is racy here: resource_manager.py:151-177)."""
        return self._lock


class TaskTableRepo(SqlDataBase):
    """taskmgr_table — one row per task (ols_core/taskMgr/utils/utils.py:29-267)."""

    def __init__(self, path: str = ":memory:", shared_key: Optional[str] = None):
        super().__init__(path, "taskmgr_table", TASK_COLUMNS, shared_key)

    def check_task_in_database(self, task_id: str) -> bool:
        return self.check_item_in_database("task_id", task_id)

    def get_task_value(self, task_id: str, item: str) -> Any:
        return self.get_item_value("task_id", task_id, item)

    def set_task_value(self, task_id: str, item: str, value: Any) -> bool:
        return self.set_item_value("task_id", task_id, item, value)

    def create_task(self, task_id: str, user_id: str = "", status: str = "UNDONE") -> bool:
        """The GUI/back-end of the reference pre-creates an UNDONE row before submitTask (task_manager.py:204-215)."""
        if self.check_task_in_database(task_id):
            return False
        return self.add_item({"task_id": task_id, "task_status": status, "user_id": user_id, "resource_occupied": 0})


class ResTableRepo(SqlDataBase):
    """resmgr_table — resource ledger rows, status 0 = held (ols_core/resourceMgr/utils/utils.py)."""

    def __init__(self, path: str = ":memory:", shared_key: Optional[str] = None):
        super().__init__(path, "resmgr_table", RES_COLUMNS, shared_key)

    def insert_data(self, task_id: str, cpu: float, mem: float, gpu: float = 0.0, detail: Optional[dict] = None) -> bool:
        return self.add_item({"task_id": task_id, "cpu": cpu, "mem": mem, "gpu": gpu, "status": 0,
                              "freeze_time": time.strftime("%Y-%m-%d %H:%M:%S"), "detail": json.dumps(detail or {})})

    def delete_by_taskid(self, task_id: str) -> bool:
        return self.delete_item("task_id", task_id)

    def get_jsons_by_item_values(self, item: str, values: Iterable[Any]) -> List[Dict[str, Any]]:
        return self.get_data_by_item_values(item, values)

    def held(self) -> List[Dict[str, Any]]:
        return self.get_data_by_item_values("status", [0])


class FlowTableRepo(SqlDataBase):
    """deviceflow_table — registry + flow state for restart recovery (deviceflow_server.py:83-164)."""

    def __init__(self, path: str = ":memory:", shared_key: Optional[str] = None):
        super().__init__(path, "deviceflow_table", FLOW_COLUMNS, shared_key)


class LogTableRepo(SqlDataBase):
    def __init__(self, path: str = ":memory:", shared_key: Optional[str] = None):
        super().__init__(path, "log_table", LOG_COLUMNS, shared_key)
