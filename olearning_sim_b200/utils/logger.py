"""Structured logger with the reference's 4-field call signature.

``Logger().info(task_id, system_name, module_name, message)`` — same signature as ols_core/simu_log.py:13-172, which
writes every call synchronously to a file *and* to MySQL ``log_table`` through ``pandas.to_sql`` and swallows all
failures.  Here the sinks are stderr, an optional JSONL file and an optional ``LogTableRepo``; table writes are
batched on a background thread so logging never blocks the round loop.
"""
from __future__ import annotations

import json
import queue
import sys
import threading
import time
from typing import Any, Optional

_LEVELS = {"INFO": 20, "WARNING": 30, "ERROR": 40}


class Logger:
    _instance: Optional["Logger"] = None
    _guard = threading.Lock()

    def __new__(cls, *a, **k):  # process-wide singleton, like the reference's module-level `logger`
        with cls._guard:
            if cls._instance is None:
                inst = super().__new__(cls)
                inst._init()
                cls._instance = inst
            return cls._instance

    def _init(self) -> None:
        self.stderr_level = 30
        self.jsonl_path: Optional[str] = None
        self.repo = None
        self._q: "queue.Queue[Optional[dict]]" = queue.Queue(maxsize=65536)
        self._thread: Optional[threading.Thread] = None
        self._file_lock = threading.Lock()
        self.records_dropped = 0
        self.ring: list = []           # last records, for tests / status RPCs
        self.ring_size = 512

    # ---- configuration ---------------------------------------------------------------------------
    def set_logger(self, log_yaml: Optional[str] = None, jsonl_path: Optional[str] = None, repo: Any = None,
                   stderr_level: Optional[str] = None) -> None:
        """``log_yaml`` is accepted for signature parity (the reference reads a MySQL DSN from it) and ignored."""
        if jsonl_path is not None:
            self.jsonl_path = jsonl_path
        if repo is not None:
            self.repo = repo
            self._ensure_thread()
        if stderr_level is not None:
            self.stderr_level = _LEVELS.get(stderr_level.upper(), 30)

    def _ensure_thread(self) -> None:
        if self._thread is None or not self._thread.is_alive():
            self._thread = threading.Thread(target=self._drain, name="olsb-logger", daemon=True)
            self._thread.start()

    # ---- API -------------------------------------------------------------------------------------
    def info(self, task_id: str = "", system_name: str = "", module_name: str = "", message: str = "") -> None:
        self._emit("INFO", task_id, system_name, module_name, message)

    def warning(self, task_id: str = "", system_name: str = "", module_name: str = "", message: str = "") -> None:
        self._emit("WARNING", task_id, system_name, module_name, message)

    def error(self, task_id: str = "", system_name: str = "", module_name: str = "", message: str = "") -> None:
        self._emit("ERROR", task_id, system_name, module_name, message)

    def _emit(self, level: str, task_id: str, system_name: str, module_name: str, message: str) -> None:
        rec = {"ts": time.time(), "log_type": level, "task_id": task_id, "system_name": system_name,
               "module_name": module_name, "message": str(message), "other_number": ""}
        self.ring.append(rec)
        if len(self.ring) > self.ring_size:
            del self.ring[: len(self.ring) - self.ring_size]
        if _LEVELS[level] >= self.stderr_level:
            print(f"[{level}] {system_name}/{module_name} task={task_id}: {message}", file=sys.stderr)
        if self.jsonl_path:
            with self._file_lock:
                with open(self.jsonl_path, "a") as f:
                    f.write(json.dumps(rec) + "\n")
        if self.repo is not None:
            try:
                self._q.put_nowait(rec)
            except queue.Full:
                self.records_dropped += 1

    def _drain(self) -> None:
        while True:
            rec = self._q.get()
            if rec is None:
                return
            batch = [rec]
            try:
                while len(batch) < 256:
                    nxt = self._q.get_nowait()
                    if nxt is None:
                        break
                    batch.append(nxt)
            except queue.Empty:
                pass
            repo = self.repo
            if repo is None:
                continue
            for r in batch:
                try:
                    repo.add_item(r)
                except Exception:  # noqa: BLE001 — logging must never take the control plane down
                    self.records_dropped += 1

    def flush(self, timeout: float = 2.0) -> None:
        t0 = time.time()
        while not self._q.empty() and time.time() - t0 < timeout:
            time.sleep(0.01)
