"""Small infrastructure pieces with reference counterparts:

* ``QueueRepo``       — list queue with ``rpush`` / ``lpop`` (ols_core/taskMgr/utils/utils_redis.py:16-47 ``RedisRepo``;
                        its use in TaskManager is commented out in the reference).  In-process or SQLite-durable.
* ``Fragment`` + ``JsonFragmentRepo`` / ``ProtoFragmentRepo`` — consumers that parse client updates ("fragments") off a
                        topic (ols_core/ofl_commons/infrastructure/FragmentRepo/*.py; their base class is missing from the
                        reference tree).  The "proto" flavour uses a self-describing length-prefixed binary frame.
* ``LogUtils``        — rotating file logger (ols_core/utils/log_utils.py:17-45).
* ``PerformanceManager`` — ``commitJob`` (ols_core/proto/performanceService.proto:4-6; implementation absent from the
                        reference): records per-round timing reports and serves roofline summaries.
"""
from __future__ import annotations

import collections
import json
import logging
import logging.handlers
import os
import sqlite3
import struct
import threading
from dataclasses import dataclass
from typing import Any, Deque, Dict, List, Optional


class QueueRepo:
    def __init__(self, path: Optional[str] = None):
        self._mem: Dict[str, Deque[str]] = collections.defaultdict(collections.deque)
        self._lock = threading.Lock()
        self._db = None
        if path:
            self._db = sqlite3.connect(path, check_same_thread=False, isolation_level=None)
            self._db.execute("CREATE TABLE IF NOT EXISTS q (id INTEGER PRIMARY KEY AUTOINCREMENT, name TEXT, value TEXT)")

    def rpush(self, name: str, value: str) -> int:
        with self._lock:
            if self._db:
                self._db.execute("INSERT INTO q (name, value) VALUES (?, ?)", (name, value))
                return self._db.execute("SELECT COUNT(*) FROM q WHERE name=?", (name,)).fetchone()[0]
            self._mem[name].append(value)
            return len(self._mem[name])

    def lpop(self, name: str) -> Optional[str]:
        with self._lock:
            if self._db:
                row = self._db.execute("SELECT id, value FROM q WHERE name=? ORDER BY id LIMIT 1", (name,)).fetchone()
                if row is None:
                    return None
                self._db.execute("DELETE FROM q WHERE id=?", (row[0],))
                return row[1]
            return self._mem[name].popleft() if self._mem[name] else None

    def llen(self, name: str) -> int:
        with self._lock:
            if self._db:
                return self._db.execute("SELECT COUNT(*) FROM q WHERE name=?", (name,)).fetchone()[0]
            return len(self._mem[name])


class RespClient:
    """Redis over its wire protocol (RESP2) — the handful of commands the queue repo needs (ols_core/utils/redis_repo.py
    wraps ``redis.StrictRedis`` rpush / lpop / llen); no ``redis`` package required.  One request/response at a time."""

    def __init__(self, host: str = "127.0.0.1", port: int = 6379, password: Optional[str] = None, db: int = 0,
                 timeout: float = 10.0):
        import socket
        self._s = socket.create_connection((host, int(port)), timeout=timeout)
        self._f = self._s.makefile("rb")
        self._lock = threading.Lock()
        if password:
            self.command("AUTH", password)
        if db:
            self.command("SELECT", str(db))

    def _read(self):
        line = self._f.readline()
        if not line:
            raise ConnectionError("redis: connection closed")
        t, rest = line[:1], line[1:-2]
        if t == b"+":
            return rest.decode()
        if t == b"-":
            raise RuntimeError("redis: " + rest.decode())
        if t == b":":
            return int(rest)
        if t == b"$":
            n = int(rest)
            if n < 0:
                return None
            data = self._f.read(n + 2)
            return data[:-2].decode("utf-8", "replace")
        if t == b"*":
            n = int(rest)
            return None if n < 0 else [self._read() for _ in range(n)]
        raise RuntimeError(f"redis: unexpected reply {line!r}")

    def command(self, *args: str):
        parts = [a if isinstance(a, bytes) else str(a).encode() for a in args]
        with self._lock:
            self._s.sendall(b"*%d\r\n" % len(parts) + b"".join(b"$%d\r\n%s\r\n" % (len(p), p) for p in parts))
            return self._read()

    def close(self) -> None:
        try:
            self._s.close()
        except OSError:
            pass


class RedisRepo(QueueRepo):
    """The queue repo on a Redis server: ``RedisRepo(url="redis://[:password@]host:port/db")``; without ``url`` it is the
    in-process / SQLite ``QueueRepo`` (what the single-box deployment uses)."""

    def __init__(self, path: Optional[str] = None, url: Optional[str] = None):
        self._r: Optional[RespClient] = None
        if url:
            import urllib.parse
            u = urllib.parse.urlparse(url)
            self._r = RespClient(u.hostname or "127.0.0.1", u.port or 6379, u.password, int((u.path or "/0").strip("/") or 0))
            self._lock = threading.Lock()
        else:
            super().__init__(path)

    def rpush(self, name: str, value: str) -> int:
        return int(self._r.command("RPUSH", name, value)) if self._r else super().rpush(name, value)

    def lpop(self, name: str) -> Optional[str]:
        return self._r.command("LPOP", name) if self._r else super().lpop(name)

    def llen(self, name: str) -> int:
        return int(self._r.command("LLEN", name)) if self._r else super().llen(name)


@dataclass
class Fragment:
    """One client update as seen by an aggregator."""
    task_id: str
    client_id: str
    round: int
    num_samples: int
    payload: bytes = b""
    meta: Optional[Dict[str, Any]] = None


class FragmentRepo:
    def __init__(self, bus: Any, topic: str, subscription: str = "fragment_repo"):
        self._consumer = bus.subscribe(topic, subscription)

    def decode(self, raw: bytes) -> Fragment:
        raise NotImplementedError

    def encode(self, frag: Fragment) -> bytes:
        raise NotImplementedError

    def consume(self, max_fragments: int = 10, timeout_millis: int = 0) -> List[Fragment]:
        from ..deviceflow.bus import BusTimeout
        out = []
        while len(out) < max_fragments:
            try:
                m = self._consumer.receive(timeout_millis=timeout_millis)
            except BusTimeout:
                break
            try:
                out.append(self.decode(m.value()))
            finally:
                self._consumer.acknowledge(m)
        return out


class JsonFragmentRepo(FragmentRepo):
    def encode(self, f: Fragment) -> bytes:
        return json.dumps({"task_id": f.task_id, "client_id": f.client_id, "round": f.round, "num_samples": f.num_samples,
                           "payload": f.payload.hex(), "meta": f.meta or {}}).encode()

    def decode(self, raw: bytes) -> Fragment:
        d = json.loads(raw)
        return Fragment(d["task_id"], d["client_id"], int(d["round"]), int(d["num_samples"]),
                        bytes.fromhex(d.get("payload", "")), d.get("meta") or {})


class ProtoFragmentRepo(FragmentRepo):
    """Binary frame: magic 'OLSF', u32 header length, JSON header, raw payload (no protoc dependency)."""
    MAGIC = b"OLSF"

    def encode(self, f: Fragment) -> bytes:
        hdr = json.dumps({"task_id": f.task_id, "client_id": f.client_id, "round": f.round,
                          "num_samples": f.num_samples, "meta": f.meta or {}}).encode()
        return self.MAGIC + struct.pack("<I", len(hdr)) + hdr + f.payload

    def decode(self, raw: bytes) -> Fragment:
        if raw[:4] != self.MAGIC:
            raise ValueError("not a fragment frame")
        (n,) = struct.unpack("<I", raw[4:8])
        d = json.loads(raw[8:8 + n])
        return Fragment(d["task_id"], d["client_id"], int(d["round"]), int(d["num_samples"]), raw[8 + n:], d.get("meta") or {})


class LogUtils:
    def __init__(self, log_dir: str = "log", name: str = "log_simulation", max_bytes: int = 10 * 1024 * 1024, backups: int = 5):
        os.makedirs(log_dir, exist_ok=True)
        self.logger = logging.getLogger(f"olsb.{name}")
        self.logger.setLevel(logging.INFO)
        if not self.logger.handlers:
            h = logging.handlers.RotatingFileHandler(os.path.join(log_dir, f"{name}.log"), maxBytes=max_bytes, backupCount=backups)
            h.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(message)s"))
            self.logger.addHandler(h)

    def info(self, msg: str) -> None:
        self.logger.info(msg)

    def error(self, msg: str) -> None:
        self.logger.error(msg)

    def warning(self, msg: str) -> None:
        self.logger.warning(msg)


class PerformanceManager:
    """``commitJob({task_id, round, t_bcast_ms, t_local_ms, t_reduce_ms, clients, local_steps, bytes_bcast, ...})``"""

    def __init__(self, peaks: Optional[Dict[str, float]] = None, metrics: Optional[Any] = None):
        self._jobs: Dict[str, List[Dict[str, Any]]] = collections.defaultdict(list)
        self._lock = threading.Lock()
        self.peaks = peaks or {}
        self.metrics = metrics          # utils.metrics.Metrics: every committed record is also exported to Prometheus

    def commitJob(self, request: Dict[str, Any]) -> Dict[str, Any]:
        tid = request.get("task_id", "")
        if not tid:
            return {"is_success": False}
        with self._lock:
            self._jobs[tid].append(dict(request))
        if self.metrics is not None:
            self.metrics.observe_round(request)
        return {"is_success": True, "count": len(self._jobs[tid])}

    def summary(self, task_id: str) -> Dict[str, Any]:
        with self._lock:
            recs = list(self._jobs.get(task_id, []))
        if not recs:
            return {}
        tot = sum(r.get("t_round_ms", 0.0) for r in recs)
        steps = sum(r.get("client_steps", 0) for r in recs)
        out = {"rounds": len(recs), "rounds_per_s": 1e3 * len(recs) / tot if tot else None,
               "client_steps_per_s": 1e3 * steps / tot if tot else None}
        if self.peaks.get("hbm_gbs") and any("hbm_bytes" in r for r in recs):
            b = sum(r.get("hbm_bytes", 0) for r in recs)
            out["hbm_fraction_of_measured"] = (b / (tot / 1e3)) / (self.peaks["hbm_gbs"] * 1e9) if tot else None
        return out
