"""In-process message bus with Pulsar-like semantics + injectable clocks.

DeviceFlow in the reference rides on Apache Pulsar: producers publish to topics, consumers attach to a named
*subscription* (``ConsumerType.Shared``), ``receive(timeout_millis)`` raises on timeout, messages stay pending until
``acknowledge`` and are redelivered to a new consumer of the same subscription after a restart
(ols_core/deviceflow/non_grpc/bound_room.py:36-46, shelf_room.py:82-113, sorter.py:37-51).  This module provides those
semantics without a broker so Sorter / Shelf / Dispatcher behaviour is testable offline and usable in-process by the
GPU engine.  ``journal_path`` makes subscriptions durable across process restarts (JSONL append log).
"""
from __future__ import annotations

import base64
import collections
import json
import os
import threading
import time
from dataclasses import dataclass
from typing import Deque, Dict, List, Optional


class BusTimeout(Exception):
    """receive() found nothing within the timeout (pulsar's ``_pulsar.Timeout``)."""


class BusClosed(Exception):
    pass


@dataclass
class Message:
    msg_id: int
    topic: str
    payload: bytes

    def value(self) -> bytes:
        return self.payload

    def data(self) -> bytes:
        return self.payload


class Clock:
    def time(self) -> float:
        return time.time()

    def sleep(self, seconds: float) -> None:
        if seconds > 0:
            time.sleep(seconds)


class FakeClock(Clock):
    """Virtual time: sleep() advances instantly (DeviceFlow timing tests run in microseconds)."""

    def __init__(self, start: float = 0.0):
        self._t = start
        self._lock = threading.Lock()
        self.sleeps: List[float] = []

    def time(self) -> float:
        with self._lock:
            return self._t

    def sleep(self, seconds: float) -> None:
        with self._lock:
            self.sleeps.append(seconds)
            if seconds > 0:
                self._t += seconds


class _Subscription:
    def __init__(self) -> None:
        self.pending: Deque[Message] = collections.deque()
        self.inflight: Dict[int, Message] = {}
        self.cond = threading.Condition()


class Consumer:
    def __init__(self, bus: "InMemoryBus", topic: str, subscription: str):
        self._bus, self.topic, self.subscription = bus, topic, subscription
        self._sub = bus._subscription(topic, subscription)
        self._closed = False

    def receive(self, timeout_millis: Optional[int] = None) -> Message:
        if self._closed:
            raise BusClosed(self.topic)
        deadline = None if timeout_millis is None else time.monotonic() + timeout_millis / 1000.0
        sub = self._sub
        with sub.cond:
            while not sub.pending:
                if self._closed:
                    raise BusClosed(self.topic)
                if deadline is None:
                    sub.cond.wait(0.05)
                    continue
                remaining = deadline - time.monotonic()
                if remaining <= 0 or self._bus.virtual_time:
                    raise BusTimeout(self.topic)
                sub.cond.wait(remaining)
            msg = sub.pending.popleft()
            sub.inflight[msg.msg_id] = msg
            return msg

    def acknowledge(self, msg: Message) -> None:
        with self._sub.cond:
            self._sub.inflight.pop(msg.msg_id, None)
        self._bus._journal({"op": "ack", "topic": self.topic, "sub": self.subscription, "id": msg.msg_id})

    def backlog(self) -> int:
        with self._sub.cond:
            return len(self._sub.pending)

    def close(self) -> None:
        self._closed = True
        with self._sub.cond:
            # un-acked messages go back to the subscription (redelivery)
            for m in sorted(self._sub.inflight.values(), key=lambda x: x.msg_id, reverse=True):
                self._sub.pending.appendleft(m)
            self._sub.inflight.clear()
            self._sub.cond.notify_all()


class Producer:
    def __init__(self, bus: "InMemoryBus", topic: str):
        self._bus, self.topic = bus, topic
        self.sent = 0

    def send(self, payload: bytes) -> int:
        if isinstance(payload, str):
            payload = payload.encode("utf-8")
        self.sent += 1
        return self._bus._publish(self.topic, payload)

    def close(self) -> None:
        pass


class InMemoryBus:
    def __init__(self, journal_path: Optional[str] = None, virtual_time: bool = False):
        self._topics: Dict[str, Dict[str, _Subscription]] = {}
        self._retained: Dict[str, List[Message]] = {}
        self._lock = threading.RLock()
        self._next_id = 0
        self.virtual_time = virtual_time       # True: receive() never blocks (FakeClock tests)
        self._journal_path = journal_path
        self._jlock = threading.Lock()
        if journal_path and os.path.exists(journal_path):
            self._replay()

    # ---- topology --------------------------------------------------------------------------------
    def _subscription(self, topic: str, name: str) -> _Subscription:
        with self._lock:
            subs = self._topics.setdefault(topic, {})
            if name not in subs:
                subs[name] = _Subscription()
                self._journal({"op": "sub", "topic": topic, "sub": name})
            return subs[name]

    def create_producer(self, topic: str) -> Producer:
        with self._lock:
            self._topics.setdefault(topic, {})
        return Producer(self, topic)

    def subscribe(self, topic: str, subscription: str) -> Consumer:
        return Consumer(self, topic, subscription)

    def delete_topic(self, topic: str) -> None:
        with self._lock:
            self._topics.pop(topic, None)
        self._journal({"op": "del", "topic": topic})

    def topics(self) -> List[str]:
        with self._lock:
            return sorted(self._topics)

    def backlog(self, topic: str, subscription: str) -> int:
        with self._lock:
            sub = self._topics.get(topic, {}).get(subscription)
        if sub is None:
            return 0
        with sub.cond:
            return len(sub.pending) + len(sub.inflight)

    # ---- data path -------------------------------------------------------------------------------
    def _publish(self, topic: str, payload: bytes, msg_id: Optional[int] = None, journal: bool = True) -> int:
        with self._lock:
            if msg_id is None:
                msg_id = self._next_id
            self._next_id = max(self._next_id, msg_id + 1)
            subs = list(self._topics.setdefault(topic, {}).values())
        msg = Message(msg_id, topic, payload)
        for sub in subs:
            with sub.cond:
                sub.pending.append(msg)
                sub.cond.notify()
        if journal:
            self._journal({"op": "pub", "topic": topic, "id": msg_id, "b64": base64.b64encode(payload).decode()})
        return msg_id

    # ---- durability ------------------------------------------------------------------------------
    def _journal(self, rec: dict) -> None:
        if not self._journal_path:
            return
        with self._jlock:
            with open(self._journal_path, "a") as f:
                f.write(json.dumps(rec) + "\n")

    def _replay(self) -> None:
        path, self._journal_path = self._journal_path, None      # do not re-journal while replaying
        acked: Dict[tuple, set] = {}
        ops = []
        with open(path) as f:
            for line in f:
                try:
                    ops.append(json.loads(line))
                except ValueError:
                    continue
        for o in ops:
            if o["op"] == "ack":
                acked.setdefault((o["topic"], o["sub"]), set()).add(o["id"])
        for o in ops:
            if o["op"] == "sub":
                self._subscription(o["topic"], o["sub"])
            elif o["op"] == "del":
                self._topics.pop(o["topic"], None)
            elif o["op"] == "pub":
                self._publish(o["topic"], base64.b64decode(o["b64"]), msg_id=o["id"], journal=False)
        for (topic, name), ids in acked.items():
            sub = self._topics.get(topic, {}).get(name)
            if sub:
                sub.pending = collections.deque(m for m in sub.pending if m.msg_id not in ids)
        self._journal_path = path
