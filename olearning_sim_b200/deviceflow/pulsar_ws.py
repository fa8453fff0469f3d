"""Pulsar over its WebSocket API — a broker transport that needs no ``pulsar-client`` wheel.

A Pulsar broker (or its WebSocket proxy) serves, next to the binary protocol on :6650, a WebSocket API on :8080:

    producer   ws://host:8080/ws/v2/producer/persistent/<tenant>/<ns>/<topic>
                   -> {"payload": base64, "context": "<n>"}          <- {"result": "ok", "messageId": "...", "context": "<n>"}
    consumer   ws://host:8080/ws/v2/consumer/persistent/<tenant>/<ns>/<topic>/<subscription>?subscriptionType=Shared
                   <- {"messageId": "...", "payload": base64, ...}    -> {"messageId": "..."}           (acknowledgement)

It is the API the reference's own outbound producer speaks (ols_core/deviceflow/non_grpc/message_producer.py:58-78).
``PulsarWsClient`` wraps it in the ``pulsar.Client`` surface that ``PulsarBus`` (deviceflow/pulsar_bus.py) is written
against — ``create_producer(topic).send(bytes)``, ``subscribe(topic, subscription_name, consumer_type)`` with
``receive(timeout_millis)`` / ``acknowledge`` / ``close``, messages with ``data()`` / ``message_id()`` — so
``PulsarBus("ws://broker:8080")`` puts DeviceFlow's inbound room, shelves and dispatchers on a real broker with nothing but
the ``websockets`` package.  Un-acknowledged messages are redelivered by the broker when the consumer reconnects.

tests/test_round2_parity.py runs it against an in-process WebSocket server that implements these two endpoints (per-
subscription cursors, acks, redelivery).
"""
from __future__ import annotations

import base64
import json
import threading
from typing import Any, Dict, Optional


class Timeout(Exception):
    """no message within ``timeout_millis`` (the name PulsarBus maps to BusTimeout, like ``_pulsar.Timeout``)"""


class AlreadyClosed(Exception):
    pass


class _WsMessage:
    def __init__(self, mid: str, data: bytes, properties: Optional[Dict[str, str]] = None):
        self._id, self._d, self._p = mid, data, properties or {}

    def data(self) -> bytes:
        return self._d

    def message_id(self) -> str:
        return self._id

    def properties(self) -> Dict[str, str]:
        return self._p


def _topic_path(topic: str) -> str:
    """persistent://tenant/ns/topic -> persistent/tenant/ns/topic (bare names land in public/default)"""
    if "://" in topic:
        scheme, rest = topic.split("://", 1)
        return f"{scheme}/{rest}"
    return "persistent/public/default/" + topic if "/" not in topic else "persistent/public/" + topic


class _WsProducer:
    def __init__(self, url: str, timeout: float):
        from websockets.sync.client import connect
        self._ws, self._n, self._timeout, self._lock = connect(url, open_timeout=timeout), 0, timeout, threading.Lock()

    def send(self, payload: bytes, properties: Optional[Dict[str, str]] = None) -> str:
        if isinstance(payload, str):
            payload = payload.encode("utf-8")
        with self._lock:
            self._n += 1
            frame: Dict[str, Any] = {"payload": base64.b64encode(payload).decode("ascii"), "context": str(self._n)}
            if properties:
                frame["properties"] = properties
            self._ws.send(json.dumps(frame))
            ack = json.loads(self._ws.recv(timeout=self._timeout))
        if ack.get("result") != "ok":
            raise RuntimeError(f"Pulsar producer: {ack.get('errorMsg') or ack}")
        return ack.get("messageId", "")

    def close(self) -> None:
        try:
            self._ws.close()
        except Exception:  # noqa: BLE001
            pass


class _WsConsumer:
    def __init__(self, url: str, timeout: float):
        from websockets.sync.client import connect
        self._ws, self._closed = connect(url, open_timeout=timeout), False

    def receive(self, timeout_millis: Optional[int] = None) -> _WsMessage:
        if self._closed:
            raise AlreadyClosed("consumer closed")
        try:
            raw = self._ws.recv(timeout=None if timeout_millis is None else max(timeout_millis, 0) / 1000.0)
        except TimeoutError:
            raise Timeout("Pulsar error: TimeOut") from None
        m = json.loads(raw)
        return _WsMessage(m["messageId"], base64.b64decode(m.get("payload", "")), m.get("properties"))

    def acknowledge(self, msg: _WsMessage) -> None:
        self._ws.send(json.dumps({"messageId": msg.message_id()}))

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            try:
                self._ws.close()
            except Exception:  # noqa: BLE001
                pass


class PulsarWsClient:
    """``service_url``: ``ws://host:8080`` or ``wss://...`` of the broker's WebSocket service; ``token`` (optional) is sent
    as the ``token`` query parameter, the form the WebSocket API accepts for JWT authentication."""

    class ConsumerType:
        Shared, Exclusive, Failover = "Shared", "Exclusive", "Failover"

    def __init__(self, service_url: str, token: Optional[str] = None, timeout: float = 10.0):
        self.base, self.token, self.timeout = service_url.rstrip("/"), token, timeout
        self._handles = []

    def _url(self, kind: str, topic: str, tail: str = "", query: str = "") -> str:
        q = [x for x in (query, f"token={self.token}" if self.token else "") if x]
        return f"{self.base}/ws/v2/{kind}/{_topic_path(topic)}{tail}" + ("?" + "&".join(q) if q else "")

    def create_producer(self, topic: str, **_kw) -> _WsProducer:
        p = _WsProducer(self._url("producer", topic), self.timeout)
        self._handles.append(p)
        return p

    def subscribe(self, topic: str, subscription_name: str, consumer_type: Any = "Shared", **_kw) -> _WsConsumer:
        ctype = getattr(consumer_type, "name", None) or str(consumer_type)
        c = _WsConsumer(self._url("consumer", topic, "/" + subscription_name, f"subscriptionType={ctype}"), self.timeout)
        self._handles.append(c)
        return c

    def close(self) -> None:
        for h in self._handles:
            h.close()
        self._handles.clear()
