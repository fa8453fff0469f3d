"""Apache Pulsar transport for DeviceFlow — the broker the reference runs on.

``PulsarBus`` has the interface of ``InMemoryBus`` (``create_producer`` / ``subscribe`` / ``delete_topic``; consumers with
``receive(timeout_millis)`` / ``acknowledge`` / ``close``), so ``DeviceFlowService(bus=PulsarBus("pulsar://host:6650"))``
puts the inbound room, the per-flow shelves and the dispatchers on a real broker:

* inbound topic + single ``sorter`` subscription            ols_core/deviceflow/non_grpc/bound_room.py:22-64, sorter.py:27-51
* one shelf topic per flow, subscription ``shelf_room``       ols_core/deviceflow/non_grpc/shelf_room.py:82-131
* outbound producer of the cloud aggregator                  ols_core/deviceflow/non_grpc/message_producer.py:42-56

Subscriptions are ``Shared`` like the reference's, receive timeouts map to ``BusTimeout``; un-acked messages are
redelivered by the broker after a restart, which is what the service's restart recovery relies on.  The ``pulsar`` client
is imported lazily (it is not installable offline).  ``PulsarBus("ws://host:8080")`` needs no wheel at all: it talks to
the broker's WebSocket API through ``deviceflow/pulsar_ws.py``; ``client=`` injects any object with the ``pulsar.Client``
surface.  The tests drive the whole sorter -> shelf -> dispatcher path both through an in-memory fake client and through
the WebSocket client against an in-process broker emulation.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

from .bus import BusClosed, BusTimeout


class _PulsarMessage:
    def __init__(self, raw: Any, topic: str):
        self.raw, self.topic = raw, topic
        self.msg_id = raw.message_id() if hasattr(raw, "message_id") else None

    def value(self) -> bytes:
        return self.raw.data()

    def data(self) -> bytes:
        return self.raw.data()


class PulsarConsumer:
    def __init__(self, bus: "PulsarBus", topic: str, subscription: str, raw: Any):
        self._bus, self.topic, self.subscription, self._raw = bus, topic, subscription, raw
        self._closed = False

    def receive(self, timeout_millis: Optional[int] = None) -> _PulsarMessage:
        if self._closed:
            raise BusClosed(self.topic)
        try:
            raw = self._raw.receive(timeout_millis=timeout_millis) if timeout_millis is not None else self._raw.receive()
        except Exception as e:  # noqa: BLE001 - the client raises _pulsar.Timeout (a plain Exception subclass)
            if type(e).__name__ in ("Timeout", "TimeOut") or "timeout" in str(e).lower():
                raise BusTimeout(self.topic) from None
            if type(e).__name__ in ("AlreadyClosed", "ConsumerClosed"):
                raise BusClosed(self.topic) from None
            raise
        return _PulsarMessage(raw, self.topic)

    def acknowledge(self, msg: _PulsarMessage) -> None:
        self._raw.acknowledge(msg.raw)

    def backlog(self) -> int:
        return -1                                    # not observable from a client

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            try:
                self._raw.close()
            except Exception:  # noqa: BLE001
                pass


class PulsarProducer:
    def __init__(self, topic: str, raw: Any):
        self.topic, self._raw, self.sent = topic, raw, 0

    def send(self, payload: bytes) -> Any:
        if isinstance(payload, str):
            payload = payload.encode("utf-8")
        self.sent += 1
        return self._raw.send(payload)

    def close(self) -> None:
        try:
            self._raw.close()
        except Exception:  # noqa: BLE001
            pass


class PulsarBus:
    virtual_time = False

    def __init__(self, service_url: str = "pulsar://127.0.0.1:6650", client: Any = None, tenant_prefix: str = "persistent://public/"):
        if client is None and service_url.startswith(("ws://", "wss://")):
            # the broker's WebSocket API (deviceflow/pulsar_ws.py): no pulsar-client wheel needed
            from .pulsar_ws import PulsarWsClient
            client = PulsarWsClient(service_url)
            self._shared = PulsarWsClient.ConsumerType.Shared
        elif client is None:
            try:
                import pulsar  # type: ignore
            except ImportError as e:
                raise RuntimeError("pulsar:// URLs need the pulsar-client package, which is not installed on this box; use the "
                                   "broker's WebSocket service (ws://host:8080), InMemoryBus, or inject a client") from e
            client = pulsar.Client(service_url)
            self._shared = pulsar.ConsumerType.Shared
        else:
            self._shared = getattr(getattr(client, "ConsumerType", None), "Shared", "Shared")
        self.service_url, self._client, self._prefix = service_url, client, tenant_prefix
        self._producers: Dict[str, PulsarProducer] = {}
        self._consumers: List[PulsarConsumer] = []

    def _full(self, topic: str) -> str:
        """shelf topics are "<namespace>/<flow_id>", the inbound topic is a bare name or already fully qualified"""
        if "://" in topic:
            return topic
        return self._prefix + (topic if "/" in topic else f"default/{topic}")

    def create_producer(self, topic: str) -> PulsarProducer:
        p = self._producers.get(topic)
        if p is None:
            p = self._producers[topic] = PulsarProducer(topic, self._client.create_producer(self._full(topic)))
        return p

    def subscribe(self, topic: str, subscription: str) -> PulsarConsumer:
        raw = self._client.subscribe(self._full(topic), subscription_name=subscription, consumer_type=self._shared)
        c = PulsarConsumer(self, topic, subscription, raw)
        self._consumers.append(c)
        return c

    def delete_topic(self, topic: str) -> None:
        """Topic deletion is a broker-admin operation; closing our handles is what a client can do (the reference's
        ``remove_shelf`` only closes its producer / consumer too: shelf_room.py:58-70)."""
        p = self._producers.pop(topic, None)
        if p is not None:
            p.close()
        for c in [c for c in self._consumers if c.topic == topic]:
            c.close()
            self._consumers.remove(c)

    def topics(self) -> List[str]:
        return sorted(set(self._producers) | {c.topic for c in self._consumers})

    def backlog(self, topic: str, subscription: str) -> int:
        return -1

    def close(self) -> None:
        for c in list(self._consumers):
            c.close()
        for p in list(self._producers.values()):
            p.close()
        try:
            self._client.close()
        except Exception:  # noqa: BLE001
            pass
