"""Inbound / outbound rooms, shelf room, message envelope helpers, producers, exceptions, observer.

Counterparts (all under ols_core/deviceflow/non_grpc/): ``bound_room.py`` (InboundRoom/OutboundRoom :36-64),
``shelf_room.py`` (one topic per flow "{namespace}/{flow_id}", subscription "shelf_room" :8-131),
``utils/message.py`` (:4-19), ``message_producer.py`` (Pulsar producer :42-56, WebSocket producer wrapping the
payload as {"payload": base64} :59-78), ``utils/exception.py`` and ``utils/observer.py``.
"""
from __future__ import annotations

import base64
import json
import threading
from abc import ABC, abstractmethod
from typing import Any, Callable, Dict, List, Optional

from .bus import Consumer, InMemoryBus, Producer

SHELF_SUBSCRIPTION = "shelf_room"


# ---- exceptions (utils/exception.py) -------------------------------------------------------------------
class DeviceFlowError(Exception):
    def __init__(self, message: str = "", **ctx: Any):
        super().__init__(message)
        self.message, self.ctx = message, ctx


class InvalidStrategyNameException(DeviceFlowError): ...
class AddShelfException(DeviceFlowError): ...
class NotifyDispatchException(DeviceFlowError): ...
class ShelfNotExistException(DeviceFlowError): ...
class MarkNotifyStartException(DeviceFlowError): ...
class MarkNotifyCompleteException(DeviceFlowError): ...
class GetConsumerException(DeviceFlowError): ...


# ---- observer (utils/observer.py; InboundRoom is a Subject, as in the reference) ------------------------
class Observer(ABC):
    @abstractmethod
    def update(self, subject: "Subject", *args: Any, **kwargs: Any) -> None: ...


class Subject:
    def __init__(self) -> None:
        self._observers: List[Observer] = []

    def attach(self, o: Observer) -> None:
        if o not in self._observers:
            self._observers.append(o)

    def detach(self, o: Observer) -> None:
        if o in self._observers:
            self._observers.remove(o)

    def notify(self, *args: Any, **kwargs: Any) -> None:
        for o in list(self._observers):
            o.update(self, *args, **kwargs)


# ---- message envelope (utils/message.py) ---------------------------------------------------------------
def make_message(routing_key: str, compute_resource: str, message: str) -> bytes:
    """The envelope user operators publish: {"routing_key","compute_resource","message"}"""
    return json.dumps({"routing_key": routing_key, "compute_resource": compute_resource, "message": message}).encode()


def parse_envelope(raw: bytes) -> Dict[str, Any]:
    """One parse for all three fields (the reference re-parses the JSON once per field)."""
    d = json.loads(raw)
    return {"routing_key": d["routing_key"], "compute_resource": d["compute_resource"], "message": d["message"]}


def parse_routing_key(raw: bytes) -> str:
    return parse_envelope(raw)["routing_key"]


def parse_compute_resource(raw: bytes) -> str:
    return parse_envelope(raw)["compute_resource"]


def parse_message_content(raw: bytes) -> str:
    return parse_envelope(raw)["message"]


# ---- rooms ---------------------------------------------------------------------------------------------
class InboundRoom(Subject):
    def __init__(self, bus: InMemoryBus, url: str = "mem://local", topic: str = "deviceflow_inbound"):
        super().__init__()
        self._bus, self._url, self._topic = bus, url, topic
        bus.create_producer(topic)

    def get_consumer(self, subscription: str) -> Consumer:
        try:
            return self._bus.subscribe(self._topic, subscription)
        except Exception as e:  # noqa: BLE001
            raise GetConsumerException(str(e), topic=self._topic) from e

    def get_producer(self) -> Producer:
        return self._bus.create_producer(self._topic)

    def get_availabl_source_url(self) -> str:      # (sic) reference spelling kept for API parity
        return self._url

    def get_availabl_source_topic(self) -> str:
        return self._topic

    get_available_source_url = get_availabl_source_url
    get_available_source_topic = get_availabl_source_topic


class OutboundRoom:
    def __init__(self, bus: InMemoryBus, url: str = "mem://local", topic: str = "deviceflow_outbound"):
        self._bus, self._url, self._topic = bus, url, topic

    def get_producer(self) -> Producer:
        return self._bus.create_producer(self._topic)

    def get_consumer(self, subscription: str) -> Consumer:
        return self._bus.subscribe(self._topic, subscription)


class Shelf:
    """Durable per-(task, operator, round) buffer of client updates."""

    def __init__(self, bus: InMemoryBus, namespace: str, flow_id: str):
        self.flow_id = flow_id
        self.topic = f"{namespace}/{flow_id}"
        self._bus = bus
        self._consumer = bus.subscribe(self.topic, SHELF_SUBSCRIPTION)   # subscribe first: nothing is lost
        self._producer = bus.create_producer(self.topic)

    def put(self, payload: bytes) -> None:
        self._producer.send(payload)

    def get_consumer(self) -> Consumer:
        return self._consumer

    def close(self) -> None:
        self._consumer.close()
        self._producer.close()


class ShelfRoom:
    def __init__(self, bus: InMemoryBus, url: str = "mem://local", namespace: str = "shelf_room"):
        self._bus, self._url, self._namespace = bus, url, namespace
        self._shelves: Dict[str, Shelf] = {}
        self._lock = threading.RLock()

    def add_shelf(self, flow_id: str) -> bool:
        with self._lock:
            if flow_id in self._shelves:
                return True
            try:
                self._shelves[flow_id] = Shelf(self._bus, self._namespace, flow_id)
            except Exception as e:  # noqa: BLE001
                raise AddShelfException(str(e), flow_id=flow_id) from e
            return True

    def has_shelf(self, flow_id: str) -> bool:
        with self._lock:
            return flow_id in self._shelves

    def remove_shelf(self, flow_id: str) -> bool:
        with self._lock:
            sh = self._shelves.pop(flow_id, None)
        if sh is None:
            return False
        sh.close()
        self._bus.delete_topic(sh.topic)
        return True

    def put_on_shelf(self, flow_id: str, payload: bytes) -> None:
        with self._lock:
            sh = self._shelves.get(flow_id)
        if sh is None:
            raise ShelfNotExistException(f"no shelf for flow {flow_id}", flow_id=flow_id)
        sh.put(payload)

    def get_consumer(self, flow_id: str) -> Consumer:
        with self._lock:
            sh = self._shelves.get(flow_id)
        if sh is None:
            raise ShelfNotExistException(f"no shelf for flow {flow_id}", flow_id=flow_id)
        return sh.get_consumer()


# ---- outbound producers (message_producer.py) -------------------------------------------------------------
class MessageProducerBase(ABC):
    @abstractmethod
    def get_producer(self, outbound_params: Dict[str, Any]) -> None: ...

    @abstractmethod
    def send(self, message: bytes) -> None: ...

    def close(self) -> None:
        pass


class PulsarClientProducer(MessageProducerBase):
    """outbound_service = {"pulsar_client": {"url", "topic"}} — here the URL selects a bus from a registry."""

    buses: Dict[str, InMemoryBus] = {}

    def __init__(self, default_bus: Optional[InMemoryBus] = None):
        self._default, self._producer = default_bus, None

    def get_producer(self, outbound_params: Dict[str, Any]) -> None:
        url = outbound_params.get("url", "")
        bus = self.buses.get(url)
        if bus is None and url.startswith(("pulsar://", "pulsar+ssl://")):
            from .pulsar_bus import PulsarBus           # a real broker URL: connect once, keep the bus for later flows
            bus = self.buses[url] = PulsarBus(url)
        bus = bus or self._default
        if bus is None:
            raise DeviceFlowError(f"no bus registered for url {url!r}")
        self._producer = bus.create_producer(outbound_params["topic"])

    def send(self, message: bytes) -> None:
        self._producer.send(message)


class WebsocketProducer(MessageProducerBase):
    """Wraps each message as {"payload": base64(msg)} (Pulsar WebSocket producer format).  The transport is a
    callable registered per URL (tests / in-process aggregator) or, if none, a real ``websockets`` connection."""

    transports: Dict[str, Callable[[str], None]] = {}

    def __init__(self) -> None:
        self._send: Optional[Callable[[str], None]] = None
        self._ws = None

    def get_producer(self, outbound_params: Dict[str, Any]) -> None:
        url = outbound_params["url"]
        if url in self.transports:
            self._send = self.transports[url]
            return
        from websockets.sync.client import connect        # lazy: needs a live endpoint
        self._ws = connect(url)
        self._send = self._ws.send

    def send(self, message: bytes) -> None:
        if isinstance(message, str):
            message = message.encode("utf-8")
        self._send(json.dumps({"payload": base64.b64encode(message).decode("utf-8")}))

    def close(self) -> None:
        if self._ws is not None:
            self._ws.close()


class CallbackProducer(MessageProducerBase):
    """outbound_service = {"callback": {"name"}} — B200 extension: deliver forwarded updates straight to an
    in-process sink (the GPU aggregator turns them into participation-mask bits)."""

    sinks: Dict[str, Callable[[bytes], None]] = {}

    def __init__(self) -> None:
        self._sink: Optional[Callable[[bytes], None]] = None

    def get_producer(self, outbound_params: Dict[str, Any]) -> None:
        self._sink = self.sinks[outbound_params["name"]]

    def send(self, message: bytes) -> None:
        self._sink(message)


def make_outbound_producer(outbound_service: Dict[str, Dict[str, Any]], default_bus: Optional[InMemoryBus] = None):
    """dispatcher.py:65-77 ``get_producer``"""
    if "pulsar_client" in outbound_service:
        p: MessageProducerBase = PulsarClientProducer(default_bus)
        p.get_producer(outbound_service["pulsar_client"])
        return p
    if "websocket" in outbound_service:
        p = WebsocketProducer()
        p.get_producer(outbound_service["websocket"])
        return p
    if "callback" in outbound_service:
        p = CallbackProducer()
        p.get_producer(outbound_service["callback"])
        return p
    return None
