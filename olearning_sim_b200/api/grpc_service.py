"""gRPC façade with the reference's service and method names, and ``SimulatorSession`` to host them.

Services / methods (ols_core/proto/*.proto):

    TaskMgr                          submitTask stopTask getTaskStatus getTaskQueue changeScheduler
    ResourceMgr                      getClusterAvailableResource getClusterResourceDetail requestClusterResource
                                     releaseClusterResource getClusterTotalResource getResource requestResource
                                     releaseResource getVMResource getPhoneResource getTotalResource
    TaskOperatorOrientedDeviceFlow   GetDeviceflowPulsarClient GetDeviceflowWebsocket NotifyStart NotifyComplete
                                     RegisterTask UnRegisterTask GetTotalComputeResources CheckDeviceflowDispatchFinished
    RayClusterMgr (-> GpuClusterMgr) createRayCluster deleteRayCluster modifyRayCluster queryRayCluster
    PerformanceMgr                   commitJob

Two codecs share every method path (``/TaskMgr/submitTask``, ``/deviceflow.TaskOperatorOrientedDeviceFlow/NotifyStart`` ...):

* **protobuf** — the reference's messages with their field numbers (``api/wire.py`` builds the descriptors without
  ``protoc``): what a client generated from ols_core/proto/*.proto sends.  This is the default codec.
* **JSON** — one UTF-8 JSON object per message; chosen when the call carries the metadata ``x-olsb-codec: json`` (``JsonStub``
  sends it) or the payload starts like a JSON document.

Same 100 MiB message cap (ols_core/config/config.conf:2-3), insecure port like the reference (ols_core/simu_session.py:
31-47).  ``JsonStub`` / ``ProtoStub`` are the matching clients.
"""
from __future__ import annotations

import json
import time
from concurrent import futures
from dataclasses import asdict, is_dataclass
from typing import Any, Callable, Dict, Optional

import grpc

MAX_MESSAGE = 100 * 1024 * 1024
_OPTS = [("grpc.max_send_message_length", MAX_MESSAGE), ("grpc.max_receive_message_length", MAX_MESSAGE)]


def _ser(obj: Any) -> bytes:
    def default(o):
        if is_dataclass(o):
            return asdict(o)
        if hasattr(o, "name") and hasattr(o, "value"):
            return o.name
        return str(o)
    return json.dumps(obj, default=default).encode("utf-8")


def _de(b: bytes) -> Any:
    return json.loads(b.decode("utf-8")) if b else {}


JSON_CODEC_MD = ("x-olsb-codec", "json")


def _raw(b: bytes) -> bytes:
    return b


def _generic(service: str, methods: Dict[str, Callable[[Dict[str, Any]], Any]], mount_as: Optional[str] = None):
    from . import wire

    def wrap(name, fn):
        def handler(payload: bytes, context):
            md = dict(context.invocation_metadata() or ())
            use_json = md.get(JSON_CODEC_MD[0]) == JSON_CODEC_MD[1] or wire.looks_like_json(payload or b"") or \
                name not in wire.SERVICES.get(service, {})
            try:
                if use_json:
                    return _ser(fn(_de(payload) or {}))
                return wire.encode_response(service, name, fn(wire.decode_request(service, name, payload or b"") or {}))
            except Exception as e:  # noqa: BLE001
                context.set_code(grpc.StatusCode.INTERNAL)
                context.set_details(f"{type(e).__name__}: {e}")
                return _ser({"is_success": False, "error": str(e)}) if use_json else b""
        return grpc.unary_unary_rpc_method_handler(handler, request_deserializer=_raw, response_serializer=_raw)
    return grpc.method_handlers_generic_handler(mount_as or service, {m: wrap(m, f) for m, f in methods.items()})


# ---- servicers ---------------------------------------------------------------------------------------------
def task_mgr_handler(tm: Any):
    return _generic("TaskMgr", {
        "submitTask": lambda r: {"is_success": bool(tm.submitTask(r.get("task", r)))},
        "stopTask": lambda r: {"is_success": bool(tm.stopTask(r["taskID"]))},
        "getTaskStatus": lambda r: {"taskStatus": tm.getTaskStatus(r["taskID"])},
        "getTaskQueue": lambda r: {"tasks": tm.getTaskQueue()},
        "changeScheduler": lambda r: {"is_success": bool(tm.changeScheduler(r.get("strategy")))},
    })


def resource_mgr_handler(rm: Any):
    return _generic("ResourceMgr", {
        "getClusterAvailableResource": lambda r: rm.getClusterAvailableResource(),
        "getClusterResourceDetail": lambda r: {"details": rm.getClusterResourceDetail()},
        "requestClusterResource": lambda r: {"status": rm.requestClusterResource(r["task_id"], r.get("cpu", 0.0), r.get("mem", 0.0), r.get("gpu", 0.0)).name},
        "releaseClusterResource": lambda r: {"status": rm.releaseClusterResource(r["task_id"]).name},
        "getClusterTotalResource": lambda r: rm.getClusterTotalResource(),
        "getResource": lambda r: rm.getResource(),
        "requestResource": lambda r: dict(zip(("is_success", "status"), rm.requestResource(
            r["task_id"], r.get("user_id", ""), r.get("cpu", 0.0), r.get("mem", 0.0), r.get("gpu", 0.0), r.get("phones"), r.get("type", 0)))),
        "releaseResource": lambda r: dict(zip(("is_success", "status"), rm.releaseResource(
            r["task_id"], r.get("user_id", ""), r.get("phones"), r.get("type", 0)))),
        "getVMResource": lambda r: rm.getVMResource(),
        "getPhoneResource": lambda r: rm.getPhoneResource(),
        "getTotalResource": lambda r: rm.getTotalResource(),
    })


def deviceflow_handler(df: Any, mount_as: Optional[str] = None):
    from ..deviceflow.service import NotifyRequest

    def req(r):
        return NotifyRequest(r["task_id"], r["routing_key"], r["compute_resource"], r.get("strategy", ""),
                             r.get("outbound_service", {}) or {})
    return _generic("TaskOperatorOrientedDeviceFlow", {
        "GetDeviceflowPulsarClient": lambda r: df.GetDeviceflowPulsarClient(),
        "GetDeviceflowWebsocket": lambda r: df.GetDeviceflowWebsocket(),
        "NotifyStart": lambda r: {"is_success": bool(df.NotifyStart(req(r)))},
        "NotifyComplete": lambda r: {"is_success": bool(df.NotifyComplete(req(r)))},
        "RegisterTask": lambda r: {"is_success": bool(df.RegisterTask(r["task_id"], r.get("total_compute_resources", [])))},
        "UnRegisterTask": lambda r: {"is_success": bool(df.UnRegisterTask(r["task_id"]))},
        "GetTotalComputeResources": lambda r: {"total_compute_resources": df.GetTotalComputeResources(r["task_id"])},
        "CheckDeviceflowDispatchFinished": lambda r: {"is_success": bool(df.CheckDeviceflowDispatchFinished(r["task_id"]))},
        # extension (JSON codec): the GPU engine reports what its on-device dispatch of a round did
        "ReportDeviceDispatch": lambda r: {"is_success": bool(df.ReportDeviceDispatch(
            r["task_id"], r["flow_id"], r.get("compute_resource", "logical_simulation"), r.get("summary", {})))},
    }, mount_as=mount_as)


def cluster_mgr_handler(cm: Any):
    return _generic("RayClusterMgr", {
        "createRayCluster": lambda r: cm.createRayCluster(r),
        "deleteRayCluster": lambda r: cm.deleteRayCluster(r),
        "modifyRayCluster": lambda r: cm.modifyRayCluster(r),
        "queryRayCluster": lambda r: cm.queryRayCluster(r),
    })


def phone_mgr_handler(pm: Any):
    """Phone-side ``TaskManager`` service (ols_core/proto/phoneMgr.proto:53-66) — in the reference only the client stubs
    exist and the server is an external product; mounting the in-process phone manager behind the same method names
    lets a separate ResourceManager / TaskManager process reach it through ``RemotePhoneMgr``."""
    return _generic("TaskManager", {
        "submitTask": lambda r: {"is_success": bool(pm.submitTask(r.get("task", r)))},
        "getDeviceAvailableResource": lambda r: {"resources": pm.getDeviceAvailableResource()},
        "requestDeviceResource": lambda r: {"is_success": bool(pm.requestDeviceResource(r["task_id"], r.get("user_id", ""),
                                                                                    r.get("phones", {})))},
        "releaseDeviceResource": lambda r: {"is_success": bool(pm.releaseDeviceResource(r["task_id"]))},
        "stopDevice": lambda r: {"is_success": bool(pm.stopDevice(r["task_id"]))},
        "getDeviceTaskStatus": lambda r: pm.getDeviceTaskStatus(r["task_id"]),
    })


def phone_manager_handler(pm: Any):
    """``PhoneManager`` (proto only in the reference): device attach / local inventory."""
    return _generic("PhoneManager", {
        "connectMSPDevices": lambda r: {"is_success": bool(pm.connectMSPDevices(r.get("user_id", ""), r.get("devices", {})))},
        "getLocalDeviceResource": lambda r: {"resources": pm.getLocalDeviceResource()},
    })


class RemotePhoneMgr:
    """Client-side adapter with ``FakePhoneMgr``'s method surface over gRPC (what the reference's ResourceMgr / TaskMgr hold
    as phone-manager stubs: resource_manager.py:29-47, task_manager.py:36-60)."""

    def __init__(self, address: str):
        self._stub = JsonStub(connect(address), "TaskManager")

    def getDeviceAvailableResource(self):
        return self._stub.getDeviceAvailableResource({})["resources"]

    def requestDeviceResource(self, task_id, user_id, phones):
        return self._stub.requestDeviceResource({"task_id": task_id, "user_id": user_id, "phones": phones})["is_success"]

    def releaseDeviceResource(self, task_id):
        return self._stub.releaseDeviceResource({"task_id": task_id})["is_success"]

    def submitTask(self, device_task):
        return self._stub.submitTask({"task": device_task})["is_success"]

    def stopDevice(self, task_id):
        return self._stub.stopDevice({"task_id": task_id})["is_success"]

    def getDeviceTaskStatus(self, task_id):
        return self._stub.getDeviceTaskStatus({"task_id": task_id})


class RemoteResourceMgr:
    """Client-side adapter with the part of ``ResourceManager`` the task side uses (``getResource`` / ``requestResource`` /
    ``releaseResource`` and the ``phone_mgr`` handle), over gRPC — the reference's TaskMgr and TaskScheduler hold exactly such
    stubs to a ResourceMgr in another process (task_scheduler.py:40-75, task_manager.py:36-60; address = ``[grpc] resmgr``)."""

    def __init__(self, address: str, phone_mgr: Any = None, phone_address: Optional[str] = None):
        self.address = address
        self._stub = JsonStub(connect(address), "ResourceMgr")
        self.phone_mgr = phone_mgr if phone_mgr is not None else RemotePhoneMgr(phone_address or address)

    def getResource(self) -> Dict[str, Any]:                      # noqa: N802 - RPC names of the reference
        return self._stub.getResource({})

    def getClusterAvailableResource(self) -> Dict[str, Any]:      # noqa: N802
        return self._stub.getClusterAvailableResource({})

    def requestResource(self, task_id: str, user_id: str = "", cpu: float = 0.0, mem: float = 0.0, gpu: float = 0.0,   # noqa: N802
                        phones: Optional[Dict[str, int]] = None, type: int = 0):                                       # noqa: A002
        r = self._stub.requestResource({"task_id": task_id, "user_id": user_id, "cpu": cpu, "mem": mem, "gpu": gpu,
                                        "phones": phones or {}, "type": type})
        return bool(r["is_success"]), r.get("status", "")

    def releaseResource(self, task_id: str, user_id: str = "", phones: Optional[Dict[str, int]] = None, type: int = 0):   # noqa: N802,A002
        r = self._stub.releaseResource({"task_id": task_id, "user_id": user_id, "phones": phones or {}, "type": type})
        return bool(r["is_success"]), r.get("status", "")


class RemoteDeviceFlow:
    """Client-side adapter with ``DeviceFlowService``'s method surface over gRPC — what a job process holds when the service
    lives in another process (the reference's ``Deviceflow`` wrapper around the generated stub, utils_run_task.py:53-144:
    ``start`` -> NotifyStart, ``stop`` -> NotifyComplete, inbound endpoint getters)."""

    def __init__(self, address: str):
        self.address = address
        self._stub = JsonStub(connect(address), "TaskOperatorOrientedDeviceFlow")

    @staticmethod
    def _req(req: Any) -> Dict[str, Any]:
        return {"task_id": req.task_id, "routing_key": req.routing_key, "compute_resource": req.compute_resource,
                "strategy": getattr(req, "strategy", ""), "outbound_service": getattr(req, "outbound_service", {}) or {}}

    def NotifyStart(self, req: Any) -> bool:                      # noqa: N802 - RPC names of the reference
        return bool(self._stub.NotifyStart(self._req(req))["is_success"])

    def NotifyComplete(self, req: Any) -> bool:                   # noqa: N802
        return bool(self._stub.NotifyComplete(self._req(req))["is_success"])

    def RegisterTask(self, task_id: str, total_compute_resources) -> bool:      # noqa: N802
        return bool(self._stub.RegisterTask({"task_id": task_id, "total_compute_resources": list(total_compute_resources)})["is_success"])

    def UnRegisterTask(self, task_id: str) -> bool:               # noqa: N802
        return bool(self._stub.UnRegisterTask({"task_id": task_id})["is_success"])

    def GetTotalComputeResources(self, task_id: str):            # noqa: N802
        return self._stub.GetTotalComputeResources({"task_id": task_id})["total_compute_resources"]

    def GetDeviceflowPulsarClient(self) -> Dict[str, Any]:        # noqa: N802
        return self._stub.GetDeviceflowPulsarClient({})

    def GetDeviceflowWebsocket(self) -> Dict[str, Any]:           # noqa: N802
        return self._stub.GetDeviceflowWebsocket({})

    def CheckDeviceflowDispatchFinished(self, task_id: str) -> bool:            # noqa: N802
        return bool(self._stub.CheckDeviceflowDispatchFinished({"task_id": task_id})["is_success"])

    def ReportDeviceDispatch(self, task_id: str, flow_id: str, compute_resource: str, summary: Dict[str, Any]) -> bool:   # noqa: N802
        return bool(self._stub.ReportDeviceDispatch({"task_id": task_id, "flow_id": flow_id, "compute_resource": compute_resource,
                                                     "summary": summary})["is_success"])


def performance_mgr_handler(pm: Any):
    return _generic("PerformanceMgr", {"commitJob": lambda r: pm.commitJob(r)})


class JsonStub:
    """Client for any of the services above:  ``JsonStub(channel, "TaskMgr").submitTask({...})``"""

    def __init__(self, channel: grpc.Channel, service: str):
        self._channel, self._service = channel, service

    def __getattr__(self, method: str):
        call = self._channel.unary_unary(f"/{self._service}/{method}", request_serializer=_ser, response_deserializer=_de)
        return lambda request=None, timeout=30: call(request or {}, timeout=timeout, metadata=(JSON_CODEC_MD,))


class ProtoStub:
    """Client speaking the reference's protobuf messages — what code generated from ols_core/proto/*.proto does:
    ``ProtoStub(channel, "TaskMgr").submitTask(wire.task_json_to_pb(task_json))`` returns an ``OperationStatus``."""

    def __init__(self, channel: grpc.Channel, service: str):
        from . import wire
        self._channel, self._service, self._wire = channel, service, wire
        pkg = wire.SERVICE_PACKAGE.get(service)
        self._path = f"{pkg}.{service}" if pkg else service

    def __getattr__(self, method: str):
        req_t, resp_t = self._wire.SERVICES[self._service][method]
        resp_cls = self._wire.msg(resp_t)
        call = self._channel.unary_unary(f"/{self._path}/{method}", request_serializer=lambda m: m.SerializeToString(),
                                         response_deserializer=resp_cls.FromString)
        return lambda request=None, timeout=30: call(request if request is not None else self._wire.msg(req_t)(), timeout=timeout)


def connect(address: str) -> grpc.Channel:
    return grpc.insecure_channel(address, options=_OPTS)


class SimulatorSession:
    """``SimulatorSession(port, services).start()`` — mounts servicers on one gRPC server.
    ``svc``: 0 all / 1 task / 2 resource / 3 cluster / 4 performance / 5 deviceflow / 6 phone manager
    (simu_session.py:34-46)."""

    def __init__(self, port: int = 60061, task_manager: Any = None, resource_manager: Any = None, deviceflow: Any = None,
                 cluster_manager: Any = None, performance_manager: Any = None, svc: int = 0, max_workers: int = 10,
                 host: str = "127.0.0.1", session_sleep_time: float = 10.0, phone_manager: Any = None):
        self.server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers), options=_OPTS)
        mounted = []
        if task_manager is not None and svc in (0, 1):
            self.server.add_generic_rpc_handlers((task_mgr_handler(task_manager),)); mounted.append("TaskMgr")
        if resource_manager is not None and svc in (0, 2):
            self.server.add_generic_rpc_handlers((resource_mgr_handler(resource_manager),)); mounted.append("ResourceMgr")
        if cluster_manager is not None and svc in (0, 3):
            self.server.add_generic_rpc_handlers((cluster_mgr_handler(cluster_manager),)); mounted.append("RayClusterMgr")
        if performance_manager is not None and svc in (0, 4):
            self.server.add_generic_rpc_handlers((performance_mgr_handler(performance_manager),)); mounted.append("PerformanceMgr")
        if phone_manager is not None and svc in (0, 6):
            # the phone-side service is named "TaskManager" in the reference's proto (the task service is "TaskMgr")
            self.server.add_generic_rpc_handlers((phone_mgr_handler(phone_manager),)); mounted.append("TaskManager(phone)")
            self.server.add_generic_rpc_handlers((phone_manager_handler(phone_manager),)); mounted.append("PhoneManager")
        if deviceflow is not None and svc in (0, 5):
            self.server.add_generic_rpc_handlers((deviceflow_handler(deviceflow),)); mounted.append("TaskOperatorOrientedDeviceFlow")
            # the reference's proto declares `package deviceflow;` — generated clients use the qualified path
            self.server.add_generic_rpc_handlers((deviceflow_handler(deviceflow, "deviceflow.TaskOperatorOrientedDeviceFlow"),))
        self.mounted = mounted
        self.port = self.server.add_insecure_port(f"{host}:{port}")
        self.address = f"{host}:{self.port}"
        if deviceflow is not None and svc in (0, 5) and task_manager is not None and hasattr(task_manager, "set_job_context"):
            # job processes (SubprocessJobExecutor / torchrun ranks) reach the DeviceFlow service through this address
            task_manager.set_job_context("deviceflow_address", self.address)
        self._sleep = session_sleep_time

    def start(self, block: bool = False) -> "SimulatorSession":
        self.server.start()
        if block:
            try:
                while True:
                    time.sleep(self._sleep)
            except KeyboardInterrupt:
                self.stop()
        return self

    def stop(self, grace: Optional[float] = 0.5) -> None:
        self.server.stop(grace)
