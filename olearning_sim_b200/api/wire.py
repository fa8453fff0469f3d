"""Protobuf wire contract of the reference's gRPC services, built without ``protoc``.

The reference ships six ``.proto`` files and their generated ``*_pb2`` modules (ols_core/proto/taskService.proto:4-211,
resourceService.proto:6-104, deviceflow.proto:6-72, phoneMgr.proto:8-107, rayclusterService.proto:4-38,
performanceService.proto:4-6).  ``grpc_tools`` is not available offline, so the same messages — same names, field numbers
and types, i.e. the same bytes on the wire — are declared here as a table and turned into message classes through
``descriptor_pb2`` + ``message_factory``.  ``api/grpc_service.py`` uses them as its second codec (JSON stays the first):
a request whose first byte is not a JSON opener is parsed as the method's protobuf request type, converted to the
dictionary the servicer expects, and the result is sent back as the method's protobuf response type — an existing
GUI / PhoneMgr client built from the reference's protos can talk to this server unchanged.

``task_json_to_pb`` / ``task_pb_to_json`` map the canonical task JSON to ``TaskConfig`` and back
(the proto twin renames a few fields: ``useController`` = ``use_gradient_house``, ``strategyBehaviorController`` =
``strategy_gradient_house``, ...: ols_core/taskMgr/utils/utils.py:831-1197).
"""
from __future__ import annotations

import json
from typing import Any, Callable, Dict, List, Optional, Tuple

from google.protobuf import descriptor_pb2, descriptor_pool, empty_pb2, json_format, message_factory

F = descriptor_pb2.FieldDescriptorProto
_T = {"string": F.TYPE_STRING, "int32": F.TYPE_INT32, "bool": F.TYPE_BOOL, "float": F.TYPE_FLOAT}

# file -> (package, dependencies, enums, messages).  A field is (name, number, type[, "repeated"]); message / enum types
# are written by name.
_FILES: Dict[str, Tuple[str, List[str], Dict[str, List[Tuple[str, int]]], Dict[str, List[tuple]]]] = {
    "taskService.proto": ("", ["google/protobuf/empty.proto"], {
        "FileTransferType": [("FILE", 0), ("HTTP", 1), ("S3", 2), ("MINIO", 3)],
        "TaskStatusEnum": [("SUCCEEDED", 0), ("PENDING", 1), ("RUNNING", 2), ("STOPPED", 3), ("FAILED", 4), ("MISSING", 5),
                           ("UNDONE", 6), ("QUEUED", 7)],
    }, {
        "TaskConfig": [("userID", 1, "string"), ("taskID", 2, "TaskID"), ("target", 3, "Target"), ("operatorFlow", 4, "OperatorFlow"),
                       ("logicalSimulation", 5, "LogicalSimulation"), ("deviceSimulation", 6, "DeviceSimulation")],
        "Target": [("targetData", 1, "TargetData", "repeated"), ("priority", 2, "int32")],
        "TargetData": [("dataName", 1, "string"), ("dataPath", 2, "string"), ("dataSplitType", 3, "bool"),
                       ("dataTransferType", 4, "FileTransferType"), ("taskType", 5, "string"),
                       ("totalSimulation", 6, "TotalSimulation"), ("allocation", 7, "Allocation")],
        "TotalSimulation": [("deviceTotalSimulation", 1, "string", "repeated"), ("numTotalSimulation", 2, "int32", "repeated"),
                            ("dynamicNumTotalSimulation", 3, "int32", "repeated")],
        "Allocation": [("optimization", 1, "bool"), ("allocationLogicalSimulation", 2, "int32", "repeated"),
                       ("allocationDeviceSimulation", 3, "int32", "repeated"), ("runningResponse", 4, "RunningResponse")],
        "RunningResponse": [("deviceRunningResponse", 1, "string", "repeated"), ("numRunningResponse", 2, "int32", "repeated")],
        "OperatorFlow": [("flowSetting", 1, "FlowSetting"), ("operator", 2, "Operator", "repeated")],
        "FlowSetting": [("round", 1, "int32"), ("startCondition", 2, "OperatorFlowCondition"),
                        ("stopCondition", 3, "OperatorFlowCondition")],
        "OperatorFlowCondition": [("logicalSimulationStrategy", 1, "StrategyCondition"),
                                  ("deviceSimulationStrategy", 2, "StrategyCondition")],
        "StrategyCondition": [("strategyCondition", 1, "string"), ("waitInterval", 2, "int32"), ("totalTimeout", 3, "int32")],
        "Operator": [("name", 1, "string"), ("operationBehaviorController", 2, "OperationBehaviorController"),
                     ("input", 3, "string", "repeated"), ("useData", 4, "bool"), ("model", 5, "Model"),
                     ("logicalSimulationOperatorInfo", 6, "OperatorSimulationInfo"),
                     ("deviceSimulationOperatorInfo", 7, "OperatorSimulationInfo")],
        "OperationBehaviorController": [("useController", 1, "bool"), ("strategyBehaviorController", 2, "string"),
                                        ("outboundService", 3, "string")],
        "Model": [("useModel", 1, "bool"), ("modelForTrain", 2, "bool"), ("modelTransferType", 3, "FileTransferType"),
                  ("modelPath", 4, "string"), ("modelUpdateStyle", 5, "string")],
        "OperatorSimulationInfo": [("operatorTransferType", 1, "FileTransferType"), ("operatorCodePath", 2, "string"),
                                   ("operatorEntryFile", 3, "string"), ("operatorParams", 4, "string")],
        "LogicalSimulation": [("computationUnit", 1, "ComputationUnit"),
                              ("resourceRequestLogicalSimulation", 2, "ResourceRequest", "repeated")],
        "ComputationUnit": [("devicesUnit", 1, "string", "repeated"), ("unitSetting", 2, "UnitSetting", "repeated")],
        "UnitSetting": [("numCpus", 1, "int32")],
        "ResourceRequest": [("dataNameResourceRequest", 1, "string"), ("deviceResourceRequest", 2, "string", "repeated"),
                            ("numResourceRequest", 3, "int32", "repeated")],
        "DeviceSimulation": [("resourceRequestDeviceSimulation", 1, "ResourceRequest", "repeated")],
        "TaskID": [("taskID", 1, "string")],
        "TaskQueue": [("tasks", 1, "TaskID", "repeated")],
        "TaskStatus": [("taskStatus", 1, "TaskStatusEnum")],
        "OperationStatus": [("is_success", 1, "bool")],
        "Scheduler": [("scheduler", 1, "string")],
        "DeviceTaskConfig": [("userID", 1, "string"), ("taskID", 2, "string"), ("deviceTargetData", 3, "DeviceTargetData", "repeated"),
                             ("deviceOperatorFlow", 4, "DeviceOperatorFlow"),
                             ("resourceRequestDeviceSimulation", 5, "ResourceRequest", "repeated")],
        "DeviceTargetData": [("dataName", 1, "string"), ("dataPath", 2, "string"), ("dataSplitType", 3, "bool"),
                             ("dataTransferType", 4, "FileTransferType"), ("taskType", 5, "string"),
                             ("deviceTotalSimulation", 6, "DeviceTotalSimulation"), ("runningResponse", 7, "RunningResponse")],
        "DeviceTotalSimulation": [("devices", 1, "string", "repeated"), ("nums", 2, "int32", "repeated")],
        "DeviceOperatorFlow": [("deviceFlowSetting", 1, "DeviceFlowSetting"), ("deviceOperator", 2, "DeviceOperator", "repeated")],
        "DeviceFlowSetting": [("round", 1, "int32"), ("deviceStartCondition", 2, "StrategyCondition"),
                              ("deviceStopCondition", 3, "StrategyCondition")],
        "DeviceOperator": [("name", 1, "string"), ("operationBehaviorController", 2, "OperationBehaviorController"),
                           ("input", 3, "string", "repeated"), ("useData", 4, "bool"), ("model", 5, "Model"),
                           ("deviceSimulationOperatorInfo", 6, "OperatorSimulationInfo")],
    }),
    "phoneMgr.proto": ("", ["google/protobuf/empty.proto", "taskService.proto"], {}, {
        "DeviceTaskResult": [("isFinished", 1, "bool"), ("maxRound", 2, "int32"), ("round", 3, "int32"), ("operator", 4, "string"),
                             ("deviceDataStatus", 5, "DeviceDataStatus", "repeated")],
        "DeviceDataStatus": [("name", 1, "string"), ("deviceType", 2, "string", "repeated"), ("successNum", 3, "int32", "repeated"),
                             ("failedNum", 4, "int32", "repeated")],
        "AllUsersDeviceAvailableResource": [("userDeviceAvailableResource", 1, "UserDeviceAvailableResource", "repeated")],
        "UserDeviceAvailableResource": [("userID", 1, "string"), ("deviceResourceInfo", 2, "DeviceResourceInfo", "repeated")],
        "DeviceResource": [("taskID", 1, "string"), ("userID", 2, "string"), ("deviceResourceInfo", 3, "DeviceResourceInfo", "repeated")],
        "DeviceResourceInfo": [("phoneType", 1, "string"), ("num", 2, "int32")],
        "ActionStatus": [("isSuccess", 1, "bool")],
        "MSPDevicesConfig": [("mspDeviceInfo", 1, "MSPDeviceInfo", "repeated")],
        "MSPDeviceInfo": [("deviceID", 1, "string"), ("deviceIP", 2, "string"), ("devicePort", 3, "string"), ("pubNetIP", 4, "string"),
                          ("pubNetPort", 5, "string")],
        "DevicesInfo": [("deviceInfo", 1, "DeviceInfo", "repeated")],
        "DeviceInfo": [("serialNumber", 1, "string"), ("model", 2, "string"), ("cpuCoreNumber", 3, "int32"), ("memory", 4, "int32")],
    }),
    "resourceService.proto": ("", ["google/protobuf/empty.proto", "phoneMgr.proto"], {
        "ServerClusterResStatus": [("SUCCESS", 0), ("NOT_ENOUGH_RES", 1), ("TASKID_NULL", 2), ("TASK_REPEAT", 3),
                                   ("TASKID_NOT_FOUND", 4), ("FAIL", 5)],
    }, {
        "Resource": [("clusterRes", 1, "ServerClusterResource"), ("phoneRes", 2, "AllUsersDeviceAvailableResource")],
        "ServerClusterResource": [("cores", 1, "float"), ("mem", 2, "float")],
        "VMClusterResource": [("availablePhones", 1, "int32")],
        "PhoneClusterResource": [("avaiablePhones", 1, "int32")],
        "ResourceMessage": [("server", 1, "ServerClusterResource"), ("VM", 2, "VMClusterResource"), ("phone", 3, "PhoneClusterResource")],
        "ResReq": [("type", 1, "int32"), ("clusterReq", 2, "ServerClusterReq"), ("phoneReq", 3, "DeviceResource")],
        "ResRelease": [("type", 1, "int32"), ("taskId", 2, "string")],
        "ServerClusterReq": [("taskId", 1, "string"), ("cores", 2, "float"), ("mem", 3, "float")],
        "RequestStatus": [("clusterStatus", 1, "ServerClusterStatus"), ("phoneStatus", 2, "ActionStatus")],
        "ReleaseStatus": [("clusterStatus", 1, "ServerClusterStatus"), ("phoneStatus", 2, "ActionStatus")],
        "ServerClusterStatus": [("status", 1, "ServerClusterResStatus")],
        "ServerClusterDetail": [("detail", 1, "string")],
    }),
    "deviceflow.proto": ("deviceflow", ["google/protobuf/empty.proto"], {}, {
        "NotifyRequest": [("task_id", 1, "string"), ("routing_key", 2, "string"), ("outbound_service", 3, "OutboundService"),
                          ("strategy", 4, "string"), ("compute_resource", 5, "string")],
        "NofifyCompleteRequest": [("task_id", 1, "string"), ("routing_key", 2, "string"), ("compute_resource", 3, "string")],
        "OutboundService": [("pulsar_client", 1, "PulsarClient"), ("websocket", 2, "Websocket")],
        "PulsarClient": [("url", 1, "string"), ("topic", 2, "string")],
        "Websocket": [("url", 1, "string")],
        "OperationResponse": [("is_success", 1, "bool")],
        "TaskInfo": [("total_compute_resources", 1, "string", "repeated")],
        "RegisterRequest": [("task_id", 1, "string"), ("total_compute_resources", 2, "string", "repeated")],
        "UnRegisterRequest": [("task_id", 1, "string")],
        "TotalComputeResourcesRequest": [("task_id", 1, "string")],
        "TotalComputeResourcesResponse": [("total_compute_resources", 1, "string", "repeated")],
        "CheckDeviceflowDispatchRequest": [("task_id", 1, "string")],
    }),
    "rayclusterService.proto": ("", ["google/protobuf/empty.proto"], {}, {
        "RayClusterActResponse": [("ok", 1, "bool")],
        "RayClusterParam": [("k8s_namespace", 1, "string"), ("ray_label", 2, "string")],
        "RayClusterQueryResult": [("json_data", 1, "string")],
        "RayClusterModifyParam": [("k8s_namespace", 1, "string"), ("ray_name", 2, "string"), ("ray_label", 3, "string"),
                                  ("worker_min_replicas", 4, "int32"), ("worker_replicas", 5, "int32"),
                                  ("worker_max_replicas", 6, "int32")],
    }),
}

# service -> method -> (request message, response message); "Empty" = google.protobuf.Empty
SERVICES: Dict[str, Dict[str, Tuple[str, str]]] = {
    "TaskMgr": {"submitTask": ("TaskConfig", "OperationStatus"), "stopTask": ("TaskID", "OperationStatus"),
                "getTaskStatus": ("TaskID", "TaskStatus"), "getTaskQueue": ("Empty", "TaskQueue"),
                "changeScheduler": ("Scheduler", "OperationStatus")},
    "ResourceMgr": {"getClusterAvailableResource": ("Empty", "ServerClusterResource"),
                    "getClusterResourceDetail": ("Empty", "ServerClusterDetail"),
                    "requestClusterResource": ("ServerClusterReq", "ServerClusterStatus"),
                    "releaseClusterResource": ("ServerClusterReq", "ServerClusterStatus"),
                    "getClusterTotalResource": ("Empty", "ServerClusterResource"),
                    "getVMResource": ("Empty", "VMClusterResource"), "getPhoneResource": ("Empty", "PhoneClusterResource"),
                    "getTotalResource": ("Empty", "ResourceMessage"), "getResource": ("Empty", "Resource"),
                    "requestResource": ("ResReq", "RequestStatus"), "releaseResource": ("ResRelease", "ReleaseStatus")},
    "TaskOperatorOrientedDeviceFlow": {
        "GetDeviceflowPulsarClient": ("Empty", "deviceflow.PulsarClient"), "GetDeviceflowWebsocket": ("Empty", "deviceflow.Websocket"),
        "NotifyStart": ("deviceflow.NotifyRequest", "deviceflow.OperationResponse"),
        "NotifyComplete": ("deviceflow.NofifyCompleteRequest", "deviceflow.OperationResponse"),
        "RegisterTask": ("deviceflow.RegisterRequest", "deviceflow.OperationResponse"),
        "UnRegisterTask": ("deviceflow.UnRegisterRequest", "deviceflow.OperationResponse"),
        "GetTotalComputeResources": ("deviceflow.TotalComputeResourcesRequest", "deviceflow.TotalComputeResourcesResponse"),
        "CheckDeviceflowDispatchFinished": ("deviceflow.CheckDeviceflowDispatchRequest", "deviceflow.OperationResponse")},
    "RayClusterMgr": {"createRayCluster": ("Empty", "RayClusterActResponse"), "deleteRayCluster": ("RayClusterParam", "RayClusterActResponse"),
                      "modifyRayCluster": ("RayClusterModifyParam", "RayClusterActResponse"),
                      "queryRayCluster": ("RayClusterParam", "RayClusterQueryResult")},
    "PerformanceMgr": {"commitJob": ("TaskConfig", "OperationStatus")},
    "TaskManager": {"submitTask": ("DeviceTaskConfig", "ActionStatus"),
                    "getDeviceAvailableResource": ("Empty", "AllUsersDeviceAvailableResource"),
                    "requestDeviceResource": ("DeviceResource", "ActionStatus"), "releaseDeviceResource": ("TaskID", "ActionStatus"),
                    "stopDevice": ("TaskID", "ActionStatus"), "getDeviceTaskStatus": ("TaskID", "DeviceTaskResult")},
    "PhoneManager": {"connectMSPDevices": ("MSPDevicesConfig", "Empty"), "getLocalDeviceResource": ("Empty", "DevicesInfo")},
}
# services the reference declares inside a proto package: clients address them as /<package>.<Service>/<method>
SERVICE_PACKAGE = {"TaskOperatorOrientedDeviceFlow": "deviceflow"}

_pool = descriptor_pool.DescriptorPool()
_classes: Dict[str, Any] = {}


def _build() -> None:
    if _classes:
        return
    _pool.AddSerializedFile(empty_pb2.DESCRIPTOR.serialized_pb)
    known: Dict[str, str] = {}                      # short type name -> fully qualified
    for fname, (pkg, deps, enums, msgs) in _FILES.items():
        fd = descriptor_pb2.FileDescriptorProto(name=fname, package=pkg, syntax="proto3", dependency=deps)
        q = (lambda n, p=pkg: f".{p}.{n}" if p else f".{n}")
        for en, vals in enums.items():
            e = fd.enum_type.add(name=en)
            for vn, vv in vals:
                e.value.add(name=vn, number=vv)
            known[en] = q(en)
        for mn in msgs:
            known.setdefault(mn, q(mn)) if not pkg else None
        local = {mn: q(mn) for mn in msgs}
        for mn, fields in msgs.items():
            m = fd.message_type.add(name=mn)
            for f in fields:
                name, num, typ = f[0], f[1], f[2]
                fld = m.field.add(name=name, number=num, json_name=name,
                                  label=F.LABEL_REPEATED if len(f) > 3 else F.LABEL_OPTIONAL)
                if typ in _T:
                    fld.type = _T[typ]
                else:
                    full = local.get(typ) or known.get(typ)
                    if full is None:
                        raise KeyError(f"{fname}: unknown type {typ}")
                    is_enum = any(typ in en for _, (_, _, en, _) in _FILES.items())
                    fld.type = F.TYPE_ENUM if is_enum else F.TYPE_MESSAGE
                    fld.type_name = full
        _pool.Add(fd)
        for mn in msgs:
            full = f"{pkg}.{mn}" if pkg else mn
            _classes[full] = message_factory.GetMessageClass(_pool.FindMessageTypeByName(full))
    _classes["Empty"] = empty_pb2.Empty


def msg(name: str):
    """Message class by name (``"TaskConfig"``, ``"deviceflow.NotifyRequest"``, ``"Empty"``)."""
    _build()
    return _classes[name]


def to_dict(m: Any) -> Dict[str, Any]:
    return json_format.MessageToDict(m, preserving_proto_field_name=True, always_print_fields_with_no_presence=True)


def from_dict(name: str, d: Dict[str, Any]):
    return json_format.ParseDict(d, msg(name)(), ignore_unknown_fields=True)


def looks_like_json(b: bytes) -> bool:
    """JSON documents start with one of `{ [ " digit - t f n` (after whitespace); none of those bytes is a valid first
    tag of the request messages above (0x7b = field 15 / start-group, ...), so the two codecs cannot be confused."""
    s = b.lstrip()[:1]
    return bool(s) and s in b'{["-0123456789tfn'


# ------------------------------------------------------------------------------------------------------------
# task JSON <-> TaskConfig
# ------------------------------------------------------------------------------------------------------------
def _ftt(v: Any) -> str:
    from .schema import FileTransferType
    return FileTransferType.parse(v if v not in ("", None) else 0).name


def _cond(c: Dict[str, Any]) -> Dict[str, Any]:
    return {"strategyCondition": c.get("strategy", ""), "waitInterval": int(c.get("wait_interval", 0)),
            "totalTimeout": int(c.get("total_timeout", 0))}


def _cond_back(c: Dict[str, Any]) -> Dict[str, Any]:
    return {"strategy": c.get("strategyCondition", ""), "wait_interval": c.get("waitInterval", 0), "total_timeout": c.get("totalTimeout", 0)}


def _sim_info(i: Dict[str, Any]) -> Dict[str, Any]:
    return {"operatorTransferType": _ftt(i.get("operator_transfer_type", 0)), "operatorCodePath": i.get("operator_code_path", ""),
            "operatorEntryFile": i.get("operator_entry_file", ""), "operatorParams": i.get("operator_params", "") or ""}


def _sim_info_back(i: Dict[str, Any]) -> Dict[str, Any]:
    return {"operator_transfer_type": i.get("operatorTransferType", "FILE"), "operator_code_path": i.get("operatorCodePath", ""),
            "operator_entry_file": i.get("operatorEntryFile", ""), "operator_params": i.get("operatorParams", "")}


def _rr(r: Dict[str, Any]) -> Dict[str, Any]:
    return {"dataNameResourceRequest": r.get("name", ""), "deviceResourceRequest": list(r.get("devices", [])),
            "numResourceRequest": [int(n) for n in r.get("num_request", [])]}


def _rr_back(r: Dict[str, Any]) -> Dict[str, Any]:
    return {"name": r.get("dataNameResourceRequest", ""), "devices": list(r.get("deviceResourceRequest", [])),
            "num_request": list(r.get("numResourceRequest", []))}


def task_json_to_dict(tj: Dict[str, Any]) -> Dict[str, Any]:
    """canonical task JSON -> the dict form of ``TaskConfig`` (proto field names)"""
    of = tj.get("operatorflow", {})
    fs = of.get("flow_setting", {})
    ls, ds = tj.get("logical_simulation", {}), tj.get("device_simulation", {})
    cu = ls.get("computation_unit", {})
    return {
        "userID": tj.get("user_id", ""), "taskID": {"taskID": tj.get("task_id", "")},
        "target": {"priority": int(tj.get("target", {}).get("priority", 0)), "targetData": [{
            "dataName": d.get("name", ""), "dataPath": d.get("data_path", ""), "dataSplitType": bool(d.get("data_split_type", False)),
            "dataTransferType": _ftt(d.get("data_transfer_type", 0)), "taskType": d.get("task_type", ""),
            "totalSimulation": {"deviceTotalSimulation": list(d.get("total_simulation", {}).get("devices", [])),
                                "numTotalSimulation": [int(n) for n in d.get("total_simulation", {}).get("nums", [])],
                                "dynamicNumTotalSimulation": [int(n) for n in d.get("total_simulation", {}).get("dynamic_nums", [])]},
            "allocation": {"optimization": bool(d.get("allocation", {}).get("optimization", False)),
                           "allocationLogicalSimulation": [int(n) for n in d.get("allocation", {}).get("logical_simulation", [])],
                           "allocationDeviceSimulation": [int(n) for n in d.get("allocation", {}).get("device_simulation", [])],
                           "runningResponse": {
                               "deviceRunningResponse": list(d.get("allocation", {}).get("running_response", {}).get("devices", [])),
                               "numRunningResponse": [int(n) for n in d.get("allocation", {}).get("running_response", {}).get("nums", [])]}},
        } for d in tj.get("target", {}).get("data", [])]},
        "operatorFlow": {
            "flowSetting": {"round": int(fs.get("round", 0)),
                            "startCondition": {"logicalSimulationStrategy": _cond(fs.get("start", {}).get("logical_simulation", {})),
                                               "deviceSimulationStrategy": _cond(fs.get("start", {}).get("device_simulation", {}))},
                            "stopCondition": {"logicalSimulationStrategy": _cond(fs.get("stop", {}).get("logical_simulation", {})),
                                              "deviceSimulationStrategy": _cond(fs.get("stop", {}).get("device_simulation", {}))}},
            "operator": [{
                "name": o.get("name", ""), "input": list(o.get("input", [])), "useData": bool(o.get("use_data", False)),
                "operationBehaviorController": {
                    "useController": bool(o.get("operation_behavior_controller", {}).get("use_gradient_house", False)),
                    "strategyBehaviorController": o.get("operation_behavior_controller", {}).get("strategy_gradient_house", "") or "",
                    "outboundService": o.get("operation_behavior_controller", {}).get("outbound_service", "") or ""},
                "model": {"useModel": bool(o.get("model", {}).get("use_model", False)),
                          "modelForTrain": bool(o.get("model", {}).get("model_for_train", False)),
                          "modelTransferType": _ftt(o.get("model", {}).get("model_transfer_type", 0)),
                          "modelPath": o.get("model", {}).get("model_path", ""),
                          "modelUpdateStyle": o.get("model", {}).get("model_update_style", "")},
                "logicalSimulationOperatorInfo": _sim_info(o.get("logical_simulation", {})),
                "deviceSimulationOperatorInfo": _sim_info(o.get("device_simulation", {})),
            } for o in of.get("operators", [])]},
        "logicalSimulation": {"computationUnit": {"devicesUnit": list(cu.get("devices", [])),
                                                  "unitSetting": [{"numCpus": int(s.get("num_cpus", 1))} for s in cu.get("setting", [])]},
                              "resourceRequestLogicalSimulation": [_rr(r) for r in ls.get("resource_request", [])]},
        "deviceSimulation": {"resourceRequestDeviceSimulation": [_rr(r) for r in ds.get("resource_request", [])]},
    }


def task_json_to_pb(tj: Dict[str, Any]):
    return from_dict("TaskConfig", task_json_to_dict(tj))


def task_pb_to_json(m: Any) -> Dict[str, Any]:
    d = to_dict(m) if not isinstance(m, dict) else m
    fs = d.get("operatorFlow", {}).get("flowSetting", {})
    sc, ec = fs.get("startCondition", {}), fs.get("stopCondition", {})
    cu = d.get("logicalSimulation", {}).get("computationUnit", {})
    return {
        "user_id": d.get("userID", ""), "task_id": d.get("taskID", {}).get("taskID", ""),
        "target": {"priority": d.get("target", {}).get("priority", 0), "data": [{
            "name": t.get("dataName", ""), "data_path": t.get("dataPath", ""), "data_split_type": t.get("dataSplitType", False),
            "data_transfer_type": t.get("dataTransferType", "FILE"), "task_type": t.get("taskType", ""),
            "total_simulation": {"devices": list(t.get("totalSimulation", {}).get("deviceTotalSimulation", [])),
                                 "nums": list(t.get("totalSimulation", {}).get("numTotalSimulation", [])),
                                 "dynamic_nums": list(t.get("totalSimulation", {}).get("dynamicNumTotalSimulation", []))},
            "allocation": {"optimization": t.get("allocation", {}).get("optimization", False),
                           "logical_simulation": list(t.get("allocation", {}).get("allocationLogicalSimulation", [])),
                           "device_simulation": list(t.get("allocation", {}).get("allocationDeviceSimulation", [])),
                           "running_response": {
                               "devices": list(t.get("allocation", {}).get("runningResponse", {}).get("deviceRunningResponse", [])),
                               "nums": list(t.get("allocation", {}).get("runningResponse", {}).get("numRunningResponse", []))}},
        } for t in d.get("target", {}).get("targetData", [])]},
        "operatorflow": {
            "flow_setting": {"round": fs.get("round", 0),
                             "start": {"logical_simulation": _cond_back(sc.get("logicalSimulationStrategy", {})),
                                       "device_simulation": _cond_back(sc.get("deviceSimulationStrategy", {}))},
                             "stop": {"logical_simulation": _cond_back(ec.get("logicalSimulationStrategy", {})),
                                      "device_simulation": _cond_back(ec.get("deviceSimulationStrategy", {}))}},
            "operators": [{
                "name": o.get("name", ""), "input": list(o.get("input", [])),
                "operation_behavior_controller": {
                    "use_gradient_house": o.get("operationBehaviorController", {}).get("useController", False),
                    "strategy_gradient_house": o.get("operationBehaviorController", {}).get("strategyBehaviorController", ""),
                    "outbound_service": o.get("operationBehaviorController", {}).get("outboundService", "")},
                "use_data": o.get("useData", False),
                "model": {"use_model": o.get("model", {}).get("useModel", False), "model_for_train": o.get("model", {}).get("modelForTrain", False),
                          "model_transfer_type": o.get("model", {}).get("modelTransferType", "FILE"),
                          "model_path": o.get("model", {}).get("modelPath", ""),
                          "model_update_style": o.get("model", {}).get("modelUpdateStyle", "")},
                "logical_simulation": _sim_info_back(o.get("logicalSimulationOperatorInfo", {})),
                "device_simulation": _sim_info_back(o.get("deviceSimulationOperatorInfo", {})),
            } for o in d.get("operatorFlow", {}).get("operator", [])]},
        "logical_simulation": {"computation_unit": {"devices": list(cu.get("devicesUnit", [])),
                                                    "setting": [{"num_cpus": s.get("numCpus", 1)} for s in cu.get("unitSetting", [])]},
                               "resource_request": [_rr_back(r) for r in d.get("logicalSimulation", {}).get("resourceRequestLogicalSimulation", [])]},
        "device_simulation": {"resource_request": [_rr_back(r) for r in d.get("deviceSimulation", {}).get("resourceRequestDeviceSimulation", [])]},
    }


# ------------------------------------------------------------------------------------------------------------
# per-method adapters: protobuf request dict -> servicer request dict, servicer result -> protobuf response dict
# ------------------------------------------------------------------------------------------------------------
def _phones(infos: List[Dict[str, Any]]) -> Dict[str, int]:
    return {i.get("phoneType", ""): int(i.get("num", 0)) for i in infos or []}


def _phone_res(avail: Dict[str, Dict[str, int]]) -> Dict[str, Any]:
    return {"userDeviceAvailableResource": [{"userID": u, "deviceResourceInfo": [{"phoneType": t, "num": int(n)} for t, n in v.items()]}
                                            for u, v in (avail or {}).items()]}


def _cluster(r: Dict[str, Any]) -> Dict[str, Any]:
    return {"cores": float(r.get("cpu", 0.0)), "mem": float(r.get("mem", 0.0))}


REQ: Dict[Tuple[str, str], Callable[[Dict[str, Any]], Dict[str, Any]]] = {
    ("TaskMgr", "submitTask"): lambda d: {"task": task_pb_to_json(d)},
    ("PerformanceMgr", "commitJob"): lambda d: task_pb_to_json(d),
    ("TaskMgr", "changeScheduler"): lambda d: {"strategy": d.get("scheduler", "")},
    ("ResourceMgr", "requestClusterResource"): lambda d: {"task_id": d.get("taskId", ""), "cpu": d.get("cores", 0.0), "mem": d.get("mem", 0.0)},
    ("ResourceMgr", "releaseClusterResource"): lambda d: {"task_id": d.get("taskId", "")},
    ("ResourceMgr", "requestResource"): lambda d: {
        "task_id": d.get("clusterReq", {}).get("taskId") or d.get("phoneReq", {}).get("taskID", ""),
        "user_id": d.get("phoneReq", {}).get("userID", ""), "cpu": d.get("clusterReq", {}).get("cores", 0.0),
        "mem": d.get("clusterReq", {}).get("mem", 0.0), "phones": _phones(d.get("phoneReq", {}).get("deviceResourceInfo")),
        "type": d.get("type", 0)},
    ("ResourceMgr", "releaseResource"): lambda d: {"task_id": d.get("taskId", ""), "type": d.get("type", 0)},
    ("TaskManager", "submitTask"): lambda d: {"task": d},
    ("TaskManager", "requestDeviceResource"): lambda d: {"task_id": d.get("taskID", ""), "user_id": d.get("userID", ""),
                                                         "phones": _phones(d.get("deviceResourceInfo"))},
    ("TaskManager", "releaseDeviceResource"): lambda d: {"task_id": d.get("taskID", "")},
    ("TaskManager", "stopDevice"): lambda d: {"task_id": d.get("taskID", "")},
    ("TaskManager", "getDeviceTaskStatus"): lambda d: {"task_id": d.get("taskID", "")},
    ("PhoneManager", "connectMSPDevices"): lambda d: {"user_id": "msp", "devices": {i.get("deviceID", ""): 1 for i in d.get("mspDeviceInfo", [])}},
}


def _status_pair(r: Dict[str, Any]) -> Dict[str, Any]:
    st = r.get("status", "SUCCESS")
    ok = bool(r.get("is_success", False))
    cluster = st if st in ("SUCCESS", "NOT_ENOUGH_RES", "TASKID_NULL", "TASK_REPEAT", "TASKID_NOT_FOUND", "FAIL") else ("SUCCESS" if ok else "FAIL")
    return {"clusterStatus": {"status": cluster}, "phoneStatus": {"isSuccess": ok or st != "PHONE_NOT_ENOUGH_RES" and cluster == "SUCCESS"}}


def _device_result(r: Dict[str, Any]) -> Dict[str, Any]:
    return {"isFinished": bool(r.get("is_finished", False)), "round": int(r.get("device_round", 0) or 0),
            "operator": r.get("device_operator", "") or "",
            "deviceDataStatus": [{"name": x.get("name", ""), "deviceType": list(x.get("simulation_target", {}).get("devices", [])),
                                  "successNum": list(x.get("simulation_target", {}).get("success_num", [])),
                                  "failedNum": list(x.get("simulation_target", {}).get("failed_num", []))}
                                 for x in r.get("device_result", []) or []]}


RESP: Dict[Tuple[str, str], Callable[[Any], Dict[str, Any]]] = {
    ("TaskMgr", "getTaskQueue"): lambda r: {"tasks": [{"taskID": t if isinstance(t, str) else t.get("task_id", "")} for t in r.get("tasks", [])]},
    ("ResourceMgr", "getClusterAvailableResource"): _cluster,
    ("ResourceMgr", "getClusterTotalResource"): _cluster,
    ("ResourceMgr", "getClusterResourceDetail"): lambda r: {"detail": json.dumps(r.get("details", []), default=lambda o: getattr(o, "__dict__", str(o)))},
    ("ResourceMgr", "getResource"): lambda r: {"clusterRes": _cluster(r.get("cluster", {})), "phoneRes": _phone_res(r.get("phone", {}))},
    ("ResourceMgr", "getVMResource"): lambda r: {"availablePhones": 0},
    ("ResourceMgr", "getPhoneResource"): lambda r: {"avaiablePhones": int(sum(sum(v.values()) for v in r.values()) if r else 0)},
    ("ResourceMgr", "getTotalResource"): lambda r: {"server": _cluster(r.get("cluster", {})), "VM": {"availablePhones": 0},
                                                   "phone": {"avaiablePhones": int(sum(sum(v.values()) for v in r.get("phone", {}).values()))}},
    ("ResourceMgr", "requestResource"): _status_pair,
    ("ResourceMgr", "releaseResource"): _status_pair,
    ("TaskManager", "submitTask"): lambda r: {"isSuccess": bool(r.get("is_success"))},
    ("TaskManager", "requestDeviceResource"): lambda r: {"isSuccess": bool(r.get("is_success"))},
    ("TaskManager", "releaseDeviceResource"): lambda r: {"isSuccess": bool(r.get("is_success"))},
    ("TaskManager", "stopDevice"): lambda r: {"isSuccess": bool(r.get("is_success"))},
    ("TaskManager", "getDeviceAvailableResource"): lambda r: _phone_res(r.get("resources", {})),
    ("TaskManager", "getDeviceTaskStatus"): _device_result,
    ("PhoneManager", "connectMSPDevices"): lambda r: {},
    ("PhoneManager", "getLocalDeviceResource"): lambda r: {"deviceInfo": [{"serialNumber": f"{u}:{t}", "model": t, "cpuCoreNumber": 0, "memory": 0}
                                                                         for u, v in r.get("resources", {}).items() for t in v]},
}


def decode_request(service: str, method: str, payload: bytes) -> Optional[Dict[str, Any]]:
    """protobuf request bytes -> the dictionary the JSON servicer of (service, method) expects; None if unknown method"""
    spec = SERVICES.get(service, {}).get(method)
    if spec is None:
        return None
    m = msg(spec[0])()
    m.ParseFromString(payload)
    d = to_dict(m)
    f = REQ.get((service, method))
    return f(d) if f else d


def encode_response(service: str, method: str, result: Any) -> bytes:
    spec = SERVICES[service][method]
    f = RESP.get((service, method))
    d = f(result if isinstance(result, dict) else {"value": result}) if f else (result if isinstance(result, dict) else {})
    return from_dict(spec[1], d).SerializeToString()
