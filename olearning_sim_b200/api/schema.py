"""TaskConfig — the canonical task description, as typed dataclasses with a lossless JSON round-trip.

The JSON shape is the reference's task schema (``json2taskconfig`` / ``taskconfig2json`` in
ols_core/taskMgr/utils/utils.py:831-1197; proto twin ols_core/proto/taskService.proto:4-125) so an existing
task file can be submitted unchanged.  Differences: plain dataclasses instead of generated protobuf classes,
enum values looked up by table (never ``eval``), and an optional ``engine`` block (ignored by the reference)
that selects the GPU-native FL operator instead of an opaque script.
"""
from __future__ import annotations

import copy
import json
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Any, Dict, List, Optional


class FileTransferType(IntEnum):
    """ols_core/proto/taskService.proto:131-136"""
    FILE = 0
    HTTP = 1
    S3 = 2
    MINIO = 3

    @classmethod
    def parse(cls, v: Any) -> "FileTransferType":
        if isinstance(v, cls):
            return v
        if isinstance(v, int) and not isinstance(v, bool):
            return cls(v)
        if isinstance(v, str) and v.upper() in cls.__members__:
            return cls[v.upper()]
        raise ValueError(f"invalid FileTransferType {v!r}")


class TaskStatus:
    """ols_core/taskMgr/task_manager.py:41-49 / proto TaskStatusEnum (taskService.proto:138-147)"""
    SUCCEEDED = "SUCCEEDED"
    PENDING = "PENDING"
    RUNNING = "RUNNING"
    STOPPED = "STOPPED"
    FAILED = "FAILED"
    MISSING = "MISSING"
    UNDONE = "UNDONE"
    QUEUED = "QUEUED"
    ALL = ("SUCCEEDED", "PENDING", "RUNNING", "STOPPED", "FAILED", "MISSING", "UNDONE", "QUEUED")
    TERMINAL = ("SUCCEEDED", "STOPPED", "FAILED")


@dataclass
class TotalSimulation:
    devices: List[str] = field(default_factory=list)
    nums: List[int] = field(default_factory=list)
    dynamic_nums: List[int] = field(default_factory=list)


@dataclass
class RunningResponse:
    devices: List[str] = field(default_factory=list)
    nums: List[int] = field(default_factory=list)


@dataclass
class Allocation:
    optimization: bool = False
    logical_simulation: List[int] = field(default_factory=list)
    device_simulation: List[int] = field(default_factory=list)
    running_response: RunningResponse = field(default_factory=RunningResponse)


@dataclass
class TargetData:
    name: str = ""
    data_path: str = ""
    data_split_type: bool = False
    data_transfer_type: FileTransferType = FileTransferType.FILE
    task_type: str = "classification"
    total_simulation: TotalSimulation = field(default_factory=TotalSimulation)
    allocation: Allocation = field(default_factory=Allocation)


@dataclass
class Target:
    priority: int = 0
    data: List[TargetData] = field(default_factory=list)


@dataclass
class ConditionStrategy:
    strategy: str = ""
    wait_interval: int = 0
    total_timeout: int = 0


@dataclass
class Condition:
    logical_simulation: ConditionStrategy = field(default_factory=ConditionStrategy)
    device_simulation: ConditionStrategy = field(default_factory=ConditionStrategy)


@dataclass
class FlowSetting:
    round: int = 1
    start: Condition = field(default_factory=Condition)
    stop: Condition = field(default_factory=Condition)


@dataclass
class BehaviorController:
    use_gradient_house: bool = False
    strategy_gradient_house: str = ""
    outbound_service: str = ""


@dataclass
class ModelSpec:
    use_model: bool = False
    model_for_train: bool = True
    model_transfer_type: FileTransferType = FileTransferType.FILE
    model_path: str = ""
    model_update_style: str = ""


@dataclass
class OperatorInfo:
    operator_transfer_type: FileTransferType = FileTransferType.FILE
    operator_code_path: str = ""
    operator_entry_file: str = ""
    operator_params: str = ""


@dataclass
class Operator:
    name: str = ""
    input: List[str] = field(default_factory=list)
    operation_behavior_controller: BehaviorController = field(default_factory=BehaviorController)
    use_data: bool = True
    model: ModelSpec = field(default_factory=ModelSpec)
    logical_simulation: OperatorInfo = field(default_factory=OperatorInfo)
    device_simulation: OperatorInfo = field(default_factory=OperatorInfo)


@dataclass
class OperatorFlowSpec:
    flow_setting: FlowSetting = field(default_factory=FlowSetting)
    operators: List[Operator] = field(default_factory=list)


@dataclass
class UnitSetting:
    num_cpus: int = 1
    # B200 extension: fraction of one GPU's SMs / HBM a compute unit of this class may use (ledger accounting)
    gpu_fraction: float = 0.0


@dataclass
class ComputationUnit:
    devices: List[str] = field(default_factory=list)
    setting: List[UnitSetting] = field(default_factory=list)


@dataclass
class ResourceRequest:
    name: str = ""
    devices: List[str] = field(default_factory=list)
    num_request: List[int] = field(default_factory=list)


@dataclass
class LogicalSimulation:
    computation_unit: ComputationUnit = field(default_factory=ComputationUnit)
    resource_request: List[ResourceRequest] = field(default_factory=list)


@dataclass
class DeviceSimulation:
    resource_request: List[ResourceRequest] = field(default_factory=list)


@dataclass
class TaskConfig:
    user_id: str = ""
    task_id: str = ""
    target: Target = field(default_factory=Target)
    operatorflow: OperatorFlowSpec = field(default_factory=OperatorFlowSpec)
    logical_simulation: LogicalSimulation = field(default_factory=LogicalSimulation)
    device_simulation: DeviceSimulation = field(default_factory=DeviceSimulation)
    # optional, engine-specific (GPU-native operator selection); passes through untouched
    engine: Dict[str, Any] = field(default_factory=dict)

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def from_json(obj: Any) -> "TaskConfig":
        return json2taskconfig(obj)

    def to_json(self) -> Dict[str, Any]:
        return taskconfig2json(self)

    def to_json_str(self) -> str:
        return json.dumps(self.to_json())

    def operator_names(self) -> List[str]:
        return [op.name for op in self.operatorflow.operators]

    def data_names(self) -> List[str]:
        return [d.name for d in self.target.data]

    def uses_deviceflow(self) -> bool:
        return any(op.operation_behavior_controller.use_gradient_house for op in self.operatorflow.operators)


# ----------------------------------------------------------------------------------------------------
def _g(d: Any, key: str, default: Any) -> Any:
    if not isinstance(d, dict):
        return default
    v = d.get(key, default)
    return default if v is None else v


def _cond(d: Any) -> Condition:
    def one(x):
        return ConditionStrategy(strategy=_g(x, "strategy", ""), wait_interval=_g(x, "wait_interval", 0),
                                 total_timeout=_g(x, "total_timeout", 0))
    return Condition(logical_simulation=one(_g(d, "logical_simulation", {})),
                     device_simulation=one(_g(d, "device_simulation", {})))


def _opinfo(d: Any) -> OperatorInfo:
    return OperatorInfo(operator_transfer_type=FileTransferType.parse(_g(d, "operator_transfer_type", "FILE")),
                        operator_code_path=_g(d, "operator_code_path", ""),
                        operator_entry_file=_g(d, "operator_entry_file", ""),
                        operator_params=_as_json_str(_g(d, "operator_params", "")))


def _as_json_str(v: Any) -> str:
    """The schema stores nested structures as JSON *strings*; accept dicts for convenience."""
    if isinstance(v, (dict, list)):
        return json.dumps(v)
    return v


def json2taskconfig(obj: Any) -> TaskConfig:
    """Task JSON (str or dict) -> TaskConfig.  Unknown keys are ignored, missing keys take schema defaults.
    Raises ValueError on malformed enum names (the reference would raise from ``eval``)."""
    if isinstance(obj, (str, bytes)):
        obj = json.loads(obj)
    if not isinstance(obj, dict):
        raise ValueError("task description must be a JSON object")
    tc = TaskConfig(user_id=_g(obj, "user_id", ""), task_id=_g(obj, "task_id", ""))
    tgt = _g(obj, "target", {})
    tc.target.priority = _g(tgt, "priority", 0)
    for d in _g(tgt, "data", []):
        ts = _g(d, "total_simulation", {})
        al = _g(d, "allocation", {})
        rr = _g(al, "running_response", {})
        tc.target.data.append(TargetData(
            name=_g(d, "name", ""), data_path=_g(d, "data_path", ""), data_split_type=_g(d, "data_split_type", False),
            data_transfer_type=FileTransferType.parse(_g(d, "data_transfer_type", "FILE")),
            task_type=_g(d, "task_type", "classification"),
            total_simulation=TotalSimulation(list(_g(ts, "devices", [])), list(_g(ts, "nums", [])),
                                             list(_g(ts, "dynamic_nums", []))),
            allocation=Allocation(optimization=_g(al, "optimization", False),
                                  logical_simulation=list(_g(al, "logical_simulation", [])),
                                  device_simulation=list(_g(al, "device_simulation", [])),
                                  running_response=RunningResponse(list(_g(rr, "devices", [])),
                                                                   list(_g(rr, "nums", []))))))
    of = _g(obj, "operatorflow", {})
    fs = _g(of, "flow_setting", {})
    tc.operatorflow.flow_setting = FlowSetting(round=_g(fs, "round", 1), start=_cond(_g(fs, "start", {})),
                                               stop=_cond(_g(fs, "stop", {})))
    for o in _g(of, "operators", []):
        bc = _g(o, "operation_behavior_controller", {})
        m = _g(o, "model", {})
        tc.operatorflow.operators.append(Operator(
            name=_g(o, "name", ""), input=list(_g(o, "input", [])),
            operation_behavior_controller=BehaviorController(
                use_gradient_house=_g(bc, "use_gradient_house", False),
                strategy_gradient_house=_as_json_str(_g(bc, "strategy_gradient_house", "")),
                outbound_service=_as_json_str(_g(bc, "outbound_service", ""))),
            use_data=_g(o, "use_data", True),
            model=ModelSpec(use_model=_g(m, "use_model", False), model_for_train=_g(m, "model_for_train", True),
                            model_transfer_type=FileTransferType.parse(_g(m, "model_transfer_type", "FILE")),
                            model_path=_g(m, "model_path", ""), model_update_style=_g(m, "model_update_style", "")),
            logical_simulation=_opinfo(_g(o, "logical_simulation", {})),
            device_simulation=_opinfo(_g(o, "device_simulation", {}))))
    ls = _g(obj, "logical_simulation", {})
    cu = _g(ls, "computation_unit", {})
    tc.logical_simulation.computation_unit = ComputationUnit(
        devices=list(_g(cu, "devices", [])),
        setting=[UnitSetting(num_cpus=_g(s, "num_cpus", 1), gpu_fraction=_g(s, "gpu_fraction", 0.0))
                 for s in _g(cu, "setting", [])])
    tc.logical_simulation.resource_request = [
        ResourceRequest(_g(r, "name", ""), list(_g(r, "devices", [])), list(_g(r, "num_request", [])))
        for r in _g(ls, "resource_request", [])]
    ds = _g(obj, "device_simulation", {})
    tc.device_simulation.resource_request = [
        ResourceRequest(_g(r, "name", ""), list(_g(r, "devices", [])), list(_g(r, "num_request", [])))
        for r in _g(ds, "resource_request", [])]
    tc.engine = copy.deepcopy(_g(obj, "engine", {}))
    return tc


def taskconfig2json(tc: TaskConfig) -> Dict[str, Any]:
    def cond(c: Condition):
        one = lambda s: {"strategy": s.strategy, "wait_interval": s.wait_interval, "total_timeout": s.total_timeout}
        return {"logical_simulation": one(c.logical_simulation), "device_simulation": one(c.device_simulation)}

    def opinfo(i: OperatorInfo):
        return {"operator_transfer_type": FileTransferType(i.operator_transfer_type).name,
                "operator_code_path": i.operator_code_path, "operator_entry_file": i.operator_entry_file,
                "operator_params": i.operator_params}

    out: Dict[str, Any] = {
        "user_id": tc.user_id, "task_id": tc.task_id,
        "target": {"priority": tc.target.priority, "data": [{
            "name": d.name, "data_path": d.data_path, "data_split_type": d.data_split_type,
            "data_transfer_type": FileTransferType(d.data_transfer_type).name, "task_type": d.task_type,
            "total_simulation": {"devices": list(d.total_simulation.devices), "nums": list(d.total_simulation.nums),
                                 "dynamic_nums": list(d.total_simulation.dynamic_nums)},
            "allocation": {"optimization": d.allocation.optimization,
                           "logical_simulation": list(d.allocation.logical_simulation),
                           "device_simulation": list(d.allocation.device_simulation),
                           "running_response": {"devices": list(d.allocation.running_response.devices),
                                                "nums": list(d.allocation.running_response.nums)}}}
            for d in tc.target.data]},
        "operatorflow": {
            "flow_setting": {"round": tc.operatorflow.flow_setting.round,
                             "start": cond(tc.operatorflow.flow_setting.start),
                             "stop": cond(tc.operatorflow.flow_setting.stop)},
            "operators": [{
                "name": o.name, "input": list(o.input),
                "operation_behavior_controller": {
                    "use_gradient_house": o.operation_behavior_controller.use_gradient_house,
                    "strategy_gradient_house": o.operation_behavior_controller.strategy_gradient_house,
                    "outbound_service": o.operation_behavior_controller.outbound_service},
                "use_data": o.use_data,
                "model": {"use_model": o.model.use_model, "model_for_train": o.model.model_for_train,
                          "model_transfer_type": FileTransferType(o.model.model_transfer_type).name,
                          "model_path": o.model.model_path, "model_update_style": o.model.model_update_style},
                "logical_simulation": opinfo(o.logical_simulation),
                "device_simulation": opinfo(o.device_simulation)} for o in tc.operatorflow.operators]},
        "logical_simulation": {
            "computation_unit": {"devices": list(tc.logical_simulation.computation_unit.devices),
                                 "setting": [{"num_cpus": s.num_cpus, **({"gpu_fraction": s.gpu_fraction}
                                                                         if s.gpu_fraction else {})}
                                             for s in tc.logical_simulation.computation_unit.setting]},
            "resource_request": [{"name": r.name, "devices": list(r.devices), "num_request": list(r.num_request)}
                                 for r in tc.logical_simulation.resource_request]},
        "device_simulation": {
            "resource_request": [{"name": r.name, "devices": list(r.devices), "num_request": list(r.num_request)}
                                 for r in tc.device_simulation.resource_request]},
    }
    if tc.engine:
        out["engine"] = copy.deepcopy(tc.engine)
    return out
