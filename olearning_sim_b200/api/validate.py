"""Three-pass validation of a TaskConfig: types -> values -> cross-field relationships.

Rule-for-rule counterpart of ``ValidateParameters`` (ols_core/taskMgr/utils/utils.py:269-829): ``validate_type``
(:283-399), ``validate_correctness`` (:401-554), ``validate_relationship`` (:556-811).  Each pass returns
``(ok, message)`` and stops at the first violated rule, like the reference's assert chain.
"""
from __future__ import annotations

import json
import os
import re
from typing import Any, List, Tuple

from .schema import FileTransferType, Operator, TaskConfig


class ValidationError(AssertionError):
    pass


def _req(cond: bool, msg: str) -> None:
    if not cond:
        raise ValidationError(msg)


def is_in_ascii(s: str) -> bool:
    """printable ASCII only (32..126) — ols_core/taskMgr/utils/utils.py:273-274"""
    return all(31 < ord(ch) < 127 for ch in s)


def is_file_ext(s: str, ext: str) -> bool:
    return os.path.splitext(s)[1] == ext


def _all(kind, seq) -> bool:
    if kind is int:
        return all(isinstance(e, int) and not isinstance(e, bool) for e in seq)
    return all(isinstance(e, kind) for e in seq)


def _rearrange(devices: List[str], keys: List[str], values: List[Any], default: Any) -> List[Any]:
    d = dict(zip(keys, values))
    return [d.get(dev, default) for dev in devices]


_URL_DIR = re.compile(r"^[A-Za-z0-9_./:\-]+$")


class ValidateParameters:
    # ------------------------------------------------------------------------------------------
    def validate_type(self, t: TaskConfig) -> Tuple[bool, str]:
        try:
            _req(isinstance(t.user_id, str), "userID should be str")
            _req(isinstance(t.task_id, str), "taskID should be str")
            for i, d in enumerate(t.target.data):
                _req(isinstance(d.name, str), f"The name of No.{i} data in target should be str")
                n = d.name
                _req(isinstance(d.data_path, str), f"data_name={n} dataPath should be str")
                _req(isinstance(d.data_split_type, bool), f"data_name={n} dataSplitType should be bool")
                _req(isinstance(d.data_transfer_type, int), f"data_name={n} dataTransferType should be int")
                _req(isinstance(d.task_type, str), f"data_name={n} taskType should be str")
                _req(_all(str, d.total_simulation.devices), f"elements in data_name={n} totalSimulation.devices should be str")
                _req(_all(int, d.total_simulation.nums), f"elements in data_name={n} totalSimulation.nums should be int")
                _req(_all(int, d.total_simulation.dynamic_nums), f"elements in data_name={n} totalSimulation.dynamic_nums should be int")
                _req(isinstance(d.allocation.optimization, bool), f"data_name={n} allocation.optimization should be bool")
                _req(_all(int, d.allocation.logical_simulation), f"elements in data_name={n} allocation.logical_simulation should be int")
                _req(_all(int, d.allocation.device_simulation), f"elements in data_name={n} allocation.device_simulation should be int")
                _req(_all(str, d.allocation.running_response.devices), f"elements in data_name={n} runningResponse.devices should be str")
                _req(_all(int, d.allocation.running_response.nums), f"elements in data_name={n} runningResponse.nums should be int")
            _req(isinstance(t.target.priority, int) and not isinstance(t.target.priority, bool), "target.priority should be int")
            fs = t.operatorflow.flow_setting
            _req(isinstance(fs.round, int) and not isinstance(fs.round, bool), "operatorFlow.flowSetting.round should be int")
            for nm, c in (("start", fs.start), ("stop", fs.stop)):
                for side, s in (("logical", c.logical_simulation), ("device", c.device_simulation)):
                    _req(isinstance(s.strategy, str), f"flowSetting.{nm}.{side}.strategy should be str")
                    _req(isinstance(s.wait_interval, int), f"flowSetting.{nm}.{side}.waitInterval should be int")
                    _req(isinstance(s.total_timeout, int), f"flowSetting.{nm}.{side}.totalTimeout should be int")
            for i, o in enumerate(t.operatorflow.operators):
                _req(isinstance(o.name, str), f"The name of No.{i} operator in operatorFlow should be str")
                n = o.name
                bc = o.operation_behavior_controller
                _req(isinstance(bc.use_gradient_house, bool), f"operator={n} useController should be bool")
                _req(isinstance(bc.strategy_gradient_house, str), f"operator={n} strategyBehaviorController should be str")
                _req(isinstance(bc.outbound_service, str), f"operator={n} outboundService should be str")
                _req(_all(str, o.input), f"elements in operator={n} input should be str")
                _req(isinstance(o.use_data, bool), f"operator={n} useData should be bool")
                _req(isinstance(o.model.use_model, bool), f"operator={n} model.useModel should be bool")
                _req(isinstance(o.model.model_for_train, bool), f"operator={n} model.modelForTrain should be bool")
                _req(isinstance(o.model.model_transfer_type, int), f"operator={n} model.modelTransferType should be int")
                _req(isinstance(o.model.model_path, str), f"operator={n} model.modelPath should be str")
                _req(isinstance(o.model.model_update_style, str), f"operator={n} model.modelUpdateStyle should be str")
                for side, info in (("logical", o.logical_simulation), ("device", o.device_simulation)):
                    _req(isinstance(info.operator_transfer_type, int), f"operator={n} {side}.operatorTransferType should be int")
                    _req(isinstance(info.operator_code_path, str), f"operator={n} {side}.operatorCodePath should be str")
                    _req(isinstance(info.operator_entry_file, str), f"operator={n} {side}.operatorEntryFile should be str")
                    _req(isinstance(info.operator_params, str), f"operator={n} {side}.operatorParams should be str")
            cu = t.logical_simulation.computation_unit
            _req(_all(str, cu.devices), "elements in logicalSimulation.computationUnit.devicesUnit should be str")
            _req(all(isinstance(s.num_cpus, int) and not isinstance(s.num_cpus, bool) for s in cu.setting),
                 "numCpus in logicalSimulation.computationUnit.unitSetting should be int")
            for side, reqs in (("logicalSimulation", t.logical_simulation.resource_request),
                               ("deviceSimulation", t.device_simulation.resource_request)):
                for i, r in enumerate(reqs):
                    _req(isinstance(r.name, str), f"The name of No.{i} resource_request in {side} should be str")
                    _req(_all(str, r.devices), f"elements in {side} resource_request={r.name} devices should be str")
                    _req(_all(int, r.num_request), f"elements in {side} resource_request={r.name} num_request should be int")
        except ValidationError as e:
            return False, str(e)
        return True, "Pass"

    # ------------------------------------------------------------------------------------------
    def validate_correctness(self, t: TaskConfig) -> Tuple[bool, str]:
        try:
            _req(t.user_id != "", "userID should not be empty")
            _req(is_in_ascii(t.user_id), f"userID={t.user_id} contains illegal characters.")
            _req(t.task_id != "", "taskID should not be empty")
            _req(is_in_ascii(t.task_id), f"taskID={t.task_id} contains illegal characters.")
            valid_ft = set(int(x) for x in FileTransferType)
            for i, d in enumerate(t.target.data):
                _req(d.name != "", f"The name of No.{i} data in target should not be empty")
                _req(is_in_ascii(d.name), f"The name of No.{i} data in target = {d.name} contains illegal characters.")
                n = d.name
                if d.data_path:
                    _req(is_in_ascii(d.data_path), f"target data_name = {n}, dataPath={d.data_path} contains illegal characters.")
                    _req(is_file_ext(d.data_path, ".zip") or bool(_URL_DIR.match(d.data_path)),
                         f"target data_name = {n}, dataPath={d.data_path} should be a .zip file or a directory path.")
                _req(int(d.data_transfer_type) in valid_ft, f"The value of target_data.dataTransferType={d.data_transfer_type} is not valid.")
                _req(is_in_ascii(d.task_type), f"target data_name = {n}, taskType = {d.task_type} contains illegal characters.")
                devs = list(d.total_simulation.devices)
                _req(len(devs) > 0, f"target data_name = {n}, totalSimulation.devices length should larger than 0")
                _req(len(devs) == len(set(devs)), f"target data_name = {n}, totalSimulation.devices={devs} contains repeated elements!")
                _req(all(is_in_ascii(e) for e in devs), f"target data_name = {n}, elements in totalSimulation.devices={devs} contain illegal characters.")
                _req(all(e > 0 for e in d.total_simulation.nums), f"target data_name = {n}, elements in totalSimulation.nums={d.total_simulation.nums} should be larger than 0.")
                _req(all(e >= 0 for e in d.total_simulation.dynamic_nums), f"target data_name = {n}, elements in totalSimulation.dynamic_nums={d.total_simulation.dynamic_nums} should not be less than 0.")
                _req(all(e >= 0 for e in d.allocation.logical_simulation), f"target data_name = {n}, elements in allocation.logical_simulation should not be less than 0.")
                _req(all(e >= 0 for e in d.allocation.device_simulation), f"target data_name = {n}, elements in allocation.device_simulation should not be less than 0.")
                rr = d.allocation.running_response
                _req(all(is_in_ascii(e) for e in rr.devices), f"target data_name = {n}, elements in runningResponse.devices contain illegal characters.")
                _req(len(rr.devices) == len(set(rr.devices)), f"target data_name = {n}, runningResponse.devices={rr.devices} contains repeated elements!")
                _req(all(e >= 0 for e in rr.nums), f"target data_name = {n}, elements in runningResponse.nums should not be less than 0.")
            _req(0 <= t.target.priority <= 10, f"target.priority={t.target.priority} should be in range from 0 to 10")
            fs = t.operatorflow.flow_setting
            _req(fs.round > 0, f"operatorFlow.flowSetting.round={fs.round} should be larger than 0")
            for nm, c in (("startCondition", fs.start), ("stopCondition", fs.stop)):
                for side, s in (("logicalSimulationStrategy", c.logical_simulation), ("deviceSimulationStrategy", c.device_simulation)):
                    _req(is_in_ascii(s.strategy), f"flowSetting.{nm}.{side}.strategyCondition contains illegal characters.")
                    _req(s.wait_interval >= 0, f"flowSetting.{nm}.{side}.waitInterval = {s.wait_interval} should not be less than 0")
                    _req(s.total_timeout >= 0, f"flowSetting.{nm}.{side}.totalTimeout = {s.total_timeout} should not be less than 0")
            for i, o in enumerate(t.operatorflow.operators):
                _req(o.name != "", f"The name of No.{i} operator in operatorFlow should not be empty")
                _req(is_in_ascii(o.name), f"The name of No.{i} operator in operatorFlow = {o.name} contains illegal characters.")
                _req(" " not in o.name, f"The name of No.{i} operator in operatorFlow = {o.name} includes spaces, and that's not allowed.")
                n = o.name
                bc = o.operation_behavior_controller
                _req(is_in_ascii(bc.strategy_gradient_house), f"operator={n} strategyBehaviorController contains illegal characters.")
                _req(is_in_ascii(bc.outbound_service), f"operator={n} outboundService contains illegal characters.")
                _req(all(is_in_ascii(e) for e in o.input), f"elements in operator_name={n} input = {o.input} contain illegal characters.")
                _req(int(o.model.model_transfer_type) in valid_ft, f"operator_name={n}, the value of model.modelTransferType is not valid.")
                _req(is_in_ascii(o.model.model_path), f"operator_name={n}, model.modelPath contains illegal characters.")
                _req(is_in_ascii(o.model.model_update_style), f"operator_name={n}, model.modelUpdateStyle contains illegal characters.")
                li = o.logical_simulation
                _req(int(li.operator_transfer_type) in valid_ft, f"operator_name={n}, logicalSimulationOperatorInfo.operatorTransferType is not valid.")
                if li.operator_code_path != "":
                    _req(is_in_ascii(li.operator_code_path), f"logicalSimulationOperatorInfo.operatorCodePath of {n} contains illegal characters.")
                    _req(os.path.isdir(os.path.abspath(li.operator_code_path)) or is_file_ext(li.operator_code_path, ".zip")
                         or li.operator_code_path.startswith("builtin:"),
                         f"logicalSimulationOperatorInfo.operatorCodePath of {n} = {li.operator_code_path} should be an existing directory, a .zip file or a builtin: operator.")
                if li.operator_entry_file != "":
                    _req(is_in_ascii(li.operator_entry_file), f"logicalSimulationOperatorInfo.operatorEntryFile of {n} contains illegal characters.")
                    _req(is_file_ext(li.operator_entry_file, ".py"), f"logicalSimulationOperatorInfo.operatorEntryFile of {n} = {li.operator_entry_file} should be a .py file.")
                if li.operator_params:
                    _req(self._is_json(li.operator_params), f"operator={n} logicalSimulationOperatorInfo.operatorParams should be a json string")
                di = o.device_simulation
                _req(int(di.operator_transfer_type) in valid_ft, f"operator_name={n}, deviceSimulationOperatorInfo.operatorTransferType is not valid.")
                if di.operator_code_path != "":
                    _req(is_in_ascii(di.operator_code_path), f"deviceSimulationOperatorInfo.operatorCodePath of {n} contains illegal characters.")
                    _req(is_file_ext(di.operator_code_path, ".apk"), f"deviceSimulationOperatorInfo.operatorCodePath of {n} = {di.operator_code_path} should be a .apk file.")
                if di.operator_entry_file != "":
                    _req(is_in_ascii(di.operator_entry_file), f"deviceSimulationOperatorInfo.operatorEntryFile of {n} contains illegal characters.")
                    _req(is_file_ext(di.operator_entry_file, ".apk"), f"deviceSimulationOperatorInfo.operatorEntryFile of {n} = {di.operator_entry_file} should be a .apk file.")
                if di.operator_params:
                    _req(self._is_json(di.operator_params), f"operator={n} deviceSimulationOperatorInfo.operatorParams should be a json string")
            cu = t.logical_simulation.computation_unit
            _req(len(cu.devices) == len(set(cu.devices)), f"logicalSimulation, computationUnit.devicesUnit={cu.devices} contains repeated elements!")
            _req(all(is_in_ascii(e) for e in cu.devices), "logicalSimulation, elements in computationUnit.devicesUnit contain illegal characters.")
            _req(all(s.num_cpus >= 1 for s in cu.setting), "element.numCpus in logicalSimulation.computationUnit.unitSetting should not be less than 1.")
            for side, reqs in (("logicalSimulation", t.logical_simulation.resource_request),
                               ("deviceSimulation", t.device_simulation.resource_request)):
                for i, r in enumerate(reqs):
                    _req(r.name != "", f"The name of No.{i} resource_request in {side} should not be empty")
                    _req(is_in_ascii(r.name), f"The name of No.{i} resource_request in {side} = {r.name} contains illegal characters.")
                    _req(len(r.devices) == len(set(r.devices)), f"In {side}, resource_request.name = {r.name} devices = {r.devices} contains repeated elements!")
                    _req(all(is_in_ascii(e) for e in r.devices), f"In {side}, elements in resource_request.name={r.name} devices contain illegal characters.")
                    _req(all(e >= 0 for e in r.num_request), f"In {side}, elements in resource_request_name={r.name} num_request should not be less than 0.")
        except ValidationError as e:
            return False, str(e)
        return True, "Pass"

    @staticmethod
    def _is_json(s: str) -> bool:
        try:
            json.loads(s)
            return True
        except (ValueError, TypeError):
            return False

    # ------------------------------------------------------------------------------------------
    def validate_relationship(self, t: TaskConfig) -> Tuple[bool, str]:
        try:
            for d in t.target.data:
                n = d.name
                if d.data_path and FileTransferType(d.data_transfer_type).name not in ("MINIO", "FILE"):
                    _req(is_file_ext(d.data_path, ".zip"), f"target data_name = {n}, dataPath={d.data_path} should be a .zip file for HTTP/S3 transfer.")
                devices, nums, dyn = d.total_simulation.devices, d.total_simulation.nums, d.total_simulation.dynamic_nums
                _req(len(devices) == len(nums) == len(dyn), f"target, data_name={n}, devices, nums and dynamic_nums in total_simulation should have the same length")
                _req(all(nums[i] > dyn[i] for i in range(len(nums))), f"target, data_name={n}, elements in nums={nums} should be larger than dynamic_nums={dyn}")
                rr = d.allocation.running_response
                _req(set(rr.devices).issubset(devices), f"target, data_name={n}, devices in runningResponse should be a subset of devices in totalSimulation")
                _req(len(rr.devices) == len(rr.nums), f"target, data_name={n}, devices, nums in allocation.runningResponse should have the same length")
                rr_nums = _rearrange(devices, rr.devices, rr.nums, 0)
                _req(all(rr_nums[i] <= nums[i] for i in range(len(nums))), f"target, data_name={n}, nums in runningResponse ({rr_nums}) should not be larger than nums in totalSimulation ({nums})")
                if not d.allocation.optimization:
                    al = list(d.allocation.logical_simulation) or [0] * len(nums)
                    ad = list(d.allocation.device_simulation) or [0] * len(nums)
                    _req(len(al) == len(nums) and len(al) == len(ad), f"target, data_name={n}, allocation.logical_simulation, allocation.device_simulation and totalSimulation.nums should have the same length")
                    _req(all(nums[i] == al[i] + ad[i] for i in range(len(nums))), f"target, data_name={n}, nums should equal logical_simulation + device_simulation per device class when optimization is false")
                    _req(all(ad[i] >= rr_nums[i] for i in range(len(nums))), f"target, data_name={n}, allocation.device_simulation should not be less than runningResponse nums")
            fs = t.operatorflow.flow_setting
            _req(fs.start.logical_simulation.wait_interval <= fs.start.logical_simulation.total_timeout
                 and fs.start.device_simulation.wait_interval <= fs.start.device_simulation.total_timeout
                 and fs.stop.logical_simulation.wait_interval <= fs.stop.logical_simulation.total_timeout
                 and fs.stop.device_simulation.wait_interval <= fs.stop.device_simulation.total_timeout,
                 "waitInterval should not be larger than totalTimeout in flowSetting start/stop conditions")
            seen: List[str] = []
            for o in t.operatorflow.operators:
                n = o.name
                _req(n not in seen, f"operator_name = {n} is repeated in operatorFlow")
                if o.operation_behavior_controller.use_gradient_house:
                    _req(o.operation_behavior_controller.strategy_gradient_house != "", f"operator_name = {n}, strategyBehaviorController should not be empty when useController == True")
                if o.input:
                    _req(set(o.input).issubset(set(seen)), f"operator_name = {n}, input = {o.input} should be a subset of the previous operators {seen}")
                if o.model.use_model:
                    _req(o.model.model_path != "", f"operator_name = {n}, modelPath should not be empty when useModel == True")
                cp = o.logical_simulation.operator_code_path
                if cp != "" and os.path.isdir(os.path.abspath(cp)):
                    _req(FileTransferType(o.logical_simulation.operator_transfer_type).name == "FILE", f"operator={n}, if operatorCodePath is a dir, only FILE is allowed as operatorTransferType")
                _req(not (o.logical_simulation.operator_code_path == "" and o.device_simulation.operator_code_path == ""),
                     f"operator_name = {n}, operatorCodePath in logical_simulation and device_simulation should not both be empty")
                _req(not (o.logical_simulation.operator_entry_file == "" and o.device_simulation.operator_entry_file == ""),
                     f"operator_name = {n}, operatorEntryFile in logical_simulation and device_simulation should not both be empty")
                seen.append(n)
            data_names = t.data_names()
            lnames = [r.name for r in t.logical_simulation.resource_request]
            dnames = [r.name for r in t.device_simulation.resource_request]
            _req(set(data_names).issubset(set(lnames)) and set(data_names).issubset(set(dnames)),
                 "resource request in both logical simulation and device simulation should cover the target data.")
            cu = t.logical_simulation.computation_unit
            _req(len(cu.devices) == len(cu.setting), "logicalSimulation, computationUnit.devicesUnit and unitSetting should have the same length")
            all_devs = sorted({dev for d in t.target.data for dev in d.total_simulation.devices})
            _req(set(all_devs).issubset(set(cu.devices)), f"In logicalSimulation, all devices within total_simulation = {all_devs} should be included in computationUnit.devicesUnit={cu.devices}.")
            by_name = {d.name: d for d in t.target.data}
            for r in t.logical_simulation.resource_request:
                _req(r.name in data_names, f"In logicalSimulation, resource_request.name = {r.name} should be one of the target data names {data_names}")
                _req(len(r.devices) == len(r.num_request), f"In logicalSimulation, resource_request.name = {r.name}, devices and num_request should have the same length")
                d = by_name[r.name]
                if not d.allocation.optimization:
                    devices = d.total_simulation.devices
                    al = _rearrange(devices, devices, list(d.allocation.logical_simulation) or [0] * len(devices), 0)
                    req = _rearrange(devices, r.devices, r.num_request, 0)
                    for i, dev in enumerate(devices):
                        if al[i] == 0:
                            _req(req[i] >= 0, f"In logicalSimulation, data_name = {d.name}, num_request of {dev} should not be less than 0")
                        else:
                            _req(req[i] > 0, f"In logicalSimulation, data_name = {d.name}, num_request of {dev} should be larger than 0 when its logical allocation is {al[i]}")
            for r in t.device_simulation.resource_request:
                _req(r.name in data_names, f"In deviceSimulation, resource_request.name = {r.name} should be one of the target data names {data_names}")
                _req(len(r.devices) == len(r.num_request), f"In deviceSimulation, resource_request.name = {r.name}, devices and num_request should have the same length")
                d = by_name[r.name]
                devices = d.total_simulation.devices
                rr_nums = _rearrange(devices, d.allocation.running_response.devices, d.allocation.running_response.nums, 0)
                req = _rearrange(devices, r.devices, r.num_request, 0)
                if not d.allocation.optimization:
                    ad = _rearrange(devices, devices, list(d.allocation.device_simulation) or [0] * len(devices), 0)
                    for i, dev in enumerate(devices):
                        if ad[i] == 0:
                            continue
                        if ad[i] == rr_nums[i]:
                            _req(req[i] >= rr_nums[i], f"In deviceSimulation, data_name = {d.name}, num_request of {dev} should not be less than runningResponse nums {rr_nums[i]}")
                        else:
                            _req(req[i] >= 1, f"In deviceSimulation, data_name = {d.name}, num_request of {dev} should be at least 1 when its device allocation is {ad[i]}")
                            _req(req[i] > rr_nums[i], f"In deviceSimulation, data_name = {d.name}, num_request of {dev} should be larger than runningResponse nums {rr_nums[i]}")
                else:
                    for i, dev in enumerate(devices):
                        if rr_nums[i] > 0:
                            _req(req[i] > rr_nums[i], f"In deviceSimulation, data_name = {d.name}, num_request of {dev} should be larger than runningResponse nums {rr_nums[i]}")
        except ValidationError as e:
            return False, str(e)
        return True, "Pass"

    # ------------------------------------------------------------------------------------------
    def validate_task_parameters(self, t: TaskConfig) -> Tuple[bool, str]:
        for fn in (self.validate_type, self.validate_correctness, self.validate_relationship):
            ok, msg = fn(t)
            if not ok:
                return False, f"[{fn.__name__}]: {msg}"
        return True, "Pass"
