"""TaskConfig JSON round-trip + the three validation passes (reference: taskMgr/utils/utils.py:269-1197)."""
import copy
import json

import pytest

from olearning_sim_b200.api.schema import FileTransferType, TaskConfig, json2taskconfig, taskconfig2json
from olearning_sim_b200.api.validate import ValidateParameters
from tests.helpers import make_task_json


def test_json_roundtrip_is_lossless():
    tj = make_task_json(nums=(8, 4), devices=("High", "Low"), dynamic=(1, 0), logical=(6, 4))
    tc = json2taskconfig(json.dumps(tj))
    assert isinstance(tc, TaskConfig)
    assert taskconfig2json(tc) == tj
    assert taskconfig2json(json2taskconfig(taskconfig2json(tc))) == tj
    assert tc.target.data[0].data_transfer_type == FileTransferType.FILE
    assert tc.operator_names() == ["training"] and tc.data_names() == ["data_0"]


def test_enum_by_table_not_eval():
    tj = make_task_json()
    tj["target"]["data"][0]["data_transfer_type"] = "__import__('os').system('true')"
    with pytest.raises(ValueError):
        json2taskconfig(tj)
    tj["target"]["data"][0]["data_transfer_type"] = "minio"
    assert json2taskconfig(tj).target.data[0].data_transfer_type == FileTransferType.MINIO


def _v(tj):
    return ValidateParameters().validate_task_parameters(json2taskconfig(tj))


def test_valid_task_passes():
    ok, msg = _v(make_task_json())
    assert ok, msg


@pytest.mark.parametrize("mutate,needle", [
    (lambda t: t.update(user_id=""), "userID should not be empty"),
    (lambda t: t.update(task_id="täsk"), "illegal characters"),
    (lambda t: t["target"].update(priority=11), "priority"),
    (lambda t: t["operatorflow"]["flow_setting"].update(round=0), "round"),
    (lambda t: t["target"]["data"][0]["total_simulation"].update(nums=[0]), "larger than 0"),
    (lambda t: t["target"]["data"][0]["total_simulation"].update(dynamic_nums=[8]), "larger than dynamic_nums"),
    (lambda t: t["target"]["data"][0]["total_simulation"].update(devices=["High", "High"], nums=[4, 4], dynamic_nums=[0, 0]), "repeated"),
    (lambda t: t["target"]["data"][0]["allocation"].update(logical_simulation=[5]), "logical_simulation + device_simulation"),
    (lambda t: t["operatorflow"]["operators"][0].update(name="has space"), "spaces"),
    (lambda t: t["operatorflow"]["operators"][0].update(input=["later_op"]), "subset of the previous operators"),
    (lambda t: t["operatorflow"]["operators"][0]["logical_simulation"].update(operator_params="{not json"), "json string"),
    (lambda t: t["operatorflow"]["operators"][0]["logical_simulation"].update(operator_entry_file="train.sh"), ".py"),
    (lambda t: t["operatorflow"]["operators"][0]["operation_behavior_controller"].update(use_gradient_house=True), "strategyBehaviorController should not be empty"),
    (lambda t: t["operatorflow"]["flow_setting"]["start"]["logical_simulation"].update(wait_interval=5, total_timeout=1), "waitInterval"),
    (lambda t: t["logical_simulation"]["computation_unit"].update(devices=["Middle"]), "computationUnit"),
    (lambda t: t["logical_simulation"]["resource_request"][0].update(name="other"), "cover the target data"),
    (lambda t: t["logical_simulation"]["resource_request"][0].update(num_request=[0]), "should be larger than 0"),
    (lambda t: t["logical_simulation"]["computation_unit"]["setting"][0].update(num_cpus=0), "numCpus"),
    (lambda t: t["operatorflow"]["operators"][0]["model"].update(use_model=True), "modelPath"),
    (lambda t: t["target"]["data"][0].update(data_transfer_type="S3", data_path="bucket/dir"), ".zip"),
])
def test_each_rule_rejects(mutate, needle):
    tj = make_task_json()
    mutate(tj)
    ok, msg = _v(tj)
    assert not ok and needle in msg, msg


def test_type_pass_catches_wrong_types():
    tc = json2taskconfig(make_task_json())
    tc.target.priority = "3"
    ok, msg = ValidateParameters().validate_type(tc)
    assert not ok and "priority" in msg
    tc = json2taskconfig(make_task_json())
    tc.target.data[0].total_simulation.nums = [True]
    ok, msg = ValidateParameters().validate_type(tc)
    assert not ok


def test_device_side_rules():
    # phones requested must exceed running_response when phones are allocated
    tj = make_task_json(nums=(8,), logical=(4,), device=(4,), phone_request=(1,))
    tj["target"]["data"][0]["allocation"]["running_response"] = {"devices": ["High"], "nums": [1]}
    ok, msg = _v(tj)
    assert not ok and "runningResponse" in msg
    tj["device_simulation"]["resource_request"][0]["num_request"] = [2]
    tj["operatorflow"]["operators"][0]["device_simulation"].update(operator_code_path="app.apk", operator_entry_file="app.apk")
    ok, msg = _v(tj)
    assert ok, msg


def test_optimization_skips_user_split_rule():
    tj = make_task_json(optimization=True)
    tj["target"]["data"][0]["allocation"].update(logical_simulation=[], device_simulation=[])
    ok, msg = _v(tj)
    assert ok, msg
