"""Shared builders for tests."""
import json


def make_task_json(task_id="t1", user_id="u1", rounds=2, nums=(8,), devices=("High",), dynamic=(0,), priority=3,
                   logical=None, device=None, operator_params=None, use_df=False, strategy=None, num_request=None,
                   phone_request=None, code_path="builtin:fedavg", entry="train.py", optimization=False,
                   model="mlp2", engine=None, outbound=None):
    nums, devices, dynamic = list(nums), list(devices), list(dynamic)
    logical = list(logical) if logical is not None else list(nums)
    device = list(device) if device is not None else [n - l for n, l in zip(nums, logical)]
    op_params = operator_params if operator_params is not None else {
        "model": model, "local_steps": 1, "batch": 8, "shard": 8, "lr": 0.1, "device": "cpu"}
    return {
        "user_id": user_id, "task_id": task_id,
        "target": {"priority": priority, "data": [{
            "name": "data_0", "data_path": "", "data_split_type": False, "data_transfer_type": "FILE",
            "task_type": "classification",
            "total_simulation": {"devices": devices, "nums": nums, "dynamic_nums": dynamic},
            "allocation": {"optimization": optimization, "logical_simulation": logical, "device_simulation": device,
                           "running_response": {"devices": [], "nums": []}}}]},
        "operatorflow": {
            "flow_setting": {"round": rounds,
                             "start": {"logical_simulation": {"strategy": "", "wait_interval": 0, "total_timeout": 0},
                                       "device_simulation": {"strategy": "", "wait_interval": 0, "total_timeout": 0}},
                             "stop": {"logical_simulation": {"strategy": "", "wait_interval": 0, "total_timeout": 0},
                                      "device_simulation": {"strategy": "", "wait_interval": 0, "total_timeout": 0}}},
            "operators": [{
                "name": "training", "input": [],
                "operation_behavior_controller": {"use_gradient_house": use_df,
                                                  "strategy_gradient_house": json.dumps(strategy) if strategy else "",
                                                  "outbound_service": json.dumps(outbound) if outbound else ""},
                "use_data": True,
                "model": {"use_model": False, "model_for_train": True, "model_transfer_type": "FILE", "model_path": "",
                          "model_update_style": ""},
                "logical_simulation": {"operator_transfer_type": "FILE", "operator_code_path": code_path,
                                       "operator_entry_file": entry, "operator_params": json.dumps(op_params)},
                "device_simulation": {"operator_transfer_type": "FILE", "operator_code_path": "",
                                      "operator_entry_file": "", "operator_params": ""}}]},
        "logical_simulation": {"computation_unit": {"devices": devices, "setting": [{"num_cpus": 1} for _ in devices]},
                               "resource_request": [{"name": "data_0", "devices": devices,
                                                     "num_request": list(num_request) if num_request else [1] * len(devices)}]},
        "device_simulation": {"resource_request": [{"name": "data_0", "devices": devices,
                                                    "num_request": list(phone_request) if phone_request else [0] * len(devices)}]},
        **({"engine": engine} if engine else {}),
    }


REAL_TIME = {"real_time_dispatch": {"use_strategy": True, "dispatch_batch_sizes": [3],
                                    "drop_simulation": {"drop_probability": 0.0}},
             "flow_dispatch": {"use_strategy": False}}
