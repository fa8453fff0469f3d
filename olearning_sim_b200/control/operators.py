"""Operator contract.

Two kinds of operator can appear in a task's ``operatorflow.operators``:

1. **Script operators** — the reference's contract, kept for API parity (CPU): a directory with an entry file that
   is started once per simulated device as ``python3 <entry> --params <file.json>`` and exits 0 on success.
   ``OperatorABC`` / ``OperatorRunScriptABC`` are the base classes user code derives from
   (ols_core/taskMgr/base/base_operator.py:7-135, base_operator_run_script.py:5-32); the per-device params schema
   documented there (:12-53) is what ``construct_params_for_phones`` builds.
2. **Built-in GPU operators** — ``operator_code_path = "builtin:<name>"``: the whole client population of the job is
   trained by the FedEngine on the GPUs (no subprocess, no file exchange).  ``BUILTIN_OPERATORS`` maps names to
   factories taking the operator's ``operator_params`` JSON.
"""
from __future__ import annotations

import argparse
import json
from abc import ABC, abstractmethod
from typing import Any, Callable, Dict, Optional


class OperatorABC(ABC):
    """Per-device params (``--params`` JSON), as built by the actor for every simulated phone:

        {"task_id", "operator_name", "round" (current round index), "device_type", "split_index" (client index
         inside the dataset = its non-IID shard id), "data": {"name", "path", "split_type", "dataconfig"},
         "model": {"use_model", "for_train", "path"}, "operator_params": {...user json...},
         "deviceflow": {"use": bool, "routing_key", "inbound": {"url","topic"} | {"websocket_url"},
                        "compute_resource": "logical_simulation"}}
    """

    def __init__(self) -> None:
        self.params: Dict[str, Any] = {}

    def get_params(self, argv: Optional[list] = None) -> Dict[str, Any]:
        ap = argparse.ArgumentParser()
        ap.add_argument("--params", type=str, required=True, help="path of the per-device params json")
        args, _ = ap.parse_known_args(argv)
        with open(args.params) as f:
            self.params = json.load(f)
        return self.params

    @abstractmethod
    def construct_run_params(self) -> Dict[str, Any]: ...

    @abstractmethod
    def construct_run_script(self) -> str: ...

    @abstractmethod
    def run(self) -> int: ...


class OperatorRunScriptABC(ABC):
    """Base of the script that actually does the local work (train / upload ...)."""

    def __init__(self) -> None:
        self.args: Dict[str, Any] = {}

    def get_args(self, argv: Optional[list] = None) -> Dict[str, Any]:
        ap = argparse.ArgumentParser()
        ap.add_argument("--params", type=str, required=True)
        a, _ = ap.parse_known_args(argv)
        with open(a.params) as f:
            self.args = json.load(f)
        return self.args

    @abstractmethod
    def execution(self) -> bool: ...

    def run(self, argv: Optional[list] = None) -> int:
        self.get_args(argv)
        return 0 if self.execution() else 1


def publish_update(params: Dict[str, Any], message: str, bus: Any = None) -> None:
    """Helper for script operators: publish this device's update to the DeviceFlow inbound topic in the envelope
    {"routing_key","compute_resource","message"} (ols_core/deviceflow/non_grpc/utils/message.py:4-19)."""
    from ..deviceflow.rooms import make_message
    df = params.get("deviceflow", {})
    if not df.get("use"):
        return
    payload = make_message(df["routing_key"], df.get("compute_resource", "logical_simulation"), message)
    if bus is None:
        raise RuntimeError("no bus handle: in-process operators receive the bus, subprocess operators use the journal file")
    bus.create_producer(df["inbound"]["topic"]).send(payload)


# ---- built-in GPU operators ------------------------------------------------------------------------------
BUILTIN_OPERATORS: Dict[str, Callable[[Dict[str, Any]], Dict[str, Any]]] = {}


def register_builtin(name: str):
    def deco(fn):
        BUILTIN_OPERATORS[name] = fn
        return fn
    return deco


@register_builtin("fedavg")
def _fedavg(p: Dict[str, Any]) -> Dict[str, Any]:
    return {"algorithm": "fedavg", **p}


@register_builtin("fedprox")
def _fedprox(p: Dict[str, Any]) -> Dict[str, Any]:
    return {"algorithm": "fedprox", "mu": p.get("mu", 0.01), **p}


@register_builtin("evaluate")
def _evaluate(p: Dict[str, Any]) -> Dict[str, Any]:
    return {"algorithm": "evaluate", **p}


def parse_builtin(code_path: str) -> Optional[str]:
    return code_path.split(":", 1)[1] if code_path.startswith("builtin:") else None
