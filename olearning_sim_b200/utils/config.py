"""Typed configuration tree with layered sources: defaults < file (INI / YAML / JSON) < environment < CLI overrides.

The reference scatters configuration over four mechanisms: an INI read with ``configparser`` whose string
values keep their quotes (``[1:-1]`` slicing, ols_core/taskMgr/task_manager.py:162-167), one YAML per concern re-read
on every getter call (ols_core/deviceflow/utils/config_parser.py:4-81) with ``${ENV}`` placeholders that are never
expanded, per-task JSON, and CLI flags; and it opens MySQL at import time.  Here there is one dataclass tree, nothing
is read at import, ``${ENV}`` placeholders *are* expanded, and the reference's section / key names are accepted so
an existing ``config.conf`` / ``deviceflow_config.yaml`` loads unchanged.
"""
from __future__ import annotations

import configparser
import json
import os
import re
from dataclasses import dataclass, field, fields, is_dataclass
from typing import Any, Dict, List, Optional

import yaml


@dataclass
class GrpcConfig:
    max_send_message_length: int = 100 * 1024 * 1024
    max_receive_message_length: int = 100 * 1024 * 1024
    resmgr: str = "127.0.0.1:60062"
    phonemgr: str = "127.0.0.1:60063"
    taskmgr_port: int = 60061
    resmgr_port: int = 60062
    max_workers: int = 10


@dataclass
class RuntimeConfig:
    session_sleep_time: float = 10.0


@dataclass
class ClusterConfig:
    gpus: int = 0                     # 0 = autodetect
    job_executor: str = "inprocess"   # inprocess | subprocess
    k8s_namespace: str = "default"
    ray_name: str = "olsb-demo1"
    ray_label: str = "demo1-cluster"
    worker_replicas: int = 4
    worker_minReplicas: int = 1
    worker_maxReplicas: int = 8


@dataclass
class TaskMgrSection:
    scheduler_sleep_time: float = 5.0
    release_sleep_time: float = 10.0
    interrupt_sleep_time: float = 300.0
    interrupt_queue_time: float = 3600.0
    interrupt_running_time: float = 172800.0


@dataclass
class DeviceflowSection:
    grpc_service_address: str = "127.0.0.1"
    grpc_service_port: int = 50051
    max_workers: int = 10
    shelfroom_url: str = "mem://local"
    shelfroom_namespace: str = "public/shelf_room"
    source_url: str = "mem://local"
    source_topic: str = "persistent://public/default/deviceflow_inbound"
    source_websocket_url: str = ""
    journal_path: str = ""


@dataclass
class StorageConfig:
    task_table: str = ":memory:"
    res_table: str = ":memory:"
    deviceflow_table: str = ":memory:"
    log_table: str = ""
    log_jsonl: str = ""
    checkpoint_dir: str = ""
    blob_root: str = "/tmp/olsb_blobs"


@dataclass
class HybridCostModel:
    alpha: float = 3.5      # logical per round-unit seconds   (utils_runner.py:941)
    beta: float = 0.14      # phone per round-unit seconds     (:942)
    lam: float = 8.808      # phone fixed start-up seconds     (:943)


@dataclass
class OlsbConfig:
    grpc: GrpcConfig = field(default_factory=GrpcConfig)
    runtime: RuntimeConfig = field(default_factory=RuntimeConfig)
    cluster: ClusterConfig = field(default_factory=ClusterConfig)
    taskMgr: TaskMgrSection = field(default_factory=TaskMgrSection)
    deviceflow: DeviceflowSection = field(default_factory=DeviceflowSection)
    storage: StorageConfig = field(default_factory=StorageConfig)
    hybrid: HybridCostModel = field(default_factory=HybridCostModel)


_ENV = re.compile(r"\$\{([A-Za-z_][A-Za-z0-9_]*)(?::-([^}]*))?\}")


def expand_env(v: Any) -> Any:
    if isinstance(v, str):
        return _ENV.sub(lambda m: os.environ.get(m.group(1), m.group(2) or ""), v)
    if isinstance(v, dict):
        return {k: expand_env(x) for k, x in v.items()}
    if isinstance(v, list):
        return [expand_env(x) for x in v]
    return v


def _coerce(value: Any, target: Any) -> Any:
    if isinstance(value, str):
        value = value.strip()
        if len(value) >= 2 and value[0] == value[-1] and value[0] in "'\"":
            value = value[1:-1]                      # quoted INI strings
    if isinstance(target, bool):
        return value if isinstance(value, bool) else str(value).lower() in ("1", "true", "yes", "on")
    if isinstance(target, int) and not isinstance(target, bool):
        return int(float(value))
    if isinstance(target, float):
        return float(value)
    return value


def _apply(cfg: Any, data: Dict[str, Any]) -> None:
    for f in fields(cfg):
        if f.name not in data:
            continue
        cur = getattr(cfg, f.name)
        if is_dataclass(cur) and isinstance(data[f.name], dict):
            _apply(cur, data[f.name])
        else:
            setattr(cfg, f.name, _coerce(data[f.name], cur))


def _read_file(path: str) -> Dict[str, Any]:
    ext = os.path.splitext(path)[1].lower()
    if ext in (".yaml", ".yml"):
        with open(path) as f:
            d = yaml.safe_load(f) or {}
        # a bare deviceflow_config.yaml (flat keys) is mapped into the deviceflow section
        if "grpc_service_address" in d or "source_topic" in d:
            return {"deviceflow": d}
        return d
    if ext == ".json":
        with open(path) as f:
            return json.load(f)
    cp = configparser.ConfigParser()
    cp.read(path)
    out: Dict[str, Any] = {s: dict(cp.items(s)) for s in cp.sections()}
    if "ray_cluster" in out:                           # reference section name
        out.setdefault("cluster", {}).update(out.pop("ray_cluster"))
    return out


def load_config(paths: Optional[List[str]] = None, env_prefix: str = "OLSB_", overrides: Optional[List[str]] = None) -> OlsbConfig:
    """defaults < files (in order) < ``OLSB_SECTION__KEY=value`` environment < ``section.key=value`` overrides"""
    cfg = OlsbConfig()
    for p in paths or []:
        _apply(cfg, expand_env(_read_file(p)))
    env: Dict[str, Dict[str, Any]] = {}
    for k, v in os.environ.items():
        if k.startswith(env_prefix) and "__" in k:
            sec, key = k[len(env_prefix):].split("__", 1)
            names = {f.name.lower(): f.name for f in fields(cfg)}
            if sec.lower() in names:
                env.setdefault(names[sec.lower()], {})[key.lower()] = v
    _apply(cfg, env)
    ov: Dict[str, Dict[str, Any]] = {}
    for item in overrides or []:
        lhs, rhs = item.split("=", 1)
        sec, key = lhs.split(".", 1)
        ov.setdefault(sec, {})[key] = rhs
    _apply(cfg, ov)
    return cfg
