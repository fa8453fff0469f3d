"""Getter-style access to ``deviceflow_config.yaml``.

The reference reads one key per helper and re-parses the YAML file on every call
(ols_core/deviceflow/utils/config_parser.py:4-81).  The same getters exist here for code written against them; the file is
parsed once per (path, mtime), ``${ENV}`` placeholders are expanded, and a missing key falls back to the typed defaults of
``utils.config.DeviceflowSection`` instead of raising ``KeyError`` deep inside a service thread.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Tuple

import yaml

from ..utils.config import DeviceflowSection, expand_env

_cache: Dict[Tuple[str, float], Dict[str, Any]] = {}
_DEFAULTS = DeviceflowSection()


def _load(path: str) -> Dict[str, Any]:
    key = (os.path.abspath(path), os.path.getmtime(path))
    if key not in _cache:
        with open(path, encoding="utf-8") as f:
            _cache[key] = expand_env(yaml.safe_load(f.read()) or {})
    return _cache[key]


def _get(path: str, key: str, default: Any) -> Any:
    return _load(path).get(key, default)


def get_grpc_service_address(path: str) -> str:
    return _get(path, "grpc_service_address", _DEFAULTS.grpc_service_address)


def get_port(path: str) -> int:
    return int(_get(path, "port", _DEFAULTS.grpc_service_port))


def get_max_workers(path: str) -> int:
    return int(_get(path, "max_workers", _DEFAULTS.max_workers))


def get_pulsar_url_of_outbound_room(path: str) -> str:
    return _get(path, "service_url", _DEFAULTS.source_url)


def get_topic_url_of_outbound_room(path: str) -> str:
    return _get(path, "topic_url", "")


def get_subscription_of_outbound_room(path: str) -> str:
    return _get(path, "subscription", "deviceflow_outbound")


def get_source_url(path: str) -> str:
    return _get(path, "source_url", _DEFAULTS.source_url)


def get_source_topic(path: str) -> str:
    return _get(path, "source_topic", _DEFAULTS.source_topic)


def get_shelfroom_url(path: str) -> str:
    return _get(path, "shelfroom_url", _DEFAULTS.shelfroom_url)


def get_shelfroom_namespace(path: str) -> str:
    return _get(path, "shelfroom_namespace", _DEFAULTS.shelfroom_namespace)
