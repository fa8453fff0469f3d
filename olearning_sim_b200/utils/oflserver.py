"""Function-style storage / messaging helpers keyed by a ``manager_config`` file.

The reference's operators and actors call a flat helper module instead of the repo classes
(ols_core/taskMgr/utils/utils_oflserver.py:11-186): every helper re-reads a YAML ``manager_config`` and builds a fresh
S3 / MinIO / Pulsar client.  The same names exist here so operator code written against that module keeps working; the
differences are

* ``manager_config`` may be a path (YAML / JSON), an already-parsed ``dict`` or empty (environment: ``OLSB_STORE_ROOT``);
  parsed files and the repos built from them are cached per path+mtime instead of being rebuilt on every call;
* a config with ``{"file": {"root": ...}}`` (or no object-store section at all) maps every helper onto
  ``LocalFileRepo`` so the helpers work on a box without S3 / MinIO;
* credentials come from the config file or ``${ENV}`` placeholders only, never from source.
"""
from __future__ import annotations

import fnmatch
import json
import os
from typing import Any, Dict, List, Optional, Tuple, Union

import yaml

from .config import expand_env
from .file_repo import FileRepo, LocalFileRepo, MinioFileRepo, S3FileRepo

ConfigLike = Union[str, Dict[str, Any], None]
_cache: Dict[Tuple[str, float], Dict[str, Any]] = {}
_repos: Dict[Tuple[str, str], FileRepo] = {}


def _load(manager_config: ConfigLike) -> Dict[str, Any]:
    if isinstance(manager_config, dict):
        return expand_env(manager_config)
    if not manager_config:
        return {"file": {"root": os.environ.get("OLSB_STORE_ROOT", "/tmp/olsb_store")}}
    key = (os.path.abspath(manager_config), os.path.getmtime(manager_config))
    if key not in _cache:
        with open(manager_config) as f:
            text = f.read()
        doc = json.loads(text) if manager_config.endswith(".json") else yaml.safe_load(text)
        _cache[key] = expand_env(doc or {})
    return _cache[key]


def _repo(manager_config: ConfigLike, kind: str) -> FileRepo:
    """kind: 's3' | 'minio'.  Falls back to the local store when the section is absent."""
    cfg = _load(manager_config)
    ident = (json.dumps(cfg, sort_keys=True, default=str), kind)
    if ident in _repos:
        return _repos[ident]
    sect = cfg.get(kind) or cfg.get(kind.upper())
    if sect and kind == "s3":
        repo: FileRepo = S3FileRepo(sect["endpoint_url"], sect["access_key"], sect["secret_key"], sect["bucket"],
                                    sect.get("region"))
    elif sect and kind == "minio":
        repo = MinioFileRepo(sect["endpoint"], sect["access_key"], sect["secret_key"], sect["bucket"],
                             bool(sect.get("secure", False)))
    else:
        f = cfg.get("file", {})
        repo = LocalFileRepo(f.get("root", os.environ.get("OLSB_STORE_ROOT", "/tmp/olsb_store")), f.get("bucket", kind))
    _repos[ident] = repo
    return repo


# ---- S3 (utils_oflserver.py:40-133) ----------------------------------------------------------------------------------
def getS3FileRepo(manager_config: ConfigLike = "") -> FileRepo:
    return _repo(manager_config, "s3")


def getS3SelectionFiles(manager_config: ConfigLike = "", jsonfiles_path: str = "") -> List[str]:
    """Object names of the per-client selection configs ``{jsonfiles_path}/*.json`` (utils_oflserver.py:47-84)."""
    names = _repo(manager_config, "s3").listdir_object_name(jsonfiles_path.rstrip("/") + "/" if jsonfiles_path else "")
    return sorted(n for n in names if fnmatch.fnmatch(n, f"{jsonfiles_path.rstrip('/')}/*.json"))


def downloadS3File(manager_config: ConfigLike = "", local_file: str = "", taget_file: str = "") -> bool:
    """Skip the transfer when ``local_file`` already exists (utils_oflserver.py:100-103)."""
    return _repo(manager_config, "s3").download_if_missing(taget_file, local_file)


def uploadS3File(manager_config: ConfigLike = "", local_file: str = "", taget_file: str = "") -> bool:
    return _repo(manager_config, "s3").upload_file(local_file, taget_file)


def deleteS3File(manager_config: ConfigLike = "", taget_file: str = "") -> bool:
    return _repo(manager_config, "s3").delete_file(taget_file)


# ---- MinIO (utils_oflserver.py:135-186) ------------------------------------------------------------------------------
def downloadMinioFile(manager_config: ConfigLike = "", local_file: str = "", taget_file: str = "") -> bool:
    return _repo(manager_config, "minio").download_if_missing(taget_file, local_file)


def uploadMinioFile(manager_config: ConfigLike = "", local_file: str = "", taget_file: str = "") -> bool:
    return _repo(manager_config, "minio").upload_file(local_file, taget_file)


def listdirMinioFile(manager_config: ConfigLike = "", taget_dir: str = "") -> List[str]:
    return _repo(manager_config, "minio").listdir_object_name(taget_dir)


# ---- message bus / task table (utils_oflserver.py:11-38) -------------------------------------------------------------
def getPulsarObjects(manager_config: ConfigLike = "", bus: Optional[Any] = None):
    """(bus, producer) on the configured gradient-house topic.  The reference returns a Pulsar client + producer; here
    the bus is the in-process journaled bus unless the caller passes its own (deviceflow/bus.py)."""
    from ..deviceflow.bus import InMemoryBus
    cfg = _load(manager_config)
    sect = cfg.get("pulsar", cfg.get("bus", {}))
    bus = bus or InMemoryBus(journal_path=sect.get("journal"))
    return bus, bus.create_producer(sect.get("topic", "gradient_house"))


def get_gradient_house_pulsar(config: ConfigLike = "") -> Dict[str, str]:
    """Inbound endpoint clients publish their updates to: ``{"url", "topic"}`` (utils_oflserver.py:32-38)."""
    sect = _load(config).get("pulsar", _load(config).get("bus", {}))
    return {"url": sect.get("url", "mem://local"), "topic": sect.get("topic", "gradient_house")}


def getSqlTaskRepo(repo_config: ConfigLike = ""):
    """Task table repo from a repo config (``{"database": path}``; default: in-memory) — utils_oflserver.py:11-23."""
    from ..control.repo import TaskTableRepo
    cfg = _load(repo_config) if repo_config else {}
    return TaskTableRepo(cfg.get("database") or cfg.get("sqlite", {}).get("path") or ":memory:")
