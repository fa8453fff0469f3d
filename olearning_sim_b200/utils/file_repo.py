"""Blob stores for data / model / operator-code transfer (FILE, HTTP, S3, MINIO).

Counterparts: ``S3FileRepo`` / ``MinioFileRepo`` (ols_core/ofl_commons/infrastructure/FileRepo/{s3,minio}_file_repo.py
— whose base class ``file_repo.FileRepo`` is missing from the reference tree), the helper functions of
ols_core/taskMgr/utils/utils_oflserver.py:11-186 (download/upload/delete, skip the download when the local file
already exists :100-103,148-151), and the ``FileTransferType`` dispatch in ols_core/taskMgr/utils/utils_run_task.py:
174-397.  ``LocalFileRepo`` is a directory-backed bucket store so every code path is exercisable offline; the S3 / MinIO
classes bind lazily to ``boto3`` / ``minio`` and raise a clear error when those clients are not installed.
"""
from __future__ import annotations

import os
import shutil
import urllib.request
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional


class FileRepo(ABC):
    @abstractmethod
    def upload_file(self, local_path: str, object_name: str) -> bool: ...

    @abstractmethod
    def download_file(self, object_name: str, local_path: str) -> bool: ...

    @abstractmethod
    def delete_file(self, object_name: str) -> bool: ...

    def download_payload(self, object_name: str) -> Optional[bytes]:
        tmp = f"/tmp/olsb_payload_{os.getpid()}_{abs(hash(object_name))}"
        try:
            if not self.download_file(object_name, tmp):
                return None
            with open(tmp, "rb") as f:
                return f.read()
        finally:
            if os.path.exists(tmp):
                os.remove(tmp)

    def listdir_object_name(self, prefix: str) -> List[str]:
        raise NotImplementedError

    def download_if_missing(self, object_name: str, local_path: str) -> bool:
        """utils_oflserver.py:100-103 — an existing local copy short-circuits the transfer."""
        if os.path.exists(local_path):
            return True
        os.makedirs(os.path.dirname(local_path) or ".", exist_ok=True)
        return self.download_file(object_name, local_path)


class LocalFileRepo(FileRepo):
    """``root/bucket/object`` on the local filesystem."""

    def __init__(self, root: str, bucket: str = "default"):
        self.base = os.path.join(root, bucket)
        os.makedirs(self.base, exist_ok=True)

    def _p(self, object_name: str) -> str:
        p = os.path.normpath(os.path.join(self.base, object_name.lstrip("/")))
        if not p.startswith(os.path.normpath(self.base)):
            raise ValueError(f"object name escapes the bucket: {object_name}")
        return p

    def upload_file(self, local_path, object_name) -> bool:
        dst = self._p(object_name)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(local_path, dst)
        return True

    def download_file(self, object_name, local_path) -> bool:
        src = self._p(object_name)
        if not os.path.isfile(src):
            return False
        os.makedirs(os.path.dirname(local_path) or ".", exist_ok=True)
        shutil.copyfile(src, local_path)
        return True

    def delete_file(self, object_name) -> bool:
        p = self._p(object_name)
        if os.path.isfile(p):
            os.remove(p)
            return True
        return False

    def listdir_object_name(self, prefix: str = "") -> List[str]:
        out = []
        root = self._p(prefix) if prefix else self.base
        for base, _, files in os.walk(root if os.path.isdir(root) else os.path.dirname(root)):
            for f in files:
                rel = os.path.relpath(os.path.join(base, f), self.base)
                if rel.startswith(prefix.lstrip("/")):
                    out.append(rel)
        return sorted(out)


class HttpFileRepo(FileRepo):
    def __init__(self, base_url: str = "", timeout: float = 60.0):
        self.base_url, self.timeout = base_url.rstrip("/"), timeout

    def download_file(self, object_name, local_path) -> bool:
        url = object_name if object_name.startswith(("http://", "https://")) else f"{self.base_url}/{object_name.lstrip('/')}"
        try:
            with urllib.request.urlopen(url, timeout=self.timeout) as r, open(local_path, "wb") as f:
                shutil.copyfileobj(r, f)
            return True
        except Exception:  # noqa: BLE001
            return False

    def upload_file(self, local_path, object_name) -> bool:
        return False        # the reference's HTTP transport is download-only as well

    def delete_file(self, object_name) -> bool:
        return False


class S3FileRepo(FileRepo):
    """S3-compatible object store (ofl_commons/.../s3_file_repo.py:1-64).  Transport: an injected ``client``, else ``boto3``
    when it is installed, else the built-in Signature-V4 REST client (utils/s3_rest.py) — the same five calls either way."""

    def __init__(self, endpoint_url: str = "", access_key: str = "", secret_key: str = "", bucket: str = "",
                 region: Optional[str] = None, client: Any = None):
        if client is None:
            try:
                import boto3  # type: ignore
                client = boto3.client("s3", endpoint_url=endpoint_url, aws_access_key_id=access_key,
                                      aws_secret_access_key=secret_key, region_name=region)
            except ImportError:
                from .s3_rest import S3RestClient
                client = S3RestClient(endpoint_url or f"https://s3.{region or 'us-east-1'}.amazonaws.com", access_key, secret_key,
                                      region or "us-east-1")
        self._s3, self._bucket = client, bucket

    def upload_file(self, local_path, object_name) -> bool:
        try:
            self._s3.upload_file(local_path, self._bucket, object_name)
            return True
        except Exception:  # noqa: BLE001
            return False

    def download_file(self, object_name, local_path) -> bool:
        try:
            os.makedirs(os.path.dirname(local_path) or ".", exist_ok=True)
            self._s3.download_file(self._bucket, object_name, local_path)
            return True
        except Exception:  # noqa: BLE001
            if os.path.exists(local_path) and os.path.getsize(local_path) == 0:
                os.remove(local_path)
            return False

    def delete_file(self, object_name) -> bool:
        try:
            self._s3.delete_object(Bucket=self._bucket, Key=object_name)
            return True
        except Exception:  # noqa: BLE001
            return False

    def download_payload(self, object_name):
        try:
            return self._s3.get_object(Bucket=self._bucket, Key=object_name)["Body"].read()
        except Exception:  # noqa: BLE001
            return None

    def listdir_object_name(self, prefix: str) -> List[str]:
        out, token = [], None
        while True:
            kw = {"Bucket": self._bucket, "Prefix": prefix}
            if token:
                kw["ContinuationToken"] = token
            page = self._s3.list_objects_v2(**kw)
            out += [o["Key"] for o in page.get("Contents", [])]
            token = page.get("NextContinuationToken")
            if not page.get("IsTruncated") or not token:
                return sorted(out)


class MinioFileRepo(FileRepo):
    """MinIO (ofl_commons/.../minio_file_repo.py:1-89): an injected ``client``, else the ``minio`` package, else the built-in
    Signature-V4 REST client (MinIO speaks the S3 protocol)."""

    def __init__(self, endpoint: str = "", access_key: str = "", secret_key: str = "", bucket: str = "",
                 secure: bool = False, client: Any = None):
        if client is None:
            try:
                from minio import Minio  # type: ignore
                client = Minio(endpoint, access_key=access_key, secret_key=secret_key, secure=secure)
            except ImportError:
                from .s3_rest import S3RestClient
                url = endpoint if "://" in endpoint else ("https://" if secure else "http://") + endpoint
                client = S3RestClient(url, access_key, secret_key)
        self._c, self._bucket = client, bucket

    def upload_file(self, local_path, object_name) -> bool:
        try:
            self._c.fput_object(self._bucket, object_name, local_path)
            return True
        except Exception:  # noqa: BLE001
            return False

    def download_file(self, object_name, local_path) -> bool:
        try:
            os.makedirs(os.path.dirname(local_path) or ".", exist_ok=True)
            self._c.fget_object(self._bucket, object_name, local_path)
            return True
        except Exception:  # noqa: BLE001
            return False

    def delete_file(self, object_name) -> bool:
        try:
            self._c.remove_object(self._bucket, object_name)
            return True
        except Exception:  # noqa: BLE001
            return False

    def listdir_object_name(self, prefix: str) -> List[str]:
        return sorted(o.object_name for o in self._c.list_objects(self._bucket, prefix=prefix, recursive=True))


def get_file_repo(transfer_type: Any, config: Optional[Dict[str, Any]] = None) -> Optional[FileRepo]:
    """FileTransferType -> repo (config = the ``manager_config.yaml`` sections of the reference)."""
    name = getattr(transfer_type, "name", str(transfer_type)).upper()
    cfg = config or {}
    if name == "FILE":
        return LocalFileRepo(cfg.get("root", "/"), cfg.get("bucket", ""))
    if name == "HTTP":
        return HttpFileRepo(cfg.get("base_url", ""))
    if name == "S3":
        c = cfg.get("s3", cfg)
        return S3FileRepo(c.get("endpoint_url", ""), c.get("access_key", ""), c.get("secret_key", ""), c.get("bucket", ""),
                          client=c.get("client"))
    if name == "MINIO":
        c = cfg.get("minio", cfg)
        return MinioFileRepo(c.get("endpoint", ""), c.get("access_key", ""), c.get("secret_key", ""), c.get("bucket", ""),
                             client=c.get("client"))
    raise ValueError(f"unknown transfer type {transfer_type!r}")


# ---- utils_oflserver.py function-style helpers ----------------------------------------------------------------
def downloadFile(repo: FileRepo, object_name: str, local_path: str) -> bool:
    return repo.download_if_missing(object_name, local_path)


def uploadFile(repo: FileRepo, local_path: str, object_name: str) -> bool:
    return repo.upload_file(local_path, object_name)


def deleteFile(repo: FileRepo, object_name: str) -> bool:
    return repo.delete_file(object_name)


def listdirFile(repo: FileRepo, prefix: str) -> List[str]:
    return repo.listdir_object_name(prefix)
