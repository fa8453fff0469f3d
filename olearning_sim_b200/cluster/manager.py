"""Compute-cluster manager: create / modify / delete / query the worker pool a task's jobs run on.

Counterpart of ols_core/rayclusterMgr/ (1 326 LoC): ``RayClusterManager`` servicer (kuberay_cluster_manager.py:59-225),
``RayClusterApi`` — K8s CustomObjects CRUD on ``ray.io/v1alpha1 rayclusters`` (kuberay_cluster_api.py:37-304),
``ClusterBuilder`` / ``Director`` — fluent CR builder with small / medium / large presets
(kuberay_cluster_builder.py:41-303), ``ClusterUtils`` — worker-group edits (kuberay_cluster_utils.py:25-467) and the
constants / defaults model (kuberay_cluster_constants.py:5-26).

On a B200 box the "cluster" is a head (the control-plane process) plus worker groups of GPU ranks (one process per
GPU).  The same resource document shape is kept (``metadata`` / ``spec.headGroupSpec`` / ``spec.workerGroupSpecs`` with
``replicas`` / ``minReplicas`` / ``maxReplicas``) so cluster descriptions remain portable; the API backend is pluggable:
``LocalClusterApi`` keeps the documents in a registry and clamps replicas to the GPUs actually present, a Kubernetes
backend can be dropped in where the ``kubernetes`` client is installed.
"""
from __future__ import annotations

import copy
import json
import threading
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

from ..utils.logger import Logger

log = Logger()          # (the reference uses an undefined name `log` in two modules; here it exists)

GROUP, VERSION, PLURAL, KIND = "olsb200.io", "v1", "gpuclusters", "GpuCluster"


@dataclass
class ClusterDefaults:
    """kuberay_cluster_constants.RayCluster defaults, re-expressed for GPU workers."""
    k8s_namespace: str = "default"
    ray_name: str = "olsb-demo1"
    ray_label: str = "demo1-cluster"
    head_image: str = "olsb200:latest"
    head_cpu_request: str = "2"
    head_cpu_limit: str = "3"
    head_mem_request: str = "1G"
    head_mem_limit: str = "2G"
    worker_image: str = "olsb200:latest"
    worker_cpu_request: str = "2"
    worker_cpu_limit: str = "3"
    worker_mem_request: str = "1G"
    worker_mem_limit: str = "2G"
    worker_gpus: int = 1
    worker_replicas: int = 4
    worker_minReplicas: int = 1
    worker_maxReplicas: int = 8


class ClusterBuilder:
    """Fluent builder of the cluster resource document."""

    def __init__(self) -> None:
        self.cluster: Dict[str, Any] = {}
        self.succeeded = False

    def build_meta(self, name: str, k8s_namespace: str = "default", labels: Optional[dict] = None, version: str = "1.0"):
        self.cluster = {"apiVersion": f"{GROUP}/{VERSION}", "kind": KIND,
                        "metadata": {"name": name, "namespace": k8s_namespace, "labels": dict(labels or {})},
                        "spec": {"version": version, "workerGroupSpecs": []}}
        self.succeeded = bool(name)
        return self

    def build_head(self, image: str = "olsb200:latest", service_type: str = "ClusterIP", cpu_requests: str = "2",
                   memory_requests: str = "1G", cpu_limits: Optional[str] = None, memory_limits: Optional[str] = None,
                   start_params: Optional[dict] = None):
        if not self.cluster:
            self.succeeded = False
            return self
        self.cluster["spec"]["headGroupSpec"] = {
            "serviceType": service_type, "startParams": dict(start_params or {"dashboard-host": "127.0.0.1"}),
            "template": {"image": image, "resources": {"requests": {"cpu": cpu_requests, "memory": memory_requests},
                                                       "limits": {"cpu": cpu_limits or cpu_requests,
                                                                  "memory": memory_limits or memory_requests}}}}
        return self

    def build_worker(self, group_name: str, image: str = "olsb200:latest", replicas: int = 1, min_replicas: int = 0,
                     max_replicas: int = 8, cpu_requests: str = "1", memory_requests: str = "1G", gpus: int = 1,
                     cpu_limits: Optional[str] = None, memory_limits: Optional[str] = None):
        if not self.cluster or not group_name or not (min_replicas <= replicas <= max_replicas):
            self.succeeded = False
            return self
        self.cluster["spec"]["workerGroupSpecs"].append({
            "groupName": group_name, "replicas": replicas, "minReplicas": min_replicas, "maxReplicas": max_replicas,
            "template": {"image": image, "resources": {
                "requests": {"cpu": cpu_requests, "memory": memory_requests, "nvidia.com/gpu": gpus},
                "limits": {"cpu": cpu_limits or cpu_requests, "memory": memory_limits or memory_requests,
                           "nvidia.com/gpu": gpus}}}})
        return self

    def get_cluster(self) -> Dict[str, Any]:
        spec = self.cluster.get("spec", {})
        self.succeeded = self.succeeded and "headGroupSpec" in spec
        return copy.deepcopy(self.cluster)


class Director:
    """Canned cluster sizes (kuberay_cluster_builder.py:188-303)."""

    def __init__(self) -> None:
        self.builder = ClusterBuilder()

    def _build(self, name, ns, labels, head_cpu, head_mem, replicas, wcpu, wmem):
        return (self.builder.build_meta(name, ns, labels).build_head(cpu_requests=head_cpu, memory_requests=head_mem)
                .build_worker("workers", replicas=replicas, min_replicas=0, max_replicas=max(replicas, 8),
                              cpu_requests=wcpu, memory_requests=wmem).get_cluster())

    def build_basic_cluster(self, name, k8s_namespace="default", labels=None):
        return (self.builder.build_meta(name, k8s_namespace, labels).build_head().get_cluster())

    def build_small_cluster(self, name, k8s_namespace="default", labels=None):
        return self._build(name, k8s_namespace, labels, "2", "2G", 1, "2", "8G")

    def build_medium_cluster(self, name, k8s_namespace="default", labels=None):
        return self._build(name, k8s_namespace, labels, "4", "4G", 4, "4", "32G")

    def build_large_cluster(self, name, k8s_namespace="default", labels=None):
        return self._build(name, k8s_namespace, labels, "8", "8G", 8, "8", "64G")


class ClusterUtils:
    """Worker-group edit helpers; every function returns (document, succeeded)."""

    @staticmethod
    def _find(cluster: Dict[str, Any], group_name: str) -> Optional[Dict[str, Any]]:
        for g in cluster.get("spec", {}).get("workerGroupSpecs", []):
            if g.get("groupName") == group_name:
                return g
        return None

    def update_worker_group_replicas(self, cluster, group_name, max_replicas, min_replicas, replicas) -> Tuple[dict, bool]:
        c = copy.deepcopy(cluster)
        g = self._find(c, group_name)
        if g is None or min(max_replicas, min_replicas, replicas) < 0 or not (min_replicas <= replicas <= max_replicas):
            return c, False
        g.update(replicas=replicas, minReplicas=min_replicas, maxReplicas=max_replicas)
        return c, True

    def update_worker_group_resources(self, cluster, group_name, cpu_requests=None, memory_requests=None, gpus=None,
                                      cpu_limits=None, memory_limits=None) -> Tuple[dict, bool]:
        c = copy.deepcopy(cluster)
        g = self._find(c, group_name)
        if g is None:
            return c, False
        res = g["template"]["resources"]
        for key, req, lim in (("cpu", cpu_requests, cpu_limits), ("memory", memory_requests, memory_limits), ("nvidia.com/gpu", gpus, gpus)):
            if req is not None:
                res["requests"][key] = req
                res["limits"][key] = lim if lim is not None else req
        return c, True

    def duplicate_worker_group(self, cluster, group_name, new_group_name) -> Tuple[dict, bool]:
        c = copy.deepcopy(cluster)
        g = self._find(c, group_name)
        if g is None or self._find(c, new_group_name) is not None:
            return c, False
        ng = copy.deepcopy(g)
        ng["groupName"] = new_group_name
        c["spec"]["workerGroupSpecs"].append(ng)
        return c, True

    def delete_worker_group(self, cluster, group_name) -> Tuple[dict, bool]:
        c = copy.deepcopy(cluster)
        groups = c.get("spec", {}).get("workerGroupSpecs", [])
        n = len(groups)
        c["spec"]["workerGroupSpecs"] = [g for g in groups if g.get("groupName") != group_name]
        return c, len(c["spec"]["workerGroupSpecs"]) != n

    def is_valid_name(self, name: str) -> bool:
        import re
        return bool(re.fullmatch(r"[a-z0-9]([-a-z0-9]*[a-z0-9])?", name or "")) and len(name) <= 63

    def is_valid_label(self, label: str) -> bool:
        return self.is_valid_name(label.replace(".", "-").replace("_", "-").lower()) if label else False


class LocalClusterApi:
    """In-memory CRUD with the KubeRay python-client method names (create / get / list / patch / delete)."""

    def __init__(self, gpu_count: Optional[int] = None):
        self._items: Dict[Tuple[str, str], Dict[str, Any]] = {}
        self._lock = threading.Lock()
        if gpu_count is None:
            try:
                import torch
                gpu_count = torch.cuda.device_count() if torch.cuda.is_available() else 0
            except Exception:  # noqa: BLE001
                gpu_count = 0
        self.gpu_count = gpu_count

    def _status(self, doc: Dict[str, Any]) -> Dict[str, Any]:
        want = sum(g.get("replicas", 0) * g["template"]["resources"]["requests"].get("nvidia.com/gpu", 0)
                   for g in doc["spec"].get("workerGroupSpecs", []))
        return {"state": "ready" if want <= max(self.gpu_count, 0) or self.gpu_count == 0 else "pending",
                "desiredGpus": want, "availableGpus": self.gpu_count}

    def create_ray_cluster(self, body: Dict[str, Any], k8s_namespace: Optional[str] = None) -> Optional[Dict[str, Any]]:
        ns = k8s_namespace or body["metadata"].get("namespace", "default")
        key = (ns, body["metadata"]["name"])
        with self._lock:
            if key in self._items:
                return None
            doc = copy.deepcopy(body)
            doc["status"] = self._status(doc)
            self._items[key] = doc
            return copy.deepcopy(doc)

    def get_ray_cluster(self, name: str, k8s_namespace: str = "default") -> Optional[Dict[str, Any]]:
        with self._lock:
            d = self._items.get((k8s_namespace, name))
            return copy.deepcopy(d) if d else None

    def list_ray_clusters(self, k8s_namespace: str = "default", label_selector: str = "") -> Dict[str, Any]:
        want = dict(kv.split("=", 1) for kv in label_selector.split(",") if "=" in kv)
        with self._lock:
            items = [copy.deepcopy(d) for (ns, _), d in self._items.items() if ns == k8s_namespace
                     and all(d["metadata"].get("labels", {}).get(k) == v for k, v in want.items())]
        return {"items": items}

    def patch_ray_cluster(self, name: str, ray_patch: Dict[str, Any], k8s_namespace: str = "default") -> bool:
        with self._lock:
            key = (k8s_namespace, name)
            if key not in self._items:
                return False
            doc = copy.deepcopy(ray_patch)
            doc["status"] = self._status(doc)
            self._items[key] = doc
            return True

    def delete_ray_cluster(self, name: str, k8s_namespace: str = "default") -> bool:
        with self._lock:
            return self._items.pop((k8s_namespace, name), None) is not None


class KubeRestCustomObjects:
    """Dependency-free client for the custom-objects endpoints of a Kubernetes API server (the five verbs of
    ``kubernetes.client.CustomObjectsApi`` that KubeClusterApi uses), over ``requests``:

        POST / GET(list)   /apis/{group}/{version}/namespaces/{ns}/{plural}
        GET / PATCH / DELETE  .../{plural}/{name}          (PATCH as application/merge-patch+json)

    ``from_env()`` picks up the in-cluster service account (token + CA under /var/run/secrets/kubernetes.io/serviceaccount,
    KUBERNETES_SERVICE_HOST/PORT) or ``KUBE_API_URL`` [+ ``KUBE_TOKEN``, ``KUBE_CA_FILE``].  Non-2xx answers raise
    ``KubeApiError`` (status + the server's Status message) — KubeClusterApi turns them into logged failures, like the
    reference does with ``ApiException`` (kuberay_cluster_api.py:60-75)."""

    SA_DIR = "/var/run/secrets/kubernetes.io/serviceaccount"

    def __init__(self, base_url: str, token: Optional[str] = None, verify: Any = True, timeout: float = 10.0):
        import requests
        self.base, self.timeout = base_url.rstrip("/"), timeout
        self.s = requests.Session()
        self.s.verify = verify
        if token:
            self.s.headers["Authorization"] = f"Bearer {token}"

    @classmethod
    def from_env(cls) -> "KubeRestCustomObjects":
        import os
        url = os.environ.get("KUBE_API_URL")
        token, ca = os.environ.get("KUBE_TOKEN"), os.environ.get("KUBE_CA_FILE")
        if not url and os.environ.get("KUBERNETES_SERVICE_HOST"):
            url = f"https://{os.environ['KUBERNETES_SERVICE_HOST']}:{os.environ.get('KUBERNETES_SERVICE_PORT', '443')}"
            try:
                with open(os.path.join(cls.SA_DIR, "token")) as f:
                    token = token or f.read().strip()
            except OSError:
                pass
            ca = ca or (os.path.join(cls.SA_DIR, "ca.crt") if os.path.exists(os.path.join(cls.SA_DIR, "ca.crt")) else None)
        if not url:
            raise RuntimeError("no Kubernetes API server: set KUBE_API_URL (and KUBE_TOKEN) or run inside a pod")
        return cls(url, token, ca if ca else True)

    def _url(self, group, version, namespace, plural, name=None) -> str:
        u = f"{self.base}/apis/{group}/{version}/namespaces/{namespace}/{plural}"
        return f"{u}/{name}" if name else u

    def _do(self, method: str, url: str, **kw) -> Dict[str, Any]:
        r = self.s.request(method, url, timeout=self.timeout, **kw)
        if r.status_code // 100 != 2:
            try:
                msg = r.json().get("message", r.text)
            except ValueError:
                msg = r.text
            raise KubeApiError(r.status_code, msg)
        return r.json() if r.content else {}

    def create_namespaced_custom_object(self, group, version, namespace, plural, body):
        return self._do("POST", self._url(group, version, namespace, plural), json=body)

    def get_namespaced_custom_object(self, group, version, namespace, plural, name):
        return self._do("GET", self._url(group, version, namespace, plural, name))

    def list_namespaced_custom_object(self, group, version, namespace, plural, label_selector=""):
        return self._do("GET", self._url(group, version, namespace, plural),
                        params={"labelSelector": label_selector} if label_selector else None)

    def patch_namespaced_custom_object(self, group, version, namespace, plural, name, body):
        return self._do("PATCH", self._url(group, version, namespace, plural, name), data=json.dumps(body),
                        headers={"Content-Type": "application/merge-patch+json"})

    def delete_namespaced_custom_object(self, group, version, namespace, plural, name):
        return self._do("DELETE", self._url(group, version, namespace, plural, name))


class KubeApiError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{status}: {message}")
        self.status = status


class KubeClusterApi:
    """The same five calls against a Kubernetes API server (custom objects of ``GROUP/VERSION``, plural ``PLURAL``) — the
    counterpart of the reference's KubeRay client (kuberay_cluster_api.py:37-304).  Transport: the official ``kubernetes``
    client when it is installed, otherwise the built-in REST client ``KubeRestCustomObjects`` (same verbs over
    ``requests``); ``api`` may also be injected.  Tested end to end against an in-process HTTP API server
    (tests/test_infra.py)."""

    def __init__(self, api: Optional[Any] = None, group: str = GROUP, version: str = VERSION, plural: str = PLURAL):
        if api is None:
            try:
                from kubernetes import client, config          # type: ignore
                try:
                    config.load_incluster_config()
                except Exception:  # noqa: BLE001 - outside a pod: fall back to the user's kubeconfig
                    config.load_kube_config()
                api = client.CustomObjectsApi()
            except ImportError:
                api = KubeRestCustomObjects.from_env()
        self.api, self.group, self.version, self.plural = api, group, version, plural

    def _call(self, fn, *a, **kw):
        try:
            return fn(*a, **kw)
        except Exception as e:  # noqa: BLE001 - ApiException and transport errors: log and report failure
            log.error("", "cluster", "KubeClusterApi", f"{getattr(fn, '__name__', fn)} failed: {e}")
            return None

    def create_ray_cluster(self, body: Dict[str, Any], k8s_namespace: Optional[str] = None) -> Optional[Dict[str, Any]]:
        ns = k8s_namespace or body["metadata"].get("namespace", "default")
        return self._call(self.api.create_namespaced_custom_object, group=self.group, version=self.version,
                          namespace=ns, plural=self.plural, body=body)

    def get_ray_cluster(self, name: str, k8s_namespace: str = "default") -> Optional[Dict[str, Any]]:
        return self._call(self.api.get_namespaced_custom_object, group=self.group, version=self.version,
                          namespace=k8s_namespace, plural=self.plural, name=name)

    def list_ray_clusters(self, k8s_namespace: str = "default", label_selector: str = "") -> Dict[str, Any]:
        out = self._call(self.api.list_namespaced_custom_object, group=self.group, version=self.version,
                         namespace=k8s_namespace, plural=self.plural, label_selector=label_selector)
        return out if out is not None else {"items": []}

    def patch_ray_cluster(self, name: str, ray_patch: Dict[str, Any], k8s_namespace: str = "default") -> bool:
        return self._call(self.api.patch_namespaced_custom_object, group=self.group, version=self.version,
                          namespace=k8s_namespace, plural=self.plural, name=name, body=ray_patch) is not None

    def delete_ray_cluster(self, name: str, k8s_namespace: str = "default") -> bool:
        return self._call(self.api.delete_namespaced_custom_object, group=self.group, version=self.version,
                          namespace=k8s_namespace, plural=self.plural, name=name) is not None


class GpuClusterManager:
    """Servicer with the reference's four RPC names (kept verbatim for API parity)."""

    def __init__(self, api: Optional[Any] = None, defaults: Optional[ClusterDefaults] = None):
        self._api = api or LocalClusterApi()
        self._d = defaults or ClusterDefaults()
        self._builder = ClusterBuilder()
        self._utils = ClusterUtils()

    def _build(self, ns: str, name: str, label: str) -> Dict[str, Any]:
        d = self._d
        return (self._builder.build_meta(name=name, k8s_namespace=ns, labels={label: "yes"})
                .build_head(image=d.head_image, cpu_requests=d.head_cpu_request, memory_requests=d.head_mem_request,
                            cpu_limits=d.head_cpu_limit, memory_limits=d.head_mem_limit)
                .build_worker(group_name="workers", image=d.worker_image, replicas=d.worker_replicas,
                              min_replicas=d.worker_minReplicas, max_replicas=d.worker_maxReplicas,
                              cpu_requests=d.worker_cpu_request, memory_requests=d.worker_mem_request, gpus=d.worker_gpus,
                              cpu_limits=d.worker_cpu_limit, memory_limits=d.worker_mem_limit).get_cluster())

    def createRayCluster(self, request: Optional[Dict[str, Any]] = None) -> Dict[str, bool]:
        try:
            cluster = self._build(self._d.k8s_namespace, self._d.ray_name, self._d.ray_label)
            if not self._builder.succeeded:
                log.error("", "ClusterMgr", "manager", "[createRayCluster]: error building the cluster, aborting")
                return {"ok": False}
            return {"ok": self._api.create_ray_cluster(body=cluster) is not None}
        except Exception as e:  # noqa: BLE001
            log.error("", "ClusterMgr", "manager", f"[createRayCluster]: {e}")
            return {"ok": False}

    def modifyRayCluster(self, request: Dict[str, Any]) -> Dict[str, bool]:
        try:
            ns, name, label = request.get("k8s_namespace", ""), request.get("ray_name", ""), request.get("ray_label", "")
            mn, cur, mx = (request.get(k, -1) for k in ("worker_min_replicas", "worker_replicas", "worker_max_replicas"))
            if not ns or not name or not label or min(mn, cur, mx) < 0:
                return {"ok": False}
            existing = self._api.get_ray_cluster(name, ns) or self._build(ns, name, label)
            patched, ok = self._utils.update_worker_group_replicas(existing, "workers", mx, mn, cur)
            if not ok:
                return {"ok": False}
            return {"ok": bool(self._api.patch_ray_cluster(name=name, ray_patch=patched, k8s_namespace=ns))}
        except Exception as e:  # noqa: BLE001
            log.error("", "ClusterMgr", "manager", f"[modifyRayCluster]: {e}")
            return {"ok": False}

    def deleteRayCluster(self, request: Dict[str, Any]) -> Dict[str, bool]:
        try:
            ns = request.get("k8s_namespace", "")
            if not ns:
                return {"ok": False}
            label = request.get("ray_label", "")
            lst = self._api.list_ray_clusters(ns, f"{label}=yes" if label else "")
            for c in lst.get("items", []):
                self._api.delete_ray_cluster(c["metadata"]["name"], c["metadata"]["namespace"])
            return {"ok": True}
        except Exception as e:  # noqa: BLE001
            log.error("", "ClusterMgr", "manager", f"[deleteRayCluster]: {e}")
            return {"ok": False}

    def queryRayCluster(self, request: Dict[str, Any]) -> Dict[str, str]:
        ns = request.get("k8s_namespace", "")
        if not ns:
            return {"json_data": ""}
        label = request.get("ray_label", "")
        return {"json_data": json.dumps(self._api.list_ray_clusters(ns, f"{label}=yes" if label else ""))}


RayClusterManager = GpuClusterManager      # reference name
