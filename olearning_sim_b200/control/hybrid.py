"""Hybrid (server-simulated vs real-phone) allocation of a dataset's clients.

Counterpart of ``HybridOptimizer`` / ``auto_allocation_hybrid_task`` (ols_core/taskMgr/utils/utils_runner.py:23-176,
939-1022).  The reference formulates, per device class i,

    minimise  z   s.t.   z >= alpha * ceil(x_i * k_i / f_i),   z >= beta * ceil((N_i - q_i - x_i) / m_i) + lambda,
    0 <= x_i <= N_i - q_i integer

with alpha=3.5 s (logical per round-unit), beta=0.14 s (phone per round-unit), lambda=8.808 s (phone start-up) and
solves it with PuLP/CBC.  The classes only interact through the shared makespan ``z``, so the optimum is the max of
the per-class optima and each class is a 1-D problem: the logical term is non-decreasing in x, the phone term
non-increasing, so an exact binary search (plus a local scan for ties) replaces the ILP solver.  Among equally good
x we return the smallest makespan, then the largest x (keeps phones free) — any minimiser is a valid ILP answer.
"""
from __future__ import annotations

import copy
import math
import os
from typing import Any, Dict, List, Tuple

ALPHA, BETA, LAMBDA = 3.5, 0.14, 8.808


def rearrange_by_devices(devices: List[str], keys: List[str], values: List[Any], default: Any) -> List[Any]:
    d = dict(zip(keys, values))
    return [d.get(dev, default) for dev in devices]


def _cost(x: int, N: int, f: int, k: int, m: int, q: int, alpha: float, beta: float, lam: float) -> float:
    t_log = alpha * math.ceil(x * k / f)
    t_ph = beta * math.ceil((N - q - x) / m) + lam
    return max(t_log, t_ph)


def solve_class(N: int, f: int, k: int, m: int, q: int, alpha: float = ALPHA, beta: float = BETA,
                lam: float = LAMBDA) -> Tuple[int, float]:
    """Exact argmin over x in [0, N-q] of max(alpha*ceil(x*k/f), beta*ceil((N-q-x)/m)+lam)."""
    hi = N - q
    if hi <= 0:
        return 0, _cost(0, N, f, k, m, q, alpha, beta, lam)
    # largest x with logical term <= phone term (monotone crossing)
    lo_x, hi_x = 0, hi
    while lo_x < hi_x:
        mid = (lo_x + hi_x + 1) // 2
        if alpha * math.ceil(mid * k / f) <= beta * math.ceil((N - q - mid) / m) + lam:
            lo_x = mid
        else:
            hi_x = mid - 1
    best_x, best_c = 0, float("inf")
    # the optimum sits at the crossing; ceil() plateaus can shift it by up to one plateau (f/k or m wide)
    span = max(2, f // max(k, 1) + 2, m + 2)
    for x in range(max(0, lo_x - span), min(hi, lo_x + span) + 1):
        c = _cost(x, N, f, k, m, q, alpha, beta, lam)
        if c < best_c - 1e-12 or (abs(c - best_c) <= 1e-12 and x > best_x):
            best_x, best_c = x, c
    return best_x, best_c


def brute_force_class(N: int, f: int, k: int, m: int, q: int, alpha: float = ALPHA, beta: float = BETA,
                      lam: float = LAMBDA) -> float:
    return min(_cost(x, N, f, k, m, q, alpha, beta, lam) for x in range(0, max(N - q, 0) + 1))


def auto_allocation_hybrid_task(data_dict: Dict[str, List[int]]) -> Tuple[List[int], List[int]]:
    """data_dict = {N, f, k, m, q} per device class -> (allocation_logical, allocation_device)."""
    N, f, k, m, q = (list(data_dict[x]) for x in ("N", "f", "k", "m", "q"))
    n = len(N)
    logical, device = [0] * n, [0] * n
    for i in range(n):
        if f[i] == 0:
            device[i] = N[i]
        elif m[i] == 0:
            logical[i] = N[i]
        else:
            x, _ = solve_class(N[i], f[i], max(k[i], 1), m[i], q[i])
            logical[i], device[i] = x, N[i] - x
    return logical, device


class HybridOptimizer:
    """Fills ``allocation`` (when optimization is on) and the per-side data paths of every target dataset."""

    def __init__(self, task_json: Dict[str, Any], data_folder: str = "data_split"):
        self._task_json = task_json
        self._data_folder = data_folder

    def fix_data_parameters(self) -> List[Dict[str, Any]]:
        tj = self._task_json
        datas = tj.get("target", {}).get("data", [])
        cu = tj.get("logical_simulation", {}).get("computation_unit", {})
        rl = {r.get("name"): r for r in tj.get("logical_simulation", {}).get("resource_request", [])}
        rd = {r.get("name"): r for r in tj.get("device_simulation", {}).get("resource_request", [])}
        out = []
        for d in datas:
            d = copy.deepcopy(d)
            if d.get("allocation", {}).get("optimization", False):
                d = self.hybrid_optimization(d, cu, rl.get(d.get("name"), {}), rd.get(d.get("name"), {}))
            else:
                d = self.hybrid_setting_by_user(d)
            out.append(d)
        return out

    def _set_paths(self, d: Dict[str, Any], logical_num: int, device_num: int) -> Dict[str, Any]:
        name, path = d.get("name", ""), d.get("data_path", "")
        folder = os.path.dirname(path)
        if logical_num == 0 and device_num == 0:
            d.update(data_path_logical="", data_path_device="")
        elif device_num == 0:
            d.update(data_path_logical=path, data_path_device="")
        elif logical_num == 0:
            d.update(data_path_logical="", data_path_device=path)
        else:
            d.update(data_path_logical=f"{folder}/{self._data_folder}/{name}_logical.zip",
                     data_path_device=f"{folder}/{self._data_folder}/{name}_device.zip")
        return d

    def hybrid_setting_by_user(self, d: Dict[str, Any]) -> Dict[str, Any]:
        al = d.get("allocation", {})
        return self._set_paths(d, sum(al.get("logical_simulation", [])), sum(al.get("device_simulation", [])))

    def hybrid_optimization(self, d, computation_unit, logical_resource, device_resource) -> Dict[str, Any]:
        devices = d.get("total_simulation", {}).get("devices", [])
        nums = list(d.get("total_simulation", {}).get("nums", []))
        rr = d.get("allocation", {}).get("running_response", {})
        q = rearrange_by_devices(devices, rr.get("devices", []), rr.get("nums", []), 0)
        k = [s.get("num_cpus", 0) if isinstance(s, dict) else 0 for s in rearrange_by_devices(
            devices, computation_unit.get("devices", []), computation_unit.get("setting", []), {"num_cpus": 0})]
        f = rearrange_by_devices(devices, logical_resource.get("devices", []), logical_resource.get("num_request", []), 0)
        m = rearrange_by_devices(devices, device_resource.get("devices", []), device_resource.get("num_request", []), 0)
        if sum(f) == 0:
            logical, device = [0] * len(nums), nums
        elif sum(m) == 0:
            logical, device = nums, [0] * len(nums)
        else:
            logical, device = auto_allocation_hybrid_task({"N": nums, "f": f, "k": k, "m": m, "q": q})
        d.setdefault("allocation", {})["logical_simulation"] = logical
        d["allocation"]["device_simulation"] = device
        return self._set_paths(d, sum(logical), sum(device) + sum(rr.get("nums", [])))
