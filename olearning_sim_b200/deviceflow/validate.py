"""Validation of DeviceFlow requests and strategy JSON.

Rules from ols_core/deviceflow/utils/validate_parameters.py:12-233 (exactly one of real_time/flow, exactly one of
specific_timing/specific_interval, time type / zone, drop settings, sizes, monotone intervals, evaluable rate
functions) — with the rate functions checked by ``safe_expr`` instead of ``eval``.
"""
from __future__ import annotations

import json
from datetime import datetime
from enum import Enum
from typing import Any, Dict, Tuple
from zoneinfo import ZoneInfo, ZoneInfoNotFoundError

from . import safe_expr

TIME_FORMAT = "%Y-%m-%d %H:%M:%S"


class ComputeResource(Enum):
    logical_simulation = 1
    device_simulation = 2


class StrategyTimeType(Enum):
    absolute = 1
    relative = 2


def _increasing(flat) -> bool:
    """[[a,b],[c,d]]: a<b (strict inside an interval), b<=c (intervals may touch)"""
    for i in range(len(flat) - 1):
        if i % 2 == 0:
            if flat[i] >= flat[i + 1]:
                return False
        elif flat[i] > flat[i + 1]:
            return False
    return True


class ValidateParameters:
    def check_params_of_notify_start(self, request: Any) -> Tuple[bool, str]:
        try:
            strategy = json.loads(request.strategy)
        except Exception:  # noqa: BLE001
            return False, "strategy not json format"
        if request.compute_resource not in ComputeResource.__members__:
            return False, "compute resource error"
        return self.check_strategy(strategy)

    def check_strategy(self, strategy: Dict[str, Any]) -> Tuple[bool, str]:
        if not isinstance(strategy, dict):
            return False, "strategy not json object"
        rt = strategy.get("real_time_dispatch", {})
        fd = strategy.get("flow_dispatch", {})
        use_rt, use_fd = bool(rt.get("use_strategy", False)), bool(fd.get("use_strategy", False))
        if use_rt == use_fd:
            return False, "Must use one strategy"
        return self.check_real_time_dispatch_strategy(rt) if use_rt else self.check_flow_dispatch_strategy(fd)

    def check_real_time_dispatch_strategy(self, rt: Dict[str, Any]) -> Tuple[bool, str]:
        p = rt.get("drop_simulation", {}).get("drop_probability", -1)
        if p != -1 and not (isinstance(p, (int, float)) and 0 <= p <= 1):
            return False, "drop probability must in [0,1]"
        sizes = rt.get("dispatch_batch_sizes", [])
        if not all(isinstance(s, int) and not isinstance(s, bool) and s > 0 for s in sizes):
            return False, "dispatch batch sizes must be positive integers"
        return True, "Pass"

    def check_flow_dispatch_strategy(self, fd: Dict[str, Any]) -> Tuple[bool, str]:
        total = fd.get("total_dispatch_amount", -1)
        st, si = fd.get("specific_timing", {}), fd.get("specific_interval", {})
        use_t, use_i = bool(st.get("use", False)), bool(si.get("use", False))
        if use_t == use_i:
            return False, "Must use one specific strategy"
        sec = st if use_t else si
        time_type, time_zone = sec.get("time_type", ""), sec.get("time_zone", "")
        if time_type == "":
            return False, "time type error"
        if time_type not in StrategyTimeType.__members__:
            return False, "time type error, absolute or relative need"
        if time_type == "absolute":
            if time_zone == "":
                return False, "time zone error"
            try:
                ZoneInfo(time_zone)
            except (ZoneInfoNotFoundError, ValueError):
                return False, "time zone error, format must be a known IANA zone"
        dp = sec.get("drop_simulation", {}).get("drop_probability", [])
        da = sec.get("drop_simulation", {}).get("drop_amounts", [])
        if dp and da:
            return False, "drop probability and drop amounts can't be set at the same time"
        if dp:
            if any(not 0 <= p <= 1 for p in dp):
                return False, "drop probability must in [0,1]"
        elif da and total < sum(da):
            return False, "drop amounts sum > total dispatch amount"
        try:
            if use_t:
                amounts = st.get("amounts", [])
                timings_list = [st.get("timings", [])] if time_type == "relative" else st.get("timings", [])
                for timings in timings_list:
                    if len(amounts) != len(timings):
                        return False, "amounts and timings must have the same size"
                    if dp and len(amounts) != len(dp):
                        return False, "amounts, timings and drop_probability must have the same size"
                    if da and len(amounts) != len(da):
                        return False, "amounts, timings and drop_amounts must have the same size"
                    if total != sum(amounts):
                        return False, "amounts not equal total dispatch amount"
                    if time_type == "absolute":
                        for t in timings:
                            try:
                                datetime.strptime(t, TIME_FORMAT)
                            except (ValueError, TypeError):
                                return False, "absolute time format error, must %Y-%m-%d %H:%M:%S"
                    else:
                        for t in timings:
                            if not isinstance(t, (int, float)) or isinstance(t, bool):
                                return False, "relative time format error, must figure"
                            if t < 0:
                                return False, "relative time format error, must >= 0"
            else:
                intervals_list = [si.get("intervals", [])] if time_type == "relative" else si.get("intervals", [])
                domains = si.get("dispatch_rules", {}).get("domains", [])
                functions = si.get("dispatch_rules", {}).get("functions", [])
                for intervals in intervals_list:
                    flat = [x for iv in intervals for x in iv]
                    if time_type == "absolute":
                        try:
                            stamps = [datetime.strptime(t, TIME_FORMAT).timestamp() for t in flat]
                        except (ValueError, TypeError):
                            return False, "absolute time format error, must %Y-%m-%d %H:%M:%S"
                        if not _increasing(stamps):
                            return False, "absolute time value error"
                    else:
                        if not all(isinstance(x, int) and not isinstance(x, bool) for x in flat):
                            return False, "relative time format error, must be integer seconds"
                        if not all(x >= 0 for x in flat[:-1]):
                            return False, "relative time format error, must >= 0"
                        if not _increasing(flat):
                            return False, "relative time value error"
                    if not len(intervals) == len(domains) == len(functions):
                        return False, "intervals, domains and functions must have the same size"
                    if dp and len(intervals) != len(dp):
                        return False, "intervals, domains, functions and drop_probability must have the same size"
                    if da and len(intervals) != len(da):
                        return False, "intervals, domains, functions and drop_amounts must have the same size"
                    for dom, fn in zip(domains, functions):
                        if dom[0] >= dom[1]:
                            return False, "domains right value must be greater than the left value"
                        try:
                            float(safe_expr.evaluate(fn, dom[0]))
                        except Exception:  # noqa: BLE001 — includes UnsafeExpression
                            return False, "domains or functions error, variable must be t"
        except Exception as e:  # noqa: BLE001
            return False, f"{e}"
        return True, "Pass"
