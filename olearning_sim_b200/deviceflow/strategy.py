"""DeviceFlow dispatch strategies: strategy JSON -> (dispatch_timing[], dispatch_amount[], drop_lists[][]).

Semantics follow ols_core/deviceflow/non_grpc/strategy.py line by line:

* real-time   (:19-31)   batch sizes cycling + i.i.d. drop probability
* specific_timing (:73-162)   scripted (delta-t, amount) slots; absolute times -> deltas from "now" in a time zone,
                              sorted, past slots discarded
* specific_interval (:166-273, :314-445)  rate curves f_j(t): each 1-second bin integrated with a 100-segment
                              trapezoid counting only positive area, interval share by area with the last interval
                              taking the remainder, error-diffusion rounding inside an interval
* drops (:275-311, :386-423)  ``drop_probability`` Bernoulli per message index / ``drop_amounts`` exact counts sampled
                              without replacement and bucketed per slot

Differences (deliberate): rate functions go through ``safe_expr`` instead of ``eval``; the wall clock and the RNG
are injectable (``now=``, ``rng=``) so schedules are reproducible; ``zoneinfo`` replaces ``pytz``.
"""
from __future__ import annotations

import copy
import json
import random
from datetime import datetime
from typing import Any, Callable, Dict, List, Optional, Tuple
from zoneinfo import ZoneInfo

import numpy as np

from . import safe_expr

AREA_CALCULATION_NUM = 100
DATE_FORMAT = "%Y-%m-%d %H:%M:%S"
Schedule = Tuple[List[float], List[int], List[List[int]]]
_EMPTY: Schedule = ([], [], [])


def _loads(strategy: Any) -> Dict[str, Any]:
    return json.loads(strategy) if isinstance(strategy, (str, bytes)) else dict(strategy)


def _round_of(flow_id: str) -> int:
    """flow_id = f"{task_id}_{operator_name}_{round_idx}" (ols_core/taskMgr/run_task.py:240)"""
    return int(flow_id.rsplit("_", 1)[1])


class Strategy:
    def __init__(self, rng: Optional[random.Random] = None, now: Optional[Callable[[], datetime]] = None):
        self.rng = rng or random.Random()
        self.now = now or datetime.now

    # ---- real time -------------------------------------------------------------------------------
    @staticmethod
    def check_real_time_dispatch(strategy: Any) -> bool:
        return bool(_loads(strategy).get("real_time_dispatch", {}).get("use_strategy", False))

    @staticmethod
    def real_time_strategy_analysis(strategy: Any) -> Tuple[List[int], float]:
        rt = _loads(strategy).get("real_time_dispatch", {})
        return list(rt.get("dispatch_batch_sizes", [])), rt.get("drop_simulation", {}).get("drop_probability", 0)

    # ---- flow ------------------------------------------------------------------------------------
    @staticmethod
    def flow_strategy_analysis(strategy: Any, flow_id: str, rng: Optional[random.Random] = None,
                               now: Optional[Callable[[], datetime]] = None) -> Schedule:
        fd = _loads(strategy).get("flow_dispatch", {})
        if not fd.get("use_strategy", False):
            return _EMPTY
        total = fd.get("total_dispatch_amount", 0)
        if total <= 0:
            return _EMPTY
        st = fd.get("specific_timing", {})
        si = fd.get("specific_interval", {})
        use_t, use_i = bool(st.get("use", False)), bool(si.get("use", False))
        if use_t == use_i:
            return _EMPTY
        s = Strategy(rng, now)
        if use_t:
            return s._specific_timing_analysis(copy.deepcopy(st), flow_id)
        return s._specific_interval_analysis(total, copy.deepcopy(si), flow_id)

    def _now_in_zone(self, time_zone: str) -> Tuple[datetime, float]:
        """current time as a naive datetime in `time_zone` truncated to seconds + the sub-second part (2 decimals)"""
        cur = self.now()
        frac = round(float(f"0.{cur.microsecond}"), 2)      # reference quirk: "0.<microsecond>" without zero padding
        aware = cur.astimezone() if cur.tzinfo is None else cur
        local = aware.astimezone(ZoneInfo(time_zone)).replace(tzinfo=None, microsecond=0)
        return local, frac

    def _specific_timing_analysis(self, st: Dict[str, Any], flow_id: str) -> Schedule:
        time_type = st.get("time_type", "relative")
        time_zone = st.get("time_zone", "Asia/Shanghai")
        if time_type == "relative":
            timing = list(st.get("timings", []))
        else:
            try:
                timing = list(st.get("timings", [])[_round_of(flow_id)])
            except (IndexError, ValueError, TypeError):
                return _EMPTY
        amount = [int(a) for a in st.get("amounts", [])]
        if len(timing) != len(amount) or not timing:
            return _EMPTY
        drop = st.get("drop_simulation", {})
        if drop:
            if len(drop) != 1:
                return _EMPTY
            drops = self._generate_drop_simulation_list(amount, drop)
        else:
            drops = [[] for _ in amount]
        if time_type == "absolute":
            ts = [datetime.strptime(t, DATE_FORMAT) for t in timing]
            cur, frac = self._now_in_zone(time_zone)
            secs = [(t - cur).total_seconds() for t in ts]
            order = sorted(range(len(secs)), key=lambda i: secs[i])
            secs, amount, drops = [secs[i] for i in order], [amount[i] for i in order], [drops[i] for i in order]
            first = next((i for i, s in enumerate(secs) if s >= 0), -1)
            if first == -1:
                return _EMPTY
            secs, amount, drops = secs[first:], amount[first:], drops[first:]
            timing = [secs[0] - frac] + [secs[i] - secs[i - 1] for i in range(1, len(secs))]
        return timing, amount, drops

    def _specific_interval_analysis(self, total: int, si: Dict[str, Any], flow_id: str) -> Schedule:
        time_type = si.get("time_type", "relative")
        time_zone = si.get("time_zone", "Asia/Shanghai")
        if time_type == "relative":
            intervals = si.get("intervals", [])
        else:
            try:
                intervals = si.get("intervals", [])[_round_of(flow_id)]
            except (IndexError, ValueError, TypeError):
                return _EMPTY
        rules = si.get("dispatch_rules", {})
        domains, functions = rules.get("domains", []), rules.get("functions", [])
        drop = si.get("drop_simulation", {})
        if len(intervals) != len(domains) or len(domains) != len(functions) or not intervals:
            return _EMPTY
        if drop and len(drop) != 1:
            return _EMPTY
        if time_type == "relative":
            return self._get_interval_params(total, intervals, domains, functions, drop)
        # absolute -> relative seconds, first interval starting at 0
        abs_iv = intervals
        rel: List[List[int]] = []
        for i, (a, b) in enumerate(abs_iv):
            start, end = datetime.strptime(a, DATE_FORMAT), datetime.strptime(b, DATE_FORMAT)
            if i == 0:
                s0 = 0
            else:
                prev_end = datetime.strptime(abs_iv[i - 1][1], DATE_FORMAT)
                s0 = int((start - prev_end).total_seconds()) + rel[i - 1][1]
            rel.append([s0, int((end - start).total_seconds()) + s0])
        timing, amount, drops = self._get_interval_params(total, rel, domains, functions, drop)
        if not timing:
            return _EMPTY
        cur, frac = self._now_in_zone(time_zone)
        start0 = datetime.strptime(abs_iv[0][0], DATE_FORMAT)
        timing[0] = int((start0 - cur).total_seconds()) - frac
        cum = [timing[0]]
        for i in range(1, len(timing)):
            cum.append(cum[-1] + timing[i])
        first = next((i for i, s in enumerate(cum) if s >= 0), -1)
        if first == -1:
            return _EMPTY
        if first > 0:
            timing, amount, drops = timing[first:], amount[first:], drops[first:]
            timing[0] = cum[first]
        return timing, amount, drops

    # ---- drops -----------------------------------------------------------------------------------
    def _generate_drop_simulation_list(self, dispatch_amount: List[int], drop: Dict[str, Any]) -> List[List[int]]:
        out: List[List[int]] = []
        if "drop_probability" in drop:
            for p, amount in zip(drop.get("drop_probability", []), dispatch_amount):
                if p == 0:
                    idx: List[int] = []
                elif p == 1:
                    idx = list(range(int(amount)))
                elif 0 < p < 1:
                    idx = [i for i in range(int(amount)) if self.rng.random() < p]
                else:
                    idx = []
                out.append(idx)
            return out
        if "drop_amounts" in drop:
            for d, amount in zip(drop.get("drop_amounts", []), dispatch_amount):
                if d == 0:
                    idx = []
                elif 0 < d and int(d) < int(amount):
                    idx = sorted(self.rng.sample(range(int(amount)), int(d)))
                else:
                    idx = list(range(int(amount)))
                out.append(idx)
            return out
        return []

    # ---- rate curves -----------------------------------------------------------------------------
    @staticmethod
    def interval_areas(interval: List[int], domain: List[float], func: str) -> Tuple[List[int], List[float]]:
        """per 1-second bin: positive-area trapezoid integral of f over the bin mapped onto the domain"""
        f = safe_expr.compile_expr(func)
        t_int = list(range(interval[0], interval[1] + 1))
        length, dlen = interval[1] - interval[0], domain[1] - domain[0]
        t_dom = [domain[0] + dlen / length * (t - t_int[0]) for t in t_int]
        areas = []
        for i in range(len(t_dom) - 1):
            ts = np.linspace(t_dom[i], t_dom[i + 1], num=AREA_CALCULATION_NUM + 1).tolist()
            ys = [f(t) for t in ts]
            area = 0.0
            for j in range(1, len(ys)):
                a = 0.5 * (ys[j] + ys[j - 1]) * (1 / AREA_CALCULATION_NUM)
                if a > 0:
                    area += a
            areas.append(area)
        return t_int[:-1], areas

    @staticmethod
    def diffuse(send_nums: List[float]) -> List[int]:
        """error-diffusion rounding (strategy.py:370-382): carry the residual, emit 0 while round(tmp) <= 0"""
        out, carry = [], 0.0
        for s in send_nums:
            tmp = carry + s
            r = round(tmp)
            if r > 0:
                out.append(r)
                carry = tmp - r
            else:
                out.append(0)
                carry = tmp
        return out

    @staticmethod
    def interval_samples(interval: List[int], domain: List[float], func: str) -> np.ndarray:
        """f sampled at the 101 trapezoid nodes of every 1-second bin: float64 [bins, 101] — the input of the on-device
        integrator (``rate_curve_area_kernel``), which reproduces ``interval_areas`` bit for bit."""
        f = safe_expr.compile_expr(func)
        t_int = list(range(interval[0], interval[1] + 1))
        length, dlen = interval[1] - interval[0], domain[1] - domain[0]
        t_dom = [domain[0] + dlen / length * (t - t_int[0]) for t in t_int]
        out = np.empty((len(t_dom) - 1, AREA_CALCULATION_NUM + 1), np.float64)
        for i in range(len(t_dom) - 1):
            ts = np.linspace(t_dom[i], t_dom[i + 1], num=AREA_CALCULATION_NUM + 1).tolist()
            out[i] = [f(t) for t in ts]
        return out

    def _interval_sends(self, total: int, intervals, domains, functions,
                        area_fn: Optional[Callable[[np.ndarray], np.ndarray]] = None):
        """deterministic part of the rate-curve analysis: per-second send counts of every interval"""
        t_list, area_list = [], []
        for iv, dom, fn in zip(intervals, domains, functions):
            if area_fn is None:
                t, a = self.interval_areas(iv, dom, fn)
            else:
                t = list(range(iv[0], iv[1]))
                a = [float(v) for v in area_fn(self.interval_samples(iv, dom, fn))]
            t_list.append(t)
            area_list.append(a)
        totals = [sum(a) for a in area_list]
        total_area = sum(totals)
        if total_area <= 0:
            return None
        amounts = [round(t / total_area * total) for t in totals]
        amounts[-1] = total - sum(amounts[:-1])
        sends = [self.diffuse([a / totals[i] * amounts[i] for a in area_list[i]]) if totals[i] > 0
                 else [0] * len(area_list[i]) for i in range(len(amounts))]
        return t_list, sends

    def _interval_finish(self, t_list, sends, drop: Dict[str, Any]) -> Schedule:
        """random part: expand the per-interval drop rule to per-second slots and draw the drop lists"""
        drop = copy.deepcopy(drop)
        if "drop_probability" in drop:
            dp = drop.get("drop_probability", [])
            drop["drop_probability"] = [dp[i] for i, el in enumerate(sends) for _ in el]
        elif "drop_amounts" in drop:
            da = drop.get("drop_amounts", [])
            new: List[int] = []
            for i, el in enumerate(sends):
                el_sum, d = sum(el), da[i]
                if d == 0:
                    new.extend([0] * len(el))
                elif d == el_sum:
                    new.extend(el)
                elif 0 < d < el_sum:
                    chosen = set(self.rng.sample(range(el_sum), d))
                    k = -1
                    for n in el:
                        cnt = 0
                        for _ in range(n):
                            k += 1
                            if k in chosen:
                                cnt += 1
                        new.append(cnt)
                # d > el_sum: the reference appends nothing (lists fall out of step and zip() truncates)
            drop["drop_amounts"] = new
        origin = [t for ts in t_list for t in ts]
        amount = [n for el in sends for n in el]
        timing = [origin[0]] + [origin[i] - origin[i - 1] for i in range(1, len(origin))]
        drops = self._generate_drop_simulation_list(amount, drop) if drop else [[] for _ in amount]
        return timing, amount, drops

    def _get_interval_params(self, total: int, intervals, domains, functions, drop: Dict[str, Any]) -> Schedule:
        det = self._interval_sends(total, intervals, domains, functions)
        if det is None:
            return _EMPTY
        return self._interval_finish(det[0], det[1], drop)

    # ---- per-task plan: parse + integrate once, draw drops per round --------------------------------------
    @staticmethod
    def is_round_invariant(strategy: Any) -> bool:
        """True for relative-time flow strategies: slots and amounts are the same in every round."""
        fd = _loads(strategy).get("flow_dispatch", {})
        st, si = fd.get("specific_timing", {}), fd.get("specific_interval", {})
        sec = st if st.get("use", False) else si
        return bool(fd.get("use_strategy", False)) and sec.get("time_type", "relative") == "relative"

    @staticmethod
    def flow_plan(strategy: Any, flow_id: str = "plan_x_0",
                  area_fn: Optional[Callable[[np.ndarray], np.ndarray]] = None) -> Dict[str, Any]:
        """Round-invariant part of a relative-time flow strategy ({"timing", "amount", ...}); ``area_fn`` integrates the
        rate-curve samples (the GPU engine passes the on-device integrator).  ``draw_drops`` adds the per-round part; the
        pair reproduces ``flow_strategy_analysis`` exactly, including the order in which the RNG is consumed."""
        fd = _loads(strategy).get("flow_dispatch", {})
        empty = {"kind": "empty", "timing": [], "amount": []}
        total = fd.get("total_dispatch_amount", 0)
        st, si = fd.get("specific_timing", {}), fd.get("specific_interval", {})
        use_t, use_i = bool(st.get("use", False)), bool(si.get("use", False))
        if not fd.get("use_strategy", False) or total <= 0 or use_t == use_i:
            return empty
        s = Strategy()
        if use_t:
            timing, amount = list(st.get("timings", [])), [int(a) for a in st.get("amounts", [])]
            drop = st.get("drop_simulation", {})
            if len(timing) != len(amount) or not timing or (drop and len(drop) != 1):
                return empty
            return {"kind": "timing", "timing": timing, "amount": amount, "drop": copy.deepcopy(drop)}
        intervals = si.get("intervals", [])
        rules = si.get("dispatch_rules", {})
        domains, functions = rules.get("domains", []), rules.get("functions", [])
        drop = si.get("drop_simulation", {})
        if len(intervals) != len(domains) or len(domains) != len(functions) or not intervals or (drop and len(drop) != 1):
            return empty
        det = s._interval_sends(total, intervals, domains, functions, area_fn)
        if det is None:
            return empty
        t_list, sends = det
        origin = [t for ts in t_list for t in ts]
        return {"kind": "interval", "t_list": t_list, "sends": sends, "drop": copy.deepcopy(drop),
                "amount": [n for el in sends for n in el],
                "timing": [origin[0]] + [origin[i] - origin[i - 1] for i in range(1, len(origin))]}

    @staticmethod
    def draw_drops(plan: Dict[str, Any], rng: Optional[random.Random] = None) -> List[List[int]]:
        s = Strategy(rng)
        if plan["kind"] == "timing":
            return s._generate_drop_simulation_list(plan["amount"], plan["drop"]) if plan["drop"] else [[] for _ in plan["amount"]]
        if plan["kind"] == "interval":
            return s._interval_finish(plan["t_list"], plan["sends"], plan["drop"])[2]
        return []


# --------------------------------------------------------------------------------------------------------
# GPU mapping helpers: a schedule -> per-client arrival ticks, so a DeviceFlow strategy becomes the
# `arrive_tick` / `max_tick` inputs of deviceflow_mask_kernel (message k of the shelf = k-th finishing client).
# --------------------------------------------------------------------------------------------------------
def schedule_to_ticks(amount: List[int], drops: List[List[int]], num_messages: int) -> np.ndarray:
    """tick[k] = slot index in which message k (arrival order) is forwarded, -1 if dropped or never sent
    (messages beyond sum(amount) are discarded by ``clean_remain_message``: dispatcher.py:244-252)."""
    ticks = np.full(num_messages, -1, np.int32)
    k = 0
    for slot, n in enumerate(amount):
        dl = set(drops[slot]) if slot < len(drops) else set()
        for i in range(n):
            if k >= num_messages:
                return ticks
            ticks[k] = -1 if i in dl else slot
            k += 1
    return ticks
