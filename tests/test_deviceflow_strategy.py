"""DeviceFlow strategy semantics — golden vectors from the formulas in SURVEY §5.9
(reference: deviceflow/non_grpc/strategy.py, deviceflow/utils/validate_parameters.py)."""
import json
import math
import random
from datetime import datetime, timedelta

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from olearning_sim_b200.deviceflow import safe_expr
from olearning_sim_b200.deviceflow.strategy import Strategy, schedule_to_ticks
from olearning_sim_b200.deviceflow.validate import ValidateParameters


def flow(section, total, **kw):
    return {"real_time_dispatch": {"use_strategy": False},
            "flow_dispatch": {"use_strategy": True, "total_dispatch_amount": total, **{section: {"use": True, **kw}}}}


# ---- safe expressions ---------------------------------------------------------------------------------
def test_safe_expr_accepts_reference_style_functions():
    assert safe_expr.evaluate("math.sin(t)+1", 0.0) == 1.0
    assert safe_expr.evaluate("np.abs(t) * 2 + math.pi", -1.0) == pytest.approx(2 + math.pi)
    assert safe_expr.evaluate("1 if t < 3 else 5", 4) == 5
    assert safe_expr.evaluate("max(0, 10 - t**2)", 2) == 6


@pytest.mark.parametrize("bad", ["__import__('os').system('id')", "open('/etc/passwd').read()", "t.__class__",
                                 "(lambda: 1)()", "[x for x in (1,2)]", "math.sin", "os.getcwd()", "eval('1')",
                                 "t; import os", "'a'*3", "9**9**9"])
def test_safe_expr_rejects_everything_else(bad):
    with pytest.raises(Exception):
        safe_expr.evaluate(bad, 1.0)


# ---- real time ------------------------------------------------------------------------------------------
def test_real_time_analysis():
    s = {"real_time_dispatch": {"use_strategy": True, "dispatch_batch_sizes": [5, 7], "drop_simulation": {"drop_probability": 0.25}}}
    assert Strategy.check_real_time_dispatch(json.dumps(s))
    assert Strategy.real_time_strategy_analysis(json.dumps(s)) == ([5, 7], 0.25)
    assert not Strategy.check_real_time_dispatch(json.dumps({"flow_dispatch": {"use_strategy": True}}))


# ---- specific_timing ------------------------------------------------------------------------------------
def test_relative_timing_passthrough_and_drop_amounts():
    s = flow("specific_timing", 10, time_type="relative", timings=[0, 2, 5], amounts=[3, 3, 4],
             drop_simulation={"drop_amounts": [0, 1, 4]})
    t, a, d = Strategy.flow_strategy_analysis(json.dumps(s), "task_op_0", rng=random.Random(1))
    assert t == [0, 2, 5] and a == [3, 3, 4]
    assert d[0] == [] and len(d[1]) == 1 and 0 <= d[1][0] < 3 and d[2] == [0, 1, 2, 3]


def test_drop_probability_extremes():
    s = flow("specific_timing", 6, time_type="relative", timings=[0, 1], amounts=[3, 3],
             drop_simulation={"drop_probability": [0, 1]})
    _, _, d = Strategy.flow_strategy_analysis(s, "x_y_0")
    assert d == [[], [0, 1, 2]]


def test_absolute_timing_sorted_past_slots_discarded_and_round_indexed():
    now = datetime(2030, 1, 1, 12, 0, 0, 500000)
    fmt = "%Y-%m-%d %H:%M:%S"
    base = now.replace(microsecond=0)
    r0 = [(base + timedelta(seconds=x)).strftime(fmt) for x in (30, -10, 10)]     # one slot already in the past
    r1 = [(base + timedelta(seconds=x)).strftime(fmt) for x in (5, 6, 7)]
    s = flow("specific_timing", 6, time_type="absolute", time_zone="UTC", timings=[r0, r1], amounts=[1, 2, 3])
    import datetime as _dt
    aware_now = lambda: now.replace(tzinfo=_dt.timezone.utc)
    t, a, _ = Strategy.flow_strategy_analysis(s, "task_op_0", now=aware_now)
    assert a == [3, 1]                      # sorted by time: (-10 -> dropped), 10 -> amount 3, 30 -> amount 1
    assert t == [pytest.approx(10 - 0.5), 20]
    t1, a1, _ = Strategy.flow_strategy_analysis(s, "task_op_1", now=aware_now)
    assert a1 == [1, 2, 3] and t1[1:] == [1, 1]
    assert Strategy.flow_strategy_analysis(s, "task_op_2", now=aware_now) == ([], [], [])   # no list for round 2


# ---- specific_interval ------------------------------------------------------------------------------------
def test_constant_rate_interval_distributes_evenly():
    s = flow("specific_interval", 100, time_type="relative", intervals=[[0, 10]],
             dispatch_rules={"domains": [[0, 1]], "functions": ["1"]})
    t, a, d = Strategy.flow_strategy_analysis(s, "t_o_0")
    assert a == [10] * 10 and t == [0] + [1] * 9 and d == [[]] * 10


def test_error_diffusion_rounding_conserves_total():
    assert Strategy.diffuse([0.4, 0.4, 0.4, 0.4, 0.4]) == [0, 1, 0, 1, 0]
    assert sum(Strategy.diffuse([2.5] * 4)) == 10


def test_sine_curve_matches_reference_formula():
    total = 1000
    s = flow("specific_interval", total, time_type="relative", intervals=[[0, 10], [12, 16]],
             dispatch_rules={"domains": [[0, 6.28], [0, 1]], "functions": ["math.sin(t)+1", "2*t"]})
    t, a, _ = Strategy.flow_strategy_analysis(s, "t_o_0")
    # independent re-computation of the reference algorithm
    def areas(iv, dom, f):
        out = []
        n = iv[1] - iv[0]
        for i in range(n):
            ts = np.linspace(dom[0] + (dom[1] - dom[0]) / n * i, dom[0] + (dom[1] - dom[0]) / n * (i + 1), 101)
            ys = [f(x) for x in ts]
            out.append(sum(max(0.0, 0.5 * (ys[j] + ys[j - 1]) * 0.01) for j in range(1, 101)))
        return out
    a0, a1 = areas([0, 10], [0, 6.28], lambda x: math.sin(x) + 1), areas([12, 16], [0, 1], lambda x: 2 * x)
    tot = sum(a0) + sum(a1)
    n0 = round(sum(a0) / tot * total)
    assert sum(a[:10]) == n0 and sum(a[10:]) == total - n0 and sum(a) == total
    assert t == [0] + [1] * 9 + [3] + [1] * 3          # gap between the intervals becomes the delta of the first slot
    assert a[:10] == Strategy.diffuse([x / sum(a0) * n0 for x in a0])


def test_negative_lobes_contribute_no_area():
    t, areas = Strategy.interval_areas([0, 4], [0, 2 * math.pi], "math.sin(t)")
    assert areas[0] > 0 and areas[1] > 0 and areas[2] == pytest.approx(0, abs=1e-3) and areas[3] == pytest.approx(0, abs=1e-3)


def test_interval_drop_amounts_bucketed_per_slot():
    s = flow("specific_interval", 40, time_type="relative", intervals=[[0, 4]],
             dispatch_rules={"domains": [[0, 1]], "functions": ["1"]}, drop_simulation={"drop_amounts": [7]})
    _, a, d = Strategy.flow_strategy_analysis(s, "t_o_0", rng=random.Random(5))
    assert a == [10, 10, 10, 10] and sum(len(x) for x in d) == 7
    assert all(0 <= i < 10 for slot in d for i in slot)


def test_exactly_one_of_timing_or_interval():
    s = {"flow_dispatch": {"use_strategy": True, "total_dispatch_amount": 5,
                           "specific_timing": {"use": True, "timings": [0], "amounts": [5]},
                           "specific_interval": {"use": True}}}
    assert Strategy.flow_strategy_analysis(s, "a_b_0") == ([], [], [])


@settings(max_examples=40, deadline=None)
@given(total=st.integers(1, 5000), length=st.integers(1, 30), amp=st.floats(0.1, 5), phase=st.floats(0, 6))
def test_interval_total_is_conserved(total, length, amp, phase):
    s = flow("specific_interval", total, time_type="relative", intervals=[[0, length]],
             dispatch_rules={"domains": [[0, 6.28]], "functions": [f"{amp}*math.sin(t+{phase})+{amp}+0.01"]})
    t, a, d = Strategy.flow_strategy_analysis(s, "t_o_0")
    assert len(t) == len(a) == len(d) == length
    assert abs(sum(a) - total) <= 1 and all(x >= 0 for x in a)   # diffusion can strand < 1 message of residual


def test_schedule_to_ticks():
    ticks = schedule_to_ticks([2, 3], [[1], []], 7)
    assert ticks.tolist() == [0, -1, 1, 1, 1, -1, -1]          # dropped index, then leftovers discarded


# ---- validation --------------------------------------------------------------------------------------------
V = ValidateParameters()


def test_validate_requires_exactly_one_strategy():
    assert V.check_strategy({"real_time_dispatch": {"use_strategy": True}, "flow_dispatch": {"use_strategy": True}}) == (False, "Must use one strategy")
    assert V.check_strategy({})[0] is False
    assert V.check_strategy({"real_time_dispatch": {"use_strategy": True, "drop_simulation": {"drop_probability": 1.5}}})[0] is False
    assert V.check_strategy({"real_time_dispatch": {"use_strategy": True, "dispatch_batch_sizes": [4]}})[0] is True


@pytest.mark.parametrize("kw,ok", [
    (dict(time_type="relative", timings=[0, 1], amounts=[2, 3]), True),
    (dict(time_type="relative", timings=[0, 1], amounts=[2, 2]), False),            # sum != total
    (dict(time_type="relative", timings=[0, -1], amounts=[2, 3]), False),
    (dict(time_type="relative", timings=[0], amounts=[2, 3]), False),
    (dict(time_type="sometimes", timings=[0, 1], amounts=[2, 3]), False),
    (dict(time_type="absolute", time_zone="Mars/Olympus", timings=[["2030-01-01 00:00:00", "2030-01-01 00:00:01"]], amounts=[2, 3]), False),
    (dict(time_type="absolute", time_zone="Asia/Shanghai", timings=[["2030-01-01 00:00:00", "2030-01-01 00:00:01"]], amounts=[2, 3]), True),
    (dict(time_type="absolute", time_zone="UTC", timings=[["01/01/2030", "x"]], amounts=[2, 3]), False),
    (dict(time_type="relative", timings=[0, 1], amounts=[2, 3], drop_simulation={"drop_probability": [0.1, 0.2], "drop_amounts": [1, 1]}), False),
    (dict(time_type="relative", timings=[0, 1], amounts=[2, 3], drop_simulation={"drop_amounts": [3, 3]}), False),
    (dict(time_type="relative", timings=[0, 1], amounts=[2, 3], drop_simulation={"drop_probability": [0.1]}), False),
])
def test_validate_timing(kw, ok):
    res, msg = V.check_strategy(flow("specific_timing", 5, **kw))
    assert res is ok, msg


@pytest.mark.parametrize("kw,ok", [
    (dict(time_type="relative", intervals=[[0, 5], [5, 9]], dispatch_rules={"domains": [[0, 1], [0, 2]], "functions": ["t", "math.sin(t)+1"]}), True),
    (dict(time_type="relative", intervals=[[1, 1], [2, 3]], dispatch_rules={"domains": [[0, 1], [0, 2]], "functions": ["t", "t"]}), False),
    (dict(time_type="relative", intervals=[[0, 5]], dispatch_rules={"domains": [[1, 0]], "functions": ["t"]}), False),
    (dict(time_type="relative", intervals=[[0, 5]], dispatch_rules={"domains": [[0, 1]], "functions": ["__import__('os')"]}), False),
    (dict(time_type="relative", intervals=[[0, 5]], dispatch_rules={"domains": [[0, 1]], "functions": ["x+1"]}), False),
    (dict(time_type="relative", intervals=[[0, 5]], dispatch_rules={"domains": [[0, 1], [0, 1]], "functions": ["t"]}), False),
])
def test_validate_interval(kw, ok):
    res, msg = V.check_strategy(flow("specific_interval", 5, **kw))
    assert res is ok, msg
