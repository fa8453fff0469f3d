"""DeviceFlow strategies and availability traces driving the engine's per-round masks (CPU backend; the native and
hybrid backends draw the same masks from the CUDA twin of the Philox generator — tests/test_gpu_kernels.py)."""
import json
import math

import torch

from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine

BASE = dict(model="mlp2", num_clients=40, local_steps=1, batch=8, shard=8, lr=0.1, device="cpu", backend="torch", seed=3)


def _engine(**kw):
    e = FedEngine(FedConfig(**{**BASE, **kw}))
    e.load_synthetic()
    return e


def test_online_trace_scales_participation_per_round():
    """diurnal trace: p_online(t) = 0.9 * (0.5 + 0.5*cos(t*pi/4)) -> many clients at t=0, (almost) none at t=4"""
    e = _engine(p_online=0.9, online_trace="0.5+0.5*math.cos(t*math.pi/4)")
    parts = [e.run_round(r).participants for r in range(9)]
    assert parts[0] >= 30 and parts[4] == 0 and parts[8] >= 30
    assert parts[2] < parts[0] and parts[6] < parts[8]
    # deterministic: a second engine with the same seed sees the same trace
    e2 = _engine(p_online=0.9, online_trace="0.5+0.5*math.cos(t*math.pi/4)")
    assert [e2.run_round(r).participants for r in range(9)] == parts


def test_real_time_strategy_drop_probability_one_keeps_the_model():
    st = {"real_time_dispatch": {"use_strategy": True, "dispatch_batch_sizes": [4], "drop_simulation": {"drop_probability": 1.0}},
          "flow_dispatch": {"use_strategy": False}}
    e = _engine(deviceflow_strategy=json.dumps(st))
    g0 = e.global_model().clone()
    r = e.run_round(0)
    assert r.participants == 40 and r.dropped == 40 and r.arrived == 0
    assert torch.equal(e.global_model(), g0)                     # nothing reached the aggregator: model unchanged


def test_flow_strategy_quota_and_drop_list_decide_arrivals():
    """specific timing: 3 slots forwarding 10 + 10 + 5 messages, the first two of slot 1 dropped; the 40 - 25 clients
    beyond the total are discarded as leftovers (late)"""
    st = {"real_time_dispatch": {"use_strategy": False},
          "flow_dispatch": {"use_strategy": True, "total_dispatch_amount": 25,
                            "specific_timing": {"use": True, "time_type": "relative", "timings": [0, 1, 2], "amounts": [10, 10, 5],
                                                "drop_simulation": {"drop_amounts": [0, 2, 0]}}}}
    e = _engine(deviceflow_strategy=json.dumps(st))
    r = e.run_round(0)
    assert r.participants == 40
    assert r.arrived == 23 and r.late == 17 and r.dropped == 0          # 2 dropped in flight count with the undelivered
    assert r.arrived + r.late + r.dropped == r.participants


def test_trace_and_strategy_compose_and_loss_stays_finite():
    st = {"real_time_dispatch": {"use_strategy": True, "dispatch_batch_sizes": [8], "drop_simulation": {"drop_probability": 0.25}},
          "flow_dispatch": {"use_strategy": False}}
    e = _engine(p_online=0.8, online_trace="0.6+0.4*math.sin(t/2)", deviceflow_strategy=json.dumps(st), hetero_steps=True,
                local_steps=3, shard=24)
    for r in range(4):
        res = e.run_round(r)
        assert res.arrived + res.dropped + res.late == res.participants <= 40
        assert math.isfinite(res.loss) and res.client_steps <= res.participants * 3
