"""DeviceFlow data plane: bus semantics, sorter gating, real-time / flow dispatch, teardown, restart recovery.
(reference: deviceflow/grpc_service/deviceflow_server.py, non_grpc/{sorter,dispatcher,shelf_room,registry}.py)"""
import base64
import json
import os

import pytest

from olearning_sim_b200.control.repo import FlowTableRepo
from olearning_sim_b200.deviceflow.bus import BusTimeout, FakeClock, InMemoryBus
from olearning_sim_b200.deviceflow.rooms import (CallbackProducer, WebsocketProducer, make_message, parse_envelope,
                                                 ShelfNotExistException, ShelfRoom)
from olearning_sim_b200.deviceflow.service import DeviceFlowConfig, DeviceFlowService, NotifyRequest

OUT = {"pulsar_client": {"url": "mem://local", "topic": "agg_in"}}


def rt_strategy(sizes, p=0.0):
    return json.dumps({"real_time_dispatch": {"use_strategy": True, "dispatch_batch_sizes": sizes,
                                              "drop_simulation": {"drop_probability": p}},
                       "flow_dispatch": {"use_strategy": False}})


def flow_strategy(timings, amounts, drop=None):
    st = {"use": True, "time_type": "relative", "timings": timings, "amounts": amounts}
    if drop:
        st["drop_simulation"] = drop
    return json.dumps({"real_time_dispatch": {"use_strategy": False},
                       "flow_dispatch": {"use_strategy": True, "total_dispatch_amount": sum(amounts), "specific_timing": st}})


def make(seed=1, repo=None, bus=None, clock=None):
    bus = bus or InMemoryBus(virtual_time=True)
    svc = DeviceFlowService(bus=bus, repo=repo or FlowTableRepo(), clock=clock or FakeClock(), seed=seed,
                            dispatcher_timeout_millis=0)
    return svc, bus


def publish(svc, bus, flow_id, n, cr="logical_simulation", start=0):
    prod = bus.create_producer(svc.config.source_topic)
    for i in range(start, start + n):
        prod.send(make_message(flow_id, cr, f"update-{i}"))


def drain(consumer):
    out = []
    while True:
        try:
            m = consumer.receive(timeout_millis=0)
        except BusTimeout:
            return out
        consumer.acknowledge(m)
        out.append(m.value().decode())


# ---- bus ----------------------------------------------------------------------------------------------------
def test_bus_shared_subscription_ack_and_redelivery():
    bus = InMemoryBus(virtual_time=True)
    c = bus.subscribe("t", "s")
    p = bus.create_producer("t")
    for i in range(3):
        p.send(f"m{i}".encode())
    m0 = c.receive(0)
    c.acknowledge(m0)
    m1 = c.receive(0)             # not acked
    c.close()                     # un-acked message goes back to the subscription
    c2 = bus.subscribe("t", "s")
    assert [m.value() for m in (c2.receive(0), c2.receive(0))] == [b"m1", b"m2"]
    with pytest.raises(BusTimeout):
        c2.receive(0)


def test_bus_journal_survives_restart(tmp_path):
    j = str(tmp_path / "bus.jsonl")
    bus = InMemoryBus(journal_path=j, virtual_time=True)
    c = bus.subscribe("t", "durable")
    p = bus.create_producer("t")
    for i in range(4):
        p.send(f"m{i}".encode())
    c.acknowledge(c.receive(0))
    bus2 = InMemoryBus(journal_path=j, virtual_time=True)       # "process restart"
    c2 = bus2.subscribe("t", "durable")
    assert [c2.receive(0).value() for _ in range(3)] == [b"m1", b"m2", b"m3"]


def test_envelope_and_websocket_payload_format():
    env = parse_envelope(make_message("t_op_3", "device_simulation", "hello"))
    assert env == {"routing_key": "t_op_3", "compute_resource": "device_simulation", "message": "hello"}
    got = []
    WebsocketProducer.transports["ws://agg"] = got.append
    wp = WebsocketProducer()
    wp.get_producer({"url": "ws://agg"})
    wp.send(b"abc")
    assert json.loads(got[0]) == {"payload": base64.b64encode(b"abc").decode()}


def test_shelf_room_errors():
    room = ShelfRoom(InMemoryBus(virtual_time=True))
    with pytest.raises(ShelfNotExistException):
        room.put_on_shelf("nope", b"x")
    assert room.add_shelf("f") and room.has_shelf("f") and room.remove_shelf("f") and not room.has_shelf("f")


# ---- service ------------------------------------------------------------------------------------------------
def test_register_validates_compute_resources():
    svc, _ = make()
    assert not svc.RegisterTask("t", ["gpu_simulation"])
    assert svc.RegisterTask("t", ["logical_simulation"])
    assert svc.GetTotalComputeResources("t") == ["logical_simulation"]
    assert svc.GetDeviceflowPulsarClient()["topic"] == svc.config.source_topic


def test_notify_start_requires_registration_and_valid_strategy():
    svc, _ = make()
    req = NotifyRequest("t", "t_training_0", "logical_simulation", rt_strategy([2]), OUT)
    assert not svc.NotifyStart(req)                       # not registered
    svc.RegisterTask("t", ["logical_simulation"])
    assert not svc.NotifyStart(NotifyRequest("t", "t_training_0", "logical_simulation", "{bad json", OUT))
    assert not svc.NotifyStart(NotifyRequest("t", "t_training_0", "tpu", rt_strategy([2]), OUT))
    assert svc.NotifyStart(req)


def test_sorter_discards_before_start_and_unknown_resources():
    svc, bus = make()
    svc.RegisterTask("t", ["logical_simulation"])
    fid = "t_training_0"
    publish(svc, bus, fid, 2)                             # before NotifyStart -> discarded
    svc.pump()
    assert svc.stats()["sorter"] == {"shelved": 0, "discarded": 2}
    svc.NotifyStart(NotifyRequest("t", fid, "logical_simulation", rt_strategy([10]), OUT))
    publish(svc, bus, fid, 3)
    publish(svc, bus, fid, 1, cr="device_simulation")     # this side never called NotifyStart
    svc.pump()
    assert svc.stats()["sorter"] == {"shelved": 3, "discarded": 3}


def test_real_time_batches_cycle_and_flush_on_release():
    svc, bus = make()
    out = bus.subscribe("agg_in", "agg")
    svc.RegisterTask("t", ["logical_simulation"])
    fid = "t_training_0"
    svc.NotifyStart(NotifyRequest("t", fid, "logical_simulation", rt_strategy([2, 3]), OUT))
    publish(svc, bus, fid, 6)
    svc.pump()
    # batches of 2 then 3 are forwarded, 1 message stays buffered inside the dispatcher until release
    assert drain(out) == [f"update-{i}" for i in range(5)]
    assert not svc.CheckDeviceflowDispatchFinished("t")
    svc.NotifyComplete(NotifyRequest("t", fid, "logical_simulation"))
    # dispatcher already consumed msg 5 into its partial batch in the previous dispatch call; re-arm and flush
    svc._flow[fid]["to_dispatch"] = False
    svc.pump()
    assert svc.CheckDeviceflowDispatchFinished("t")
    assert fid not in svc.stats()["flows"]


def test_real_time_drop_probability_one_drops_everything():
    svc, bus = make()
    out = bus.subscribe("agg_in", "agg")
    svc.RegisterTask("t", ["logical_simulation"])
    fid = "t_training_1"
    svc.NotifyStart(NotifyRequest("t", fid, "logical_simulation", rt_strategy([2], p=1.0), OUT))
    publish(svc, bus, fid, 4)
    svc.NotifyComplete(NotifyRequest("t", fid, "logical_simulation"))
    svc.pump()
    assert drain(out) == []
    assert svc.CheckDeviceflowDispatchFinished("t")


def test_flow_mode_waits_for_all_sides_then_follows_schedule_and_discards_leftovers():
    clock = FakeClock()
    svc, bus = make(clock=clock)
    out = bus.subscribe("agg_in", "agg")
    svc.RegisterTask("t", ["logical_simulation", "device_simulation"])
    fid = "t_upload_0"
    strat = flow_strategy([1, 2, 5], [2, 1, 2], drop={"drop_amounts": [0, 0, 1]})
    assert svc.NotifyStart(NotifyRequest("t", fid, "logical_simulation", strat, OUT))
    assert svc.NotifyStart(NotifyRequest("t", fid, "device_simulation", strat, OUT))
    publish(svc, bus, fid, 4)
    publish(svc, bus, fid, 4, cr="device_simulation", start=4)
    svc.NotifyComplete(NotifyRequest("t", fid, "logical_simulation"))
    svc.pump()
    assert drain(out) == []                               # device side has not completed yet -> nothing dispatched
    svc.NotifyComplete(NotifyRequest("t", fid, "device_simulation"))
    svc.pump()
    got = drain(out)
    assert len(got) == 4                                  # 5 scheduled, 1 dropped, 3 leftovers discarded
    assert got[:3] == ["update-0", "update-1", "update-2"]
    assert clock.sleeps == [1, 2, 5]
    assert svc.CheckDeviceflowDispatchFinished("t")


def test_inconsistent_second_notify_start_tears_the_flow_down():
    svc, _ = make()
    svc.RegisterTask("t", ["logical_simulation", "device_simulation"])
    fid = "t_op_0"
    assert svc.NotifyStart(NotifyRequest("t", fid, "logical_simulation", rt_strategy([2]), OUT))
    assert not svc.NotifyStart(NotifyRequest("t", fid, "device_simulation", rt_strategy([3]), OUT))
    assert fid not in svc.stats()["flows"]


def test_callback_outbound_sink():
    got = []
    CallbackProducer.sinks["gpu_agg"] = got.append
    svc, bus = make()
    svc.RegisterTask("t", ["logical_simulation"])
    fid = "t_training_0"
    svc.NotifyStart(NotifyRequest("t", fid, "logical_simulation", rt_strategy([1]), {"callback": {"name": "gpu_agg"}}))
    publish(svc, bus, fid, 3)
    svc.pump()
    assert [g.decode() for g in got] == ["update-0", "update-1", "update-2"]


def test_unregister_removes_flows_and_rows():
    repo = FlowTableRepo()
    svc, bus = make(repo=repo)
    svc.RegisterTask("t", ["logical_simulation"])
    svc.NotifyStart(NotifyRequest("t", "t_a_0", "logical_simulation", rt_strategy([1]), OUT))
    assert repo.all_rows()
    assert svc.UnRegisterTask("t")
    assert repo.all_rows() == [] and svc.stats()["flows"] == []


def test_restart_recovery_rebuilds_registry_flows_and_dispatchers(tmp_path):
    db = str(tmp_path / "df.sqlite")
    journal = str(tmp_path / "bus.jsonl")
    bus = InMemoryBus(journal_path=journal, virtual_time=True)
    svc, _ = make(repo=FlowTableRepo(db), bus=bus)
    svc.RegisterTask("t", ["logical_simulation"])
    fid = "t_training_0"
    svc.NotifyStart(NotifyRequest("t", fid, "logical_simulation", rt_strategy([2]), OUT))
    publish(svc, bus, fid, 3)
    svc._sorter.drain()                                   # shelved but not dispatched, then the process "dies"
    svc.shutdown()
    bus2 = InMemoryBus(journal_path=journal, virtual_time=True)
    out = bus2.subscribe("agg_in", "agg")
    svc2, _ = make(repo=FlowTableRepo(db, shared_key="restart"), bus=bus2)
    assert svc2.stats()["registered"] == ["t"] and svc2.stats()["flows"] == [fid]
    svc2.NotifyComplete(NotifyRequest("t", fid, "logical_simulation"))
    svc2.pump()
    assert sorted(drain(out)) == ["update-0", "update-1", "update-2"]
    assert svc2.CheckDeviceflowDispatchFinished("t")


def test_websocket_producer_against_a_live_websocket_server():
    """the WebSocket outbound path end to end: a real ``websockets`` server on localhost receives what the dispatcher's
    producer sends, in the Pulsar-WebSocket envelope {"payload": base64(message)}"""
    import base64
    import json as _json
    import threading
    from websockets.sync.server import serve

    got, done = [], threading.Event()

    def handler(ws):
        for msg in ws:
            got.append(_json.loads(msg))
            if len(got) == 3:
                done.set()

    with serve(handler, "127.0.0.1", 0) as server:
        t = threading.Thread(target=server.serve_forever, daemon=True)
        t.start()
        port = server.socket.getsockname()[1]
        wp = WebsocketProducer()
        wp.get_producer({"url": f"ws://127.0.0.1:{port}/ws/v2/producer/persistent/public/default/out"})
        for i in range(3):
            wp.send(make_message(f"t1_train_0_dev{i}", "logical_simulation", f"update {i}"))
        assert done.wait(5.0)
        wp.close()
        server.shutdown()
    bodies = [base64.b64decode(g["payload"]) for g in got]
    assert all(b"dev" in b for b in bodies) and len({bytes(b) for b in bodies}) == 3
