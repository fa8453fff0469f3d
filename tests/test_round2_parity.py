"""Round-2 parity items: protobuf wire contract, dashboard tables, hybrid splitter over object stores, script-actor data /
model fetch, object-store adapters against fake clients, cached flow plans, per-class accounting, subprocess executor."""
import io
import json
import os
import random
import sys
import time
import zipfile

import numpy as np
import pytest

from olearning_sim_b200.api import wire
from olearning_sim_b200.api.schema import TaskStatus
from tests.helpers import make_task_json


# ---------------------------------------------------------------------------------------------------------
# protobuf wire contract (ols_core/proto/*.proto)
# ---------------------------------------------------------------------------------------------------------
def _varint_fields(b: bytes):
    """minimal wire-format walk: (field number, wire type) of the top-level fields"""
    out, i = [], 0
    while i < len(b):
        key, shift = 0, 0
        while True:
            c = b[i]; i += 1
            key |= (c & 0x7F) << shift; shift += 7
            if not c & 0x80:
                break
        fn, wt = key >> 3, key & 7
        out.append((fn, wt))
        if wt == 0:
            while b[i] & 0x80:
                i += 1
            i += 1
        elif wt == 2:
            ln, shift = 0, 0
            while True:
                c = b[i]; i += 1
                ln |= (c & 0x7F) << shift; shift += 7
                if not c & 0x80:
                    break
            i += ln
        elif wt == 5:
            i += 4
        elif wt == 1:
            i += 8
    return out


def test_taskconfig_uses_the_reference_field_numbers_and_round_trips():
    strategy = {"real_time_dispatch": {"use_strategy": True, "dispatch_batch_sizes": [3], "drop_simulation": {"drop_probability": 0.1}},
                "flow_dispatch": {"use_strategy": False}}
    tj = make_task_json(task_id="w1", nums=(4, 2), devices=("High", "Low"), dynamic=(1, 0), use_df=True, strategy=strategy,
                        num_request=(2, 1))
    m = wire.task_json_to_pb(tj)
    b = m.SerializeToString()
    # taskService.proto:4-11 — userID=1 string, taskID=2 msg, target=3, operatorFlow=4, logicalSimulation=5, deviceSimulation=6
    assert _varint_fields(b) == [(1, 2), (2, 2), (3, 2), (4, 2), (5, 2), (6, 2)]
    assert m.taskID.taskID == "w1" and m.target.targetData[0].totalSimulation.numTotalSimulation == [4, 2]
    op = m.operatorFlow.operator[0]
    assert op.operationBehaviorController.useController and json.loads(op.operationBehaviorController.strategyBehaviorController) == strategy
    back = wire.task_pb_to_json(wire.msg("TaskConfig").FromString(b))
    want = {k: v for k, v in tj.items() if k != "engine"}
    assert json.loads(json.dumps(back)) == json.loads(json.dumps(want))
    # enums travel as numbers on the wire and names in JSON
    assert wire.msg("TaskStatus")(taskStatus=7).SerializeToString() == b"\x08\x07"
    assert wire.to_dict(wire.msg("TaskStatus").FromString(b"\x08\x07")) == {"taskStatus": "QUEUED"}
    # deviceflow.proto lives in package `deviceflow`
    assert wire.msg("deviceflow.NotifyRequest").DESCRIPTOR.full_name == "deviceflow.NotifyRequest"


def test_grpc_server_speaks_protobuf_to_reference_clients():
    from olearning_sim_b200.api.grpc_service import JsonStub, ProtoStub, SimulatorSession, connect
    from olearning_sim_b200.control.jobs import InProcessJobExecutor
    from olearning_sim_b200.control.repo import FlowTableRepo, ResTableRepo, TaskTableRepo
    from olearning_sim_b200.control.resource_manager import FakePhoneMgr, ResourceManager
    from olearning_sim_b200.control.task_manager import TaskManager
    from olearning_sim_b200.deviceflow.bus import FakeClock, InMemoryBus
    from olearning_sim_b200.deviceflow.service import DeviceFlowService
    from tests.test_control_plane import StubRunner
    StubRunner.gate = None
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 4, "mem": 8, "gpu": 0}, phone_mgr=FakePhoneMgr({"u1": {"High": 2}}))
    df = DeviceFlowService(bus=InMemoryBus(virtual_time=True), repo=FlowTableRepo(), clock=FakeClock())
    tm = TaskManager(repo=TaskTableRepo(), resource_manager=rm, deviceflow=df, executor=InProcessJobExecutor(runner_factory=StubRunner))
    sess = SimulatorSession(port=0, task_manager=tm, resource_manager=rm, deviceflow=df).start()
    try:
        ch = connect(sess.address)
        t, r, d = ProtoStub(ch, "TaskMgr"), ProtoStub(ch, "ResourceMgr"), ProtoStub(ch, "TaskOperatorOrientedDeviceFlow")
        ok = t.submitTask(wire.task_json_to_pb(make_task_json(task_id="p1")))
        assert ok.is_success is True
        assert [x.taskID for x in t.getTaskQueue().tasks] == ["p1"]                       # Empty request, protobuf reply
        assert wire.to_dict(t.getTaskStatus(wire.msg("TaskID")(taskID="p1"))) == {"taskStatus": "QUEUED"}
        assert t.getTaskStatus(wire.msg("TaskID")(taskID="nope")).taskStatus == 5            # MISSING
        res = r.getResource()
        assert res.clusterRes.cores == 4 and res.phoneRes.userDeviceAvailableResource[0].deviceResourceInfo[0].phoneType == "High"
        st = r.requestClusterResource(wire.msg("ServerClusterReq")(taskId="x", cores=9, mem=1))
        assert wire.to_dict(st) == {"status": "NOT_ENOUGH_RES"}
        rq = wire.from_dict("ResReq", {"type": 0, "clusterReq": {"taskId": "x", "cores": 1, "mem": 1},
                                       "phoneReq": {"taskID": "x", "userID": "u1", "deviceResourceInfo": [{"phoneType": "High", "num": 1}]}})
        got = r.requestResource(rq)
        assert got.clusterStatus.status == 0 and got.phoneStatus.isSuccess
        assert r.releaseResource(wire.msg("ResRelease")(type=0, taskId="x")).clusterStatus.status == 0
        # DeviceFlow under its package-qualified path
        assert d.RegisterTask(wire.from_dict("deviceflow.RegisterRequest", {"task_id": "p1", "total_compute_resources": ["logical_simulation"]})).is_success
        assert list(d.GetTotalComputeResources(wire.msg("deviceflow.TotalComputeResourcesRequest")(task_id="p1")).total_compute_resources) == ["logical_simulation"]
        nr = wire.from_dict("deviceflow.NotifyRequest", {"task_id": "p1", "routing_key": "p1_training_0", "compute_resource": "logical_simulation",
                                                         "strategy": json.dumps({"real_time_dispatch": {"use_strategy": True, "dispatch_batch_sizes": [1],
                                                                                                        "drop_simulation": {"drop_probability": 0.0}},
                                                                                 "flow_dispatch": {"use_strategy": False}}),
                                                         "outbound_service": {"pulsar_client": {"url": "mem://local", "topic": "agg"}}})
        assert d.NotifyStart(nr).is_success
        assert d.GetDeviceflowPulsarClient().topic
        # the JSON codec still works on the same server
        assert JsonStub(ch, "TaskMgr").getTaskQueue() == {"tasks": ["p1"]}
    finally:
        sess.stop()


# ---------------------------------------------------------------------------------------------------------
# dashboard tables (dispatcher.py:254-395, task_runner.py:226-256)
# ---------------------------------------------------------------------------------------------------------
def test_dashboard_curves_follow_the_reference_rules():
    from olearning_sim_b200.deviceflow import dashboard as db
    timing, amount = [3, 2, 4], [1000, 2000, 300]
    assert db.operation_rows(timing, amount, "relative") == [(0, 0), (3, 1000), (5, 2000), (9, 300)]
    assert db.operation_rows(timing, amount, "absolute") == [(0, 1000), (2, 2000), (6, 300)]
    acc = dict(db.accumulated_rows(timing, amount, "relative"))
    assert acc[0] == 0 and acc[1] == 0 and acc[2] == 0          # silence before the first slot
    assert acc[3] == 750 and acc[4] == 1000                      # a 1000-message burst ramps at 750 / s
    assert acc[5] == 1750 and acc[6] == 2500 and acc[7] == 3000 and acc[8] == 3000
    assert acc[9] == 3300 and max(acc) == 10                     # tail of max(2, (9-5)//2) seconds after the last slot
    st = {"flow_dispatch": {"use_strategy": True, "total_dispatch_amount": 3300,
                            "specific_timing": {"use": True, "time_type": "relative", "timings": timing, "amounts": amount}}}
    repo = db.DashboardRepo()
    assert repo.write_dispatch_curves("t", timing, amount, json.dumps(st))
    assert repo.curve("t", "operation")[1] == (3.0, 1000) and repo.curve("t")[-1][1] == 3300
    repo.write_dispatch_curves("t", [1], [5], json.dumps(st))                           # replace, not append
    assert len(repo.curve("t", "operation")) == 2
    assert not repo.write_dispatch_curves("t", [1], [5], json.dumps({"real_time_dispatch": {"use_strategy": True}}))
    tj = make_task_json(task_id="tt", nums=(4, 2), devices=("High", "Low"), dynamic=(0, 0))
    assert db.task_target_row(tj) == {"high": 4, "low": 2}
    repo.write_task_target(tj)
    assert repo.target.get_item_value("task_id", "tt", "high") in (4, "4")


def test_dispatcher_and_task_runner_write_the_dashboard():
    from olearning_sim_b200.control.jobs import InProcessJobExecutor
    from olearning_sim_b200.control.repo import FlowTableRepo, ResTableRepo, TaskTableRepo
    from olearning_sim_b200.control.resource_manager import FakePhoneMgr, ResourceManager
    from olearning_sim_b200.control.task_manager import TaskManager
    from olearning_sim_b200.deviceflow.bus import FakeClock, InMemoryBus
    from olearning_sim_b200.deviceflow.dashboard import DashboardRepo
    from olearning_sim_b200.deviceflow.service import DeviceFlowService
    dash = DashboardRepo()
    st = {"real_time_dispatch": {"use_strategy": False},
          "flow_dispatch": {"use_strategy": True, "total_dispatch_amount": 8,
                            "specific_timing": {"use": True, "time_type": "relative", "timings": [1, 2], "amounts": [5, 3]}}}
    df = DeviceFlowService(bus=InMemoryBus(virtual_time=True), repo=FlowTableRepo(), clock=FakeClock(), seed=0,
                           dispatcher_timeout_millis=0, dashboard=dash)
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 8, "mem": 32, "gpu": 0}, phone_mgr=FakePhoneMgr())
    tm = TaskManager(repo=TaskTableRepo(), resource_manager=rm, deviceflow=df, executor=InProcessJobExecutor())
    tj = make_task_json(task_id="dsh", rounds=1, use_df=True, strategy=st, nums=(8,),
                        operator_params={"model": "mlp2", "local_steps": 1, "batch": 8, "shard": 8, "lr": 0.1, "device": "cpu"})
    assert tm.submitTask(tj)
    tm.schedule_once(inline=True)
    assert tm.wait_for("dsh", 120) == TaskStatus.SUCCEEDED
    assert dash.target.get_item_value("task_id", "dsh", "high") in (8, "8")             # task_target row
    # the GPU/CPU engine applied the 5 + 3 schedule itself and reported it: curves are on the dashboard, flow is closed
    assert dash.curve("dsh", "operation") == [(0.0, 0), (1.0, 5), (3.0, 3)]
    rep = df.stats()["device_dispatched"]["dsh_training_0"]
    assert rep["participants"] == 8 and rep["forwarded"] == 8 and rep["late"] == 0
    assert df.CheckDeviceflowDispatchFinished("dsh")


# ---------------------------------------------------------------------------------------------------------
# object stores against fake clients; hybrid splitter through them
# ---------------------------------------------------------------------------------------------------------
class FakeBoto:
    """the subset of the boto3 S3 client the adapter uses"""

    def __init__(self):
        self.objects = {}

    def upload_file(self, local, bucket, key):
        self.objects[(bucket, key)] = open(local, "rb").read()

    def download_file(self, bucket, key, local):
        if (bucket, key) not in self.objects:
            raise KeyError(key)
        with open(local, "wb") as f:
            f.write(self.objects[(bucket, key)])

    def delete_object(self, Bucket, Key):
        self.objects.pop((Bucket, Key))

    def get_object(self, Bucket, Key):
        return {"Body": io.BytesIO(self.objects[(Bucket, Key)])}

    def list_objects_v2(self, Bucket, Prefix, ContinuationToken=None):
        keys = sorted(k for (b, k) in self.objects if b == Bucket and k.startswith(Prefix))
        start = int(ContinuationToken or 0)
        page = keys[start:start + 2]
        more = start + 2 < len(keys)
        return {"Contents": [{"Key": k} for k in page], "IsTruncated": more, **({"NextContinuationToken": str(start + 2)} if more else {})}


class FakeMinio:
    class _Obj:
        def __init__(self, n):
            self.object_name = n

    def __init__(self):
        self.objects = {}

    def fput_object(self, bucket, name, local):
        self.objects[(bucket, name)] = open(local, "rb").read()

    def fget_object(self, bucket, name, local):
        with open(local, "wb") as f:
            f.write(self.objects[(bucket, name)])

    def remove_object(self, bucket, name):
        del self.objects[(bucket, name)]

    def list_objects(self, bucket, prefix="", recursive=False):
        return [self._Obj(n) for (b, n) in sorted(self.objects) if b == bucket and n.startswith(prefix)]


@pytest.mark.parametrize("kind", ["S3", "MINIO"])
def test_object_store_adapters_against_fake_clients(tmp_path, kind):
    from olearning_sim_b200.utils.file_repo import get_file_repo
    client = FakeBoto() if kind == "S3" else FakeMinio()
    repo = get_file_repo(kind, {"bucket": "bk", "client": client})
    src = tmp_path / "a.bin"
    src.write_bytes(b"hello")
    for name in ("d/a.bin", "d/b.bin", "d/c.bin", "e/z.bin"):
        assert repo.upload_file(str(src), name)
    assert repo.listdir_object_name("d/") == ["d/a.bin", "d/b.bin", "d/c.bin"]           # S3: paginated listing
    assert repo.download_file("d/b.bin", str(tmp_path / "out" / "b.bin")) and (tmp_path / "out" / "b.bin").read_bytes() == b"hello"
    assert repo.download_payload("d/a.bin") == b"hello"
    assert not repo.download_file("missing", str(tmp_path / "m.bin"))
    assert repo.delete_file("d/a.bin") and not repo.delete_file("d/a.bin")
    assert repo.download_if_missing("d/c.bin", str(tmp_path / "out" / "b.bin"))          # local copy short-circuits


def _make_zip(path, files):
    with zipfile.ZipFile(path, "w") as z:
        for name, data in files.items():
            z.writestr(name, data)


def test_hybrid_splitter_ratio_split_through_an_object_store(tmp_path):
    """unsplit data (one shared folder, CSV contract): rows are split by device_ratio, both halves re-zipped and uploaded"""
    from olearning_sim_b200.control.task_runner import HybridDataSplitter
    from olearning_sim_b200.utils.file_repo import S3FileRepo
    client = FakeBoto()
    repo = S3FileRepo(bucket="bk", client=client)
    feats = "f0,f1\n" + "\n".join(f"{i},{i * 2}" for i in range(40)) + "\n"
    labs = "y\n" + "\n".join(str(i % 3) for i in range(40)) + "\n"
    z = tmp_path / "d.zip"
    _make_zip(z, {"dataconfig.json": "{\"k\": 1}", "shared/train_features.csv": feats, "shared/train_labels.csv": labs,
                  "shared/test_features.csv": "f0,f1\n1,2\n", "shared/test_labels.csv": "y\n0\n"})
    repo.upload_file(str(z), "in/d.zip")
    data = {"name": "d", "data_path": "in/d.zip", "data_split_type": False, "data_transfer_type": "S3",
            "data_path_logical": "out/d_logical.zip", "data_path_device": "out/d_device.zip",
            "allocation": {"logical_simulation": [6], "device_simulation": [2], "running_response": {"devices": [], "nums": []}}}
    assert HybridDataSplitter("t", file_repo=repo).split(data)
    rows = {}
    for side in ("logical", "device"):
        zz = zipfile.ZipFile(io.BytesIO(client.objects[("bk", f"out/d_{side}.zip")]))
        names = set(zz.namelist())
        assert {"dataconfig.json", "shared/train_features.csv", "shared/train_labels.csv", "shared/test_features.csv"} <= names
        f = zz.read("shared/train_features.csv").decode().splitlines()
        l = zz.read("shared/train_labels.csv").decode().splitlines()
        assert f[0] == "f0,f1" and l[0] == "y" and len(f) == len(l)
        assert all(int(a.split(",")[0]) % 3 == int(b) for a, b in zip(f[1:], l[1:]))       # features and labels stay paired
        rows[side] = {int(a.split(",")[0]) for a in f[1:]}
        assert zz.read("shared/test_features.csv").decode() == "f0,f1\n1,2\n"            # test files copied to both sides
    assert len(rows["device"]) == 10 and len(rows["logical"]) == 30 and not (rows["device"] & rows["logical"])
    assert rows["device"] | rows["logical"] == set(range(40))


def test_hybrid_splitter_presplit_folders_and_array_shards(tmp_path):
    from olearning_sim_b200.control.task_runner import HybridDataSplitter
    z = tmp_path / "p.zip"
    _make_zip(z, {f"c{i:02d}/x.txt": str(i) for i in range(5)})
    data = {"name": "p", "data_path": str(z), "data_split_type": True, "data_transfer_type": "FILE",
            "data_path_logical": str(tmp_path / "pl.zip"), "data_path_device": str(tmp_path / "pd.zip"),
            "allocation": {"logical_simulation": [3], "device_simulation": [2]}}
    assert HybridDataSplitter("t").split(data)
    assert {n.split("/")[0] for n in zipfile.ZipFile(tmp_path / "pl.zip").namelist()} == {"c00", "c01", "c02"}
    assert {n.split("/")[0] for n in zipfile.ZipFile(tmp_path / "pd.zip").namelist()} == {"c03", "c04"}
    # built-in operators' array format
    import io as _io
    buf = _io.BytesIO()
    np.savez(buf, x=np.arange(20).reshape(20, 1), y=np.arange(20))
    z2 = tmp_path / "a.zip"
    _make_zip(z2, {"shared/data.npz": buf.getvalue()})
    d2 = {"name": "a", "data_path": str(z2), "data_split_type": False, "data_transfer_type": "FILE",
          "data_path_logical": str(tmp_path / "al.zip"), "data_path_device": str(tmp_path / "ad.zip"),
          "allocation": {"logical_simulation": [3], "device_simulation": [1]}}
    assert HybridDataSplitter("t").split(d2)
    yl = np.load(_io.BytesIO(zipfile.ZipFile(tmp_path / "al.zip").read("shared/data.npz")))["y"]
    yd = np.load(_io.BytesIO(zipfile.ZipFile(tmp_path / "ad.zip").read("shared/data.npz")))["y"]
    assert len(yd) == 5 and len(yl) == 15 and set(yl) | set(yd) == set(range(20))


# ---------------------------------------------------------------------------------------------------------
# script operators: per-actor download / unzip / split_index folders / per-round model file
# ---------------------------------------------------------------------------------------------------------
def test_script_actor_fetches_sharded_data_and_round_model_over_http(tmp_path):
    import http.server
    import socketserver
    import threading
    from olearning_sim_b200.control.run_task import RoundRunner
    www = tmp_path / "www"
    www.mkdir()
    _make_zip(www / "shards.zip", {"dataconfig.json": "{\"classes\": 3}", **{f"dev{i}/sample.txt": f"shard {i}" for i in range(4)}})
    (www / "tsk_1_result_model.bin").write_bytes(b"round-1-model")
    (www / "init_model.bin").write_bytes(b"init-model")
    handler = lambda *a, **k: http.server.SimpleHTTPRequestHandler(*a, directory=str(www), **k)  # noqa: E731
    srv = socketserver.TCPServer(("127.0.0.1", 0), handler)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    base = f"http://127.0.0.1:{srv.server_address[1]}"
    op = tmp_path / "ops" / "training"
    op.mkdir(parents=True)
    log = tmp_path / "seen.jsonl"
    (op / "train.py").write_text(
        "import json, os, sys\n"
        "p = json.load(open(sys.argv[sys.argv.index('--params') + 1]))\n"
        "rec = {'round': p['current_round'], 'idx': p['split_index'], 'shard': open(os.path.join(p['data']['data_path'], 'sample.txt')).read(),\n"
        "       'cfg': p['data']['dataconfig'], 'model': open(p['operator']['model']['model_path'], 'rb').read().decode(),\n"
        "       'n': p['simulation_num'], 'actor_n': p['actor_simulation_num'], 'save': os.path.isdir(p['actor_save_dir'])}\n"
        f"open({str(log)!r}, 'a').write(json.dumps(rec) + '\\n')\n")
    tj = make_task_json(task_id="tsk", rounds=2, nums=(4,), num_request=(2,), code_path=str(op), operator_params={})
    tj["target"]["data"][0].update(data_path=f"{base}/shards.zip", data_split_type=True, data_transfer_type="HTTP")
    tj["target"]["data"][0]["simulation_target"] = {"devices": ["High"], "nums": [4]}
    tj["operatorflow"]["operators"][0]["model"] = {"use_model": True, "model_for_train": True, "model_transfer_type": "HTTP",
                                                   "model_path": f"{base}/init_model.bin",
                                                   "model_update_style": "{task_id}_{current_round}_result_model.bin"}
    wd = tmp_path / "job"
    (wd / "training").mkdir(parents=True)
    (wd / "training" / "train.py").write_text((op / "train.py").read_text())
    try:
        assert RoundRunner(tj, working_dir=str(wd)).run()
    finally:
        srv.shutdown()
    seen = [json.loads(l) for l in log.read_text().splitlines()]
    assert len(seen) == 8
    r0 = sorted((s for s in seen if s["round"] == 0), key=lambda s: s["idx"])
    assert [s["shard"] for s in r0] == [f"shard {i}" for i in range(4)]                   # client i owns shard i
    assert all(s["cfg"] == {"classes": 3} and s["n"] == 4 and s["actor_n"] == 2 and s["save"] for s in seen)
    assert {s["model"] for s in seen if s["round"] == 0} == {"init-model"}
    assert {s["model"] for s in seen if s["round"] == 1} == {"round-1-model"}             # {task_id}_{current_round}_result_model


# ---------------------------------------------------------------------------------------------------------
# flow plan cache == per-round analysis; per-class accounting; subprocess executor context
# ---------------------------------------------------------------------------------------------------------
def test_flow_plan_plus_draws_reproduce_the_per_round_analysis():
    from olearning_sim_b200.deviceflow.strategy import Strategy
    for drop in ({"drop_probability": [0.3, 0.0]}, {"drop_amounts": [7, 2]}, {}):
        st = {"flow_dispatch": {"use_strategy": True, "total_dispatch_amount": 100,
                                "specific_interval": {"use": True, "time_type": "relative", "intervals": [[0, 6], [8, 12]],
                                                      "dispatch_rules": {"domains": [[0, 3.14], [0, 1]], "functions": ["math.sin(t)+0.2", "2*t"]},
                                                      **({"drop_simulation": drop} if drop else {})}}}
        assert Strategy.is_round_invariant(st)
        plan = Strategy.flow_plan(st)
        for r in range(3):
            timing, amount, drops = Strategy.flow_strategy_analysis(st, f"x_y_{r}", rng=random.Random(r))
            assert plan["amount"] == amount and plan["timing"] == timing
            assert Strategy.draw_drops(plan, random.Random(r)) == drops
        # an injected integrator (the GPU kernel's stand-in) sees the 101-node samples of every bin
        ref = {}

        def area_fn(y):
            assert y.shape[1] == 101 and y.dtype == np.float64
            out = []
            for row in y:
                a = 0.0
                for j in range(1, 101):
                    s = 0.5 * (row[j] + row[j - 1]) * (1 / 100)
                    if s > 0:
                        a += s
                out.append(a)
            ref["called"] = True
            return np.array(out)
        assert Strategy.flow_plan(st, area_fn=area_fn)["amount"] == plan["amount"] and ref["called"]
    absolute = {"flow_dispatch": {"use_strategy": True, "total_dispatch_amount": 1,
                                  "specific_timing": {"use": True, "time_type": "absolute", "timings": [["2030-01-01 00:00:00"]], "amounts": [1]}}}
    assert not Strategy.is_round_invariant(absolute)


def test_per_class_failures_are_charged_to_the_right_class(monkeypatch):
    """a poisoned Low client: FAILED must show up under Low, not under the first class (run_task.py:149-210)"""
    from olearning_sim_b200.control.run_task import RoundRunner
    monkeypatch.setenv("OLSB_FAULTS", "corrupt:client=5,round=0")          # global client 5 = the second Low client
    tj = make_task_json(task_id="pc", rounds=1, nums=(4, 3), devices=("High", "Low"), dynamic=(0, 1),
                        operator_params={"model": "mlp2", "local_steps": 1, "batch": 8, "shard": 8, "lr": 0.1, "device": "cpu",
                                         "device_classes": {"Low": {"p_online": 1.0, "local_steps": 1}}})
    from olearning_sim_b200.control.repo import TaskTableRepo
    tj["target"]["data"][0]["simulation_target"] = {"devices": ["High", "Low"], "nums": [4, 3]}     # as the JobSubmitter writes it
    repo = TaskTableRepo()
    repo.create_task("pc")
    rr = RoundRunner(tj, repo=repo)
    assert rr.run()
    rec = rr.history[-1]
    assert rec["class_failed"] == {"data_0/High": 0, "data_0/Low": 1} and rec["class_participants"] == {"data_0/High": 4, "data_0/Low": 3}
    got = json.loads(repo.get_task_value("pc", "logical_result"))["logical_result"][0]["simulation_target"]
    assert got == {"devices": ["High", "Low"], "success_num": [4, 2], "failed_num": [0, 1]}


def test_subprocess_executor_serialises_the_control_plane_context(tmp_path):
    """build_control_plane(job_executor='subprocess') hands live repo / deviceflow objects to the
    executor; they must not reach json.dump, and the task must run to SUCCEEDED in the child process."""
    from olearning_sim_b200 import build_control_plane
    from olearning_sim_b200.utils.config import OlsbConfig
    cfg = OlsbConfig()
    cfg.cluster.job_executor = "subprocess"
    cfg.cluster.gpus = 1
    cfg.storage.task_table = str(tmp_path / "tasks.db")
    tm, rm, df = build_control_plane(cfg)
    tj = make_task_json(task_id="sp", rounds=1, operator_params={"model": "mlp2", "local_steps": 1, "batch": 8, "shard": 8,
                                                                 "lr": 0.1, "device": "cpu", "backend": "torch"})
    assert tm.submitTask(tj)
    tm.schedule_once(inline=True)
    deadline = time.time() + 240
    st = None
    while time.time() < deadline:
        st = tm.getTaskStatus("sp")
        if st in TaskStatus.TERMINAL:
            break
        time.sleep(0.5)
    assert st == TaskStatus.SUCCEEDED, (st, tm._executor.get_job_error(tm._task_running.get("sp", {}).get("job_id", "")) if hasattr(tm, "_task_running") else "")


# ---------------------------------------------------------------------------------------------------------
# bounded expression evaluation, cross-rank resume agreement, relaunch under lock
# ---------------------------------------------------------------------------------------------------------
def test_safe_expr_is_float_only_and_maps_arithmetic_errors():
    from olearning_sim_b200.deviceflow import safe_expr as se
    t0 = time.time()
    for bad in ("(((((2)**64)**64)**64)**64)**64", "((((((3)**60)**60)**60)**60)**60)**60", "1/(t-t)", "math.log(-1)", "math.exp(10000)"):
        with pytest.raises(se.UnsafeExpression):
            se.evaluate(bad, 2.0)
    assert time.time() - t0 < 1.0                                   # no big-int blow-up
    assert se.evaluate("int(t)*2+round(2.7)", 3.9) == 9.0 and abs(se.evaluate("math.sin(t)+1", 0.5) - 1.479425538604203) < 1e-12
    assert isinstance(se.evaluate("2**10", 0.0), float)


def test_resume_round_is_the_minimum_loadable_round(tmp_path):
    import torch
    from olearning_sim_b200.engine.checkpoint import CheckpointManager
    r0 = CheckpointManager(str(tmp_path), "tk", rank=0, keep=2, asynchronous=False)
    r1 = CheckpointManager(str(tmp_path), "tk", rank=1, keep=2, asynchronous=False)
    for rd in range(3):
        r0.save(rd, {"round": rd, "engine": {"w": torch.full((2,), float(rd))}})
    for rd in range(2):                                               # rank 1 crashed before writing round 2
        r1.save(rd, {"round": rd, "engine": {"w": torch.full((2,), float(rd))}})
    assert r0.newest_loadable_round() == 2 and r1.newest_loadable_round() == 1
    agreed = min(r0.newest_loadable_round(), r1.newest_loadable_round())
    assert int(r0.load_latest(max_round=agreed)["round"]) == 1 and int(r1.load_latest(max_round=agreed)["round"]) == 1
    # a torn newest file on rank 1 falls back one more round
    with open(r1.path(1), "wb") as f:
        f.write(b"torn")
    assert r1.newest_loadable_round() == 0
    # single-rank agreement path
    assert int(r0.load_agreed(None, 1)["round"]) == 2


def test_subprocess_relaunch_happens_once_under_concurrent_polls(tmp_path, monkeypatch):
    import threading
    from olearning_sim_b200.control.jobs import JobStatus, SubprocessJobExecutor
    ex = SubprocessJobExecutor(nproc_per_node=4, workdir=str(tmp_path), max_restarts=1)
    launches = []

    class P:
        def __init__(self, rc):
            self.rc, self.pid = rc, 0

        def poll(self):
            return self.rc

    def fake_launch(job_id, path, nproc, attempt, port=None):
        launches.append((nproc, attempt))
        time.sleep(0.05)
        return P(1 if attempt == 0 else None), str(tmp_path / "x.log")
    monkeypatch.setattr(ex, "_launch", fake_launch)
    jid = ex.submit_job({"task_id": "x"}, {"repo": object(), "deviceflow": object(), "gpus": 2, "checkpoint_dir": None})
    blob = json.load(open(tmp_path / f"{jid}.json"))
    assert "repo" not in blob["context"] and "deviceflow" not in blob["context"] and launches == [(2, 0)]      # granted GPUs -> ranks
    out = []
    ths = [threading.Thread(target=lambda: out.append(ex.get_job_status(jid))) for _ in range(6)]
    [t.start() for t in ths]; [t.join() for t in ths]
    assert out == [JobStatus.RUNNING] * 6 and launches == [(2, 0), (2, 1)]                                     # exactly one relaunch


# ---------------------------------------------------------------------------------------------------------
# broker / database adapters against fake servers
# ---------------------------------------------------------------------------------------------------------
class FakePulsarClient:
    """pulsar.Client surface used by PulsarBus: topics with named Shared subscriptions, ack / redelivery, Timeout"""

    class Timeout(Exception):
        pass

    class _Msg:
        def __init__(self, mid, data):
            self._id, self._d = mid, data

        def data(self):
            return self._d

        def message_id(self):
            return self._id

    class _Producer:
        def __init__(self, c, topic):
            self.c, self.topic = c, topic

        def send(self, payload):
            self.c._n += 1
            for sub in self.c.subs.setdefault(self.topic, {}).values():
                sub["pending"].append(FakePulsarClient._Msg(self.c._n, payload))
            self.c.log.setdefault(self.topic, []).append(payload)

        def close(self):
            pass

    class _Consumer:
        def __init__(self, c, topic, name):
            self.sub = c.subs.setdefault(topic, {}).setdefault(name, {"pending": [], "inflight": {}})

        def receive(self, timeout_millis=None):
            if not self.sub["pending"]:
                raise FakePulsarClient.Timeout("Pulsar error: TimeOut")
            m = self.sub["pending"].pop(0)
            self.sub["inflight"][m.message_id()] = m
            return m

        def acknowledge(self, m):
            self.sub["inflight"].pop(m.message_id(), None)

        def close(self):
            self.sub["pending"][:0] = list(self.sub["inflight"].values())
            self.sub["inflight"].clear()

    def __init__(self):
        self.subs, self.log, self._n = {}, {}, 0

    def create_producer(self, topic):
        return self._Producer(self, topic)

    def subscribe(self, topic, subscription_name, consumer_type=None):
        return self._Consumer(self, topic, subscription_name)

    def close(self):
        pass


def test_deviceflow_service_runs_on_the_pulsar_transport():
    """sorter -> shelf -> dispatcher -> outbound, every hop through PulsarBus (fake broker client)"""
    from olearning_sim_b200.control.repo import FlowTableRepo
    from olearning_sim_b200.deviceflow.bus import BusTimeout, FakeClock
    from olearning_sim_b200.deviceflow.pulsar_bus import PulsarBus
    from olearning_sim_b200.deviceflow.rooms import PulsarClientProducer
    from olearning_sim_b200.deviceflow.service import DeviceFlowService, NotifyRequest
    client = FakePulsarClient()
    bus = PulsarBus("pulsar://fake:6650", client=client)
    bus.virtual_time = True                 # inline pump(): a receive timeout ends the dispatcher pass instead of polling
    PulsarClientProducer.buses["pulsar://fake:6650"] = bus
    df = DeviceFlowService(bus=bus, repo=FlowTableRepo(), clock=FakeClock(), seed=1, dispatcher_timeout_millis=0)
    st = json.dumps({"real_time_dispatch": {"use_strategy": True, "dispatch_batch_sizes": [2], "drop_simulation": {"drop_probability": 0.0}},
                     "flow_dispatch": {"use_strategy": False}})
    assert df.RegisterTask("pt", ["logical_simulation"])
    out = {"pulsar_client": {"url": "pulsar://fake:6650", "topic": "agg"}}
    assert df.NotifyStart(NotifyRequest("pt", "pt_training_0", "logical_simulation", st, out))
    inbound = bus.create_producer(df.config.source_topic)
    for i in range(5):
        inbound.send(json.dumps({"routing_key": "pt_training_0", "compute_resource": "logical_simulation", "message": f"u{i}"}))
    inbound.send(json.dumps({"routing_key": "other_flow_0", "compute_resource": "logical_simulation", "message": "stray"}))
    df._sorter.drain()                                              # inbound -> shelf (sorter thread in a live service)
    assert df.NotifyComplete(NotifyRequest("pt", "pt_training_0", "logical_simulation"))
    df.pump(); df.pump()
    full = bus._full("agg")
    assert [m.decode() for m in client.log[full]] == [f"u{i}" for i in range(5)]          # 2 + 2 + flushed 1
    assert df.CheckDeviceflowDispatchFinished("pt") and df.stats()["sorter"]["discarded"] == 1
    c = bus.subscribe("empty_topic", "s")
    with pytest.raises(BusTimeout):
        c.receive(timeout_millis=1)
    del PulsarClientProducer.buses["pulsar://fake:6650"]


class FakeMySql:
    """DB-API connection factory in the pymysql dialect (%s placeholders, backticks, AUTO_INCREMENT) backed by sqlite;
    the first statement after `drop_next` raises an operational error, as a dropped server connection would"""

    class OperationalError(Exception):
        pass

    def __init__(self):
        import sqlite3
        self.db = sqlite3.connect(":memory:", check_same_thread=False, isolation_level=None)
        self.drop_next = False
        self.connects = 0
        self.seen = []

    def connect(self):
        outer = self
        outer.connects += 1

        class Cur:
            def execute(self, q, params=()):
                if outer.drop_next:
                    outer.drop_next = False
                    raise FakeMySql.OperationalError("(2006, 'MySQL server has gone away')")
                assert "?" not in q and '"' not in q, q
                outer.seen.append(q)
                q2 = (q.replace("`", '"').replace("%s", "?").replace("BIGINT PRIMARY KEY AUTO_INCREMENT", "INTEGER PRIMARY KEY AUTOINCREMENT")
                      .replace("LONGTEXT", "TEXT").replace("DOUBLE", "REAL"))
                self._c = outer.db.execute(q2, params)
                self.rowcount, self.description = self._c.rowcount, self._c.description
                return self

            def fetchone(self):
                return self._c.fetchone()

            def fetchall(self):
                return self._c.fetchall()

        class Conn:
            def cursor(self):
                return Cur()

            def commit(self):
                pass

            def close(self):
                pass
        return Conn()


def test_mysql_repos_translate_dialect_and_reconnect_once():
    from olearning_sim_b200.control.mysql_repo import MySqlConnection, MySqlResTableRepo, MySqlTaskTableRepo, translate
    assert translate('SELECT "a" FROM "t" WHERE "b"=?') == "SELECT `a` FROM `t` WHERE `b`=%s"
    srv = FakeMySql()
    conn = MySqlConnection(connect=srv.connect, operational_errors=(FakeMySql.OperationalError,))
    tasks, res = MySqlTaskTableRepo(conn), MySqlResTableRepo(conn)
    assert any("AUTO_INCREMENT" in q and "`taskmgr_table`" in q for q in srv.seen)
    assert tasks.create_task("m1", "u") and not tasks.create_task("m1", "u")
    assert tasks.set_task_value("m1", "task_status", "QUEUED") and tasks.get_task_value("m1", "task_status") == "QUEUED"
    srv.drop_next = True                                                # server went away: reconnect once and retry
    assert tasks.get_task_value("m1", "task_status") == "QUEUED" and conn.reconnects == 1 and srv.connects == 2
    assert res.insert_data("m1", 2.0, 4.0, 1.0) and res.held()[0]["cpu"] == 2.0
    assert res.delete_by_taskid("m1") and res.held() == []
    # the control plane runs unchanged on top of it
    from olearning_sim_b200.control.resource_manager import FakePhoneMgr, ResourceManager
    rm = ResourceManager(res, totals={"cpu": 4, "mem": 8, "gpu": 0}, phone_mgr=FakePhoneMgr())
    assert rm.requestClusterResource("x", 3, 1).name == "SUCCESS" and rm.requestClusterResource("y", 3, 1).name == "NOT_ENOUGH_RES"
    srv.drop_next = True
    with pytest.raises(FakeMySql.OperationalError):                     # two failures in a row are not swallowed
        srv_drop_again = [True]

        class Twice:
            def __call__(self):
                c = srv.connect()
                srv.drop_next = True
                return c
        MySqlConnection(connect=Twice(), operational_errors=(FakeMySql.OperationalError,)).execute('SELECT 1')


# ---------------------------------------------------------------------------------------------------------
# communication runtime: multicast policy and the peer-loss error (host-side logic; the device paths are GPU tests)
# ---------------------------------------------------------------------------------------------------------
def test_multicast_policy_and_peer_lost_error(monkeypatch):
    from olearning_sim_b200.parallel import nvls
    from olearning_sim_b200.parallel.comm import PeerLostError
    monkeypatch.delenv("OLSB_NVLS", raising=False)
    monkeypatch.delenv("OLSB_NVLS_MIN_BYTES", raising=False)
    assert not nvls.enabled(1) and nvls.enabled(8)                     # never for a single rank
    assert nvls.min_bytes() == 4 << 20                                 # small (latency-bound) slabs stay on CUDA IPC
    monkeypatch.setenv("OLSB_NVLS_MIN_BYTES", "1024")
    assert nvls.min_bytes() == 1024
    monkeypatch.setenv("OLSB_NVLS", "1")
    assert nvls.min_bytes() == 0 and nvls.enabled(2)                   # forced on: every size
    monkeypatch.setenv("OLSB_NVLS", "0")
    assert not nvls.enabled(8)
    e = PeerLostError(3, 17)
    assert e.peer == 3 and e.round_idx == 17 and PeerLostError.exit_code == 75 and "rank 3" in str(e)


def test_results_tables_are_generated_from_raw_lines(tmp_path):
    """tools/make_results_md.py must only format what the JSON lines say (efficiency = value(N) / (N value(1)))."""
    import importlib.util
    import json
    import shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("mk", os.path.join(root, "tools", "make_results_md.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    prof = tmp_path / "profiles"
    prof.mkdir()
    for n, v in ((1, 100.0), (8, 640.0)):
        line = {"metric": "fl_client_local_steps_per_sec", "value": v, "ms_per_step": 1000.0 / v, "n_gpus": n,
                "config": {"id": 2, "model": "cnn2 (x)"}, "e2e": {"value": v * 0.9}, "vs_baseline": 10.0,
                "baseline": {"value": v / 10, "ms_per_step": 1.0, "e2e": {"value": v / 11}, "e2e_ratio": 9.9}, "clocks": {"sm_mhz": 1900}}
        (prof / f"r2_bench_c2_n{n}.json").write_text(json.dumps(line) + "\n")
    shutil.copy(os.path.join(root, "BASELINE.md"), tmp_path / "BASELINE.md")
    shutil.copy(os.path.join(root, "README.md"), tmp_path / "README.md")
    mk.ROOT, mk.P = str(tmp_path), str(prof)
    mk.main()
    txt = (prof / "r2_results.md").read_text()
    assert "| 2 | cnn2 | 8 | 640 |" in txt and "0.800" in txt             # 640 / (8 * 100)
    assert "640" in (tmp_path / "README.md").read_text() and "640" in (tmp_path / "BASELINE.md").read_text()


@pytest.mark.parametrize("name", ["task_mlp_cpu.json", "task_cnn2_gpu.json", "task_resnet18_gpu.json", "task_bert_mini_gpu.json"])
def test_example_tasks_pass_the_reference_validators_and_the_wire(name):
    """every shipped example is a valid reference-schema task: validators accept it, it survives the protobuf round trip, and
    its DeviceFlow strategy (when it has one) passes the strategy validator"""
    from olearning_sim_b200.api.validate import ValidateParameters
    from olearning_sim_b200.deviceflow.validate import ValidateParameters as DfValidate
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    task = json.load(open(os.path.join(root, "examples", name)))
    from olearning_sim_b200.api.schema import json2taskconfig
    ok, msg = ValidateParameters().validate_task_parameters(json2taskconfig(task))
    assert ok, msg
    back = wire.task_pb_to_json(wire.task_json_to_pb(task))
    assert back["task_id"] == task["task_id"] and len(back["operatorflow"]["operators"]) == len(task["operatorflow"]["operators"])
    for op in task["operatorflow"]["operators"]:
        obc = op.get("operation_behavior_controller") or {}
        if obc.get("use_gradient_house") and obc.get("strategy_gradient_house"):
            ok, msg = DfValidate().check_strategy(json.loads(obc["strategy_gradient_house"]))
            assert ok, msg
        params = op["logical_simulation"].get("operator_params")
        if params:
            assert isinstance(json.loads(params), dict)


class WsPulsarBroker:
    """In-process emulation of the two Pulsar WebSocket endpoints (producer / consumer) on a real ``websockets`` server: a
    topic is a log, a subscription a cursor plus the set of delivered-but-unacknowledged ids, which are redelivered to the
    next consumer that attaches (what a broker does after a consumer disconnects)."""

    def __init__(self):
        import threading
        from websockets.sync.server import serve
        self.lock = threading.Lock()
        self.topics, self.subs, self.n = {}, {}, 0                   # topic -> [(id, b64)], (topic, sub) -> {"cursor", "unacked"}
        self.server = serve(self._handle, "127.0.0.1", 0)
        self.port = self.server.socket.getsockname()[1]
        threading.Thread(target=self.server.serve_forever, daemon=True).start()

    def _handle(self, ws):
        import urllib.parse
        u = urllib.parse.urlparse(ws.request.path)
        parts = u.path.strip("/").split("/")                         # ws v2 producer|consumer persistent tenant ns topic [sub]
        kind, topic = parts[2], "/".join(parts[3:7])
        if kind == "producer":
            for frame in ws:
                m = json.loads(frame)
                with self.lock:
                    self.n += 1
                    mid = f"{self.n}:0"
                    self.topics.setdefault(topic, []).append((mid, m["payload"]))
                ws.send(json.dumps({"result": "ok", "messageId": mid, "context": m.get("context")}))
            return
        sub = self.subs.setdefault((topic, parts[7]), {"cursor": 0, "unacked": {}})
        assert "subscriptionType=Shared" in u.query
        with self.lock:
            redeliver = list(sub["unacked"].items())
        for mid, payload in redeliver:                                # unacknowledged from a previous connection first
            ws.send(json.dumps({"messageId": mid, "payload": payload, "properties": {}, "redeliveryCount": 1}))
        import time
        while True:
            with self.lock:
                log = self.topics.get(topic, [])
                new = log[sub["cursor"]:]
                sub["cursor"] = len(log)
                for mid, payload in new:
                    sub["unacked"][mid] = payload
            for mid, payload in new:
                ws.send(json.dumps({"messageId": mid, "payload": payload, "properties": {}}))
            try:
                ack = ws.recv(timeout=0.01)
                with self.lock:
                    sub["unacked"].pop(json.loads(ack)["messageId"], None)
            except TimeoutError:
                pass
            except Exception:                                         # connection closed
                return
            time.sleep(0.002)

    def close(self):
        self.server.shutdown()


def test_pulsar_websocket_client_against_a_broker_emulation():
    """PulsarBus("ws://...") = the broker's WebSocket API end to end: publish, Shared subscription, ack, receive timeout,
    redelivery of unacknowledged messages after a reconnect, and the DeviceFlow sorter -> shelf -> dispatcher -> outbound
    path with every hop on that transport"""
    from olearning_sim_b200.control.repo import FlowTableRepo
    from olearning_sim_b200.deviceflow.bus import BusTimeout, FakeClock
    from olearning_sim_b200.deviceflow.pulsar_bus import PulsarBus
    from olearning_sim_b200.deviceflow.rooms import PulsarClientProducer
    from olearning_sim_b200.deviceflow.service import DeviceFlowService, NotifyRequest
    broker = WsPulsarBroker()
    url = f"ws://127.0.0.1:{broker.port}"
    try:
        bus = PulsarBus(url)
        prod = bus.create_producer("jobs")
        cons = bus.subscribe("jobs", "workers")
        for i in range(3):
            prod.send(f"m{i}".encode())
        got = [cons.receive(timeout_millis=2000) for _ in range(3)]
        assert [m.data() for m in got] == [b"m0", b"m1", b"m2"]
        cons.acknowledge(got[0])
        with pytest.raises(BusTimeout):
            cons.receive(timeout_millis=50)
        time.sleep(0.1)                                               # let the ack land before the connection drops
        cons.close()
        cons2 = bus.subscribe("jobs", "workers")                      # same subscription: m1, m2 come back, m0 does not
        again = sorted(cons2.receive(timeout_millis=2000).data() for _ in range(2))
        assert again == [b"m1", b"m2"]
        with pytest.raises(BusTimeout):
            cons2.receive(timeout_millis=50)
        # ---- DeviceFlow on the WebSocket transport
        bus.virtual_time = True
        PulsarClientProducer.buses[url] = bus
        df = DeviceFlowService(bus=bus, repo=FlowTableRepo(), clock=FakeClock(), seed=1, dispatcher_timeout_millis=200)
        st = json.dumps({"real_time_dispatch": {"use_strategy": True, "dispatch_batch_sizes": [2], "drop_simulation": {"drop_probability": 0.0}},
                         "flow_dispatch": {"use_strategy": False}})
        assert df.RegisterTask("wt", ["logical_simulation"])
        agg = bus.subscribe("agg", "aggregator")                      # the cloud aggregator's subscription, before anything is sent
        assert df.NotifyStart(NotifyRequest("wt", "wt_training_0", "logical_simulation", st, {"pulsar_client": {"url": url, "topic": "agg"}}))
        inbound = bus.create_producer(df.config.source_topic)
        for i in range(5):
            inbound.send(json.dumps({"routing_key": "wt_training_0", "compute_resource": "logical_simulation", "message": f"u{i}"}))
        deadline = time.time() + 5
        while df.stats()["sorter"]["shelved"] < 5 and time.time() < deadline:
            df._sorter.drain()
        assert df.NotifyComplete(NotifyRequest("wt", "wt_training_0", "logical_simulation"))
        out = []
        deadline = time.time() + 10
        while len(out) < 5 and time.time() < deadline:
            df.pump()
            try:
                while True:
                    m = agg.receive(timeout_millis=100)
                    out.append(m.data().decode()); agg.acknowledge(m)
            except BusTimeout:
                pass
        assert sorted(out) == [f"u{i}" for i in range(5)]
        assert df.CheckDeviceflowDispatchFinished("wt")
        del PulsarClientProducer.buses[url]
        bus.close()
    finally:
        broker.close()


class WireMySqlServer:
    """A MySQL server on the wire protocol, over TCP, backed by sqlite: HandshakeV10 with a random salt, verification of the
    mysql_native_password scramble (recomputed here from the stored password), COM_QUERY -> OK / ERR / text result set with
    classic EOF framing, COM_PING, COM_QUIT.  ``drop_next`` closes the socket in the middle of the next query."""

    def __init__(self, user="ols", password="pw-1", database="virtually_phone_DB", auth="native"):
        import socket
        import sqlite3
        import threading
        self.user, self.password, self.database, self.auth = user, password, database, auth      # native | sha2_fast | sha2_full
        self.db = sqlite3.connect(":memory:", check_same_thread=False, isolation_level=None)
        self.lock, self.drop_next, self.connects, self.seen = threading.Lock(), False, 0, []
        self.sock = socket.socket()
        self.sock.bind(("127.0.0.1", 0))
        self.sock.listen(8)
        self.port = self.sock.getsockname()[1]
        threading.Thread(target=self._accept, daemon=True).start()

    def _accept(self):
        import threading
        while True:
            try:
                c, _ = self.sock.accept()
            except OSError:
                return
            threading.Thread(target=self._serve, args=(c,), daemon=True).start()

    @staticmethod
    def _send(c, seq, payload):
        c.sendall(len(payload).to_bytes(3, "little") + bytes([seq]) + payload)

    @staticmethod
    def _recv(c):
        def exact(n):
            b = b""
            while len(b) < n:
                x = c.recv(n - len(b))
                if not x:
                    raise ConnectionError
                b += x
            return b
        h = exact(4)
        return h[3], exact(int.from_bytes(h[:3], "little"))

    @staticmethod
    def _lenenc(b):
        n = len(b)
        return (bytes([n]) if n < 251 else b"\xfc" + n.to_bytes(2, "little")) + b

    def _ok(self, c, seq, affected=0, last=0):
        self._send(c, seq, b"\x00" + self._lenenc(b"")[:0] + bytes([affected if affected < 251 else 250, last if last < 251 else 250]) + b"\x02\x00\x00\x00")

    def _err(self, c, seq, code, msg):
        self._send(c, seq, b"\xff" + code.to_bytes(2, "little") + b"#HY000" + msg.encode())

    def _serve(self, c):
        import hashlib
        import os as _os
        import struct
        self.connects += 1
        salt = bytes(33 + b % 90 for b in _os.urandom(20))            # printable, no NULs (like a real server)
        caps = 0x1 | 0x4 | 0x8 | 0x200 | 0x2000 | 0x8000 | 0x20000 | 0x80000
        hs = (b"\x0a" + b"8.0.99-fake\0" + struct.pack("<I", self.connects) + salt[:8] + b"\0" + struct.pack("<H", caps & 0xFFFF) +
              b"\x2d" + struct.pack("<H", 2) + struct.pack("<H", caps >> 16) + bytes([21]) + b"\0" * 10 + salt[8:] + b"\0" +
              (b"mysql_native_password\0" if self.auth == "native" else b"caching_sha2_password\0"))
        try:
            self._send(c, 0, hs)
            seq, resp = self._recv(c)
            ccaps = struct.unpack_from("<I", resp, 0)[0]
            pos = 32
            end = resp.index(b"\0", pos)
            user = resp[pos:end].decode()
            pos = end + 1
            alen = resp[pos]
            auth = resp[pos + 1:pos + 1 + alen]
            pos += 1 + alen
            dbname = ""
            if ccaps & 0x8:
                end = resp.index(b"\0", pos)
                dbname = resp[pos:end].decode()
            if self.auth == "native":
                p1 = hashlib.sha1(self.password.encode()).digest()
                want = bytes(a ^ b for a, b in zip(p1, hashlib.sha1(salt + hashlib.sha1(p1).digest()).digest()))
                good = auth == want
            elif self.auth == "sha2_fast":                            # the password hash is in the server's cache
                p1 = hashlib.sha256(self.password.encode()).digest()
                want = bytes(a ^ b for a, b in zip(p1, hashlib.sha256(hashlib.sha256(p1).digest() + salt).digest()))
                good = auth == want
                if good:
                    seq += 1
                    self._send(c, seq, b"\x01\x03")
            else:                                                     # cold cache: full authentication over RSA
                from cryptography.hazmat.primitives import hashes, serialization
                from cryptography.hazmat.primitives.asymmetric import padding, rsa
                key = rsa.generate_private_key(public_exponent=65537, key_size=2048)
                pem = key.public_key().public_bytes(serialization.Encoding.PEM, serialization.PublicFormat.SubjectPublicKeyInfo)
                self._send(c, seq + 1, b"\x01\x04")
                seq, req = self._recv(c)
                assert req == b"\x02"
                self._send(c, seq + 1, b"\x01" + pem)
                seq, enc = self._recv(c)
                x = key.decrypt(enc, padding.OAEP(mgf=padding.MGF1(algorithm=hashes.SHA1()), algorithm=hashes.SHA1(), label=None))
                good = bytes(b ^ salt[i % len(salt)] for i, b in enumerate(x)) == self.password.encode() + b"\0"
            if user != self.user or not good:
                self._err(c, seq + 1, 1045, f"Access denied for user '{user}'")
                return c.close()
            if dbname != self.database:
                self._err(c, seq + 1, 1049, f"Unknown database '{dbname}'")
                return c.close()
            self._ok(c, seq + 1)
            while True:
                seq, pkt = self._recv(c)
                if pkt[0] == 0x01:
                    return c.close()
                if pkt[0] == 0x0e:
                    self._ok(c, 1)
                    continue
                q = pkt[1:].decode()
                if self.drop_next:
                    self.drop_next = False
                    return c.close()                                  # connection lost mid-query
                self._query(c, q)
        except (ConnectionError, OSError):
            return

    def _query(self, c, q):
        import sqlite3
        self.seen.append(q)
        up = q.strip().upper()
        if up.startswith("SET ") or up in ("COMMIT", "ROLLBACK"):
            return self._ok(c, 1)
        q2 = (q.replace("`", '"').replace("BIGINT PRIMARY KEY AUTO_INCREMENT", "INTEGER PRIMARY KEY AUTOINCREMENT")
              .replace("LONGTEXT", "TEXT").replace("DOUBLE", "REAL").replace("\\'", "''"))
        try:
            with self.lock:
                cur = self.db.execute(q2)
                rows = cur.fetchall() if cur.description else None
                aff, last = max(cur.rowcount, 0), cur.lastrowid or 0
        except sqlite3.IntegrityError as e:
            return self._err(c, 1, 1062, f"Duplicate entry: {e}")
        except sqlite3.Error as e:
            return self._err(c, 1, 1064, f"You have an error in your SQL syntax: {e}")
        if rows is None:
            return self._ok(c, 1, aff, last)
        ncol = len(cur.description)
        seq = 1
        self._send(c, seq, bytes([ncol]))
        for i, d in enumerate(cur.description):
            sample = next((r[i] for r in rows if r[i] is not None), "")
            ctype = 8 if isinstance(sample, int) else (5 if isinstance(sample, float) else 253)
            name = d[0].encode()
            col = (self._lenenc(b"def") + self._lenenc(b"db") + self._lenenc(b"t") + self._lenenc(b"t") + self._lenenc(name) + self._lenenc(name) +
                   b"\x0c" + (63 if ctype != 253 else 45).to_bytes(2, "little") + (1024).to_bytes(4, "little") + bytes([ctype]) + b"\0\0" + b"\0" + b"\0\0")
            seq += 1
            self._send(c, seq, col)
        seq += 1
        self._send(c, seq, b"\xfe\0\0\x02\0")
        for r in rows:
            seq = (seq + 1) & 0xFF
            self._send(c, seq, b"".join(b"\xfb" if v is None else self._lenenc(str(v).encode()) for v in r))
        self._send(c, (seq + 1) & 0xFF, b"\xfe\0\0\x02\0")

    def close(self):
        self.sock.close()


def test_mysql_repos_over_the_wire_protocol():
    """the table repos on the in-tree MySQL client over TCP: handshake + native-password authentication (verified by the
    server), DDL / DML / SELECT with typed result sets and NULLs, escaping of hostile strings, duplicate keys, a connection
    that dies mid-query (reconnect once and retry), wrong password, and the ResourceManager running on top"""
    from olearning_sim_b200.control.mysql_repo import MySqlConnection, MySqlResTableRepo, MySqlTaskTableRepo
    from olearning_sim_b200.utils import mysql_wire
    srv = WireMySqlServer()
    try:
        kw = dict(host="127.0.0.1", port=srv.port, user="ols", password="pw-1", database="virtually_phone_DB")
        raw = mysql_wire.connect(**kw)
        assert raw.ping()
        cur = raw.cursor()
        cur.execute("CREATE TABLE `t` (`a` BIGINT PRIMARY KEY AUTO_INCREMENT, `s` LONGTEXT, `x` DOUBLE)")
        nasty = "it's a \\\\ \"quoted\"\nline; DROP TABLE t; --"
        cur.execute("INSERT INTO `t` (`s`, `x`) VALUES (%s, %s)", (nasty, 1.5))
        cur.execute("INSERT INTO `t` (`s`, `x`) VALUES (%s, %s)", (None, None))
        cur.execute("SELECT `a`, `s`, `x` FROM `t` ORDER BY `a`")
        rows = cur.fetchall()
        assert rows[0][0] == 1 and isinstance(rows[0][0], int) and rows[0][2] == 1.5 and rows[1][1:] == (None, None)
        assert rows[0][1].startswith("it's a ") and "DROP TABLE" in rows[0][1] and [d[0] for d in cur.description] == ["a", "s", "x"]
        with pytest.raises(mysql_wire.ProgrammingError):
            cur.execute("SELEC nonsense")
        with pytest.raises(mysql_wire.ProgrammingError):
            cur.execute("INSERT INTO `t` (`a`) VALUES (%s)", (1,))           # duplicate key -> 1062
        raw.close()
        with pytest.raises(mysql_wire.OperationalError):
            mysql_wire.connect(**{**kw, "password": "wrong"})                  # 1045 from the server's scramble check
        with pytest.raises(mysql_wire.OperationalError):
            mysql_wire.connect(**{**kw, "database": "nope"})
        # ---- the repos, through MySqlConnection's pymysql-less fallback
        conn = MySqlConnection(**kw)
        tasks, res = MySqlTaskTableRepo(conn), MySqlResTableRepo(conn)
        assert any("AUTO_INCREMENT" in q and "`taskmgr_table`" in q for q in srv.seen)
        assert tasks.create_task("w1", "u") and not tasks.create_task("w1", "u")
        assert tasks.set_task_value("w1", "task_status", "QUEUED") and tasks.get_task_value("w1", "task_status") == "QUEUED"
        before = srv.connects
        srv.drop_next = True                                              # the server drops the connection mid-query
        assert tasks.get_task_value("w1", "task_status") == "QUEUED" and conn.reconnects == 1 and srv.connects == before + 1
        assert res.insert_data("w1", 2.0, 4.0, 1.0) and res.held()[0]["cpu"] == 2.0
        assert res.delete_by_taskid("w1") and res.held() == []
        from olearning_sim_b200.control.resource_manager import FakePhoneMgr, ResourceManager
        rm = ResourceManager(res, totals={"cpu": 4, "mem": 8, "gpu": 0}, phone_mgr=FakePhoneMgr())
        assert rm.requestClusterResource("x", 3, 1).name == "SUCCESS" and rm.requestClusterResource("y", 3, 1).name == "NOT_ENOUGH_RES"
    finally:
        srv.close()


@pytest.mark.parametrize("mode", ["sha2_fast", "sha2_full"])
def test_mysql_wire_caching_sha2_password(mode):
    """MySQL 8 default accounts: caching_sha2_password — the fast path (hash in the server's cache: SHA-256 scramble) and the
    full authentication (RSA public-key exchange, OAEP) — against the protocol server; a wrong password is refused in both"""
    from olearning_sim_b200.utils import mysql_wire
    srv = WireMySqlServer(auth=mode)
    try:
        kw = dict(host="127.0.0.1", port=srv.port, user="ols", password="pw-1", database="virtually_phone_DB")
        c = mysql_wire.connect(**kw)
        cur = c.cursor()
        cur.execute("SELECT 40 + 2")
        assert cur.fetchone() == (42,) and c.ping()
        c.close()
        with pytest.raises(mysql_wire.OperationalError):
            mysql_wire.connect(**{**kw, "password": "wrong"})
    finally:
        srv.close()
