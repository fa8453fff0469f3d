"""gRPC session, cluster manager, blob stores, config layering, queue / fragment repos, logger, checkpoints, gates."""
import json
import os
import threading
import time

import pytest
import torch

from olearning_sim_b200.api.grpc_service import JsonStub, SimulatorSession, connect
from olearning_sim_b200.api.schema import TaskStatus
from olearning_sim_b200.cluster.manager import ClusterBuilder, ClusterUtils, Director, GpuClusterManager, LocalClusterApi
from olearning_sim_b200.control.jobs import InProcessJobExecutor
from olearning_sim_b200.control.operatorflow import OperatorFlow, WaitingForGlobalAggregation
from olearning_sim_b200.control.repo import FlowTableRepo, LogTableRepo, ResTableRepo, SqlDataBase, TaskTableRepo
from olearning_sim_b200.control.resource_manager import FakePhoneMgr, ResourceManager
from olearning_sim_b200.control.task_manager import TaskManager
from olearning_sim_b200.deviceflow.bus import FakeClock, InMemoryBus
from olearning_sim_b200.deviceflow.service import DeviceFlowService
from olearning_sim_b200.engine.checkpoint import CheckpointManager
from olearning_sim_b200.utils.config import load_config
from olearning_sim_b200.utils.file_repo import LocalFileRepo, get_file_repo
from olearning_sim_b200.utils.logger import Logger
from olearning_sim_b200.utils.misc_repos import (Fragment, JsonFragmentRepo, PerformanceManager, ProtoFragmentRepo,
                                                 QueueRepo)
from tests.helpers import make_task_json
from tests.test_control_plane import StubRunner


# ---- gRPC façade ---------------------------------------------------------------------------------------------
def test_grpc_session_serves_all_services_with_reference_method_names():
    StubRunner.gate = None
    repo = TaskTableRepo()
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 4, "mem": 8, "gpu": 0}, phone_mgr=FakePhoneMgr({"u1": {"High": 2}}))
    df = DeviceFlowService(bus=InMemoryBus(virtual_time=True), repo=FlowTableRepo(), clock=FakeClock())
    tm = TaskManager(repo=repo, resource_manager=rm, deviceflow=df, executor=InProcessJobExecutor(runner_factory=StubRunner))
    sess = SimulatorSession(port=0, task_manager=tm, resource_manager=rm, deviceflow=df,
                            cluster_manager=GpuClusterManager(LocalClusterApi(gpu_count=8)),
                            performance_manager=PerformanceManager()).start()
    try:
        assert sess.mounted == ["TaskMgr", "ResourceMgr", "RayClusterMgr", "PerformanceMgr", "TaskOperatorOrientedDeviceFlow"]
        ch = connect(sess.address)
        t, r, d = JsonStub(ch, "TaskMgr"), JsonStub(ch, "ResourceMgr"), JsonStub(ch, "TaskOperatorOrientedDeviceFlow")
        assert t.submitTask({"task": make_task_json(task_id="g1")}) == {"is_success": True}
        assert t.getTaskQueue() == {"tasks": ["g1"]} and t.getTaskStatus({"taskID": "g1"}) == {"taskStatus": "QUEUED"}
        assert t.getTaskStatus({"taskID": "zz"})["taskStatus"] == "MISSING"
        assert r.getResource()["cluster"]["cpu"] == 4 and r.getPhoneResource() == {"u1": {"High": 2}}
        assert r.requestClusterResource({"task_id": "x", "cpu": 9})["status"] == "NOT_ENOUGH_RES"
        assert r.requestResource({"task_id": "x", "user_id": "u1", "cpu": 1, "mem": 1, "phones": {"High": 1}, "type": 0}) == {"is_success": True, "status": "SUCCESS"}
        assert r.getClusterResourceDetail()["details"][0]["task_id"] == "x"
        assert r.releaseResource({"task_id": "x", "user_id": "u1", "type": 0})["is_success"]
        assert d.RegisterTask({"task_id": "g1", "total_compute_resources": ["logical_simulation"]}) == {"is_success": True}
        assert d.GetTotalComputeResources({"task_id": "g1"}) == {"total_compute_resources": ["logical_simulation"]}
        assert "topic" in d.GetDeviceflowPulsarClient()
        tm.schedule_once(inline=True)
        deadline = time.time() + 10
        while time.time() < deadline and t.getTaskStatus({"taskID": "g1"})["taskStatus"] != "SUCCEEDED":
            time.sleep(0.05)
        assert t.getTaskStatus({"taskID": "g1"})["taskStatus"] == "SUCCEEDED"
        assert t.stopTask({"taskID": "g1"})["is_success"] and t.changeScheduler({"strategy": "default"})["is_success"]
        c = JsonStub(ch, "RayClusterMgr")
        assert c.createRayCluster() == {"ok": True}
        q = json.loads(c.queryRayCluster({"k8s_namespace": "default", "ray_label": "demo1-cluster"})["json_data"])
        assert q["items"][0]["spec"]["workerGroupSpecs"][0]["replicas"] == 4
        assert JsonStub(ch, "PerformanceMgr").commitJob({"task_id": "g1", "t_round_ms": 10.0, "client_steps": 40})["is_success"]
    finally:
        sess.stop()


# ---- cluster manager ----------------------------------------------------------------------------------------------
def test_cluster_builder_utils_and_manager_lifecycle():
    b = ClusterBuilder()
    doc = b.build_meta("c1", "ns", {"lab": "yes"}).build_head().build_worker("workers", replicas=2, min_replicas=1, max_replicas=4, gpus=1).get_cluster()
    assert b.succeeded and doc["spec"]["workerGroupSpecs"][0]["template"]["resources"]["requests"]["nvidia.com/gpu"] == 1
    assert not ClusterBuilder().build_meta("c2").build_worker("w", replicas=9, max_replicas=4).succeeded
    u = ClusterUtils()
    d2, ok = u.update_worker_group_replicas(doc, "workers", 8, 2, 6)
    assert ok and d2["spec"]["workerGroupSpecs"][0]["replicas"] == 6 and doc["spec"]["workerGroupSpecs"][0]["replicas"] == 2
    assert not u.update_worker_group_replicas(doc, "workers", 2, 1, 5)[1]
    assert u.duplicate_worker_group(doc, "workers", "w2")[1] and u.delete_worker_group(doc, "workers")[1]
    assert u.is_valid_name("ray-demo1") and not u.is_valid_name("Bad_Name")
    assert Director().build_large_cluster("big")["spec"]["workerGroupSpecs"][0]["replicas"] == 8
    api = LocalClusterApi(gpu_count=4)
    m = GpuClusterManager(api)
    assert m.createRayCluster() == {"ok": True} and m.createRayCluster() == {"ok": False}        # duplicate
    assert m.modifyRayCluster({"k8s_namespace": "default", "ray_name": "olsb-demo1", "ray_label": "demo1-cluster",
                               "worker_min_replicas": 1, "worker_replicas": 8, "worker_max_replicas": 8}) == {"ok": True}
    st = api.get_ray_cluster("olsb-demo1")["status"]
    assert st == {"state": "pending", "desiredGpus": 8, "availableGpus": 4}                        # more GPUs than the box has
    assert m.modifyRayCluster({"k8s_namespace": "", "ray_name": "x", "ray_label": "y", "worker_min_replicas": 0,
                               "worker_replicas": 0, "worker_max_replicas": 0}) == {"ok": False}
    assert m.deleteRayCluster({"k8s_namespace": "default", "ray_label": "demo1-cluster"}) == {"ok": True}
    assert json.loads(m.queryRayCluster({"k8s_namespace": "default"})["json_data"]) == {"items": []}


# ---- blob stores / config / repos -----------------------------------------------------------------------------------
def test_local_file_repo_and_skip_if_exists(tmp_path):
    repo = LocalFileRepo(str(tmp_path / "store"), "bkt")
    src = tmp_path / "m.bin"
    src.write_bytes(b"model-v1")
    assert repo.upload_file(str(src), "t1/model_0.bin")
    assert repo.listdir_object_name("t1/") == ["t1/model_0.bin"]
    assert repo.download_payload("t1/model_0.bin") == b"model-v1"
    dst = tmp_path / "dl" / "m.bin"
    assert repo.download_if_missing("t1/model_0.bin", str(dst)) and dst.read_bytes() == b"model-v1"
    dst.write_bytes(b"local-edit")
    assert repo.download_if_missing("t1/model_0.bin", str(dst)) and dst.read_bytes() == b"local-edit"   # not re-downloaded
    assert repo.delete_file("t1/model_0.bin") and not repo.download_file("t1/model_0.bin", str(dst))
    with pytest.raises(ValueError):
        repo.upload_file(str(src), "../../escape")
    assert isinstance(get_file_repo("FILE", {"root": str(tmp_path)}), LocalFileRepo)
    # boto3 absent: the built-in Signature-V4 REST client takes over (exercised over HTTP further down)
    from olearning_sim_b200.utils.s3_rest import S3RestClient
    assert isinstance(get_file_repo("S3", {"endpoint_url": "http://x", "access_key": "a", "secret_key": "b", "bucket": "c"})._s3, S3RestClient)


def test_config_layering_and_reference_ini_compat(tmp_path, monkeypatch):
    ini = tmp_path / "config.conf"
    ini.write_text("[grpc]\nresmgr = '10.0.0.1:60062'\n[taskMgr]\nscheduler_sleep_time = 1\ninterrupt_queue_time = 60\n"
                   "[ray_cluster]\nworker_replicas = 2\n")
    y = tmp_path / "deviceflow_config.yaml"
    y.write_text("grpc_service_address: ${DF_HOST:-127.0.0.9}\nsource_topic: topic_x\n")
    monkeypatch.setenv("OLSB_TASKMGR__RELEASE_SLEEP_TIME", "3")
    cfg = load_config([str(ini), str(y)], overrides=["storage.task_table=/tmp/t.sqlite", "taskMgr.scheduler_sleep_time=0.5"])
    assert cfg.grpc.resmgr == "10.0.0.1:60062"                       # quotes stripped
    assert cfg.taskMgr.scheduler_sleep_time == 0.5 and cfg.taskMgr.interrupt_queue_time == 60.0
    assert cfg.taskMgr.release_sleep_time == 3.0                     # environment layer
    assert cfg.cluster.worker_replicas == 2                          # [ray_cluster] section accepted
    assert cfg.deviceflow.grpc_service_address == "127.0.0.9" and cfg.deviceflow.source_topic == "topic_x"
    assert cfg.storage.task_table == "/tmp/t.sqlite" and cfg.hybrid.alpha == 3.5


def test_sql_database_generic_api(tmp_path):
    db = SqlDataBase(str(tmp_path / "x.sqlite"), "deviceflow_table", ["task_id", "flow_id", "flow"])
    assert db.add_item({"task_id": ["t"], "flow_id": ["f1"]}) and db.add_item({"task_id": "t", "flow_id": "f2", "flow": "{}"})
    assert db.get_values_by_conditions("flow_id", task_id="t") == ["f1", "f2"]
    assert db.set_item_value("flow_id", "f1", "flow", '{"a":1}') and db.get_item_value("flow_id", "f1", "flow") == '{"a":1}'
    assert [r["flow_id"] for r in db.get_all_items_not_None("flow")] == ["f1", "f2"]
    assert db.delete_item("flow_id", "f1") and not db.check_item_in_database("flow_id", "f1")
    with pytest.raises(KeyError):
        db.get_item_value("flow_id", "f2", "no_such_column")
    with pytest.raises(KeyError):
        db.get_values_by_conditions("flow_id", **{"task_id; DROP TABLE x": 1})      # column names are whitelisted


def test_queue_and_fragment_repos(tmp_path):
    q = QueueRepo(str(tmp_path / "q.sqlite"))
    assert q.rpush("tasks", "a") == 1 and q.rpush("tasks", "b") == 2
    q2 = QueueRepo(str(tmp_path / "q.sqlite"))                                       # durable across instances
    assert q2.lpop("tasks") == "a" and q2.lpop("tasks") == "b" and q2.lpop("tasks") is None
    bus = InMemoryBus(virtual_time=True)
    for cls in (JsonFragmentRepo, ProtoFragmentRepo):
        repo = cls(bus, f"frag_{cls.__name__}")
        prod = bus.create_producer(f"frag_{cls.__name__}")
        for i in range(3):
            prod.send(repo.encode(Fragment("t", f"c{i}", 2, 10 + i, bytes([i, 255]), {"k": i})))
        got = repo.consume(10)
        assert [(f.client_id, f.round, f.num_samples, f.payload, f.meta) for f in got] == \
            [(f"c{i}", 2, 10 + i, bytes([i, 255]), {"k": i}) for i in range(3)]


def test_logger_four_field_signature_and_sinks(tmp_path):
    lg = Logger()
    repo = LogTableRepo()
    lg.set_logger(jsonl_path=str(tmp_path / "log.jsonl"), repo=repo, stderr_level="ERROR")
    lg.info(task_id="t1", system_name="TaskMgr", module_name="test", message="hello")
    lg.error("t1", "Deviceflow", "test", "boom")
    lg.flush()
    time.sleep(0.1)
    lines = [json.loads(l) for l in open(tmp_path / "log.jsonl")]
    assert [(l["log_type"], l["system_name"], l["message"]) for l in lines[-2:]] == [("INFO", "TaskMgr", "hello"), ("ERROR", "Deviceflow", "boom")]
    assert {r["message"] for r in repo.all_rows()} >= {"hello", "boom"}
    lg.set_logger(jsonl_path="", repo=None)
    lg.jsonl_path, lg.repo = None, None


def test_performance_manager_summary():
    pm = PerformanceManager({"hbm_gbs": 1000.0})
    for r in range(4):
        pm.commitJob({"task_id": "t", "round": r, "t_round_ms": 20.0, "client_steps": 100, "hbm_bytes": 5e9})
    s = pm.summary("t")
    assert s["rounds"] == 4 and s["rounds_per_s"] == pytest.approx(50.0) and s["client_steps_per_s"] == pytest.approx(5000.0)
    assert s["hbm_fraction_of_measured"] == pytest.approx(0.25)
    assert not pm.commitJob({})["is_success"]


# ---- checkpoints ----------------------------------------------------------------------------------------------------------
def test_checkpoint_atomic_keep_and_resume(tmp_path):
    cm = CheckpointManager(str(tmp_path), "taskA", rank=0, keep=2, asynchronous=True)
    for r in range(5):
        cm.save(r, {"round": r, "engine": {"global": torch.full((4,), float(r))}})
    cm.wait()
    names = sorted(os.listdir(tmp_path))
    assert names == ["taskA_3_rank0.ckpt", "taskA_4_rank0.ckpt"]          # {task_id}_{current_round} template, newest 2 kept
    st = cm.load_latest()
    assert st["round"] == 4 and torch.equal(st["engine"]["global"], torch.full((4,), 4.0))
    open(tmp_path / "taskA_5_rank0.ckpt", "wb").write(b"torn write")      # crash mid-save of round 5
    assert CheckpointManager(str(tmp_path), "taskA").load_latest()["round"] == 4
    assert CheckpointManager(str(tmp_path), "other").load_latest() is None


def test_engine_resume_continues_from_checkpoint(tmp_path):
    from olearning_sim_b200.control.run_task import RoundRunner
    tj = make_task_json(task_id="ck", rounds=4, operator_params={"model": "mlp2", "local_steps": 1, "batch": 8, "shard": 8,
                                                                 "lr": 0.1, "device": "cpu", "backend": "torch"})
    full = RoundRunner(make_task_json(task_id="ck_full", rounds=4, operator_params=json.loads(
        tj["operatorflow"]["operators"][0]["logical_simulation"]["operator_params"])))
    assert full.run()
    # interrupted after 2 rounds, then restarted with the same checkpoint dir
    ev = threading.Event()
    r1 = RoundRunner({**tj, "operatorflow": {**tj["operatorflow"], "flow_setting": {**tj["operatorflow"]["flow_setting"], "round": 2}}},
                     checkpoint_dir=str(tmp_path))
    assert r1.run()
    r2 = RoundRunner(tj, checkpoint_dir=str(tmp_path))
    assert r2.run()
    assert [h["round"] for h in r2.history] == [2, 3]                     # rounds 0-1 were not recomputed
    assert abs(r2.history[-1]["loss"] - full.history[-1]["loss"]) < 1e-5  # same trajectory as the uninterrupted run


# ---- operatorflow gates ---------------------------------------------------------------------------------------------------
def test_gate_waiting_for_global_aggregation_with_fake_clock():
    clock = {"t": 0, "round": 7}
    def sleep(s):
        clock["t"] += s
        if clock["t"] >= 6:
            clock["round"] = 8                                            # the aggregator advances after 6 s
    WaitingForGlobalAggregation.services["tg"] = lambda: clock["round"]
    p = {"strategy": "waiting_for_global_aggregation", "wait_interval": 2, "total_timeout": 10}
    of = OperatorFlow("tg", 1, p, p, sleep=sleep)
    assert of.start(0) and of.stop(0) and clock["t"] == 6
    clock.update(t=0, round=8)
    of2 = OperatorFlow("tg", 1, p, {**p, "total_timeout": 4}, sleep=lambda s: clock.update(t=clock["t"] + s))
    assert of2.start(0) and not of2.stop(0)                               # aggregator never advances -> timeout


def test_gate_waiting_for_global_aggregation_over_a_live_selection_service():
    """the reference's selection service is a WebSocket endpoint: the gate sends one REGISTER_MESSAGE per poll and reads
    ``data.round_idx`` from the first answer — here against a real ``websockets`` server whose round advances after the
    third question"""
    import threading
    from websockets.sync.server import serve
    asked = []

    def handler(ws):
        msg = json.loads(ws.recv())
        asked.append(msg)
        rnd = 4 if len(asked) <= 3 else 5
        ws.send(json.dumps({"header": {"messageType": "REGISTER_RESPONSE"}, "data": {"task_id": msg["data"]["task_id"], "round_idx": rnd}}))

    with serve(handler, "127.0.0.1", 0) as server:
        threading.Thread(target=server.serve_forever, daemon=True).start()
        url = f"ws://127.0.0.1:{server.socket.getsockname()[1]}/channel"
        p = {"strategy": "waiting_for_global_aggregation", "wait_interval": 1, "total_timeout": 5}
        of = OperatorFlow("tws", 1, p, p, gate_params={"service_url": url}, sleep=lambda s: None)
        assert of.start(0) and of.stop(0)                              # 4 at start, 5 (= previous + 1) after two more polls
        assert len(asked) >= 4 and all(m["header"]["messageType"] == "REGISTER_MESSAGE" and m["data"]["task_id"] == "tws"
                                       and m["header"]["clientId"] == m["data"]["client_id"] for m in asked)
        server.shutdown()
    dead = OperatorFlow("tws", 1, p, {**p, "total_timeout": 1}, gate_params={"service_url": "ws://127.0.0.1:9/none", "service_timeout": 0.2},
                        sleep=lambda s: None)
    assert not dead.start(0)                                           # unreachable service -> the gate times out, no exception


def test_gate_sample_and_aggregation_markers(tmp_path):
    src = tmp_path / "all_clients"
    src.mkdir()
    for i in range(30):
        (src / f"client_{i:03d}.bin").write_bytes(b"x")
    done = []
    def aggregator(round_idx, folder):
        done.append((round_idx, len([f for f in os.listdir(folder) if f.startswith("client_")])))
        open(os.path.join(folder, "aggregation_finished.txt"), "w").write("ok")
    p = {"strategy": "sample_and_aggregation", "wait_interval": 1, "total_timeout": 3}
    of = OperatorFlow("ts", 2, p, p, gate_params={"root": str(tmp_path), "sample_num": 10, "aggregator": aggregator},
                      sleep=lambda s: None)
    assert of.start(0) and of.stop(0) and done == [(0, 10)]
    assert os.path.exists(tmp_path / "round_0" / "simulation_finished.txt")
    p2 = {"strategy": "sample_dc_and_aggregation", "wait_interval": 1, "total_timeout": 1}
    of2 = OperatorFlow("ts", 1, p2, p2, gate_params={"root": str(tmp_path), "bundle_num": 3, "files_per_bundle": 4}, sleep=lambda s: None)
    assert of2.start(1) and sorted(os.listdir(tmp_path / "round_1")) == ["bundle_0", "bundle_1", "bundle_2"]
    assert not of2.stop(1)                                                # nobody wrote aggregation_finished.txt
    with pytest.raises(ValueError):
        OperatorFlow("t", 1, {"strategy": "nope"}, {"strategy": ""})


def test_oflserver_helpers_on_local_store(tmp_path):
    """utils_oflserver-style flat helpers over the local store: upload / list selection files / cached download / delete"""
    from olearning_sim_b200.utils import oflserver as o
    cfg = {"file": {"root": str(tmp_path / "store")}, "bus": {"topic": "gh", "url": "mem://x"}}
    for i in range(3):
        f = tmp_path / f"c{i}.json"
        f.write_text("{}")
        assert o.uploadS3File(cfg, str(f), f"sel/round0/c{i}.json")
    (tmp_path / "m.bin").write_bytes(b"abc")
    assert o.uploadS3File(cfg, str(tmp_path / "m.bin"), "sel/round0/model.bin")
    assert o.getS3SelectionFiles(cfg, "sel/round0") == [f"sel/round0/c{i}.json" for i in range(3)]
    dst = tmp_path / "dl" / "model.bin"
    assert o.downloadS3File(cfg, str(dst), "sel/round0/model.bin") and dst.read_bytes() == b"abc"
    assert o.deleteS3File(cfg, "sel/round0/model.bin")
    assert o.downloadS3File(cfg, str(dst), "sel/round0/model.bin")          # local copy exists -> no transfer needed
    assert o.uploadMinioFile(cfg, str(tmp_path / "m.bin"), "models/a.bin") and o.listdirMinioFile(cfg, "models") == ["models/a.bin"]
    assert o.get_gradient_house_pulsar(cfg) == {"url": "mem://x", "topic": "gh"}
    bus, prod = o.getPulsarObjects(cfg)
    c = bus.subscribe("gh", "t")
    prod.send(b"hello")
    assert c.receive(100).data() == b"hello"
    assert o.getSqlTaskRepo().check_task_in_database("nope") is False


def test_deviceflow_config_getters(tmp_path, monkeypatch):
    from olearning_sim_b200.deviceflow import config_parser as cp
    monkeypatch.setenv("DF_HOST", "10.1.2.3")
    f = tmp_path / "deviceflow_config.yaml"
    f.write_text("grpc_service_address: ${DF_HOST}\nport: 50099\nsource_topic: t_in\nshelfroom_namespace: ns/shelf\n")
    assert cp.get_grpc_service_address(str(f)) == "10.1.2.3" and cp.get_port(str(f)) == 50099
    assert cp.get_source_topic(str(f)) == "t_in" and cp.get_shelfroom_namespace(str(f)) == "ns/shelf"
    assert cp.get_max_workers(str(f)) == 10 and cp.get_shelfroom_url(str(f)) == "mem://local"     # typed defaults


def test_kube_cluster_api_with_fake_custom_objects_api():
    """the Kubernetes backend speaks the CustomObjectsApi verbs; driven here through GpuClusterManager with a fake API"""
    from olearning_sim_b200.cluster.manager import GpuClusterManager, KubeClusterApi

    class FakeCOA:
        def __init__(self): self.objs = {}
        def create_namespaced_custom_object(self, group, version, namespace, plural, body):
            key = (namespace, body["metadata"]["name"])
            if key in self.objs: raise RuntimeError("409 AlreadyExists")
            self.objs[key] = body; return body
        def get_namespaced_custom_object(self, group, version, namespace, plural, name):
            return self.objs[(namespace, name)]
        def list_namespaced_custom_object(self, group, version, namespace, plural, label_selector=""):
            return {"items": [o for (ns, _), o in self.objs.items() if ns == namespace]}
        def patch_namespaced_custom_object(self, group, version, namespace, plural, name, body):
            self.objs[(namespace, name)] = body; return body
        def delete_namespaced_custom_object(self, group, version, namespace, plural, name):
            return self.objs.pop((namespace, name))

    api = KubeClusterApi(api=FakeCOA())
    mgr = GpuClusterManager(api=api)
    assert mgr.createRayCluster({})["ok"]
    assert not mgr.createRayCluster({})["ok"]                          # duplicate name -> API error -> reported, not raised
    d = mgr._d
    assert api.get_ray_cluster(d.ray_name, d.k8s_namespace)["spec"]["workerGroupSpecs"]
    assert len(api.list_ray_clusters(d.k8s_namespace)["items"]) == 1
    assert mgr.modifyRayCluster({"k8s_namespace": d.k8s_namespace, "ray_name": d.ray_name, "ray_label": d.ray_label,
                                 "worker_min_replicas": 1, "worker_replicas": 2, "worker_max_replicas": 4})["ok"]
    assert api.delete_ray_cluster("nope") is False and api.delete_ray_cluster(d.ray_name, d.k8s_namespace)


def test_phone_manager_service_and_remote_adapter():
    """the phone-side services over gRPC: a ResourceManager in another process can use them through RemotePhoneMgr"""
    from olearning_sim_b200.api.grpc_service import JsonStub, RemotePhoneMgr, SimulatorSession, connect
    from olearning_sim_b200.control.repo import ResTableRepo
    from olearning_sim_b200.control.resource_manager import FakePhoneMgr, ResourceManager
    pm = FakePhoneMgr({"u1": {"High": 2}})
    sess = SimulatorSession(port=0, phone_manager=pm, svc=6).start()
    try:
        assert sess.mounted == ["TaskManager(phone)", "PhoneManager"]
        assert JsonStub(connect(sess.address), "PhoneManager").connectMSPDevices({"user_id": "u1", "devices": {"Low": 3}})["is_success"]
        remote = RemotePhoneMgr(sess.address)
        assert remote.getDeviceAvailableResource() == {"u1": {"High": 2, "Low": 3}}
        rm = ResourceManager(ResTableRepo(), totals={"cpu": 4, "mem": 8, "gpu": 0}, phone_mgr=remote)
        assert rm.requestResource("t1", "u1", cpu=1, mem=1, phones={"High": 2}, type=0)[0]
        assert remote.getDeviceAvailableResource()["u1"]["High"] == 0
        assert not remote.requestDeviceResource("t2", "u1", {"High": 1})
        assert rm.releaseResource("t1", "u1", phones={"High": 2}, type=0)[0] and pm.getDeviceAvailableResource()["u1"]["High"] == 2
        assert remote.submitTask({"task_id": "t1"}) and remote.getDeviceTaskStatus("t1")["is_finished"] is False
        assert remote.stopDevice("t1") and remote.getDeviceTaskStatus("t1")["is_finished"] is True
    finally:
        sess.stop()


def test_prometheus_metrics_from_round_records():
    from olearning_sim_b200.utils.metrics import Metrics
    from olearning_sim_b200.utils.misc_repos import PerformanceManager
    m = Metrics()
    pm = PerformanceManager(metrics=m)
    for r in range(3):
        pm.commitJob({"task_id": "t9", "round": r, "t_round_ms": 20.0, "client_steps": 5120,
                      "phases_ms": {"broadcast": 0.05, "local_steps": 19.0, "reduce_update": 0.2}})
    m.observe_tasks({"RUNNING": 1, "QUEUED": 2})
    text = m.render()
    assert 'olsb_rounds_total{task_id="t9"} 3.0' in text and 'olsb_client_steps_total{task_id="t9"} 15360.0' in text
    assert 'olsb_phase_ms{phase="local_steps",task_id="t9"} 19.0' in text and 'olsb_tasks{status="QUEUED"} 2.0' in text
    assert pm.summary("t9")["rounds"] == 3


def test_optimizer_state_reshards_across_world_sizes():
    """checkpoints hold the full server-optimiser vectors; a restart on fewer ranks re-slices them"""
    import torch
    from olearning_sim_b200.engine.fed_engine import merge_opt_shards, take_opt_shard
    P = 22
    full = torch.arange(P, dtype=torch.float32)
    def shards(world):
        sl = ((P // 4 + world - 1) // world) * 4
        return sl, [take_opt_shard(full, r, sl, sl) for r in range(world)]
    sl4, s4 = shards(4)                      # shard_len 8: ranks 0..2 hold 8/8/6 values, rank 3 holds none
    assert sl4 == 8 and s4[2].tolist() == [16, 17, 18, 19, 20, 21, 0, 0] and s4[3].sum() == 0
    assert torch.equal(merge_opt_shards(s4, P), full)
    sl3, s3 = shards(3)                      # elastic shrink 4 -> 3 ranks
    assert torch.equal(merge_opt_shards(s3, P), full) and s3[1][0] == sl3
    assert torch.equal(take_opt_shard(merge_opt_shards(s4, P), 0, 24, 24)[:P], full)       # ... -> 1 rank


def test_kube_cluster_api_over_http_against_an_api_server():
    """KubeClusterApi with its built-in REST transport against an in-process HTTP server that implements the custom-objects
    endpoints of a Kubernetes API server (create 201 / duplicate 409 / get / list with labelSelector / merge-patch / delete /
    404), including bearer-token auth — the whole GpuClusterManager CRUD cycle goes over real HTTP"""
    import http.server
    import threading
    import urllib.parse
    from olearning_sim_b200.cluster.manager import GROUP, PLURAL, VERSION, GpuClusterManager, KubeClusterApi, KubeRestCustomObjects

    store, seen = {}, {"auth": set(), "ctype": set()}

    class H(http.server.BaseHTTPRequestHandler):
        def log_message(self, *a): pass

        def _send(self, code, obj):
            b = json.dumps(obj).encode()
            self.send_response(code); self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(b))); self.end_headers(); self.wfile.write(b)

        def _route(self):
            u = urllib.parse.urlparse(self.path)
            parts = u.path.strip("/").split("/")
            if parts[:3] != ["apis", GROUP, VERSION] or parts[3] != "namespaces" or parts[5] != PLURAL:
                return None
            return parts[4], (parts[6] if len(parts) > 6 else None), urllib.parse.parse_qs(u.query)

        def _handle(self, method):
            seen["auth"].add(self.headers.get("Authorization"))
            if self.headers.get("Authorization") != "Bearer sekrit":
                return self._send(401, {"kind": "Status", "message": "Unauthorized"})
            r = self._route()
            if r is None:
                return self._send(404, {"kind": "Status", "message": "no such resource"})
            ns, name, q = r
            n = int(self.headers.get("Content-Length") or 0)
            body = json.loads(self.rfile.read(n)) if n else None
            if method == "POST":
                key = (ns, body["metadata"]["name"])
                if key in store:
                    return self._send(409, {"kind": "Status", "message": f"{PLURAL} {key[1]!r} already exists"})
                store[key] = body
                return self._send(201, body)
            if method == "GET" and name is None:
                sel = dict(kv.split("=") for kv in q.get("labelSelector", [""])[0].split(",") if kv)
                items = [o for (n_, _), o in store.items() if n_ == ns and all(o["metadata"].get("labels", {}).get(k) == v for k, v in sel.items())]
                return self._send(200, {"kind": "List", "items": items})
            if (ns, name) not in store:
                return self._send(404, {"kind": "Status", "message": f"{PLURAL} {name!r} not found"})
            if method == "GET":
                return self._send(200, store[(ns, name)])
            if method == "PATCH":
                seen["ctype"].add(self.headers.get("Content-Type"))
                def merge(a, b):
                    for k, v in b.items():
                        if isinstance(v, dict) and isinstance(a.get(k), dict): merge(a[k], v)
                        else: a[k] = v
                merge(store[(ns, name)], body)
                return self._send(200, store[(ns, name)])
            if method == "DELETE":
                return self._send(200, {"kind": "Status", "status": "Success", "details": {"name": store.pop((ns, name))["metadata"]["name"]}})

        def do_GET(self): self._handle("GET")
        def do_POST(self): self._handle("POST")
        def do_PATCH(self): self._handle("PATCH")
        def do_DELETE(self): self._handle("DELETE")

    srv = http.server.ThreadingHTTPServer(("127.0.0.1", 0), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        url = f"http://127.0.0.1:{srv.server_address[1]}"
        api = KubeClusterApi(api=KubeRestCustomObjects(url, token="sekrit"))
        mgr = GpuClusterManager(api=api)
        d = mgr._d
        assert mgr.createRayCluster({})["ok"] and (d.k8s_namespace, d.ray_name) in store
        assert not mgr.createRayCluster({})["ok"]                                    # 409 from the server -> reported failure
        got = api.get_ray_cluster(d.ray_name, d.k8s_namespace)
        assert got["kind"] and got["spec"]["workerGroupSpecs"][0]["replicas"] == d.worker_replicas
        assert len(api.list_ray_clusters(d.k8s_namespace)["items"]) == 1
        assert api.list_ray_clusters(d.k8s_namespace, label_selector="no=match")["items"] == []
        assert mgr.modifyRayCluster({"k8s_namespace": d.k8s_namespace, "ray_name": d.ray_name, "ray_label": d.ray_label,
                                     "worker_min_replicas": 1, "worker_replicas": 2, "worker_max_replicas": 4})["ok"]
        assert store[(d.k8s_namespace, d.ray_name)]["spec"]["workerGroupSpecs"][0]["replicas"] == 2
        assert seen["ctype"] == {"application/merge-patch+json"}
        assert api.get_ray_cluster("nope", d.k8s_namespace) is None                  # 404 -> None, not an exception
        assert mgr.deleteRayCluster({"k8s_namespace": d.k8s_namespace, "ray_name": d.ray_name})["ok"] and not store
        assert api.delete_ray_cluster(d.ray_name, d.k8s_namespace) is False
        bad = KubeClusterApi(api=KubeRestCustomObjects(url, token="wrong"))
        assert bad.get_ray_cluster("x") is None and bad.list_ray_clusters()["items"] == []       # 401 -> failure, no raise
        assert seen["auth"] == {"Bearer sekrit", "Bearer wrong"}
    finally:
        srv.shutdown()


def test_s3_and_minio_repos_over_http_with_signature_v4(tmp_path):
    """S3FileRepo / MinioFileRepo on the built-in REST client against an in-process S3-compatible HTTP server that
    re-derives the Signature-V4 of every request from the shared secret (wrong secret -> 403 SignatureDoesNotMatch), stores
    objects, pages ListObjectsV2 (2 keys per page) and answers errors as S3 XML.  The signer itself is pinned to the AWS
    documentation test vector."""
    import hashlib
    import http.server
    import threading
    import urllib.parse
    from olearning_sim_b200.utils import s3_rest
    from olearning_sim_b200.utils.file_repo import MinioFileRepo, S3FileRepo

    got, want = s3_rest.known_answer()
    assert got == want                                              # published AWS example signature

    AK, SK, store = "AKIDEXAMPLE", "s3cr3t/Key+1", {}

    class H(http.server.BaseHTTPRequestHandler):
        def log_message(self, *a): pass

        def _xml(self, code, body):
            b = body.encode()
            self.send_response(code); self.send_header("Content-Type", "application/xml")
            self.send_header("Content-Length", str(len(b))); self.end_headers(); self.wfile.write(b)

        def _err(self, code, name, msg):
            self._xml(code, f"<?xml version='1.0'?><Error><Code>{name}</Code><Message>{msg}</Message></Error>")

        def _handle(self, method):
            u = urllib.parse.urlparse(self.path)
            n = int(self.headers.get("Content-Length") or 0)
            body = self.rfile.read(n) if n else b""
            query = {k: v[0] for k, v in urllib.parse.parse_qs(u.query, keep_blank_values=True).items()}
            # ---- verify the signature exactly as a server does: rebuild the canonical request from what arrived
            auth = self.headers.get("Authorization", "")
            try:
                cred = auth.split("Credential=")[1].split(",")[0]
                signed = auth.split("SignedHeaders=")[1].split(",")[0].split(";")
                sig = auth.split("Signature=")[1]
            except IndexError:
                return self._err(403, "AccessDenied", "no signature")
            hdrs = {h: self.headers.get(h) for h in signed}
            ph = self.headers.get("x-amz-content-sha256")
            if ph != hashlib.sha256(body).hexdigest():
                return self._err(400, "XAmzContentSHA256Mismatch", "payload hash")
            canon, _ = s3_rest.canonical_request(method, urllib.parse.unquote(u.path), query, hdrs, ph)
            exp, scope = s3_rest.signature_v4(SK, self.headers["x-amz-date"], cred.split("/")[2], "s3", canon)
            if cred != f"{AK}/{scope}" or sig != exp:
                return self._err(403, "SignatureDoesNotMatch", "signature mismatch")
            parts = urllib.parse.unquote(u.path).lstrip("/").split("/", 1)
            bucket, key = parts[0], (parts[1] if len(parts) > 1 else "")
            if bucket != "models":
                return self._err(404, "NoSuchBucket", bucket)
            if method == "PUT":
                store[key] = body
                return self._xml(200, "")
            if method == "GET" and not key:                          # ListObjectsV2, two keys per page
                keys = sorted(k for k in store if k.startswith(query.get("prefix", "")))
                start = int(query.get("continuation-token", "0"))
                page = keys[start:start + 2]
                more = start + 2 < len(keys)
                items = "".join(f"<Contents><Key>{k}</Key><Size>{len(store[k])}</Size></Contents>" for k in page)
                nxt = f"<NextContinuationToken>{start + 2}</NextContinuationToken>" if more else ""
                return self._xml(200, "<?xml version='1.0'?><ListBucketResult xmlns='http://s3.amazonaws.com/doc/2006-03-01/'>"
                                      f"<IsTruncated>{'true' if more else 'false'}</IsTruncated>{items}{nxt}</ListBucketResult>")
            if key not in store:
                return self._err(404, "NoSuchKey", key)
            if method == "GET":
                self.send_response(200); self.send_header("Content-Length", str(len(store[key]))); self.end_headers()
                return self.wfile.write(store[key])
            if method == "DELETE":
                del store[key]
                self.send_response(204); self.send_header("Content-Length", "0"); return self.end_headers()

        def do_GET(self): self._handle("GET")
        def do_PUT(self): self._handle("PUT")
        def do_DELETE(self): self._handle("DELETE")

    srv = http.server.ThreadingHTTPServer(("127.0.0.1", 0), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        ep = f"http://127.0.0.1:{srv.server_address[1]}"
        s3 = S3FileRepo(ep, AK, SK, "models")
        src = tmp_path / "m.bin"
        for i in range(5):
            src.write_bytes(b"model-%d " % i * (1000 + i))
            assert s3.upload_file(str(src), f"task 1/round_{i}/model+v{i}.bin")          # space and '+' in the key: URI encoding
        assert s3.listdir_object_name("task 1/") == sorted(f"task 1/round_{i}/model+v{i}.bin" for i in range(5))   # 3 pages
        dst = tmp_path / "out" / "m.bin"
        assert s3.download_file("task 1/round_3/model+v3.bin", str(dst)) and dst.read_bytes() == b"model-3 " * 1003
        assert s3.download_payload("task 1/round_0/model+v0.bin") == b"model-0 " * 1000
        assert not s3.download_file("task 1/nope", str(tmp_path / "nope")) and not (tmp_path / "nope").exists()
        assert s3.delete_file("task 1/round_0/model+v0.bin") and len(s3.listdir_object_name("task 1/")) == 4
        # the MinIO adapter speaks the same protocol through the same client
        mn = MinioFileRepo(ep.replace("http://", ""), AK, SK, "models")
        assert mn.upload_file(str(src), "minio/x.bin") and mn.listdir_object_name("minio/") == ["minio/x.bin"]
        assert mn.download_file("minio/x.bin", str(tmp_path / "x.bin")) and (tmp_path / "x.bin").read_bytes() == src.read_bytes()
        assert mn.delete_file("minio/x.bin") and mn.listdir_object_name("minio/") == []
        # wrong secret: every call is refused by the server and reported as a failure, not raised
        bad = S3FileRepo(ep, AK, "wrong", "models")
        assert not bad.upload_file(str(src), "k") and bad.download_payload("task 1/round_1/model+v1.bin") is None
        with pytest.raises(s3_rest.S3Error) as ei:
            s3_rest.S3RestClient(ep, AK, "wrong").get_object("models", "k")
        assert ei.value.status == 403 and ei.value.code == "SignatureDoesNotMatch"
    finally:
        srv.shutdown()


def test_redis_repo_over_the_resp_protocol():
    """RedisRepo(url=...) against an in-process RESP2 server over TCP: AUTH, SELECT, RPUSH / LPOP / LLEN, nil replies and
    error replies; without a URL it stays the in-process queue"""
    import socketserver
    import threading
    from olearning_sim_b200.utils.misc_repos import RedisRepo, RespClient

    lists, seen = {}, []

    class H(socketserver.StreamRequestHandler):
        def handle(self):
            authed = False
            while True:
                line = self.rfile.readline()
                if not line:
                    return
                n = int(line[1:])
                args = []
                for _ in range(n):
                    ln = int(self.rfile.readline()[1:])
                    args.append(self.rfile.read(ln + 2)[:-2].decode())
                cmd = args[0].upper()
                seen.append(cmd)
                if cmd == "AUTH":
                    authed = args[1] == "hunter2"
                    self.wfile.write(b"+OK\r\n" if authed else b"-WRONGPASS invalid password\r\n")
                elif not authed:
                    self.wfile.write(b"-NOAUTH Authentication required.\r\n")
                elif cmd == "SELECT":
                    self.wfile.write(b"+OK\r\n")
                elif cmd == "RPUSH":
                    lists.setdefault(args[1], []).append(args[2])
                    self.wfile.write(b":%d\r\n" % len(lists[args[1]]))
                elif cmd == "LPOP":
                    q = lists.get(args[1]) or []
                    if q:
                        v = q.pop(0).encode()
                        self.wfile.write(b"$%d\r\n%s\r\n" % (len(v), v))
                    else:
                        self.wfile.write(b"$-1\r\n")
                elif cmd == "LLEN":
                    self.wfile.write(b":%d\r\n" % len(lists.get(args[1], [])))
                else:
                    self.wfile.write(b"-ERR unknown command\r\n")

    srv = socketserver.ThreadingTCPServer(("127.0.0.1", 0), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        port = srv.server_address[1]
        r = RedisRepo(url=f"redis://:hunter2@127.0.0.1:{port}/2")
        assert r.rpush("q", "a\r\nwith crlf") == 1 and r.rpush("q", "b") == 2 and r.llen("q") == 2
        assert r.lpop("q") == "a\r\nwith crlf" and r.lpop("q") == "b" and r.lpop("q") is None and r.llen("q") == 0
        assert seen[:2] == ["AUTH", "SELECT"]
        with pytest.raises(RuntimeError):
            RespClient("127.0.0.1", port, password="wrong")
        with pytest.raises(RuntimeError):
            RespClient("127.0.0.1", port).command("LLEN", "q")          # NOAUTH
        with pytest.raises(RuntimeError):
            r._r.command("FLUSHALL")                                     # error reply surfaces as an exception
        local = RedisRepo()
        assert local.rpush("q", "x") == 1 and local.lpop("q") == "x" and local.lpop("q") is None
    finally:
        srv.shutdown()
