"""BASELINE.json config #1: FedAvg 2-layer MLP, 8 clients, 1 local epoch on CPU, world_size=2 (gloo), driven through
the real TaskManager / ResourceManager / TaskRunner / RoundRunner objects:  JSON task -> submitTask -> schedule ->
freeze -> run rounds -> getTaskStatus == SUCCEEDED with logical_result populated."""
import json
import os
import sys
import time

import pytest
import torch

from olearning_sim_b200.api.schema import TaskStatus
from olearning_sim_b200.control.jobs import InProcessJobExecutor, SubprocessJobExecutor
from olearning_sim_b200.control.repo import FlowTableRepo, ResTableRepo, TaskTableRepo
from olearning_sim_b200.control.resource_manager import FakePhoneMgr, ResourceManager
from olearning_sim_b200.control.task_manager import TaskManager
from olearning_sim_b200.deviceflow.bus import FakeClock, InMemoryBus
from olearning_sim_b200.deviceflow.service import DeviceFlowService
from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine
from tests.helpers import REAL_TIME, make_task_json

MLP = {"model": "mlp2", "local_steps": 2, "batch": 16, "shard": 32, "lr": 0.1, "device": "cpu", "backend": "torch"}


def test_config1_in_process_single_rank():
    repo = TaskTableRepo()
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 8, "mem": 32, "gpu": 0}, phone_mgr=FakePhoneMgr())
    tm = TaskManager(repo=repo, resource_manager=rm, executor=InProcessJobExecutor())
    assert tm.submitTask(make_task_json(task_id="cfg1", rounds=3, nums=(8,), operator_params=MLP))
    assert tm.schedule_once(inline=True) == "cfg1"
    assert tm.wait_for("cfg1", 120) == TaskStatus.SUCCEEDED, tm._executor.get_job_error(repo.get_task_value("cfg1", "job_id"))
    res = json.loads(repo.get_task_value("cfg1", "logical_result"))["logical_result"]
    assert res == [{"name": "data_0", "simulation_target": {"devices": ["High"], "success_num": [8], "failed_num": [0]}}]
    assert repo.get_task_value("cfg1", "logical_round") == 3 and repo.get_task_value("cfg1", "logical_operator") == "training"
    assert tm.reap_once() == ["cfg1"] and rm.getClusterAvailableResource()["cpu"] == 8


def test_config1_world_size_2_gloo_subprocess(tmp_path):
    db = str(tmp_path / "tasks.sqlite")
    repo = TaskTableRepo(db)
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 8, "mem": 32, "gpu": 0}, phone_mgr=FakePhoneMgr())
    ex = SubprocessJobExecutor(nproc_per_node=2, workdir=str(tmp_path / "jobs"), repo_path=db, master_port_base=29741)
    tm = TaskManager(repo=repo, resource_manager=rm, executor=ex, job_context={})
    tm._job_context.pop("repo", None); tm._job_context.pop("deviceflow", None)     # not picklable across processes
    assert tm.submitTask(make_task_json(task_id="cfg1w2", rounds=2, nums=(8,), operator_params=MLP, num_request=(2,)))
    assert tm.schedule_once(inline=True) == "cfg1w2"
    st = tm.wait_for("cfg1w2", 240, poll=0.5)
    job_id = repo.get_task_value("cfg1w2", "job_id")
    assert st == TaskStatus.SUCCEEDED, ex.get_job_error(job_id)
    res = json.loads(repo.get_task_value("cfg1w2", "logical_result"))["logical_result"]
    assert res[0]["simulation_target"]["success_num"] == [8]
    log = ex.get_job_error(job_id)
    assert '"ok": true' in log


def test_two_rank_result_equals_single_rank(tmp_path):
    """client sharding must not change the math: world_size=2 (gloo, spawned) == world_size=1"""
    import torch.multiprocessing as mp
    cfg = dict(model="mlp2", num_clients=6, local_steps=2, batch=8, shard=16, lr=0.1, device="cpu", backend="torch",
               p_online=0.8, drop_probability=0.2, hetero_steps=True, seed=5)
    e = FedEngine(FedConfig(**cfg)); e.load_synthetic()
    for r in range(2):
        r1 = e.run_round(r)
    g1 = e.global_model().clone()
    out = str(tmp_path / "g.pt")
    mp.spawn(_rank_main, args=(2, cfg, out, 29755), nprocs=2, join=True)
    g2, stats = torch.load(out)
    assert torch.allclose(g1, g2, atol=1e-5), float((g1 - g2).abs().max())
    assert stats == (r1.participants, r1.dropped, r1.arrived, r1.client_steps)


def _rank_main(rank, world, cfg, out, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e = FedEngine(FedConfig(**cfg), rank, world); e.load_synthetic()
    for r in range(2):
        res = e.run_round(r)
    if rank == 0:
        torch.save((e.global_model(), (res.participants, res.dropped, res.arrived, res.client_steps)), out)
    dist.destroy_process_group()


def test_task_with_deviceflow_real_time_strategy():
    """operator with use_gradient_house: RegisterTask -> NotifyStart/Complete per (operator, round) -> flows finish"""
    bus = InMemoryBus(virtual_time=True)
    df = DeviceFlowService(bus=bus, repo=FlowTableRepo(), clock=FakeClock(), seed=0, dispatcher_timeout_millis=0)
    repo = TaskTableRepo()
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 8, "mem": 32, "gpu": 0}, phone_mgr=FakePhoneMgr())
    tm = TaskManager(repo=repo, resource_manager=rm, deviceflow=df, executor=InProcessJobExecutor())
    tj = make_task_json(task_id="dft", rounds=2, operator_params=MLP, use_df=True, strategy=REAL_TIME,
                        outbound={"pulsar_client": {"url": "mem://local", "topic": "agg"}})
    assert tm.submitTask(tj)
    tm.schedule_once(inline=True)
    assert tm.wait_for("dft", 120) == TaskStatus.SUCCEEDED
    assert df.stats()["registered"] == ["dft"]
    # built-in operators apply the strategy on the device and report the dispatch counts: each (operator, round) flow is
    # closed at NotifyComplete with the numbers the dispatcher would have produced
    done = df.stats()["device_dispatched"]
    assert sorted(done) == ["dft_training_0", "dft_training_1"] and df.stats()["flows"] == []
    assert all(v["participants"] == 8 and v["forwarded"] + v["dropped"] + v["late"] == 8 for v in done.values())
    df.pump(); df.pump()
    assert df.CheckDeviceflowDispatchFinished("dft")
    assert tm.reap_once() == ["dft"]
    assert df.stats()["registered"] == []


def test_script_operator_reference_contract(tmp_path):
    """legacy operator: one python subprocess per simulated phone, exit status drives success/failed"""
    op = tmp_path / "ops" / "training"
    op.mkdir(parents=True)
    (op / "train.py").write_text(
        "import json, sys\n"
        "p = json.load(open(sys.argv[sys.argv.index('--params') + 1]))\n"
        "assert p['operator_name'] == 'training' and p['data']['name'] == 'data_0'\n"
        "sys.exit(1 if p['split_index'] == 3 else 0)\n")
    repo = TaskTableRepo()
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 8, "mem": 32, "gpu": 0}, phone_mgr=FakePhoneMgr())
    tm = TaskManager(repo=repo, resource_manager=rm, executor=InProcessJobExecutor())
    tj = make_task_json(task_id="script", rounds=1, nums=(4, 2), devices=("High", "Low"), dynamic=(1, 0), logical=(4, 2),
                        num_request=(2, 1), code_path=str(tmp_path / "ops" / "training"), operator_params={})
    assert tm.submitTask(tj)
    tm.schedule_once(inline=True)
    assert tm.wait_for("script", 120) == TaskStatus.SUCCEEDED          # 1 failure tolerated by dynamic_nums
    res = json.loads(repo.get_task_value("script", "logical_result"))["logical_result"][0]["simulation_target"]
    assert res == {"devices": ["High", "Low"], "success_num": [3, 2], "failed_num": [1, 0]}
    # same task with no tolerance -> FAILED early
    tj2 = make_task_json(task_id="script2", rounds=1, nums=(4, 2), devices=("High", "Low"), dynamic=(0, 0), logical=(4, 2),
                         num_request=(2, 1), code_path=str(tmp_path / "ops" / "training"), operator_params={})
    assert tm.submitTask(tj2)
    tm.schedule_once(inline=True)
    assert tm.wait_for("script2", 120) == TaskStatus.FAILED


def test_poisoned_client_is_masked_and_counted(monkeypatch):
    """fault injection: a client whose replica turns NaN must not poison the aggregate and counts as failed"""
    monkeypatch.setenv("OLSB_FAULTS", "corrupt:client=3,round=1")
    e = FedEngine(FedConfig(model="mlp2", num_clients=6, local_steps=1, batch=8, shard=8, lr=0.1, device="cpu", backend="torch"))
    e.load_synthetic()
    res = [e.run_round(r) for r in range(3)]
    assert [r.failed for r in res] == [0, 1, 0]
    assert all(r.loss == r.loss for r in res) and bool(torch.isfinite(e.global_model()).all())


def test_elastic_restart_after_rank_loss(tmp_path):
    """kill rank 1 in round 1 of a 2-rank job: the supervisor relaunches on 1 rank from the round-0 checkpoint and the
    task still SUCCEEDS with every client accounted for (clients are re-sharded by world size)."""
    db = str(tmp_path / "tasks.sqlite")
    repo = TaskTableRepo(db)
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 8, "mem": 32, "gpu": 0}, phone_mgr=FakePhoneMgr())
    ex = SubprocessJobExecutor(nproc_per_node=2, workdir=str(tmp_path / "jobs"), repo_path=db, master_port_base=29781,
                               max_restarts=1, shrink_on_restart=True,
                               env={"OLSB_FAULTS": "kill:rank=1,round=1", "OLSB_FAULT_MARKER": str(tmp_path / "killed")})
    tm = TaskManager(repo=repo, resource_manager=rm, executor=ex, job_context={"checkpoint_dir": str(tmp_path / "ckpt")})
    tm._job_context.pop("repo", None); tm._job_context.pop("deviceflow", None)
    assert tm.submitTask(make_task_json(task_id="elastic", rounds=3, nums=(8,), operator_params=MLP, num_request=(2,)))
    tm.schedule_once(inline=True)
    st = tm.wait_for("elastic", 300, poll=0.5)
    job_id = repo.get_task_value("elastic", "job_id")
    log = ex.get_job_error(job_id)
    assert st == TaskStatus.SUCCEEDED, log
    assert os.path.exists(tmp_path / "killed") and "attempt 1, 1 rank(s)" in log
    assert repo.get_task_value("elastic", "logical_round") == 3
    assert sorted(os.listdir(tmp_path / "ckpt"))[-1].startswith("elastic_2_")


def test_subprocess_job_reaches_deviceflow_over_grpc(tmp_path):
    """the product path: the task runs as a 2-rank process group (gloo) while DeviceFlow lives in the control-plane process —
    the job finds it through ``deviceflow_address`` (set by the gRPC session) and drives NotifyStart / ReportDeviceDispatch /
    NotifyComplete through ``RemoteDeviceFlow``, so the service closes every (operator, round) flow"""
    from olearning_sim_b200.api.grpc_service import RemoteDeviceFlow, SimulatorSession
    db = str(tmp_path / "tasks.sqlite")
    repo = TaskTableRepo(db)
    bus = InMemoryBus(virtual_time=True)
    df = DeviceFlowService(bus=bus, repo=FlowTableRepo(), clock=FakeClock(), seed=0, dispatcher_timeout_millis=0)
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 8, "mem": 32, "gpu": 0}, phone_mgr=FakePhoneMgr())
    ex = SubprocessJobExecutor(nproc_per_node=2, workdir=str(tmp_path / "jobs"), repo_path=db, master_port_base=29761)
    tm = TaskManager(repo=repo, resource_manager=rm, deviceflow=df, executor=ex)
    sess = SimulatorSession(port=0, task_manager=tm, resource_manager=rm, deviceflow=df).start()
    try:
        assert tm._job_context["deviceflow_address"] == sess.address
        # the adapter alone, over the wire
        rdf = RemoteDeviceFlow(sess.address)
        assert rdf.RegisterTask("probe", ["logical_simulation"]) and rdf.GetDeviceflowPulsarClient() == df.GetDeviceflowPulsarClient()
        assert rdf.UnRegisterTask("probe")
        tj = make_task_json(task_id="dfsub", rounds=2, nums=(8,), operator_params=MLP, num_request=(2,), use_df=True, strategy=REAL_TIME,
                            outbound={"pulsar_client": {"url": "mem://local", "topic": "agg"}})
        assert tm.submitTask(tj)
        assert tm.schedule_once(inline=True) == "dfsub"
        st = tm.wait_for("dfsub", 240, poll=0.5)
        assert st == TaskStatus.SUCCEEDED, ex.get_job_error(repo.get_task_value("dfsub", "job_id"))
        done = df.stats()["device_dispatched"]
        assert sorted(done) == ["dfsub_training_0", "dfsub_training_1"] and df.stats()["flows"] == []
        assert all(v["participants"] == 8 for v in done.values())
        df.pump(); df.pump()
        assert df.CheckDeviceflowDispatchFinished("dfsub")
    finally:
        sess.stop()


def test_split_deployment_task_service_with_remote_resource_manager_and_deviceflow():
    """the reference's topology: ResourceMgr (+ phone manager) in one server, DeviceFlow in another, the TaskMgr in a third
    party that only holds gRPC adapters.  A task submitted to the TaskMgr books and releases resources in the remote
    ledger and registers / closes its flows in the remote DeviceFlow"""
    from olearning_sim_b200 import build_task_service
    from olearning_sim_b200.api.grpc_service import SimulatorSession
    res_repo = ResTableRepo()
    rm = ResourceManager(res_repo, totals={"cpu": 8, "mem": 32, "gpu": 0}, phone_mgr=FakePhoneMgr())
    bus = InMemoryBus(virtual_time=True)
    df = DeviceFlowService(bus=bus, repo=FlowTableRepo(), clock=FakeClock(), seed=0, dispatcher_timeout_millis=0)
    s_res = SimulatorSession(port=0, resource_manager=rm, phone_manager=rm.phone_mgr, svc=0).start()      # ResourceMgr + phone services
    s_df = SimulatorSession(port=0, deviceflow=df, svc=5).start()
    try:
        assert "ResourceMgr" in s_res.mounted and s_df.mounted == ["TaskOperatorOrientedDeviceFlow"]
        tm, rrm, rdf = build_task_service(resmgr_address=s_res.address, phonemgr_address=s_res.address, deviceflow_address=s_df.address)
        assert rrm.getResource()["cluster"]["cpu"] == 8
        held = []
        orig = rm.requestClusterResource
        rm.requestClusterResource = lambda *a, **k: (held.append(a[0]), orig(*a, **k))[1]                   # observe the remote ledger
        tj = make_task_json(task_id="split", rounds=2, operator_params=MLP, use_df=True, strategy=REAL_TIME,
                            outbound={"pulsar_client": {"url": "mem://local", "topic": "agg"}})
        assert tm.submitTask(tj)
        tm.schedule_once(inline=True)
        assert tm.wait_for("split", 120) == TaskStatus.SUCCEEDED
        assert held == ["split"] and res_repo.held()                                  # booked in the remote ledger ...
        assert sorted(df.stats()["device_dispatched"]) == ["split_training_0", "split_training_1"]
        df.pump(); df.pump()
        assert tm.reap_once() == ["split"]
        assert res_repo.held() == [] and df.stats()["registered"] == []              # ... released and unregistered remotely
    finally:
        s_res.stop(); s_df.stop()
