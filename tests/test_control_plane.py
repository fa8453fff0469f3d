"""Scheduler / resource ledger / hybrid allocation / status fusion / TaskManager lifecycle.
(reference: taskMgr/{task_manager,task_scheduler,task_queue}.py, taskMgr/utils/{scheduler_strategy,utils_runner}.py,
resourceMgr/resource_manager.py)"""
import itertools
import json
import threading
import time

import pytest
from hypothesis import given, settings, strategies as st

from olearning_sim_b200.api.schema import TaskStatus, json2taskconfig
from olearning_sim_b200.control import hybrid
from olearning_sim_b200.control.jobs import InProcessJobExecutor, JobStatus
from olearning_sim_b200.control.repo import ResTableRepo, TaskTableRepo
from olearning_sim_b200.control.resource_manager import FakePhoneMgr, ResourceManager, ServerClusterResStatus
from olearning_sim_b200.control.scheduling import DefaultStrategy, TaskQueue, TaskScheduler
from olearning_sim_b200.control.task_manager import (TaskManager, TaskMgrConfig, calculate_conditions, combine_status)
from olearning_sim_b200.control.task_runner import JobSubmitter
from tests.helpers import make_task_json

S = TaskStatus


# ---- queue / strategy -----------------------------------------------------------------------------------------
def test_task_queue_is_thread_safe_and_deduplicates():
    q = TaskQueue()
    tasks = [json2taskconfig(make_task_json(task_id=f"t{i}")) for i in range(50)]
    ths = [threading.Thread(target=lambda t=t: (q.add(t), q.add(t))) for t in tasks]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert len(q) == 50 and sorted(q.getTaskIDsQueue()) == sorted(t.task_id for t in tasks)
    assert q.delete("t3") and "t3" not in q and not q.delete("t3")


def test_default_strategy_demand_feasibility_and_score():
    s = DefaultStrategy()
    t = json2taskconfig(make_task_json(nums=(8, 4), devices=("High", "Low"), dynamic=(0, 0), logical=(8, 2),
                                       num_request=(3, 2), phone_request=(0, 2)))
    t.logical_simulation.computation_unit.setting[0].num_cpus = 2
    req = s.get_task_request_resource(t)
    assert req["logical_simulation"] == {"cpu": 2 * 3 + 1 * 2, "mem": 5.0}
    assert req["device_simulation"] == {"u1": {"High": 0, "Low": 2}}
    avail = {"logical_simulation": {"cpu": 8, "mem": 5}, "device_simulation": {"u1": {"Low": 2}}}
    assert s.check_resource_availability(req, avail)
    assert not s.check_resource_availability(req, {**avail, "logical_simulation": {"cpu": 7.9, "mem": 5}})
    assert not s.check_resource_availability(req, {**avail, "device_simulation": {"u1": {"Low": 1}}})
    # score = (n-i)/n + priority/10 ; first max wins
    w = [{"task_priority": 0}, {"task_priority": 3}, {"task_priority": 10}]
    assert s.schedule_task(w) == 2
    assert s.schedule_task([{"task_priority": 5}, {"task_priority": 5}]) == 0
    assert s.schedule_task([{"task_priority": 0}, {"task_priority": 4}, {"task_priority": 0}]) == 1   # 0.667+0.4 > 1.0


# ---- resource manager ---------------------------------------------------------------------------------------------
def test_resource_ledger_codes_and_atomicity():
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 8, "mem": 16, "gpu": 1}, phone_mgr=FakePhoneMgr({"u": {"High": 3}}))
    assert rm.requestClusterResource("", 1, 1) == ServerClusterResStatus.TASKID_NULL
    assert rm.requestClusterResource("a", -1, 1) == ServerClusterResStatus.FAIL
    assert rm.requestClusterResource("a", 9, 1) == ServerClusterResStatus.NOT_ENOUGH_RES
    assert rm.requestClusterResource("a", 4, 4, 0.5) == ServerClusterResStatus.SUCCESS
    assert rm.requestClusterResource("a", 1, 1) == ServerClusterResStatus.TASK_REPEAT
    assert rm.getClusterAvailableResource() == {"cpu": 4, "mem": 12, "gpu": 0.5}
    assert rm.releaseClusterResource("zzz") == ServerClusterResStatus.TASKID_NOT_FOUND
    # 16 threads race for the remaining 4 cpus with 1-cpu requests: exactly 4 may win
    results = []
    ths = [threading.Thread(target=lambda i=i: results.append(rm.requestClusterResource(f"r{i}", 1, 1))) for i in range(16)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert results.count(ServerClusterResStatus.SUCCESS) == 4
    assert rm.getClusterAvailableResource()["cpu"] == 0
    # combined request rolls the cluster half back when phones are short
    rm.releaseClusterResource("r0")
    ok, msg = rm.requestResource("b", "u", cpu=1, mem=1, phones={"High": 5}, type=0)
    assert not ok and msg == "PHONE_NOT_ENOUGH_RES" and rm.getDetailResbyTaskId("b") is None
    assert rm.requestResource("b", "u", cpu=1, mem=1, phones={"High": 2}, type=0) == (True, "SUCCESS")
    assert rm.getResource()["phone"] == {"u": {"High": 1}}
    assert rm.releaseResource("b", "u", type=0)[0] and rm.getResource()["phone"] == {"u": {"High": 3}}


# ---- hybrid allocation ---------------------------------------------------------------------------------------------
@settings(max_examples=150, deadline=None)
@given(N=st.integers(1, 400), f=st.integers(1, 40), k=st.integers(1, 8), m=st.integers(1, 60), qf=st.floats(0, 0.5))
def test_hybrid_solver_matches_brute_force(N, f, k, m, qf):
    q = int(N * qf)
    x, c = hybrid.solve_class(N, f, k, m, q)
    assert 0 <= x <= N - q
    assert c == pytest.approx(hybrid.brute_force_class(N, f, k, m, q))


def test_hybrid_optimizer_fills_allocation_and_paths():
    tj = make_task_json(nums=(100, 50), devices=("High", "Low"), dynamic=(0, 0), logical=(0, 0), device=(0, 0),
                        num_request=(4, 0), phone_request=(10, 5), optimization=True)
    tj["target"]["data"][0]["data_path"] = "/data/cifar.zip"
    out = hybrid.HybridOptimizer(tj).fix_data_parameters()[0]
    al = out["allocation"]
    assert [a + b for a, b in zip(al["logical_simulation"], al["device_simulation"])] == [100, 50]
    assert al["logical_simulation"][1] == 0 and al["device_simulation"][1] == 50          # no logical units for "Low"
    assert out["data_path_logical"].endswith("data_split/data_0_logical.zip")
    # user split: all logical -> original path, no device path
    tj2 = make_task_json()
    tj2["target"]["data"][0]["data_path"] = "/data/x.zip"
    o2 = hybrid.HybridOptimizer(tj2).fix_data_parameters()[0]
    assert o2["data_path_logical"] == "/data/x.zip" and o2["data_path_device"] == ""


def test_job_submitter_strips_other_side():
    tj = make_task_json(nums=(8,), logical=(6,), device=(2,), phone_request=(1,))
    data = hybrid.HybridOptimizer(tj).fix_data_parameters()
    lg, dv = JobSubmitter("logical_simulation"), JobSubmitter("device_simulation")
    lg.integrate_submission_info(tj, data)
    dv.integrate_submission_info(tj, data)
    assert lg.needs_submission and dv.needs_submission
    li = lg.submission_info
    assert li["target"]["data"][0]["simulation_target"] == {"devices": ["High"], "nums": [6]}
    assert "device_simulation" not in li["operatorflow"]["operators"][0] and "device_simulation" not in li
    assert list(li["operatorflow"]["flow_setting"]["start"]) == ["logical_simulation"]
    di = dv.submission_info
    assert di["target"]["data"][0]["simulation_target"]["nums"] == [2]
    assert "logical_simulation" not in di["operatorflow"]["operators"][0]


# ---- status truth table ---------------------------------------------------------------------------------------------
def test_combine_status_truth_table_all_160_inputs():
    counts = {S.SUCCEEDED: 0, S.STOPPED: 0, S.FAILED: 0, S.RUNNING: 0}
    for ls, lrf, lstat, ds, drf, dstat in itertools.product([True, False], [True, False],
                                                            [S.SUCCEEDED, S.PENDING, S.RUNNING, S.STOPPED, S.FAILED],
                                                            [True, False], [True, False], [True, False]):
        r = combine_status(ls, lrf, lstat, ds, drf, dstat)
        if (ls and lrf) or (ds and drf):
            assert r == S.FAILED                       # "impossible" states are reported as FAILED
            continue
        counts[r] += 1
        if ls and ds:
            assert r == S.SUCCEEDED
        elif r == S.STOPPED:
            assert not ls and not lrf and lstat == S.STOPPED and not drf and dstat
        elif r == S.RUNNING:
            assert lstat in (S.PENDING, S.RUNNING) or ls
    # the reference's docstring: 90 reachable states = 10 SUCCEEDED + 2 STOPPED + 67 FAILED + 11 RUNNING
    assert counts == {S.SUCCEEDED: 10, S.STOPPED: 2, S.FAILED: 67, S.RUNNING: 11}


def _tp(max_round=2, nums=(8,), dyn=(1,)):
    return {"max_round": max_round, "operator_name_list": ["train", "upload"], "data_name_list": ["d"],
            "total_simulation": [{"name": "d", "simulation_target": {"devices": ["High"], "nums": list(nums), "dynamic_nums": list(dyn)}}]}


def _res(succ, fail):
    return [{"name": "d", "simulation_target": {"devices": ["High"], "success_num": [succ], "failed_num": [fail]}}]


def test_calculate_conditions_logical_only():
    tgt = [{"name": "d", "simulation_target": {"devices": ["High"], "nums": [8]}}]
    base = dict(logical_target=tgt, device_target=None, device_result=[], device_round=None, device_operator=None)
    # no result yet -> running
    assert calculate_conditions(_tp(), logical_result=None, logical_round=None, logical_operator=None, **base) == (False, False, True, False)
    # last round, last operator, 7 of 8 ok with 1 tolerated -> success
    assert calculate_conditions(_tp(), logical_result=_res(7, 1), logical_round=2, logical_operator="upload", **base) == (True, False, True, False)
    # same numbers but not the last operator -> not yet
    assert calculate_conditions(_tp(), logical_result=_res(7, 1), logical_round=2, logical_operator="train", **base)[0] is False
    # 2 failures > dynamic_nums 1 -> early round failure, even in round 1
    assert calculate_conditions(_tp(), logical_result=_res(6, 2), logical_round=1, logical_operator="train", **base) == (False, True, True, False)


def test_calculate_conditions_both_sides_sum_successes():
    lt = [{"name": "d", "simulation_target": {"devices": ["High"], "nums": [5]}}]
    dt = [{"name": "d", "simulation_target": {"devices": ["High"], "nums": [3]}}]
    r = calculate_conditions(_tp(dyn=(0,)), logical_target=lt, logical_result=_res(5, 0), logical_round=2, logical_operator="upload",
                             device_target=dt, device_result=_res(3, 0), device_round=2, device_operator="upload")
    assert r == (True, False, True, False)
    r = calculate_conditions(_tp(dyn=(0,)), logical_target=lt, logical_result=_res(5, 0), logical_round=2, logical_operator="upload",
                             device_target=dt, device_result=_res(2, 1), device_round=2, device_operator="upload")
    assert r[1] and r[3] and not r[0] and not r[2]          # shared failure budget exceeded -> both fail early


# ---- TaskManager lifecycle ---------------------------------------------------------------------------------------------
class StubRunner:
    """round-loop stand-in: reports success for every client, optionally blocks until released"""
    gate = None

    def __init__(self, task_json, stop_event=None, repo=None, **_):
        self.tj, self.stop, self.repo = task_json, stop_event, repo

    def run(self):
        if StubRunner.gate is not None:
            while not StubRunner.gate.is_set() and not self.stop.is_set():
                time.sleep(0.005)
        if self.stop.is_set():
            return False
        d = self.tj["target"]["data"][0]
        n = d["simulation_target"]["nums"]
        res = [{"name": d["name"], "simulation_target": {"devices": d["simulation_target"]["devices"],
                                                         "success_num": n, "failed_num": [0] * len(n)}}]
        self.repo.set_item_values("task_id", self.tj["task_id"], {
            "logical_round": self.tj["operatorflow"]["flow_setting"]["round"],
            "logical_operator": self.tj["operatorflow"]["operators"][-1]["name"],
            "logical_result": json.dumps({"logical_result": res})})
        return True


def make_tm(totals=None, **cfg):
    repo = TaskTableRepo()
    rm = ResourceManager(ResTableRepo(), totals=totals or {"cpu": 4, "mem": 64, "gpu": 0}, phone_mgr=FakePhoneMgr())
    ex = InProcessJobExecutor(runner_factory=StubRunner)
    return TaskManager(repo=repo, resource_manager=rm, executor=ex, config=TaskMgrConfig(**cfg)), repo, rm, ex


def test_submit_validates_and_requires_undone_row():
    tm, repo, _, _ = make_tm(auto_create_rows=False)
    tj = make_task_json(task_id="a")
    assert not tm.submitTask(tj)                       # no UNDONE row created by the "GUI"
    repo.create_task("a", "u1")
    assert tm.submitTask(tj)
    assert tm.getTaskQueue() == ["a"] and tm.getTaskStatus("a") == S.QUEUED
    assert not tm.submitTask(tj)                       # already QUEUED
    bad = make_task_json(task_id="b"); bad["target"]["priority"] = 99
    repo.create_task("b", "u1")
    assert not tm.submitTask(bad)
    assert tm.getTaskStatus("nope") == S.MISSING
    ts = json.loads(repo.get_task_value("a", "total_simulation"))
    assert ts["max_round"] == 2 and ts["operator_name_list"] == ["training"] and ts["data_name_list"] == ["data_0"]


def test_full_lifecycle_schedule_run_succeed_reap():
    StubRunner.gate = None
    tm, repo, rm, ex = make_tm()
    assert tm.submitTask(make_task_json(task_id="a", num_request=(2,)))
    assert tm.schedule_once(inline=True) == "a"
    assert repo.get_task_value("a", "resource_occupied") == 1 and repo.get_task_value("a", "job_id")
    assert rm.getClusterAvailableResource()["cpu"] == 2
    assert tm.wait_for("a", 10) == S.SUCCEEDED
    assert tm.reap_once() == ["a"]
    assert repo.get_task_value("a", "resource_occupied") == 0 and repo.get_task_value("a", "task_finished_time")
    assert rm.getClusterAvailableResource()["cpu"] == 4
    assert tm.getTaskStatus("a") == S.SUCCEEDED


def test_priority_and_resource_gating():
    StubRunner.gate = threading.Event()
    tm, repo, rm, _ = make_tm(totals={"cpu": 3, "mem": 64, "gpu": 0})
    tm.submitTask(make_task_json(task_id="big", num_request=(3,), priority=1))
    tm.submitTask(make_task_json(task_id="small", num_request=(1,), priority=9))
    first = tm.schedule_once(inline=True)
    assert first == "small"                            # 0.5 + 0.9 beats 1.0 + 0.1
    assert tm.schedule_once(inline=True) is None       # "big" needs 3 cpus, only 2 left
    assert tm.getTaskStatus("big") == S.QUEUED and tm.getTaskStatus("small") == S.RUNNING
    StubRunner.gate.set()
    assert tm.wait_for("small", 10) == S.SUCCEEDED
    tm.reap_once()
    assert tm.schedule_once(inline=True) == "big"
    assert tm.wait_for("big", 10) == S.SUCCEEDED
    StubRunner.gate = None


def test_stop_queued_and_running_tasks():
    StubRunner.gate = threading.Event()
    tm, repo, rm, _ = make_tm()
    tm.submitTask(make_task_json(task_id="q"))
    assert tm.stopTask("q") and tm.getTaskStatus("q") == S.STOPPED and tm.getTaskQueue() == []
    tm.submitTask(make_task_json(task_id="r"))
    tm.schedule_once(inline=True)
    assert tm.getTaskStatus("r") == S.RUNNING
    assert tm.stopTask("r")
    assert tm.wait_for("r", 10) == S.STOPPED
    assert tm.reap_once() == ["r"] and rm.getClusterAvailableResource()["cpu"] == 4
    StubRunner.gate = None


def test_watchdog_stops_stale_tasks():
    StubRunner.gate = threading.Event()
    tm, repo, _, _ = make_tm(interrupt_queue_time=10, interrupt_running_time=20)
    tm.submitTask(make_task_json(task_id="old"))
    repo.set_task_value("old", "in_queue_time", time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(time.time() - 100)))
    tm.submitTask(make_task_json(task_id="fresh"))
    assert tm.watchdog_once() == ["old"]
    assert tm.getTaskStatus("old") == S.STOPPED and tm.getTaskStatus("fresh") == S.QUEUED
    tm.schedule_once(inline=True)
    repo.set_task_value("fresh", "submit_task_time", time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(time.time() - 100)))
    assert tm.watchdog_once() == ["fresh"]
    assert tm.wait_for("fresh", 10) == S.STOPPED
    StubRunner.gate = None


def test_crash_recovery_rebuilds_queue_and_running_set(tmp_path):
    StubRunner.gate = threading.Event()
    db = str(tmp_path / "tasks.sqlite")
    repo = TaskTableRepo(db)
    rm = ResourceManager(ResTableRepo(str(tmp_path / "res.sqlite")), totals={"cpu": 2, "mem": 8, "gpu": 0}, phone_mgr=FakePhoneMgr())
    tm = TaskManager(repo=repo, resource_manager=rm, executor=InProcessJobExecutor(runner_factory=StubRunner))
    for tid in ("a", "b", "c"):
        tm.submitTask(make_task_json(task_id=tid, num_request=(2,)))
        time.sleep(1.05) if tid != "c" else None       # in_queue_time has 1 s resolution
    tm.schedule_once(inline=True)                      # "a" runs and holds both cpus
    tm.shutdown()
    # new process: same tables
    tm2 = TaskManager(repo=TaskTableRepo(db, shared_key="tm2"), resource_manager=rm,
                      executor=InProcessJobExecutor(runner_factory=StubRunner))
    assert tm2.getTaskQueue() == ["b", "c"]            # QUEUED rows re-enqueued in in_queue_time order
    assert "a" in tm2._task_running
    StubRunner.gate.set()
    StubRunner.gate = None


def test_concurrent_submit_stop_status_hammer():
    StubRunner.gate = None
    tm, repo, rm, _ = make_tm(totals={"cpu": 64, "mem": 640, "gpu": 0})
    errors = []

    def worker(i):
        try:
            tid = f"h{i}"
            assert tm.submitTask(make_task_json(task_id=tid))
            tm.getTaskStatus(tid)
            if i % 3 == 0:
                tm.stopTask(tid)
            tm.getTaskQueue()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(40)]
    sched = threading.Thread(target=lambda: [tm.schedule_once(inline=True) for _ in range(60)])
    [t.start() for t in ths]; sched.start()
    [t.join() for t in ths]; sched.join()
    assert not errors
    for _ in range(80):
        tm.schedule_once(inline=True)
    deadline = time.time() + 20
    while time.time() < deadline:
        tm.reap_once()
        st_all = [tm.getTaskStatus(f"h{i}") for i in range(40)]
        if all(s in S.TERMINAL for s in st_all):
            break
        time.sleep(0.05)
    assert all(s in S.TERMINAL for s in st_all), st_all
    tm.reap_once()
    assert rm.getClusterAvailableResource()["cpu"] == 64       # nothing leaked from the ledger
