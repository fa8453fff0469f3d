"""The reference's data contract on arrays: shared folder vs one sub-folder per client (client i owns shard i), zips,
and a task that trains on real shard folders end to end (CPU)."""
import json
import os
import zipfile

import numpy as np
import pytest

from olearning_sim_b200.engine.data import fetch_dataset, list_shard_folders, load_client_shards
from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine


def _write_shards(root, n, s=12, dim=784, classes=10, seed=0):
    rng = np.random.default_rng(seed)
    for i in range(n):
        d = os.path.join(root, f"client_{i:03d}")
        os.makedirs(d)
        y = np.full(s, i % classes, np.int64)                       # one label per client: as non-IID as it gets
        x = rng.standard_normal((s, dim)).astype(np.float32) + y[:, None]
        if i % 2:
            np.savez(os.path.join(d, "data.npz"), x=x, y=y)
        else:
            np.save(os.path.join(d, "x.npy"), x); np.save(os.path.join(d, "y.npy"), y)


def test_split_folders_client_i_owns_shard_i(tmp_path):
    _write_shards(str(tmp_path), 6)
    assert [os.path.basename(f) for f in list_shard_folders(str(tmp_path))] == [f"client_{i:03d}" for i in range(6)]
    x, y = load_client_shards(str(tmp_path), [4, 1], shard=8, split=True, total_clients=6, input_shape=(784,))
    assert x.shape == (2, 8, 784) and y.dtype == np.int32 and (y[0] == 4).all() and (y[1] == 1).all()
    x16, y16 = load_client_shards(str(tmp_path), [2], shard=16, split=True, total_clients=6)      # 12 samples -> wraps
    assert y16.shape == (1, 16) and np.array_equal(x16[0, 12:], x16[0, :4])
    with pytest.raises(ValueError, match="exactly 7 sub-folders"):
        load_client_shards(str(tmp_path), [0], 8, True, 7)


def test_shared_folder_gives_every_client_a_block(tmp_path):
    y = np.arange(40) % 10
    np.savez(tmp_path / "data.npz", x=np.arange(40 * 4, dtype=np.float32).reshape(40, 4), y=y)
    x, yy = load_client_shards(str(tmp_path), [0, 3, 4], shard=10, split=False, total_clients=5)
    assert np.array_equal(yy[1], y[30:40]) and np.array_equal(yy[2], y[0:10])                   # client 4 wraps around
    assert x.shape == (3, 10, 4)


def test_zip_is_fetched_and_unpacked_once(tmp_path):
    src = tmp_path / "src"
    _write_shards(str(src / "mydata"), 3)
    zpath = tmp_path / "mydata.zip"
    with zipfile.ZipFile(zpath, "w") as z:
        for base, _, files in os.walk(src):
            for f in files:
                full = os.path.join(base, f)
                z.write(full, os.path.relpath(full, src))
    root = fetch_dataset({"data_path": str(zpath), "data_transfer_type": "FILE"}, str(tmp_path / "work"))
    assert os.path.basename(root) == "mydata" and len(list_shard_folders(root)) == 3
    assert fetch_dataset({"data_path": str(zpath), "data_transfer_type": "FILE"}, str(tmp_path / "work")) == root
    evil = tmp_path / "evil.zip"
    with zipfile.ZipFile(evil, "w") as z:
        z.writestr("../escape.txt", "x")
    with pytest.raises(ValueError, match="escapes"):
        fetch_dataset({"data_path": str(evil)}, str(tmp_path / "work2"))


def test_engine_trains_on_real_shards_and_two_datasets(tmp_path):
    _write_shards(str(tmp_path / "a"), 4, seed=1)
    np.savez(tmp_path / "b.npz", x=0, y=0)                                  # unused: dataset b has no path -> synthetic
    cfg = FedConfig(model="mlp2", num_clients=6, local_steps=1, batch=8, shard=8, lr=0.1, device="cpu", backend="torch")
    e = FedEngine(cfg)
    e.load_datasets([{"name": "a", "data_path": str(tmp_path / "a"), "data_split_type": True, "num_clients": 4},
                     {"name": "b", "data_path": "", "num_clients": 2}], work_dir=str(tmp_path / "w"))
    assert e.host_x.shape == (6, 8, 784) and [int(e.host_y[i, 0]) for i in range(4)] == [0, 1, 2, 3]
    r = e.run_round(0)
    assert r.participants == 6 and r.loss == r.loss
    with pytest.raises(ValueError, match="describe 5 clients"):
        e.load_datasets([{"name": "a", "data_path": str(tmp_path / "a"), "data_split_type": True, "num_clients": 4},
                         {"name": "b", "data_path": "", "num_clients": 1}])


def test_task_with_split_data_runs_through_the_control_plane(tmp_path):
    from olearning_sim_b200.api.schema import TaskStatus
    from olearning_sim_b200.control.jobs import InProcessJobExecutor
    from olearning_sim_b200.control.repo import ResTableRepo, TaskTableRepo
    from olearning_sim_b200.control.resource_manager import FakePhoneMgr, ResourceManager
    from olearning_sim_b200.control.task_manager import TaskManager
    from tests.helpers import make_task_json
    _write_shards(str(tmp_path / "shards"), 5)
    tj = make_task_json(task_id="realdata", rounds=2, nums=(5,), operator_params={
        "model": "mlp2", "local_steps": 1, "batch": 8, "shard": 8, "lr": 0.1, "device": "cpu", "backend": "torch"})
    tj["target"]["data"][0].update(data_path=str(tmp_path / "shards"), data_split_type=True)
    repo = TaskTableRepo()
    rm = ResourceManager(ResTableRepo(), totals={"cpu": 8, "mem": 32, "gpu": 0}, phone_mgr=FakePhoneMgr())
    tm = TaskManager(repo=repo, resource_manager=rm, executor=InProcessJobExecutor())
    assert tm.submitTask(tj) and tm.schedule_once(inline=True) == "realdata"
    assert tm.wait_for("realdata", 120) == TaskStatus.SUCCEEDED, tm._executor.get_job_error(repo.get_task_value("realdata", "job_id"))
    res = json.loads(repo.get_task_value("realdata", "logical_result"))["logical_result"][0]["simulation_target"]
    assert res["success_num"] == [5]


def test_model_files_initial_model_and_per_round_result(tmp_path):
    """model contract: the initial global model comes from model_path, every training round publishes
    {task_id}_{round}_result_model.* (model_update_style) that a later task can start from"""
    import torch
    from olearning_sim_b200.control.run_task import RoundRunner
    from olearning_sim_b200.models.zoo import get_model
    from tests.helpers import make_task_json
    spec = get_model("mlp2")
    init = spec.init(123)
    mdir = tmp_path / "models"
    mdir.mkdir()
    torch.save({"flat": init[:spec.num_params]}, mdir / "init.pt")
    tj = make_task_json(task_id="mfile", rounds=2, nums=(4,), operator_params={
        "model": "mlp2", "local_steps": 1, "batch": 8, "shard": 8, "lr": 0.1, "device": "cpu", "backend": "torch"})
    tj["operatorflow"]["operators"][0]["model"] = {"use_model": True, "model_for_train": True, "model_transfer_type": "FILE",
                                                   "model_path": str(mdir / "init.pt"),
                                                   "model_update_style": "{task_id}_{current_round}_result_model.pt"}
    # the job JSON the TaskRunner hands to the round loop carries simulation_target per dataset
    tj["target"]["data"][0]["simulation_target"] = {"devices": ["High"], "nums": [4]}
    rr = RoundRunner(tj, working_dir=str(tmp_path / "job"))
    eng = rr._builtin_engine(tj["operatorflow"]["operators"][0])
    assert torch.allclose(eng.global_model()[:spec.num_params], init[:spec.num_params])
    assert rr.run()
    files = sorted(os.listdir(mdir))
    assert files == ["init.pt", "mfile_1_result_model.pt", "mfile_2_result_model.pt"]
    last = torch.load(mdir / "mfile_2_result_model.pt", weights_only=True)
    assert last["round"] == 1 and torch.allclose(last["flat"], eng.global_model()[:spec.num_params])
    assert not torch.allclose(last["flat"], init[:spec.num_params])


def test_dirichlet_partition_and_shard_writer_round_trip(tmp_path):
    from olearning_sim_b200.engine.data import dirichlet_partition, write_shards
    rng = np.random.default_rng(0)
    y = rng.integers(0, 10, 2000)
    x = rng.standard_normal((2000, 8)).astype(np.float32)
    parts = dirichlet_partition(y, 12, alpha=0.3, seed=1, min_size=8)
    allidx = np.concatenate(parts)
    assert len(parts) == 12 and len(allidx) == 2000 and len(np.unique(allidx)) == 2000          # a partition
    skew = np.mean([np.bincount(y[p], minlength=10).max() / len(p) for p in parts])
    iid = np.mean([np.bincount(y[p], minlength=10).max() / len(p) for p in dirichlet_partition(y, 12, alpha=100.0, seed=1)])
    assert skew > 0.35 > iid                                                                   # small alpha is label-skewed
    write_shards(str(tmp_path), [x[p] for p in parts], [y[p] for p in parts])
    xs, ys = load_client_shards(str(tmp_path), [0, 11], shard=8, split=True, total_clients=12, input_shape=(8,))
    assert np.array_equal(ys[1], y[parts[11]][:8]) and np.array_equal(xs[0], x[parts[0]][:8])
