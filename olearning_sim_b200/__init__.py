"""olearning_sim_b200 — B200-native federated-learning simulation framework with the capabilities of
opas-lab/olearning-sim (task manager, resource manager, DeviceFlow, cluster manager, operator contract, round loop).

Quick start::

    from olearning_sim_b200 import FedConfig, FedEngine
    eng = FedEngine(FedConfig(model="cnn2", num_clients=1024, local_steps=5))
    eng.load_synthetic()
    print(eng.run_round(0))

See DESIGN.md for the architecture and SURVEY.md for the mapping to the reference.
"""
__version__ = "0.1.0"

from .api.schema import TaskConfig, TaskStatus, json2taskconfig, taskconfig2json  # noqa: F401
from .engine.fed_engine import FedConfig, FedEngine, RoundResult  # noqa: F401


def build_task_service(config=None, resmgr_address=None, deviceflow_address=None, phonemgr_address=None, start_threads=False):
    """The reference's split deployment: a TaskMgr process whose ResourceMgr / phone manager / DeviceFlow live in other
    processes (``simu_session.py`` with svc 2 / 6 / 5) and are reached over gRPC — addresses from ``[grpc] resmgr`` /
    ``phonemgr`` and the ``deviceflow`` section unless given."""
    from .api.grpc_service import RemoteDeviceFlow, RemoteResourceMgr
    from .control.jobs import InProcessJobExecutor, SubprocessJobExecutor
    from .control.repo import TaskTableRepo
    from .control.task_manager import TaskManager, TaskMgrConfig
    from .utils.config import OlsbConfig
    cfg = config or OlsbConfig()
    rm = RemoteResourceMgr(resmgr_address or cfg.grpc.resmgr, phone_address=phonemgr_address or cfg.grpc.phonemgr)
    d = cfg.deviceflow
    df_addr = deviceflow_address or f"{d.grpc_service_address}:{d.grpc_service_port}"
    df = RemoteDeviceFlow(df_addr)
    ex = (SubprocessJobExecutor(nproc_per_node=max(cfg.cluster.gpus, 1), repo_path=cfg.storage.task_table)
          if cfg.cluster.job_executor == "subprocess" else InProcessJobExecutor())
    t = cfg.taskMgr
    tm = TaskManager(repo=TaskTableRepo(cfg.storage.task_table), resource_manager=rm, deviceflow=df, executor=ex,
                     config=TaskMgrConfig(t.scheduler_sleep_time, t.release_sleep_time, t.interrupt_sleep_time,
                                          t.interrupt_queue_time, t.interrupt_running_time),
                     start_threads=start_threads,
                     job_context={"checkpoint_dir": cfg.storage.checkpoint_dir or None, "deviceflow_address": df_addr})
    return tm, rm, df


def build_control_plane(config=None, start_threads=False):
    """Wire TaskManager + ResourceManager + DeviceFlow with the given ``OlsbConfig`` (in-process deployment)."""
    from .control.jobs import InProcessJobExecutor, SubprocessJobExecutor
    from .control.repo import FlowTableRepo, LogTableRepo, ResTableRepo, TaskTableRepo
    from .control.resource_manager import FakePhoneMgr, ResourceManager
    from .control.task_manager import TaskManager, TaskMgrConfig
    from .deviceflow.bus import InMemoryBus
    from .deviceflow.service import DeviceFlowConfig, DeviceFlowService
    from .utils.config import OlsbConfig
    from .utils.logger import Logger
    cfg = config or OlsbConfig()
    if cfg.storage.log_table or cfg.storage.log_jsonl:
        Logger().set_logger(jsonl_path=cfg.storage.log_jsonl or None,
                            repo=LogTableRepo(cfg.storage.log_table) if cfg.storage.log_table else None)
    rm = ResourceManager(ResTableRepo(cfg.storage.res_table), phone_mgr=FakePhoneMgr())
    d = cfg.deviceflow
    df = DeviceFlowService(bus=InMemoryBus(journal_path=d.journal_path or None), repo=FlowTableRepo(cfg.storage.deviceflow_table),
                           config=DeviceFlowConfig(d.grpc_service_address, d.grpc_service_port, d.max_workers, d.shelfroom_url,
                                                   d.shelfroom_namespace, d.source_url, d.source_topic, d.source_websocket_url))
    ex = (SubprocessJobExecutor(nproc_per_node=max(cfg.cluster.gpus, 1), repo_path=cfg.storage.task_table)
          if cfg.cluster.job_executor == "subprocess" else InProcessJobExecutor())
    t = cfg.taskMgr
    tm = TaskManager(repo=TaskTableRepo(cfg.storage.task_table), resource_manager=rm, deviceflow=df, executor=ex,
                     config=TaskMgrConfig(t.scheduler_sleep_time, t.release_sleep_time, t.interrupt_sleep_time,
                                          t.interrupt_queue_time, t.interrupt_running_time),
                     start_threads=start_threads,
                     job_context={"checkpoint_dir": cfg.storage.checkpoint_dir or None})
    if start_threads:
        df.start()
    return tm, rm, df
