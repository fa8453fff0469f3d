"""Round-loop driver of one task (the job a TaskRunner submits).

    python -m olearning_sim_b200.control.run_task --task '<json>'            (CLI parity with run_task.py:40-49,324-326)
    python -m torch.distributed.run --nproc-per-node N -m olearning_sim_b200.control.run_task --task-file f.json

Counterpart of ``RayRunner`` (ols_core/taskMgr/run_task.py:32-322) and of the Ray ``Actor``
(ols_core/taskMgr/utils/utils_run_task.py:146-577):

    for round in rounds:                      operatorflow.start() gate
      for operator in operators:              NotifyStart -> run all clients -> NotifyComplete -> analyze_results
                                              operatorflow.stop() gate

* built-in GPU operators: every client of every (dataset, device class) is a row of the FedEngine's packed tensors;
  one call trains them all.  Device classes map to heterogeneity knobs (local-step budget, online probability).
* script operators (reference contract): clients are sharded over ``num_request`` workers with the reference's
  near-even contiguous split (run_task.py:85-105); each worker runs its clients sequentially, one
  ``python3 <operator>/<entry> --params f.json`` subprocess per simulated phone (utils_run_task.py:481-514).
"""
from __future__ import annotations

import argparse
import concurrent.futures
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
from typing import Any, Dict, List, Optional, Tuple

from ..deviceflow.service import NotifyRequest
from ..parallel.comm import shard_range
from ..utils.logger import Logger
from .operatorflow import OperatorFlow
from .operators import BUILTIN_OPERATORS, parse_builtin

logger = Logger()


class Params:
    """Parsed job JSON (utils_run_task.py:20-51)."""

    def __init__(self, task_json: Dict[str, Any]):
        if isinstance(task_json, str):
            task_json = json.loads(task_json)
        self.json = task_json
        self.task_id = task_json.get("task_id", "")
        self.user_id = task_json.get("user_id", "")
        of = task_json.get("operatorflow", {})
        fs = of.get("flow_setting", {})
        self.round = int(fs.get("round", 1))
        self.start_params = fs.get("start", {}).get("logical_simulation", {}) or {}
        self.stop_params = fs.get("stop", {}).get("logical_simulation", {}) or {}
        self.operators = list(of.get("operators", []))
        self.data = list(task_json.get("target", {}).get("data", []))
        ls = task_json.get("logical_simulation", {})
        cu = ls.get("computation_unit", {})
        self.computation_unit = {d: s for d, s in zip(cu.get("devices", []), cu.get("setting", []))}
        self.resource_request = {r.get("name"): r for r in ls.get("resource_request", [])}
        self.engine = task_json.get("engine", {}) or {}


def construct_run_params(data: Dict[str, Any], resource_request: Dict[str, Any]) -> List[Dict[str, Any]]:
    """client -> worker sharding for one dataset: clients of device class d are split near-evenly over
    ``num_request[d]`` workers, contiguous ``split_index`` ranges; classes follow each other in index space."""
    out, base = [], 0
    devices = data.get("simulation_target", {}).get("devices", [])
    nums = data.get("simulation_target", {}).get("nums", [])
    req = dict(zip(resource_request.get("devices", []), resource_request.get("num_request", [])))
    for dev, n in zip(devices, nums):
        workers = int(req.get(dev, 0))
        if n > 0 and workers <= 0:
            workers = 1
        for w in range(workers):
            s, e = shard_range(n, workers, w)
            if e > s:
                out.append({"data_name": data.get("name", ""), "device_type": dev, "worker": w,
                            "split_index": list(range(base + s, base + e))})
        base += n
    return out


class ScriptActor:
    """One worker ("actor"): fetches its data and model once, then runs its simulated phones sequentially, one
    subprocess each (ols_core/taskMgr/utils/utils_run_task.py:146-577)."""

    def __init__(self, params: Params, operator: Dict[str, Any], working_dir: str, round_idx: int,
                 deviceflow_info: Dict[str, Any], stop_event: threading.Event, file_repo: Any = None,
                 actor_id: Optional[str] = None):
        self.p, self.op, self.wd, self.round, self.df, self.stop = params, operator, working_dir, round_idx, deviceflow_info, stop_event
        self.file_repo = file_repo
        import uuid
        self.actor_id = actor_id or f"actor_{uuid.uuid4().hex[:12]}"
        self.save_dir = os.path.join(working_dir, "actors", self.actor_id)

    # ---- downloads (utils_run_task.py:174-428) -------------------------------------------------------
    def _repo(self, ttype: str, cfg: Optional[Dict[str, Any]] = None):
        if self.file_repo is not None:
            return self.file_repo
        from ..utils.file_repo import get_file_repo
        return get_file_repo(ttype, cfg)

    def download_data_files(self, data: Dict[str, Any], split_index: List[int], simulation_num: int) -> Dict[str, Any]:
        """Resolve the dataset to local per-phone paths.  ``data_split_type`` false: every phone gets the one shared folder
        (or file); true: the unpacked data must hold exactly ``simulation_num`` entries, which are sorted and indexed by
        each phone's ``split_index`` — client i owns shard i (:248-263, :294-310).  ``dataconfig.json`` rides along."""
        out: Dict[str, Any] = {"local_data_path_list": [], "dataconfig": {}}
        if not self.op.get("use_data", True) or not data.get("data_path"):
            return out
        from ..engine.data import fetch_dataset
        ttype = str(data.get("data_transfer_type", "FILE")).upper()
        repo = None if ttype == "FILE" else self._repo(ttype, data.get("repo_config"))
        root = fetch_dataset(data, os.path.join(self.save_dir, "data"), file_repo=repo)
        if not root or not os.path.exists(root):
            raise FileNotFoundError(f"dataset {data.get('data_path')} is not reachable")
        cfg_path = os.path.join(root, "dataconfig.json")
        if os.path.isfile(cfg_path):
            try:
                with open(cfg_path) as f:
                    out["dataconfig"] = json.load(f)
            except ValueError:
                out["dataconfig"] = {}
        if data.get("data_split_type", False):
            entries = sorted(e for e in os.listdir(root) if e not in ("dataconfig.json", ".unpacked"))
            if len(entries) != simulation_num:
                raise ValueError(f"data_split_type is true but the data holds {len(entries)} shards for {simulation_num} simulated devices")
            out["local_data_path_list"] = [os.path.join(root, entries[i]) for i in split_index]
        else:
            entries = [e for e in os.listdir(root) if e not in ("dataconfig.json", ".unpacked")] if os.path.isdir(root) else []
            shared = os.path.join(root, entries[0]) if len(entries) == 1 and os.path.isdir(os.path.join(root, entries[0])) else root
            out["local_data_path_list"] = [shared for _ in split_index]
        return out

    def model_object_path(self) -> str:
        model = self.op.get("model", {}) or {}
        path = model.get("model_path", "")
        if model.get("use_model") and model.get("model_for_train") and self.round > 0 and model.get("model_update_style"):
            # per-round global model written by the aggregator: "{task_id}_{current_round}_result_model..." (:335-350)
            name = model["model_update_style"].replace("{task_id}", self.p.task_id).replace("{current_round}", str(self.round))
            path = os.path.join(os.path.dirname(path), name)
        return path

    def download_model_files(self) -> Dict[str, Any]:
        model = self.op.get("model", {}) or {}
        path = self.model_object_path()
        if not model.get("use_model") or not path:
            return {"local_model_path": ""}
        ttype = str(model.get("model_transfer_type", "FILE")).upper()
        if ttype == "FILE":
            return {"local_model_path": path}
        local = os.path.join(self.save_dir, "model", os.path.basename(path))
        os.makedirs(os.path.dirname(local), exist_ok=True)
        if not self._repo(ttype, model.get("repo_config")).download_file(path, local):
            raise FileNotFoundError(f"model {path} ({ttype}) could not be fetched")
        return {"local_model_path": local}

    # ---- per-phone parameters (utils_run_task.py:430-479; operator contract base_operator.py:12-53) ------------
    def construct_params_for_phones(self, run_params: Dict[str, Any], data: Dict[str, Any],
                                    downloads: Optional[Dict[str, Any]] = None) -> List[Dict[str, Any]]:
        model = dict(self.op.get("model", {}) or {})
        info = self.op.get("logical_simulation", {})
        try:
            user_params = json.loads(info.get("operator_params") or "{}")
        except ValueError:
            user_params = {}
        downloads = downloads or {}
        paths = downloads.get("data", {}).get("local_data_path_list", [])
        local_model = downloads.get("model", {}).get("local_model_path", "") or self.model_object_path()
        sim_num = sum(int(n) for n in data.get("simulation_target", {}).get("nums", []))
        phone_model = {"use_model": model.get("use_model", False), "model_path": local_model}
        phone_op = {"name": self.op.get("name", ""), "input": self.op.get("input", []),
                    "operation_behavior_controller": self.op.get("operation_behavior_controller", {}),
                    "use_data": self.op.get("use_data", True), "model": phone_model,
                    "operator_params": info.get("operator_params", "")}
        out = []
        for i, idx in enumerate(run_params["split_index"]):
            dpath = paths[i] if i < len(paths) else data.get("data_path", "")
            out.append({
                # reference contract
                "task_id": self.p.task_id, "current_round": self.round,
                "config": {"working_dir": self.wd, "deviceflow": self.df},
                "data": {"name": data.get("name", ""), "data_path": dpath, "data_split_type": data.get("data_split_type", False),
                         "task_type": data.get("task_type", ""), "dataconfig": downloads.get("data", {}).get("dataconfig", {}),
                         "path": dpath, "split_type": data.get("data_split_type", False)},
                "operator": phone_op, "simulation_num": sim_num, "split_index": idx, "actor_save_dir": self.save_dir,
                "actor_simulation_num": len(run_params["split_index"]), "params": user_params,
                # flat aliases kept for operators written against round 1 of this framework
                "operator_name": self.op.get("name", ""), "round": self.round, "device_type": run_params["device_type"],
                "model": {"use_model": model.get("use_model", False), "for_train": model.get("model_for_train", False),
                          "path": local_model},
                "operator_params": user_params, "deviceflow": self.df})
        return out

    def run(self, run_params: Dict[str, Any], data: Dict[str, Any]) -> Dict[str, Any]:
        info = self.op.get("logical_simulation", {})
        entry = os.path.join(self.wd, self.op.get("name", ""), info.get("operator_entry_file", ""))
        ok = bad = 0
        n_phones = len(run_params["split_index"])
        os.makedirs(self.save_dir, exist_ok=True)
        tmp = tempfile.mkdtemp(prefix="olsb_actor_")
        try:
            try:
                sim_num = sum(int(n) for n in data.get("simulation_target", {}).get("nums", []))
                downloads = {"data": self.download_data_files(data, run_params["split_index"], sim_num),
                             "model": self.download_model_files()}
            except Exception as e:  # noqa: BLE001 - a failed download fails every phone of this actor (:549-559)
                logger.error(self.p.task_id, "TaskMgr", "run_task", f"[Actor.run]: download failed: {e}")
                return {"data_name": run_params["data_name"], "device_type": run_params["device_type"], "success": 0,
                        "failed": n_phones}
            for i, prm in enumerate(self.construct_params_for_phones(run_params, data, downloads)):
                if self.stop.is_set():
                    break
                f = os.path.join(tmp, f"params_{i}.json")
                with open(f, "w") as fh:
                    json.dump(prm, fh)
                rc = subprocess.call([sys.executable, entry, "--params", f], cwd=os.path.dirname(entry))
                ok, bad = (ok + 1, bad) if rc == 0 else (ok, bad + 1)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
            shutil.rmtree(self.save_dir, ignore_errors=True)       # delete_tmp_results (:569-577)
        return {"data_name": run_params["data_name"], "device_type": run_params["device_type"], "success": ok, "failed": bad}


class RoundRunner:
    def __init__(self, task_json: Dict[str, Any], stop_event: Optional[threading.Event] = None, repo: Any = None,
                 deviceflow: Any = None, working_dir: Optional[str] = None, gate_params: Optional[Dict[str, Any]] = None,
                 rank: int = 0, world: int = 1, checkpoint_dir: Optional[str] = None, trace_path: Optional[str] = None,
                 performance_mgr: Any = None, file_repo: Any = None, **_: Any):
        self.params = Params(task_json)
        self.stop_event = stop_event or threading.Event()
        self.repo, self.deviceflow, self.working_dir = repo, deviceflow, working_dir or tempfile.mkdtemp(prefix="olsb_job_")
        self.rank, self.world = rank, world
        self.gate_params = gate_params or {}
        self.checkpoint_dir = checkpoint_dir or self.params.engine.get("checkpoint_dir")
        self.engines: Dict[str, Any] = {}
        self.history: List[Dict[str, Any]] = []
        trace_path = trace_path or self.params.engine.get("trace_path")
        self.trace = None
        if trace_path:
            from ..utils.tracing import TraceWriter
            self.trace = TraceWriter(trace_path)
        self.performance_mgr = performance_mgr
        self.file_repo = file_repo

    # ---- deviceflow wiring (utils_run_task.py:53-144) ----------------------------------------------------
    def _df_start(self, operator: Dict[str, Any], round_idx: int) -> Dict[str, Any]:
        bc = operator.get("operation_behavior_controller", {})
        if not bc.get("use_gradient_house") or self.deviceflow is None:
            return {"use": False}
        routing_key = f"{self.params.task_id}_{operator.get('name', '')}_{round_idx}"
        outbound = {}
        if bc.get("outbound_service"):
            try:
                outbound = json.loads(bc["outbound_service"])
            except ValueError:
                outbound = {}
        if self.rank == 0:
            ok = self.deviceflow.NotifyStart(NotifyRequest(self.params.task_id, routing_key, "logical_simulation",
                                                           bc.get("strategy_gradient_house", ""), outbound))
            if not ok:
                raise RuntimeError(f"deviceflow NotifyStart failed for {routing_key}")
        return {"use": True, "routing_key": routing_key, "compute_resource": "logical_simulation",
                "inbound": self.deviceflow.GetDeviceflowPulsarClient(),
                "websocket_url": self.deviceflow.GetDeviceflowWebsocket().get("url", "")}

    def _df_report(self, operator: Dict[str, Any], round_idx: int, res: Any) -> None:
        """Tell the DeviceFlow service what the on-device dispatch of this round did (see ReportDeviceDispatch)."""
        bc = operator.get("operation_behavior_controller", {}) or {}
        if not bc.get("use_gradient_house") or self.deviceflow is None or self.rank != 0:
            return
        if not hasattr(self.deviceflow, "ReportDeviceDispatch"):
            return
        flow_id = f"{self.params.task_id}_{operator.get('name', '')}_{round_idx}"
        summary = {"participants": res.participants, "forwarded": res.arrived, "dropped": res.dropped, "late": res.late}
        plan = res.extras.get("flow_plan")
        if plan:
            summary.update(timing=plan.get("timing", []), amount=plan.get("amount", []))
        self.deviceflow.ReportDeviceDispatch(self.params.task_id, flow_id, "logical_simulation", summary)

    def _df_stop(self, df_info: Dict[str, Any]) -> None:
        if df_info.get("use") and self.rank == 0:
            self.deviceflow.NotifyComplete(NotifyRequest(self.params.task_id, df_info["routing_key"], "logical_simulation"))

    # ---- built-in (GPU-native) operator ---------------------------------------------------------------
    def _builtin_engine(self, operator: Dict[str, Any]):
        from ..engine.fed_engine import FedConfig, FedEngine
        name = operator.get("name", "")
        if name in self.engines:
            return self.engines[name]
        info = operator.get("logical_simulation", {})
        kind = parse_builtin(info.get("operator_code_path", ""))
        if kind not in BUILTIN_OPERATORS:
            raise ValueError(f"unknown builtin operator {kind!r}")
        try:
            user = json.loads(info.get("operator_params") or "{}")
        except ValueError:
            user = {}
        cfgd = {**self.params.engine, **BUILTIN_OPERATORS[kind](user)}
        # shared engine for all operators that train the same population (e.g. "train" + "evaluate")
        share = cfgd.pop("share_engine_with", None)
        if share and share in self.engines:
            self.engines[name] = self.engines[share]
            return self.engines[name]
        # population: every client of every dataset / device class, in order
        classes: List[str] = []
        for d in self.params.data:
            st = d.get("simulation_target", {})
            for dev, n in zip(st.get("devices", []), st.get("nums", [])):
                classes += [f"{d.get('name', '')}/{dev}"] * int(n)
        fields = {k: v for k, v in cfgd.items() if k in FedConfig.__dataclass_fields__}
        fields["num_clients"] = len(classes)
        # DeviceFlow: the operator's dispatch strategy decides which client updates reach the aggregator (reference:
        # run_task.py:233-281 injects the inbound endpoint so every update traverses sorter -> shelf -> dispatcher)
        bc = operator.get("operation_behavior_controller", {}) or {}
        if bc.get("use_gradient_house") and bc.get("strategy_gradient_house") and not fields.get("deviceflow_strategy"):
            fields["deviceflow_strategy"] = bc["strategy_gradient_house"]
        cfg = FedConfig(**fields)
        eng = FedEngine(cfg, self.rank, self.world)
        eng.client_class = classes
        # device classes -> heterogeneity (stragglers / availability), e.g. {"Low": {"local_steps": 2, "p_online": 0.8}}
        import numpy as np
        dc = cfgd.get("device_classes", {}) or {}
        steps = np.full(len(classes), cfg.local_steps, np.int32)
        steps[eng.c0:eng.c1] = eng.local_steps_np
        pon = np.full(len(classes), cfg.p_online, np.float32)
        for i, cname in enumerate(classes):
            dev = cname.split("/", 1)[1]
            if dev in dc:
                steps[i] = min(int(dc[dev].get("local_steps", steps[i])), cfg.local_steps)
                pon[i] = float(dc[dev].get("p_online", pon[i]))
        if classes:
            eng.set_client_classes(classes, p_online=pon, local_steps=steps if (dc or cfg.hetero_steps) else None)
        # data: the task's datasets with the reference's shard contract (client i owns sub-folder i when
        # data_split_type is true); datasets without a data_path use the synthetic non-IID generator
        dsets = [{"name": d.get("name", ""), "data_path": d.get("data_path", ""), "data_split_type": d.get("data_split_type", False),
                  "data_transfer_type": d.get("data_transfer_type", "FILE"),
                  "num_clients": sum(int(n) for n in d.get("simulation_target", {}).get("nums", []))}
                 for d in self.params.data]
        if any(d["data_path"] for d in dsets) and cfgd.get("data", "task") != "synthetic":
            eng.load_datasets(dsets, work_dir=os.path.join(self.working_dir, "data"))
        else:
            eng.load_synthetic()
        if self.checkpoint_dir:
            from ..engine.checkpoint import CheckpointManager
            eng.ckpt = CheckpointManager(self.checkpoint_dir, self.params.task_id, rank=self.rank)
            state = eng.ckpt.load_agreed(eng.group, self.world)       # same resume round on every rank
            if state is not None:
                eng.load_state_dict(state["engine"])
                eng.resume_round = state["round"] + 1
        self._import_model(operator, eng)
        self.engines[name] = eng
        return eng

    # ---- model files (utils_run_task.py:327-397) -------------------------------------------------------
    # The reference's actors download the model file at the start of every operator; for ``model_for_train`` and
    # round > 0 the name is ``model_update_style`` with {task_id} / {current_round} substituted (default
    # "{task_id}_{current_round}_result_model.mnn"), written by the external aggregation service.  Here the engine is the
    # aggregator: the initial model is read from ``model_path`` and every training round publishes the new global model
    # under the same naming scheme, next to ``model_path`` (or in the job's working directory).
    def _model_cfg(self, operator: Dict[str, Any]) -> Dict[str, Any]:
        return operator.get("model", {}) or {}

    def _model_dir(self, m: Dict[str, Any]) -> str:
        path = m.get("model_path", "")
        d = os.path.dirname(path) if path and not path.endswith("/") else path
        return d or os.path.join(self.working_dir, "models")

    def _import_model(self, operator: Dict[str, Any], eng: Any) -> None:
        m = self._model_cfg(operator)
        path = m.get("model_path", "")
        if not (m.get("use_model") and path):
            return
        import torch
        local = path
        ttype = str(m.get("model_transfer_type", "FILE")).upper()
        if ttype != "FILE":
            from ..utils.file_repo import get_file_repo
            local = os.path.join(self.working_dir, "models", os.path.basename(path))
            os.makedirs(os.path.dirname(local), exist_ok=True)
            if not get_file_repo(ttype, m.get("repo_config")).download_if_missing(path, local):
                raise FileNotFoundError(f"model {path} ({ttype}) could not be fetched")
        if not os.path.isfile(local):
            return                                          # a directory: only the place where results are published
        obj = torch.load(local, map_location="cpu", weights_only=True)
        flat = obj["flat"] if isinstance(obj, dict) and "flat" in obj else obj
        if isinstance(flat, dict):                           # a state_dict in the zoo's parameter names
            flat = eng.spec.flatten({k: v for k, v in flat.items()})
        flat = torch.as_tensor(flat, dtype=torch.float32).reshape(-1)
        if flat.numel() < eng.spec.num_params:
            raise ValueError(f"model file has {flat.numel()} parameters, {eng.cfg.model} needs {eng.spec.num_params}")
        full = eng.spec.init(eng.cfg.seed)
        full[:eng.spec.num_params] = flat[:eng.spec.num_params]
        if eng.backend == "native":                          # native workspaces keep their own packed layout
            from ..ops.layout import CNN2
            full = CNN2.pack(eng.spec.unflatten(full)) if eng.cfg.model == "cnn2" else eng.ws.L.pack(full)
        eng.set_global_model(full)

    def _export_model(self, operator: Dict[str, Any], eng: Any, round_idx: int) -> Optional[str]:
        m = self._model_cfg(operator)
        if not (m.get("use_model") and m.get("model_for_train")):
            return None
        import torch
        style = m.get("model_update_style") or "{task_id}_{current_round}_result_model.mnn"
        name = style.replace("{task_id}", self.params.task_id).replace("{current_round}", str(round_idx + 1))
        flat = eng.spec_flat_global().detach().cpu()[:eng.spec.num_params]
        if self.rank != 0:
            return None
        out = os.path.join(self._model_dir(m), os.path.basename(name))
        os.makedirs(os.path.dirname(out), exist_ok=True)
        tmp = out + ".tmp"
        torch.save({"model": eng.cfg.model, "round": round_idx, "flat": flat}, tmp)
        os.replace(tmp, out)
        return out

    def _run_builtin(self, operator: Dict[str, Any], round_idx: int) -> List[Dict[str, Any]]:
        eng = self._builtin_engine(operator)
        kind = parse_builtin(operator.get("logical_simulation", {}).get("operator_code_path", ""))
        class_failed: Dict[str, int] = {}
        if getattr(eng, "resume_round", 0) > round_idx:
            res_failed = 0                                  # round already done before the restart
        elif kind == "evaluate":
            self.history.append({"round": round_idx, "operator": operator.get("name"), **eng.evaluate()})
            res_failed = 0
        else:
            res = eng.run_round(round_idx)
            rec = {"round": round_idx, "operator": operator.get("name"), "loss": res.loss,
                   "participants": res.participants, "dropped": res.dropped, "arrived": res.arrived,
                   "failed": res.failed, "client_steps": res.client_steps, "wall_s": res.wall_s,
                   **({"phases_ms": res.extras["phases_ms"]} if "phases_ms" in res.extras else {})}
            self.history.append(rec)
            if self.rank == 0:
                if self.trace is not None:
                    self.trace.write({"task_id": self.params.task_id, **rec})
                if self.performance_mgr is not None:
                    self.performance_mgr.commitJob({"task_id": self.params.task_id, "round": round_idx,
                                                    "t_round_ms": res.wall_s * 1e3, "client_steps": res.client_steps,
                                                    "phases_ms": res.extras.get("phases_ms", {})})
            res_failed = res.failed
            class_failed = res.extras.get("class_failed", {})
            if "class_participants" in res.extras:
                rec["class_participants"], rec["class_failed"] = res.extras["class_participants"], class_failed
            rec["late"] = res.late
            self._df_report(operator, round_idx, res)
            published = self._export_model(operator, eng, round_idx)
            if published:
                rec["model_file"] = published
            if getattr(eng, "ckpt", None) is not None:
                eng.ckpt.save(round_idx, {"engine": eng.state_dict(), "round": round_idx})
        # per (dataset, class) accounting: a client "fails" only if its computation failed (non-finite update);
        # offline / dropped-in-flight clients behaved as the trace prescribed and count as successes.  The per-class
        # failure counts are reduced on the device with the rest of the round record (round_stats_kernel).
        out: Dict[Tuple[str, str], List[int]] = {}
        for d in self.params.data:
            st = d.get("simulation_target", {})
            for dev, n in zip(st.get("devices", []), st.get("nums", [])):
                out[(d.get("name", ""), dev)] = [int(n), 0]
        for cname, nf in (class_failed or {}).items():
            key = tuple(cname.split("/", 1))
            if key in out and nf:
                out[key][0] -= min(int(nf), out[key][0])
                out[key][1] += int(nf)
        if res_failed and not class_failed and out:            # engine without class labels: job-wide count only
            key = next(iter(out))
            out[key][0] -= min(res_failed, out[key][0])
            out[key][1] += res_failed
        return [{"data_name": k[0], "device_type": k[1], "success": v[0], "failed": v[1]} for k, v in out.items()]

    # ---- script operator ----------------------------------------------------------------------------
    def _run_script(self, operator: Dict[str, Any], round_idx: int, df_info: Dict[str, Any]) -> List[Dict[str, Any]]:
        jobs = []
        for d in self.params.data:
            rr = self.params.resource_request.get(d.get("name"), {})
            for rp in construct_run_params(d, rr):
                jobs.append((rp, d))
        results = []
        with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
            futs = [pool.submit(ScriptActor(self.params, operator, self.working_dir, round_idx, df_info, self.stop_event,
                                            file_repo=self.file_repo).run, rp, d)
                    for rp, d in jobs]
            for f in futs:
                results.append(f.result())
        return results

    # ---- result accounting (run_task.py:149-210) -------------------------------------------------------
    def analyze_results(self, results: List[Dict[str, Any]], operator_name: str, round_idx: int) -> List[Dict[str, Any]]:
        logical_result = []
        for d in self.params.data:
            devices = d.get("simulation_target", {}).get("devices", [])
            succ, fail = [0] * len(devices), [0] * len(devices)
            for r in results:
                if r["data_name"] == d.get("name") and r["device_type"] in devices:
                    i = devices.index(r["device_type"])
                    succ[i] += int(r["success"])
                    fail[i] += int(r["failed"])
            logical_result.append({"name": d.get("name", ""),
                                   "simulation_target": {"devices": devices, "success_num": succ, "failed_num": fail}})
        if self.repo is not None and self.rank == 0:
            self.repo.set_item_values("task_id", self.params.task_id, {
                "logical_round": round_idx + 1, "logical_operator": operator_name,
                "logical_result": json.dumps({"logical_result": logical_result})})
        return logical_result

    # ---- main loop -----------------------------------------------------------------------------------
    def run(self) -> bool:
        p = self.params
        flow = OperatorFlow(p.task_id, p.round, p.start_params, p.stop_params, self.gate_params)
        try:
            for round_idx in range(p.round):
                if self.stop_event.is_set():
                    return False
                if not flow.start(round_idx):
                    raise RuntimeError(f"operatorflow start gate timed out in round {round_idx}")
                for op in p.operators:
                    if self.stop_event.is_set():
                        return False
                    df_info = self._df_start(op, round_idx)
                    code_path = op.get("logical_simulation", {}).get("operator_code_path", "")
                    if parse_builtin(code_path) is not None:
                        results = self._run_builtin(op, round_idx)
                    else:
                        results = self._run_script(op, round_idx, df_info)
                    self._df_stop(df_info)
                    self.analyze_results(results, op.get("name", ""), round_idx)
                if not flow.stop(round_idx) and round_idx != p.round - 1:
                    raise RuntimeError(f"operatorflow stop gate failed in round {round_idx}")
            return True
        finally:
            for e in {id(e): e for e in self.engines.values()}.values():
                try:
                    if getattr(e, "ckpt", None) is not None:
                        e.ckpt.wait()                      # the last round's checkpoint must be on disk before we return
                    e.close()
                except Exception:  # noqa: BLE001
                    pass


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", type=str, default=None, help="task json string")
    ap.add_argument("--task-file", type=str, default=None, help="file with {'task': ..., 'context': ...}")
    ap.add_argument("--profile", type=str, default=None, help="write one JSON record per round (phase timings) to this file")
    a = ap.parse_args(argv)
    ctx: Dict[str, Any] = {}
    if a.task_file:
        with open(a.task_file) as f:
            blob = json.load(f)
        task, ctx = blob["task"], blob.get("context", {})
    elif a.task:
        task = json.loads(a.task)
    else:
        ap.error("--task or --task-file required")
    from ..parallel.comm import init_distributed
    rank, world, _ = init_distributed()
    repo = None
    if ctx.get("repo_path"):
        from .repo import TaskTableRepo
        repo = TaskTableRepo(ctx["repo_path"])
    if a.profile:
        task.setdefault("engine", {})["profile"] = True
    deviceflow = None
    if ctx.get("deviceflow_address"):
        # the DeviceFlow service of the control-plane process, over gRPC (utils_run_task.py:53-144)
        from ..api.grpc_service import RemoteDeviceFlow
        deviceflow = RemoteDeviceFlow(ctx["deviceflow_address"])
    runner = RoundRunner(task, repo=repo, deviceflow=deviceflow, working_dir=ctx.get("working_dir"), rank=rank, world=world,
                         checkpoint_dir=ctx.get("checkpoint_dir"), trace_path=a.profile or ctx.get("trace_path"))
    from ..parallel.comm import PeerLostError
    try:
        ok = runner.run()
    except PeerLostError as e:
        # a peer GPU/process vanished mid-round: this rank is healthy and its model untouched; leave with EX_TEMPFAIL so the
        # job executor (control/jobs.py, max_restarts / shrink_on_restart) relaunches from the last agreed checkpoint
        print(json.dumps({"task_id": runner.params.task_id, "ok": False, "peer_lost": e.peer, "round": e.round_idx,
                          "rank": rank}), flush=True)
        os._exit(PeerLostError.exit_code)      # no collective teardown: the peer will never answer it
    if rank == 0:
        print(json.dumps({"task_id": runner.params.task_id, "ok": ok, "history": runner.history[-3:]}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
