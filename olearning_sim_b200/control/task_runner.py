"""TaskRunner — turns a scheduled TaskConfig into running jobs on both sides (servers / phones).

Counterpart of ols_core/taskMgr/task_runner.py:41-224 and the helpers in ols_core/taskMgr/utils/utils_runner.py:

    TaskRunner.submit :116-224      submit_rayjob :41-87 (here: JobExecutor)        submit_phonejob :89-114
    JobSubmitter :478-628           DeviceflowResgister :630-671                    get_operator_code :684-782
    json2deviceconfig :784-902      HybridDataSplitter :178-476                     temp_write_table :226-256 (display)
"""
from __future__ import annotations

import copy
import json
import os
import random
import shutil
import tempfile
import zipfile
from typing import Any, Dict, List, Optional, Tuple

from ..api.schema import TaskConfig, taskconfig2json
from ..utils.logger import Logger
from .hybrid import HybridOptimizer
from .jobs import JobExecutor

logger = Logger()


class JobSubmitter:
    """Builds the per-side submission JSON: strips the other side's operator info, replaces total_simulation /
    allocation by ``simulation_target = {devices, nums}`` and points ``data_path`` at that side's split."""

    def __init__(self, job_type: str):
        if job_type not in ("logical_simulation", "device_simulation"):
            raise NotImplementedError(job_type)
        self._job_type = job_type
        self._needs_submission = False
        self._submission_info: Dict[str, Any] = {}

    @property
    def needs_submission(self) -> bool:
        return self._needs_submission

    @property
    def submission_info(self) -> Dict[str, Any]:
        return self._submission_info

    def integrate_submission_info(self, task_json: Dict[str, Any], data_list: List[Dict[str, Any]]) -> None:
        logical = self._job_type == "logical_simulation"
        other = "device_simulation" if logical else "logical_simulation"

        def side_num(d):
            al = d.get("allocation", {})
            if logical:
                return sum(al.get("logical_simulation", []))
            return sum(al.get("device_simulation", [])) + sum(al.get("running_response", {}).get("nums", []))

        self._needs_submission = any(side_num(d) > 0 for d in data_list)
        if not self._needs_submission:
            self._submission_info = {}
            return
        datas = []
        for d in data_list:
            sd = copy.deepcopy(d)
            sd["data_path"] = d.get("data_path_logical" if logical else "data_path_device", "")
            sd.pop("data_path_logical", None)
            sd.pop("data_path_device", None)
            sd["simulation_target"] = {"devices": d.get("total_simulation", {}).get("devices", []),
                                       "nums": d.get("allocation", {}).get(self._job_type, [])}
            if not logical:
                sd["running_response"] = d.get("allocation", {}).get("running_response", {})
            sd.pop("total_simulation", None)
            sd.pop("allocation", None)
            datas.append(sd)
        of = task_json.get("operatorflow", {})
        fs = of.get("flow_setting", {})
        ops = []
        for o in of.get("operators", []):
            so = copy.deepcopy(o)
            so.pop(other, None)
            ops.append(so)
        info = {"user_id": task_json.get("user_id", ""), "task_id": task_json.get("task_id", ""),
                "target": {"data": datas},
                "operatorflow": {"flow_setting": {"round": fs.get("round", 0),
                                                  "start": {self._job_type: fs.get("start", {}).get(self._job_type, {})},
                                                  "stop": {self._job_type: fs.get("stop", {}).get(self._job_type, {})}},
                                 "operators": ops},
                self._job_type: task_json.get(self._job_type, {})}
        if "engine" in task_json:
            info["engine"] = copy.deepcopy(task_json["engine"])
        self._submission_info = info


class DeviceflowRegister:
    def __init__(self, deviceflow: Any):
        self._df = deviceflow

    @staticmethod
    def check_use_deviceflow(task_json: Dict[str, Any]) -> bool:
        return any(o.get("operation_behavior_controller", {}).get("use_gradient_house", False)
                   for o in task_json.get("operatorflow", {}).get("operators", []))

    def register_task(self, task_id: str, logical_simulation: bool, device_simulation: bool) -> bool:
        if self._df is None:
            return False
        res = (["logical_simulation"] if logical_simulation else []) + (["device_simulation"] if device_simulation else [])
        try:
            return bool(self._df.RegisterTask(task_id, res))
        except Exception as e:  # noqa: BLE001
            logger.error(task_id, "TaskMgr", "task_runner", f"[register_task]: {e}")
            return False


DeviceflowResgister = DeviceflowRegister   # reference spelling


def get_operator_code(operator: Dict[str, Any], dst_dir: str, file_repo: Optional[Any] = None) -> bool:
    """Fetch one logical operator's code into ``dst_dir/<name>/``: FILE directory copy, or a .zip through
    FILE / HTTP / S3 / MINIO (the remote transports go through a pluggable ``file_repo``; offline boxes only have
    FILE).  ``builtin:*`` operators (GPU-native) need no code."""
    info = operator.get("logical_simulation", {})
    path = info.get("operator_code_path", "")
    name = operator.get("name", "operator")
    if not path or path.startswith("builtin:"):
        return True
    target = os.path.join(dst_dir, name)
    ttype = str(info.get("operator_transfer_type", "FILE")).upper()
    try:
        if os.path.isdir(path):
            if ttype != "FILE":
                return False
            shutil.copytree(path, target, dirs_exist_ok=True)
            return True
        local_zip = os.path.join(dst_dir, f"{name}.zip")
        if ttype == "FILE":
            shutil.copyfile(path, local_zip)
        else:
            if file_repo is None:
                raise RuntimeError(f"no file repo configured for transfer type {ttype}")
            if not file_repo.download_file(path, local_zip):
                return False
        os.makedirs(target, exist_ok=True)
        with zipfile.ZipFile(local_zip) as z:
            for m in z.namelist():                      # zip-slip guard
                if os.path.isabs(m) or ".." in m.split("/"):
                    raise RuntimeError(f"unsafe path in operator zip: {m}")
            z.extractall(target)
        os.remove(local_zip)
        return True
    except Exception as e:  # noqa: BLE001
        logger.error("", "TaskMgr", "task_runner", f"[get_operator_code]: operator {name}: {e}")
        return False


def json2deviceconfig(device_info: Dict[str, Any]) -> Dict[str, Any]:
    """Device-side submission JSON -> the DeviceTaskConfig handed to PhoneMgr.submitTask
    (ols_core/taskMgr/utils/utils_runner.py:784-902; proto ols_core/proto/phoneMgr.proto)."""
    fs = device_info.get("operatorflow", {}).get("flow_setting", {})
    return {
        "user_id": device_info.get("user_id", ""), "task_id": device_info.get("task_id", ""),
        "round": fs.get("round", 0),
        "start": fs.get("start", {}).get("device_simulation", {}), "stop": fs.get("stop", {}).get("device_simulation", {}),
        "data": [{"name": d.get("name", ""), "data_path": d.get("data_path", ""),
                  "data_split_type": d.get("data_split_type", False),
                  "data_transfer_type": d.get("data_transfer_type", "FILE"), "task_type": d.get("task_type", ""),
                  "devices": d.get("simulation_target", {}).get("devices", []),
                  "nums": d.get("simulation_target", {}).get("nums", []),
                  "running_response": d.get("running_response", {})} for d in device_info.get("target", {}).get("data", [])],
        "operators": [{"name": o.get("name", ""), "input": o.get("input", []),
                       "behavior_controller": o.get("operation_behavior_controller", {}),
                       "use_data": o.get("use_data", True), "model": o.get("model", {}),
                       **o.get("device_simulation", {})} for o in device_info.get("operatorflow", {}).get("operators", [])],
        "resource_request": device_info.get("device_simulation", {}).get("resource_request", []),
    }


def fix_device_task_json(task_json: Dict[str, Any]) -> Dict[str, Any]:
    """utils_runner.py:905-932 — normalise enum spellings for the phone side."""
    tj = copy.deepcopy(task_json)
    for d in tj.get("target", {}).get("data", []):
        d["data_transfer_type"] = str(d.get("data_transfer_type", "FILE")).upper()
    return tj


class HybridDataSplitter:
    """Splits one dataset between the logical (server) and the device (phone) side of a hybrid task
    (ols_core/taskMgr/utils/utils_runner.py:178-476).

    The dataset is a zip — fetched through the task's transfer type (FILE / HTTP / S3 / MINIO) — holding an optional
    ``dataconfig.json`` next to the sample folders:

    * ``data_split_type = false``: ONE shared folder.  Its sample files are split row-wise with
      ``device_ratio = device / (logical + device)``: ``*_features.csv`` / ``*_labels.csv`` pairs (the reference's CSV
      contract; train files only, test files are copied to both sides, as in the reference) and ``x.npy`` / ``y.npy`` /
      ``data.npz`` arrays (the built-in GPU operators' format).  Rows are assigned with scikit-learn's
      ``train_test_split(random_state=42)`` when available — the reference's call — otherwise a seeded permutation.
    * ``data_split_type = true``: ``logical + device`` client sub-folders; the first ``logical`` go to the servers, the
      rest to the phones (:340-364).

    Both halves are re-zipped (folder entries + ``dataconfig.json``) and written to ``data_path_logical`` /
    ``data_path_device`` — uploaded through the same store for S3 / MINIO (:400-476), plain files otherwise."""

    def __init__(self, task_id: str, task_type: str = "classification", split_id: str = "", seed: int = 42,
                 file_repo: Optional[Any] = None):
        self.task_id, self.task_type, self.split_id, self.seed = task_id, task_type, split_id, seed
        self.file_repo = file_repo

    # ---- helpers -----------------------------------------------------------------------------------
    def _repo_for(self, data: Dict[str, Any]):
        if self.file_repo is not None:
            return self.file_repo
        ttype = str(data.get("data_transfer_type", "FILE")).upper()
        if ttype == "FILE":
            return None
        from ..utils.file_repo import get_file_repo
        return get_file_repo(ttype, data.get("repo_config"))

    def download_data(self, data: Dict[str, Any], tmp: str) -> Optional[str]:
        src = data.get("data_path", "")
        repo = self._repo_for(data)
        if repo is None:
            return src if os.path.exists(src) else None
        local = os.path.join(tmp, os.path.basename(src) or "data.zip")
        return local if repo.download_file(src, local) else None

    def _rows_split(self, n: int, device_ratio: float) -> Tuple[List[int], List[int]]:
        idx = list(range(n))
        try:
            from sklearn.model_selection import train_test_split
            a, b = train_test_split(idx, test_size=device_ratio, random_state=self.seed)
            return sorted(a), sorted(b)
        except Exception:  # noqa: BLE001 - sklearn missing or degenerate ratio
            rng = random.Random(self.seed)
            rng.shuffle(idx)
            k = n - int(round(n * device_ratio))
            return sorted(idx[:k]), sorted(idx[k:])

    def _split_shared_folder(self, folder: str, out_l: str, out_d: str, device_ratio: float) -> None:
        import numpy as np
        os.makedirs(out_l, exist_ok=True); os.makedirs(out_d, exist_ok=True)
        names = sorted(os.listdir(folder))
        done = set()
        # CSV pairs: <stem>_features.csv + <stem>_labels.csv share one row split (header kept on both sides)
        for f in names:
            if f.endswith("_features.csv") and f.startswith("train"):
                lab = f.replace("_features.csv", "_labels.csv")
                rows_f = open(os.path.join(folder, f)).read().splitlines()
                rows_l = open(os.path.join(folder, lab)).read().splitlines() if lab in names else None
                keep_l, keep_d = self._rows_split(len(rows_f) - 1, device_ratio)
                for dst, keep in ((out_l, keep_l), (out_d, keep_d)):
                    with open(os.path.join(dst, f), "w") as fh:
                        fh.write("\n".join([rows_f[0]] + [rows_f[1 + i] for i in keep]) + "\n")
                    if rows_l is not None:
                        with open(os.path.join(dst, lab), "w") as fh:
                            fh.write("\n".join([rows_l[0]] + [rows_l[1 + i] for i in keep]) + "\n")
                done.update({f, lab})
        # array shards of the built-in operators
        if "data.npz" in names:
            z = np.load(os.path.join(folder, "data.npz"))
            keep_l, keep_d = self._rows_split(len(z["y"]), device_ratio)
            for dst, keep in ((out_l, keep_l), (out_d, keep_d)):
                np.savez(os.path.join(dst, "data.npz"), **{k: z[k][keep] for k in z.files})
            done.add("data.npz")
        elif "x.npy" in names and "y.npy" in names:
            x, y = np.load(os.path.join(folder, "x.npy")), np.load(os.path.join(folder, "y.npy"))
            keep_l, keep_d = self._rows_split(len(y), device_ratio)
            for dst, keep in ((out_l, keep_l), (out_d, keep_d)):
                np.save(os.path.join(dst, "x.npy"), x[keep]); np.save(os.path.join(dst, "y.npy"), y[keep])
            done.update({"x.npy", "y.npy"})
        for f in names:                                    # everything else (test files, configs) goes to both sides
            if f not in done:
                src = os.path.join(folder, f)
                for dst in (out_l, out_d):
                    (shutil.copytree(src, os.path.join(dst, f), dirs_exist_ok=True) if os.path.isdir(src)
                     else shutil.copy2(src, os.path.join(dst, f)))

    @staticmethod
    def _zip_dir(root: str, dst_zip: str) -> None:
        os.makedirs(os.path.dirname(dst_zip) or ".", exist_ok=True)
        with zipfile.ZipFile(dst_zip, "w", zipfile.ZIP_DEFLATED) as z:
            for base, dirs, files in os.walk(root):
                for dname in dirs:
                    z.write(os.path.join(base, dname), os.path.relpath(os.path.join(base, dname), root))
                for f in files:
                    full = os.path.join(base, f)
                    z.write(full, os.path.relpath(full, root))

    def _publish(self, data: Dict[str, Any], local_zip: str, dst: str) -> bool:
        repo = self._repo_for(data)
        if repo is None:
            os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
            shutil.copyfile(local_zip, dst)
            return True
        return bool(repo.upload_file(local_zip, dst))

    # ---- entry -------------------------------------------------------------------------------------
    def split(self, data: Dict[str, Any], workdir: Optional[str] = None) -> bool:
        dl, dd = data.get("data_path_logical", ""), data.get("data_path_device", "")
        if not data.get("data_path") or not dl or not dd:
            return False
        al = data.get("allocation", {})
        n_log = sum(al.get("logical_simulation", []))
        n_dev = sum(al.get("device_simulation", [])) + sum(al.get("running_response", {}).get("nums", []))
        if n_log + n_dev == 0:
            return False
        tmp = workdir or tempfile.mkdtemp(prefix=f"olsb_split_{self.split_id}")
        try:
            src = self.download_data(data, tmp)
            if src is None:
                logger.error(self.task_id, "TaskMgr", "task_runner", f"[HybridDataSplitter]: download of {data.get('data_path')} failed")
                return False
            root = os.path.join(tmp, "src")
            if os.path.isdir(src):
                shutil.copytree(src, root, dirs_exist_ok=True)
            else:
                with zipfile.ZipFile(src) as z:
                    for m in z.namelist():
                        if os.path.isabs(m) or ".." in m.split("/"):
                            raise ValueError(f"zip entry escapes the dataset folder: {m}")
                    z.extractall(root)
            entries = sorted(e for e in os.listdir(root) if e != "dataconfig.json")
            out_l, out_d = os.path.join(tmp, "logical"), os.path.join(tmp, "device")
            os.makedirs(out_l); os.makedirs(out_d)
            if os.path.exists(os.path.join(root, "dataconfig.json")):
                shutil.copy2(os.path.join(root, "dataconfig.json"), out_l)
                shutil.copy2(os.path.join(root, "dataconfig.json"), out_d)
            if data.get("data_split_type", False):
                folders = [e for e in entries if os.path.isdir(os.path.join(root, e))]
                if len(folders) < n_log + n_dev:
                    logger.error(self.task_id, "TaskMgr", "task_runner",
                                 f"[HybridDataSplitter]: {len(folders)} client folders for {n_log + n_dev} simulated clients")
                    return False
                for part, dst in ((folders[:n_log], out_l), (folders[n_log:n_log + n_dev], out_d)):
                    for e in part:
                        shutil.copytree(os.path.join(root, e), os.path.join(dst, e))
            else:
                folders = [e for e in entries if os.path.isdir(os.path.join(root, e))]
                if len(folders) != 1:
                    # a flat zip (sample files at the top level) is treated as the one shared folder
                    if folders:
                        logger.error(self.task_id, "TaskMgr", "task_runner",
                                     f"[HybridDataSplitter]: unsplit data must hold exactly one sample folder, found {folders}")
                        return False
                    self._split_shared_folder(root, out_l, out_d, n_dev / (n_log + n_dev))
                else:
                    self._split_shared_folder(os.path.join(root, folders[0]), os.path.join(out_l, folders[0]),
                                              os.path.join(out_d, folders[0]), n_dev / (n_log + n_dev))
            ok = True
            for part, dst in ((out_l, dl), (out_d, dd)):
                z = os.path.join(tmp, os.path.basename(part) + ".zip")
                self._zip_dir(part, z)
                ok = self._publish(data, z, dst) and ok
            return ok
        except Exception as e:  # noqa: BLE001
            logger.error(self.task_id, "TaskMgr", "task_runner", f"[HybridDataSplitter.split]: {e}")
            return False
        finally:
            if workdir is None:
                shutil.rmtree(tmp, ignore_errors=True)


class TaskRunner:
    def __init__(self, task: TaskConfig, repo: Any, executor: JobExecutor, deviceflow: Optional[Any] = None,
                 phone_mgr: Optional[Any] = None, job_context: Optional[Dict[str, Any]] = None,
                 file_repo: Optional[Any] = None, dashboard: Optional[Any] = None):
        self._task, self._repo, self._executor = task, repo, executor
        self._deviceflow, self._phone_mgr = deviceflow, phone_mgr
        self._job_context = job_context or {}
        self._file_repo = file_repo
        self._dashboard = dashboard if dashboard is not None else getattr(deviceflow, "dashboard", None)

    def submit_job(self, submitter: JobSubmitter) -> Optional[str]:
        """``submit_rayjob``: stage operator code into a working dir, then hand the round-loop driver to the executor."""
        info = submitter.submission_info
        workdir = tempfile.mkdtemp(prefix=f"olsb_{info.get('task_id', 'task')}_")
        for op in info.get("operatorflow", {}).get("operators", []):
            if not get_operator_code(op, workdir, self._file_repo):
                shutil.rmtree(workdir, ignore_errors=True)
                return None
        ctx = dict(self._job_context)
        ctx["working_dir"] = workdir
        return self._executor.submit_job(info, ctx)

    def submit_phonejob(self, submitter: JobSubmitter) -> bool:
        if self._phone_mgr is None:
            return False
        try:
            return bool(self._phone_mgr.submitTask(json2deviceconfig(fix_device_task_json(submitter.submission_info))))
        except Exception as e:  # noqa: BLE001
            logger.error(self._task.task_id, "TaskMgr", "task_runner", f"[submit_phonejob]: {e}")
            return False

    def submit(self) -> Tuple[bool, Optional[str]]:
        tid = self._task.task_id
        tj = taskconfig2json(self._task)
        opt = HybridOptimizer(tj)
        data_list = opt.fix_data_parameters()
        # write the (possibly optimised) allocation back so status accounting sees it
        tj["target"]["data"] = [{k: v for k, v in d.items() if k not in ("data_path_logical", "data_path_device")} for d in data_list]
        def _reachable(d):
            remote = str(d.get("data_transfer_type", "FILE")).upper() != "FILE" and (self._file_repo is not None or d.get("repo_config"))
            return remote or os.path.exists(d.get("data_path", ""))
        both = [d for d in data_list if d.get("data_path_logical") and d.get("data_path_device")
                and d["data_path_logical"] != d.get("data_path") and _reachable(d)]
        for d in both:
            if not HybridDataSplitter(tid, d.get("task_type", "classification"), split_id=tid,
                                      file_repo=self._file_repo).split(d):
                logger.error(tid, "TaskMgr", "task_runner", "[submit]: hybrid data split failed")
                return False, None
        logical, device = JobSubmitter("logical_simulation"), JobSubmitter("device_simulation")
        logical.integrate_submission_info(tj, data_list)
        device.integrate_submission_info(tj, data_list)
        if DeviceflowRegister.check_use_deviceflow(tj):
            if not DeviceflowRegister(self._deviceflow).register_task(tid, logical.needs_submission, device.needs_submission):
                logger.error(tid, "TaskMgr", "task_runner", "[submit]: register task in deviceflow failed")
                return False, None
        job_id: Optional[str] = None
        if logical.needs_submission:
            # logical_target first: the job starts reporting results immediately
            target = {"logical_target": [{"name": d["name"], "simulation_target": d["simulation_target"]}
                                         for d in logical.submission_info["target"]["data"]]}
            if not self._repo.set_task_value(tid, "logical_target", json.dumps(target)):
                return False, None
            job_id = self.submit_job(logical)
            if job_id is None:
                return False, None
        if device.needs_submission:
            if not self.submit_phonejob(device):
                return False, job_id
            target = {"device_target": [{"name": d["name"], "simulation_target": d["simulation_target"]}
                                        for d in device.submission_info["target"]["data"]]}
            if not self._repo.set_task_value(tid, "device_target", json.dumps(target)):
                return False, job_id
        self.temp_write_table(data_list)
        return True, job_id

    def temp_write_table(self, data_list: List[Dict[str, Any]]) -> Dict[str, int]:
        """Dashboard row task_target(task_id, high, middle, low) (task_runner.py:226-256): written to the dashboard repo
        when the deployment has one (``DashboardRepo``; the reference uses a hard-coded MySQL DSN), always logged."""
        from ..deviceflow.dashboard import task_target_row
        tj = {"task_id": self._task.task_id, "target": {"data": data_list}}
        tot = {"high": 0, "middle": 0, "low": 0, **task_target_row(tj)}
        if self._dashboard is not None:
            try:
                self._dashboard.write_task_target(tj)
            except Exception as e:  # noqa: BLE001 - display only
                logger.warning(self._task.task_id, "TaskMgr", "task_runner", f"[task_target]: dashboard write failed: {e}")
        logger.info(self._task.task_id, "TaskMgr", "task_runner", f"[task_target]: {json.dumps(tot)}")
        return tot
