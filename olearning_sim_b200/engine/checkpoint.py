"""Round checkpoints and resume.

The reference has no model checkpointing: what survives a restart is the per-round model *file* the external
aggregator writes under ``model_update_style`` ("{task_id}_{current_round}_...", ols_core/taskMgr/utils/
utils_run_task.py:335-350) plus the progress cells ``logical_round`` / ``logical_operator`` / ``logical_result``
(ols_core/taskMgr/run_task.py:199-210), and a restarted job begins at round 0 again (run_task.py:215).

Here every rank writes ``{task_id}_{round}_rank{r}.ckpt`` (server model shard view, server optimiser state, RNG
counters = (seed, round), per-class accounting) from a background thread after staging to pinned host memory, keeps
the newest ``keep`` rounds, and ``load_latest`` + ``RoundRunner`` continue at ``round + 1``.  The file name follows the
reference's ``{task_id}_{current_round}`` template so external tooling that watches for the per-round model keeps working.
"""
from __future__ import annotations

import glob
import os
import re
import threading
from typing import Any, Dict, Optional

import torch


class CheckpointManager:
    def __init__(self, directory: str, task_id: str, rank: int = 0, keep: int = 2, every: int = 1,
                 template: str = "{task_id}_{current_round}_rank{rank}.ckpt", asynchronous: bool = True):
        self.dir, self.task_id, self.rank, self.keep, self.every = directory, task_id, rank, keep, max(1, every)
        self.template, self.asynchronous = template, asynchronous
        os.makedirs(directory, exist_ok=True)
        self._thread: Optional[threading.Thread] = None
        self._lock = threading.Lock()
        self.saved = 0

    def path(self, round_idx: int) -> str:
        name = (self.template.replace("{task_id}", self.task_id).replace("{current_round}", str(round_idx))
                .replace("{rank}", str(self.rank)))
        return os.path.join(self.dir, name)

    def _write(self, path: str, state: Dict[str, Any]) -> None:
        tmp = path + ".tmp"
        torch.save(state, tmp)
        os.replace(tmp, path)                      # atomic: a crash mid-write never corrupts the newest checkpoint
        with self._lock:
            self.saved += 1
        self._prune()

    def save(self, round_idx: int, state: Dict[str, Any]) -> Optional[str]:
        if (round_idx + 1) % self.every != 0:
            return None
        state = _to_host(state)                    # device -> (pinned) host copy happens on the caller's stream
        path = self.path(round_idx)
        self.wait()
        if self.asynchronous:
            self._thread = threading.Thread(target=self._write, args=(path, state), daemon=True)
            self._thread.start()
        else:
            self._write(path, state)
        return path

    def wait(self) -> None:
        if self._thread is not None:
            self._thread.join()
            self._thread = None

    def _rounds(self):
        pat = re.escape(self.template).replace(re.escape("{task_id}"), re.escape(self.task_id)) \
            .replace(re.escape("{current_round}"), r"(\d+)").replace(re.escape("{rank}"), str(self.rank))
        out = []
        for p in glob.glob(os.path.join(self.dir, "*")):
            m = re.fullmatch(pat, os.path.basename(p))
            if m:
                out.append((int(m.group(1)), p))
        return sorted(out)

    def _prune(self) -> None:
        rounds = self._rounds()
        for _, p in rounds[:-self.keep] if self.keep > 0 else []:
            try:
                os.remove(p)
            except OSError:
                pass

    def _load(self, p: str) -> Optional[Dict[str, Any]]:
        try:
            return torch.load(p, map_location="cpu", weights_only=True)       # tensors + plain containers only: no pickle code
        except Exception:  # noqa: BLE001 — torn file from a crash
            return None

    def load_latest(self, max_round: Optional[int] = None) -> Optional[Dict[str, Any]]:
        """Newest loadable checkpoint (of round <= ``max_round`` when given); a torn file falls back to the previous round."""
        self.wait()
        for r, p in reversed(self._rounds()):
            if max_round is not None and r > max_round:
                continue
            st = self._load(p)
            if st is not None:
                return st
        return None

    def newest_loadable_round(self) -> int:
        """Round index of the newest checkpoint of THIS rank that actually loads, -1 if none."""
        st = self.load_latest()
        return int(st["round"]) if st is not None and "round" in st else -1

    def load_agreed(self, group=None, world: int = 1) -> Optional[Dict[str, Any]]:
        """Multi-rank resume: writes are asynchronous and per rank, so a crash can leave rank 0 at round r and rank 1 at
        r-1 (or with a torn newest file).  Every rank therefore loads the MINIMUM over ranks of the newest loadable
        round (``keep >= 2`` retains it), so all ranks re-enter the collectives at the same round with the same state."""
        mine = self.newest_loadable_round()
        agreed = mine
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([mine], dtype=torch.int64)
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            agreed = int(t.item())
        if agreed < 0:
            return None
        st = self.load_latest(max_round=agreed)
        if st is None or int(st.get("round", -2)) != agreed:
            raise RuntimeError(f"rank {self.rank}: no loadable checkpoint of the agreed resume round {agreed} "
                               f"(newest here: {mine}); increase `keep`")
        return st


def _to_host(obj: Any) -> Any:
    if torch.is_tensor(obj):
        return obj.detach().to("cpu", non_blocking=False).clone()
    if isinstance(obj, dict):
        return {k: _to_host(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_host(v) for v in obj)
    return obj
