"""Environment-driven fault injection for the round loop.

``OLSB_FAULTS="kill:rank=1,round=2;corrupt:client=5,round=1;delay:round=3,seconds=0.5"``

* ``corrupt`` — poisons one client's replica with NaN before local training; the update must be masked out of the
  aggregate by the finite-check of the reduce path and counted as ``failed`` (reference analogue: a client subprocess
  exiting non-zero, ols_core/taskMgr/utils/utils_run_task.py:490-494).
* ``kill``    — the rank exits hard at the start of the round (actor / node loss); the job supervisor restarts the job
  from the last round checkpoint, optionally on fewer ranks (clients are re-sharded by world size).
* ``delay``   — sleeps at the start of the round (straggling rank / slow tick).

The reference only *simulates* faults at the message level (``drop_probability`` / ``drop_amounts``,
ols_core/deviceflow/non_grpc/strategy.py:275-311); those remain available through DeviceFlow strategies.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, field
from typing import Dict, List


@dataclass
class Fault:
    kind: str
    args: Dict[str, float] = field(default_factory=dict)


class FaultInjector:
    def __init__(self, spec: str = ""):
        self.faults: List[Fault] = []
        for item in filter(None, (spec or "").split(";")):
            kind, _, rest = item.partition(":")
            args = {}
            for kv in filter(None, rest.split(",")):
                k, _, v = kv.partition("=")
                args[k.strip()] = float(v)
            self.faults.append(Fault(kind.strip(), args))
        self.fired: List[str] = []

    @classmethod
    def from_env(cls) -> "FaultInjector":
        return cls(os.environ.get("OLSB_FAULTS", ""))

    def __bool__(self) -> bool:
        return bool(self.faults)

    def at_round_start(self, round_idx: int, rank: int) -> None:
        for f in self.faults:
            if int(f.args.get("round", -1)) != round_idx:
                continue
            if f.kind == "delay" and int(f.args.get("rank", rank)) == rank:
                time.sleep(f.args.get("seconds", 0.1))
                self.fired.append(f"delay@{round_idx}")
            elif f.kind == "kill" and int(f.args.get("rank", -1)) == rank:
                marker = os.environ.get("OLSB_FAULT_MARKER")
                if marker and os.path.exists(marker):
                    continue                       # already died once: the restarted job must make progress
                if marker:
                    open(marker, "w").write(f"rank {rank} killed at round {round_idx}\n")
                os._exit(17)

    def corrupted_clients(self, round_idx: int) -> List[int]:
        return [int(f.args["client"]) for f in self.faults
                if f.kind == "corrupt" and int(f.args.get("round", -1)) == round_idx and "client" in f.args]
