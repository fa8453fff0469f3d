"""Phase tracing of the round loop: CUDA-event timers (max over ranks), NVTX ranges, JSON trace per round.

The reference has no tracing at all (no timers, spans, NVTX or Ray timeline; ``PerformanceMgr.commitJob``
exists only as a proto).  ``PhaseTimer`` brackets the phases of a round (mask / broadcast / local steps / reduce) with CUDA
events on the launching stream — no host synchronisation is added inside the round; the events are resolved after the
round's single read-back — and emits NVTX ranges so the same phases show up in an external timeline.
"""
from __future__ import annotations

import contextlib
import json
import time
from typing import Dict, List, Optional

import torch


class PhaseTimer:
    def __init__(self, enabled: bool = True, device: Optional[torch.device] = None, nvtx: bool = True):
        self.enabled = enabled
        self.cuda = torch.cuda.is_available() and (device is None or device.type == "cuda")
        self.nvtx = nvtx and self.cuda
        self._open: List = []
        self._done: List = []
        self.last_min: Dict[str, float] = {}

    @contextlib.contextmanager
    def phase(self, name: str):
        if not self.enabled:
            yield
            return
        if self.nvtx:
            torch.cuda.nvtx.range_push(name)
        if self.cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        try:
            yield
        finally:
            if self.cuda:
                e1.record()
                self._done.append((name, e0, e1, None))
            else:
                self._done.append((name, None, None, (time.perf_counter() - t0) * 1e3))
            if self.nvtx:
                torch.cuda.nvtx.range_pop()

    def collect(self, group=None, world: int = 1) -> Dict[str, float]:
        """Resolve the events (call after the round's stream synchronisation); max over ranks per phase.  ``last_min``
        keeps the min over ranks: a phase that starts with a cross-rank barrier (the reduce) includes, on every rank but
        the slowest, the wait for that rank — its min over ranks is the time of the work itself."""
        out: Dict[str, float] = {}
        for name, e0, e1, host_ms in self._done:
            ms = host_ms if host_ms is not None else e0.elapsed_time(e1)
            out[name] = out.get(name, 0.0) + ms
        self._done.clear()
        if world > 1 and out:
            import torch.distributed as dist
            keys = sorted(out)
            t = torch.tensor([out[k] for k in keys], dtype=torch.float64, device="cuda" if self.cuda else "cpu")
            tmin = t.clone()
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN, group=group)
            out = {k: float(v) for k, v in zip(keys, t.tolist())}
            self.last_min = {k: float(v) for k, v in zip(keys, tmin.tolist())}
        else:
            self.last_min = dict(out)
        return out


class TraceWriter:
    """One JSON object per round appended to a file (``--profile`` output of the round-loop driver)."""

    def __init__(self, path: str):
        self.path = path

    def write(self, record: Dict) -> None:
        with open(self.path, "a") as f:
            f.write(json.dumps(record) + "\n")
