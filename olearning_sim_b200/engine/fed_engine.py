"""FedEngine — the public round-loop API of the simulator (one instance per rank / GPU).

    eng = FedEngine(FedConfig(model="cnn2", num_clients=1024, local_steps=5), rank, world)
    eng.load_synthetic()
    for r in range(rounds):
        res = eng.run_round(r)            # broadcast -> masked local training -> FedAvg reduce + server update

It replaces the reference's ``RayRunner.run`` round loop + ``Actor.run`` client loop
(ols_core/taskMgr/run_task.py:212-322, ols_core/taskMgr/utils/utils_run_task.py:481-577): clients are rows of packed
device tensors sharded contiguously over ranks (``shard_range`` = run_task.py:85-105), DeviceFlow's
online/offline/drop behaviour is a per-round mask (``deviceflow_mask_kernel`` / ``philox.round_mask``), the model
"download" is ``bcast_scatter`` and the update path is ``fedavg_partial`` + ``fedavg_finish``.

Backends:  ``native``  hand-written sm_100a kernels (CNN-2 family, CUDA only)
           ``torch``   generic ``torch.func`` backend for every model family; CPU (gloo) or CUDA; collectives via
                       ``torch.distributed`` — this is the library path / baseline, never the product path on GPU.
"""
from __future__ import annotations

import os

import math
import time
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from ..deviceflow import philox
from ..models.zoo import ModelSpec, get_model
from ..ops import reference
from ..ops.layout import CNN2
from ..parallel.comm import shard_range


NATIVE_FAMILIES = ("cnn2", "mlp2", "mlp", "resnet18", "bert_mini")      # hand-written sm_100a local step


@dataclass
class FedConfig:
    model: str = "cnn2"
    model_kwargs: Dict[str, Any] = field(default_factory=dict)
    num_clients: int = 1024
    local_steps: int = 5
    batch: int = 32
    shard: int = 64                  # samples per client
    lr: float = 0.05
    algorithm: str = "fedavg"        # fedavg | fedprox
    mu: float = 0.0                  # FedProx proximal coefficient
    server_opt: str = "avg"          # avg | momentum | adam
    server_lr: float = 1.0
    seed: int = 0
    p_online: float = 1.0            # per-round probability that a client is online / sampled
    drop_probability: float = 0.0    # DeviceFlow drop_simulation (update lost in flight)
    hetero_steps: bool = False       # stragglers: per-client local-step budget in [1, local_steps]
    backend: str = "auto"            # auto | native | torch
    device: str = "auto"
    use_cuda_graph: bool = True
    dynamic_nums: int = 0            # tolerated failed clients per round (reference: dynamic_nums)
    # DeviceFlow integration: the operator's strategy JSON (real_time_dispatch / flow_dispatch) drives
    # which client updates reach the aggregator each round; `online_trace` is a safe expression in t = round index
    # (e.g. "0.6+0.4*math.sin(t/4)") scaling p_online — the diurnal online/offline + access-spike trace.
    deviceflow_strategy: Optional[str] = None
    online_trace: Optional[str] = None
    wave: Optional[int] = None       # clients with resident activations at once (None = auto from free HBM)
    amp_bf16: bool = False           # hybrid / torch backends: run the library forward/backward under bf16 autocast
    torch_chunk: Optional[int] = None  # torch backend: clients per vmap chunk (None = sized from the parameter count)
    profile: bool = False            # CUDA-event phase timing (mask / broadcast / local / reduce), max over ranks

    def __post_init__(self):
        if self.algorithm == "fedprox" and self.mu <= 0:
            self.mu = 0.01
        if self.algorithm == "fedavg":
            self.mu = 0.0


@dataclass
class RoundResult:
    round: int
    participants: int
    dropped: int
    late: int
    arrived: int
    failed: int
    loss: float
    client_steps: int
    wall_s: float = 0.0
    extras: Dict[str, Any] = field(default_factory=dict)


def synthetic_client_data(model: str, client_ids: np.ndarray, shard: int, seed: int, spec: ModelSpec):
    """Synthetic non-IID shards: client i draws labels from a 2-class-dominant distribution and inputs from a
    label-dependent pattern, generated independently per *global* client id (any rank count sees the same data)."""
    n = len(client_ids)
    ncls = spec.num_classes
    labels = np.empty((n, shard), np.int32)
    if spec.input_dtype == torch.int64:
        xs = np.empty((n, shard) + spec.input_shape, np.int64)
    elif model in ("cnn2", "resnet18"):
        xs = np.empty((n, shard, 32, 32, 3), np.uint8)
    else:
        xs = np.empty((n, shard) + spec.input_shape, np.float32)
    for j, cid in enumerate(client_ids):
        rng = np.random.default_rng([seed, int(cid)])
        major = rng.choice(ncls, size=min(2, ncls), replace=False)
        probs = np.full(ncls, 0.2 / max(ncls - len(major), 1))
        probs[major] = 0.8 / len(major)
        probs /= probs.sum()
        y = rng.choice(ncls, size=shard, p=probs).astype(np.int32)
        labels[j] = y
        if xs.dtype == np.uint8:
            base = rng.integers(0, 256, size=(shard, 32, 32, 3), dtype=np.uint8)
            tint = ((y[:, None, None, None].astype(np.int32) * 23) % 97).astype(np.uint8)
            xs[j] = (base // 2 + tint)
        elif xs.dtype == np.int64:
            xs[j] = rng.integers(0, 1000, size=(shard,) + spec.input_shape) + y[:, None] * 7
        else:
            xs[j] = rng.standard_normal((shard,) + spec.input_shape).astype(np.float32) + y[:, None] * 0.3
    return xs, labels


def native_round_rec() -> int:
    from ..ops.native import ROUND_REC
    return ROUND_REC


def merge_opt_shards(parts: List[torch.Tensor], num_params: int) -> torch.Tensor:
    """Concatenate equal-length rank shards and cut the tail padding: the full [P] optimiser-state vector."""
    return torch.cat([p.reshape(-1) for p in parts])[:num_params].clone()


def take_opt_shard(full: torch.Tensor, rank: int, shard_len: int, out_len: int) -> torch.Tensor:
    """This rank's slice of a full optimiser-state vector under the CURRENT sharding, zero-padded to ``out_len``
    (the last rank's shard extends past P).  Accepts the old per-rank format only when the sharding is unchanged."""
    full = full.reshape(-1)
    out = torch.zeros(out_len, dtype=full.dtype)
    s0 = rank * shard_len
    n = max(0, min(shard_len, full.numel() - s0, out_len))
    if n > 0:
        out[:n] = full[s0:s0 + n]
    return out


class FedEngine:
    def __init__(self, cfg: FedConfig, rank: int = 0, world: int = 1, group: Optional[Any] = None):
        self.cfg, self.rank, self.world, self.group = cfg, rank, world, group
        dev = cfg.device
        if dev == "auto":
            dev = f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(dev)
        backend = cfg.backend
        native_ok = self.device.type == "cuda" and cfg.model in NATIVE_FAMILIES and cfg.batch == 32
        if backend == "auto":
            # native: hand-written kernels end to end; hybrid: library forward/backward + native FL kernels (broadcast,
            # masks, fused client SGD, P2P reduce + server optimiser); torch: pure library path (CPU / baseline)
            backend = "native" if native_ok else ("hybrid" if self.device.type == "cuda" else "torch")
        if backend == "native" and not native_ok:
            raise RuntimeError(f"the native backend needs CUDA, batch 32 and one of {NATIVE_FAMILIES}")
        if backend == "hybrid" and self.device.type != "cuda":
            raise RuntimeError("the hybrid backend needs CUDA")
        self.backend = backend
        self.uses_ws = backend in ("native", "hybrid")      # state lives in a GPU workspace with the native FL kernels
        self.spec = get_model(cfg.model, **cfg.model_kwargs)
        self.c0, self.c1 = shard_range(cfg.num_clients, world, rank)
        self.C = self.c1 - self.c0
        self.client_ids = np.arange(self.c0, self.c1)
        self.round_log: List[RoundResult] = []
        from ..utils.faults import FaultInjector
        from ..utils.tracing import PhaseTimer
        self.faults = FaultInjector.from_env()
        self.timer = PhaseTimer(enabled=cfg.profile, device=self.device)
        # per-client static attributes
        rng = np.random.default_rng([cfg.seed, 991])
        all_steps = rng.integers(1, cfg.local_steps + 1, size=cfg.num_clients).astype(np.int32) if cfg.hetero_steps \
            else np.full(cfg.num_clients, cfg.local_steps, np.int32)
        self.local_steps_np = all_steps[self.c0:self.c1]
        self.n_samples_np = np.full(self.C, cfg.shard, np.int32)
        # single source of truth for availability: one probability per GLOBAL client (every rank holds the full vector so
        # the flow schedule, which orders all participants, is computed identically everywhere)
        self.p_online_np = np.full(cfg.num_clients, cfg.p_online, np.float32)
        self.class_names: List[str] = []
        self.device_flow_ticks = True       # workspace backends: rank flow-dispatch arrivals on the GPU (False: host twin)
        self.class_of_client = np.zeros(cfg.num_clients, np.int32)
        self._flow_cache: Dict[str, Any] = {}
        if backend == "native":
            self._init_native()
        elif backend == "hybrid":
            self._init_hybrid()
        else:
            self._init_torch()

    # ==================================================================================================
    # DeviceFlow -> per-round masks
    # ==================================================================================================
    def set_client_classes(self, classes: List[str], p_online: Optional[np.ndarray] = None,
                           local_steps: Optional[np.ndarray] = None) -> None:
        """Per-client (dataset, device class) labels in global client order, with optional per-client online
        probabilities / local-step budgets (device classes -> heterogeneity).  Drives the per-class success / failed
        accounting of every round (reference: run_task.py:149-210)."""
        assert len(classes) == self.cfg.num_clients
        names: List[str] = []
        idx = np.empty(len(classes), np.int32)
        for i, c in enumerate(classes):
            if c not in names:
                names.append(c)
            idx[i] = names.index(c)
        self.class_names, self.class_of_client = names, idx
        if p_online is not None:
            self.p_online_np = np.asarray(p_online, np.float32).copy()
        if local_steps is not None:
            steps = np.minimum(np.asarray(local_steps, np.int32), self.cfg.local_steps)
            self.local_steps_np = steps[self.c0:self.c1].copy()
            self.cfg.hetero_steps = bool((steps != self.cfg.local_steps).any()) or self.cfg.hetero_steps
        if self.uses_ws:
            d = self.device
            self.local_steps_t = torch.from_numpy(self.local_steps_np).to(d)
            self.p_online_t = torch.from_numpy(self.p_online_np[self.c0:self.c1].copy()).to(d)
            self.class_id_t = torch.from_numpy(idx[self.c0:self.c1].copy()).to(d)

    def _flow_schedule(self, round_idx: int):
        """(amounts per slot, drop lists) of the operator's flow strategy for this round.  Relative-time strategies do
        not depend on the round: the JSON is parsed and the rate curves integrated ONCE per task (on the GPU through
        ``rate_curve_area_kernel`` when the workspace backend is active); only the drop draws are per round."""
        import random as _random
        from ..deviceflow.strategy import Strategy
        st = self.cfg.deviceflow_strategy
        rng = _random.Random(f"{self.cfg.seed}:{round_idx}")
        flow_id = f"engine_train_{round_idx}"
        if not Strategy.is_round_invariant(st):
            _, amount, drops = Strategy.flow_strategy_analysis(st, flow_id, rng=rng)
            return amount, drops
        plan = self._flow_cache.get("plan")
        if plan is None:
            area_fn = None
            if self.uses_ws:
                from ..ops import native

                def area_fn(y):                      # y: float64 [bins, 101] samples of the rate function
                    yt = torch.from_numpy(np.ascontiguousarray(y, np.float64)).to(self.device)
                    out = torch.empty(y.shape[0], dtype=torch.float64, device=self.device)
                    native.rate_curve_area(yt, out)
                    return out.cpu().numpy()
            plan = Strategy.flow_plan(st, flow_id, area_fn=area_fn)
            self._flow_cache["plan"] = plan
        return plan["amount"], Strategy.draw_drops(plan, rng)

    def _device_flow(self, round_idx: int, scale: float) -> bool:
        """Flow strategy of a workspace backend: fill ``self.tick_dev`` on the device (flow_*_kernel) from the per-task
        plan.  Host work per round: nothing, or — when the strategy drops messages — drawing the drop lists with the
        round's seeded RNG (the reference's algorithm) and uploading them as a per-message byte map."""
        cfg = self.cfg
        from ..deviceflow.strategy import Strategy
        st = cfg.deviceflow_strategy
        if not (self.uses_ws and st and not Strategy.check_real_time_dispatch(st) and Strategy.is_round_invariant(st)):
            return False
        from ..ops import native
        amount, drops = self._flow_schedule(round_idx)
        fc = self._flow_cache
        if "dev" not in fc:
            d, n = self.device, cfg.num_clients
            cum = np.cumsum(np.asarray(amount, np.int64)).astype(np.int32)
            fc["dev"] = {"cum": torch.from_numpy(cum).to(d) if len(cum) else torch.zeros(1, dtype=torch.int32, device=d),
                         "nslots": len(amount), "total": int(cum[-1]) if len(cum) else 0,
                         "p_all": torch.from_numpy(self.p_online_np.copy()).to(d),
                         "scratch": {"keys": torch.empty(n, dtype=torch.int32, device=d), "hist": torch.zeros(4096, dtype=torch.int32, device=d),
                                     "start": torch.zeros(4097, dtype=torch.int32, device=d), "cursor": torch.zeros(4096, dtype=torch.int32, device=d),
                                     "items": torch.empty(max(n, 1), dtype=torch.int32, device=d)},
                         "dropmap": None, "drop_host": None}
        dv = fc["dev"]
        dropmap = None
        if any(len(x) for x in drops):
            if dv["dropmap"] is None:
                dv["dropmap"] = torch.zeros(max(dv["total"], 1), dtype=torch.uint8, device=self.device)
                dv["drop_host"] = torch.zeros(max(dv["total"], 1), dtype=torch.uint8).pin_memory()
            hm = dv["drop_host"].numpy()
            hm[:] = 0
            k = 0
            for slot, nmsg in enumerate(amount):
                if slot < len(drops) and drops[slot]:
                    idx = np.asarray(drops[slot], np.int64)
                    hm[k + idx[idx < nmsg]] = 1
                k += nmsg
            dv["dropmap"].copy_(dv["drop_host"], non_blocking=True)
            dropmap = dv["dropmap"]
        native.flow_ticks(cfg.num_clients, self.c0, self.C, cfg.seed, round_idx, dv["p_all"], float(scale), dv["scratch"], dv["cum"],
                          dv["nslots"], dropmap, self.tick_dev)
        self.ws.kernel_launches += 4
        self._max_tick = max(dv["nslots"] - 1, 0)
        return True

    def _round_flow(self, round_idx: int):
        """(p_online_scale, drop_probability, arrive_tick[C_local] | None, max_tick) for this round.
        Flow strategies: the k-th *finishing* participant (order = Philox stream 2) is message k on the shelf; the
        schedule's slots / drop lists / discarded leftovers then decide whether its update reaches the aggregator."""
        cfg = self.cfg
        scale = 1.0
        if cfg.online_trace:
            from ..deviceflow import safe_expr
            scale = float(min(max(safe_expr.evaluate(cfg.online_trace, float(round_idx)), 0.0), 1.0))
        drop_p, ticks, max_tick = cfg.drop_probability, None, 0
        if cfg.deviceflow_strategy:
            from ..deviceflow.strategy import Strategy, schedule_to_ticks
            st = cfg.deviceflow_strategy
            if Strategy.check_real_time_dispatch(st):
                _, drop_p = Strategy.real_time_strategy_analysis(st)
            else:
                amount, drops = self._flow_schedule(round_idx)
                n = cfg.num_clients
                u = philox.round_uniforms(n, 0, cfg.seed, round_idx)
                part = u[0] < self.p_online_np * scale
                order = np.argsort(np.where(part, u[2], 2.0), kind="stable")        # finishing order of participants
                slot_of_msg = schedule_to_ticks(amount, drops, int(part.sum()))
                t_all = np.full(n, -1, np.int32)
                t_all[order[: int(part.sum())]] = slot_of_msg
                ticks = t_all[self.c0:self.c1]
                max_tick = max(len(amount) - 1, 0)
        return scale, float(drop_p), ticks, max_tick

    # ==================================================================================================
    # native backend
    # ==================================================================================================
    def _init_native(self) -> None:
        from ..parallel.comm import SymmetricHeap
        heap = SymmetricHeap(group=self.group, device=self.device)
        if self.cfg.model == "cnn2":
            from .cnn2_native import Cnn2Workspace
            self.ws = Cnn2Workspace(self.C, self.cfg.shard, device=self.device, heap=heap, wave=self.cfg.wave)
            self._g0 = CNN2.init(self.cfg.seed)
        elif self.cfg.model == "resnet18":
            from .resnet_native import ResNetWorkspace
            self.ws = ResNetWorkspace(self.C, self.cfg.shard, self.spec, device=self.device, heap=heap, wave=self.cfg.wave,
                                      groups=int(self.cfg.model_kwargs.get("groups", 2)))
            self._g0 = self.ws.L.pack(self.spec.init(self.cfg.seed))
        elif self.cfg.model == "bert_mini":
            from .bert_native import BertWorkspace
            self.ws = BertWorkspace(self.C, self.cfg.shard, self.spec, device=self.device, heap=heap, wave=self.cfg.wave)
            self._g0 = self.ws.L.pack(self.spec.init(self.cfg.seed))
        else:
            from .mlp_native import MlpLayout, MlpWorkspace
            ws_ = [shape for name, shape in self.spec.shapes if name.endswith(".weight")]        # fc1 .. fcN
            lay = MlpLayout(ws_[0][1], [w[0] for w in ws_[:-1]], ws_[-1][0])
            self.ws = MlpWorkspace(self.C, self.cfg.shard, lay, device=self.device, heap=heap)
            self._g0 = lay.pack(self.spec.init(self.cfg.seed))
        d = self.device
        S = self.cfg.local_steps
        self.p_online_t = torch.from_numpy(self.p_online_np[self.c0:self.c1].copy()).to(d)
        self.p_online_round = self.p_online_t.clone()
        self.class_id_t = torch.zeros(self.C, dtype=torch.int32, device=d)
        self.tick_host = torch.zeros(self.C, dtype=torch.int32).pin_memory()
        self.tick_dev = torch.zeros(self.C, dtype=torch.int32, device=d)
        self.local_steps_t = torch.from_numpy(self.local_steps_np).to(d)
        self.participate = torch.ones(self.C, dtype=torch.uint8, device=d)
        self.step_active = torch.ones(S, self.C, dtype=torch.uint8, device=d)
        self.agg_weight = torch.zeros(self.C, device=d)
        self.counters = torch.zeros(4, dtype=torch.int32, device=d)
        self.result_host = torch.zeros(native_round_rec() + 8).pin_memory()
        self.ws.set_global(self._g0)
        self.P = self.ws.P
        self._graphs: Dict[bool, Any] = {}
        self._graph_ok, self._graph_warm, self.graph_error = True, False, None

    def _init_hybrid(self) -> None:
        """Library forward/backward (chunked ``vmap(grad_and_value)``) on replicas held in a ``FlatWorkspace``."""
        from torch.func import grad_and_value, vmap
        from ..parallel.comm import SymmetricHeap
        from .flat_native import FlatWorkspace
        heap = SymmetricHeap(group=self.group, device=self.device)
        self.P = self.spec.padded_params
        self.ws = FlatWorkspace(self.C, self.P, device=self.device, heap=heap)
        self.ws.n_samples = torch.full((self.C,), self.cfg.shard, dtype=torch.int32, device=self.device)
        self._g0 = self.spec.init(self.cfg.seed)
        d, S = self.device, self.cfg.local_steps
        self.p_online_t = torch.from_numpy(self.p_online_np[self.c0:self.c1].copy()).to(d)
        self.p_online_round = self.p_online_t.clone()
        self.class_id_t = torch.zeros(self.C, dtype=torch.int32, device=d)
        self.tick_host = torch.zeros(self.C, dtype=torch.int32).pin_memory()
        self.tick_dev = torch.zeros(self.C, dtype=torch.int32, device=d)
        self.local_steps_t = torch.from_numpy(self.local_steps_np).to(d)
        self.participate = torch.ones(self.C, dtype=torch.uint8, device=d)
        self.step_active = torch.ones(S, self.C, dtype=torch.uint8, device=d)
        self.agg_weight = torch.zeros(self.C, device=d)
        self.counters = torch.zeros(4, dtype=torch.int32, device=d)
        self.result_host = torch.zeros(native_round_rec() + 8).pin_memory()
        self.ws.set_global(self._g0)
        self._graphs, self._graph_ok, self._graph_warm, self.graph_error = {}, False, False, None
        self.x = self.y = None
        spec = self.spec

        amp = self.cfg.amp_bf16

        def loss_fn(flat, x, y):
            if amp:           # fp32 master replicas, bf16 tensor-core compute, fp32 loss / gradients
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return spec.loss(flat, x, y).float()
            return spec.loss(flat, x, y)

        chunk = self.cfg.torch_chunk or max(1, min(self.C, int(20e9 / (max(self.P, 1) * 48))))
        self._chunk = chunk if chunk < self.C else None
        self._grad_loss_fn = vmap(grad_and_value(loss_fn), in_dims=(0, 0, 0), randomness="different",
                                  chunk_size=self._chunk)

    def _hybrid_local_steps(self, need_mask: bool) -> None:
        cfg, ws, B = self.cfg, self.ws, self.cfg.batch
        for s in range(cfg.local_steps):
            off = (s * B) % cfg.shard
            g, loss = self._grad_loss_fn(ws.master, self.x[:, off:off + B], self.y[:, off:off + B])
            act = self.step_active[s] if need_mask else None
            loss = loss.detach()
            bad = ~torch.isfinite(loss)
            if act is not None:
                on = act.bool()
                loss, bad = torch.where(on, loss, torch.zeros_like(loss)), bad & on
            ws.loss_sum += torch.where(bad, torch.zeros_like(loss), loss)
            ws.nonfinite |= bad.to(torch.int32)
            ws.sgd_step(g.contiguous(), lr=cfg.lr, mu=cfg.mu, active=act)

    def _native_round_body(self, round_idx: int) -> None:
        """All device work of one round on the current stream; no host synchronisation."""
        from ..ops import native
        cfg, ws = self.cfg, self.ws
        ws.reset_round_stats()
        on_device = False
        if cfg.deviceflow_strategy and self.device_flow_ticks:
            scale = 1.0
            if cfg.online_trace:
                from ..deviceflow import safe_expr
                scale = float(min(max(safe_expr.evaluate(cfg.online_trace, float(round_idx)), 0.0), 1.0))
            on_device = self._device_flow(round_idx, scale)
        if on_device:
            drop_p, ticks, max_tick = cfg.drop_probability, None, self._max_tick
        else:
            scale, drop_p, ticks, max_tick = self._round_flow(round_idx)
        sampled = bool(self.p_online_np.size and float(self.p_online_np.min()) * scale < 1.0)
        if sampled:
            self.p_online_round.copy_(self.p_online_t).mul_(scale)
        tick_t = self.tick_dev if on_device else None
        if ticks is not None:
            self.tick_host.copy_(torch.from_numpy(ticks))
            self.tick_dev.copy_(self.tick_host, non_blocking=True)
            tick_t = self.tick_dev
        native.deviceflow_mask(self.C, self.c0, cfg.seed, round_idx, self.p_online_round if sampled else None,
                               drop_p, tick_t, max_tick, ws.n_samples, self.local_steps_t if cfg.hetero_steps else None,
                               cfg.local_steps, self.participate, self.step_active, self.agg_weight, self.counters)
        ws.kernel_launches += 1
        need_mask = sampled or cfg.hetero_steps
        with self.timer.phase("broadcast"):
            ws.broadcast(self.participate if sampled else None)
        for cid in self.faults.corrupted_clients(round_idx) if self.faults else []:
            if self.c0 <= cid < self.c1:
                ws.master[cid - self.c0].fill_(float("nan"))             # poisoned replica: must be masked by the reduce
                if hasattr(ws, "master_bf16"):
                    ws.master_bf16[cid - self.c0].fill_(float("nan"))
                elif hasattr(ws, "wb"):
                    ws.wb[cid - self.c0].fill_(float("nan"))
                elif hasattr(ws, "w1_fwd"):
                    ws.w1_fwd[cid - self.c0].fill_(float("nan"))
        with self.timer.phase("local_steps"):
            if self.backend == "hybrid":
                self._hybrid_local_steps(need_mask)
            else:
                self._local_steps(need_mask)
        with self.timer.phase("reduce_update"):
            ws.reduce_and_update(self.agg_weight, opt=cfg.server_opt, server_lr=cfg.server_lr,
                                 accounting=(self.participate, self.step_active, self.counters, cfg.local_steps,
                                             self.class_id_t if self.class_names else None, len(self.class_names)))

    def _local_steps(self, need_mask: bool) -> None:
        """The S x (6 kernels x waves) local-step block.  Every kernel argument of the block is round-invariant (buffer
        addresses, step offsets, lr / mu; the masks live in fixed buffers), so after a warm-up round it is captured once
        into a CUDA graph per mask variant and replayed: one launch per round instead of 30+ host-side launches."""
        cfg, ws = self.cfg, self.ws

        def eager():
            for s in range(cfg.local_steps):
                ws.local_step(s, lr=cfg.lr, mu=cfg.mu, active=self.step_active[s] if need_mask else None)

        if not (cfg.use_cuda_graph and self._graph_ok):
            return eager()
        entry = self._graphs.get(need_mask)
        if entry is None:
            if not self._graph_warm:
                self._graph_warm = True          # first round runs eagerly (lazy attribute setup, first-use checks)
                return eager()
            try:
                before = ws.kernel_launches
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    eager()
                entry = (g, ws.kernel_launches - before)
                ws.kernel_launches = before       # capture records launches, it does not execute them
                self._graphs[need_mask] = entry
            except Exception as e:  # noqa: BLE001 - a failed capture must not take the job down
                self._graph_ok = False
                self.graph_error = f"{type(e).__name__}: {e}"
                torch.cuda.synchronize(self.device)
                return eager()
        entry[0].replay()
        ws.kernel_launches += entry[1]

    def _native_finish(self, round_idx: int) -> RoundResult:
        # the job-wide 8-float record was assembled on the device (round_stats_kernel per rank, summed over ranks
        # inside fedavg_finish_kernel): one 32-byte D2H copy per round is the "loss or metric" read-back
        self.result_host.copy_(self.ws.round_out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        vals = self.result_host.tolist()
        lost = int(vals[native_round_rec()])
        if lost:
            # a cross-GPU barrier of fedavg_finish_kernel gave up on a peer: the model was not updated, this process and its
            # CUDA context are intact — the job layer restarts the task on the surviving GPUs from the last checkpoint
            from ..parallel.comm import PeerLostError
            raise PeerLostError(lost - 1, round_idx)
        part_n = int(vals[1])
        res = RoundResult(round=round_idx, participants=int(vals[2]), dropped=int(vals[3]), late=int(vals[4]),
                          arrived=int(vals[5]), failed=int(vals[6]),
                          loss=float(vals[0] / max(part_n, 1)), client_steps=int(vals[7]))
        if self.class_names:
            res.extras["class_failed"] = {n: int(vals[8 + 2 * k]) for k, n in enumerate(self.class_names)}
            res.extras["class_participants"] = {n: int(vals[9 + 2 * k]) for k, n in enumerate(self.class_names)}
        return res

    # ==================================================================================================
    # generic torch backend (CPU/gloo or CUDA; library path)
    # ==================================================================================================
    def _init_torch(self) -> None:
        from torch.func import functional_call, grad, vmap  # noqa: F401
        d = self.device
        self.P = self.spec.padded_params
        g0 = self.spec.init(self.cfg.seed).to(d)
        self.global_w = g0.clone()
        self.master = g0[None, :].repeat(self.C, 1)
        self.server_state: Dict[str, Any] = {}
        self.loss_sum = torch.zeros(self.C, device=d)
        self.x = None
        self.y = None
        spec = self.spec

        amp = self.cfg.amp_bf16 and d.type == "cuda"

        def loss_fn(flat, x, y):
            if amp:           # library arm at the product's precision: bf16 tensor-core compute, fp32 masters / loss
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return spec.loss(flat, x, y).float()
            return spec.loss(flat, x, y)

        from torch.func import grad_and_value
        # one pass yields gradient and loss; clients are processed in chunks so activations of big models fit in HBM
        chunk = self.cfg.torch_chunk or max(1, min(self.C, int(20e9 / (max(self.P, 1) * 48))))
        self._chunk = chunk if chunk < self.C else None
        self._grad_loss_fn = vmap(grad_and_value(loss_fn), in_dims=(0, 0, 0), randomness="different",
                                  chunk_size=self._chunk)

    def _torch_round(self, round_idx: int) -> RoundResult:
        cfg = self.cfg
        scale, drop_p, ticks, max_tick = self._round_flow(round_idx)
        sampled = bool(self.p_online_np.size and float(self.p_online_np.min()) * scale < 1.0)
        m = philox.round_mask(self.C, self.c0, cfg.seed, round_idx,
                              (self.p_online_np[self.c0:self.c1] * np.float32(scale)).astype(np.float32) if sampled else None,
                              drop_p, ticks, max_tick, self.n_samples_np,
                              self.local_steps_np if cfg.hetero_steps else None, cfg.local_steps)
        d = self.device
        part = torch.from_numpy(m["participate"].astype(bool)).to(d)
        step_active = torch.from_numpy(m["step_active"].astype(bool)).to(d)
        weight = torch.from_numpy(m["agg_weight"]).to(d)
        # broadcast: refresh replicas of sampled clients
        with self.timer.phase("broadcast"):
            self.master = torch.where(part[:, None], self.global_w[None, :].expand_as(self.master), self.master)
        self.loss_sum.zero_()
        for cid in self.faults.corrupted_clients(round_idx) if self.faults else []:
            if self.c0 <= cid < self.c1:
                self.master[cid - self.c0] = float("nan")
        B = cfg.batch
        for s in range(cfg.local_steps):
            off = (s * B) % cfg.shard
            xb, yb = self.x[:, off:off + B], self.y[:, off:off + B]
            w = self.master
            g, loss = self._grad_loss_fn(w, xb, yb)
            self.loss_sum += loss.detach() * step_active[s]
            if cfg.mu > 0:
                g = g + cfg.mu * (w - self.global_w[None, :])
            self.master = torch.where(step_active[s][:, None], w - cfg.lr * g, w)
        finite = torch.isfinite(self.loss_sum)
        weight = torch.where(finite, weight, torch.zeros_like(weight))
        delta, wsum = reference.fedavg_delta(self.master, self.global_w, weight)
        pack = torch.cat([delta, torch.tensor([wsum], device=d, dtype=delta.dtype)])
        okm = finite & part
        stats = torch.tensor([float(torch.where(okm, self.loss_sum, torch.zeros_like(self.loss_sum)).sum() / cfg.local_steps), float(okm.sum()),
                              *[float(v) for v in m["counters"]], float((~finite).sum()), float(step_active.sum())],
                             dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(pack, group=self.group)
            st = stats.to(d) if d.type == "cuda" else stats
            dist.all_reduce(st, group=self.group)
            stats = st.cpu()
        self.global_w = reference.server_update(self.global_w, pack[:-1], float(pack[-1]), cfg.server_opt,
                                                self.server_state, cfg.server_lr)
        n = max(int(stats[1]), 1)
        res = RoundResult(round=round_idx, participants=int(stats[2]), dropped=int(stats[3]), late=int(stats[4]),
                          arrived=int(stats[5]), failed=int(stats[6]), loss=float(stats[0]) / n,
                          client_steps=int(stats[7]))
        if self.class_names:
            K = len(self.class_names)
            cls = torch.from_numpy(self.class_of_client[self.c0:self.c1].astype(np.int64)).to(d)
            cnt = torch.stack([torch.bincount(cls[~finite], minlength=K), torch.bincount(cls[part], minlength=K)]).double()
            if self.world > 1:
                dist.all_reduce(cnt, group=self.group)
            cnt = cnt.cpu()
            res.extras["class_failed"] = {nm: int(cnt[0, k]) for k, nm in enumerate(self.class_names)}
            res.extras["class_participants"] = {nm: int(cnt[1, k]) for k, nm in enumerate(self.class_names)}
        return res

    # ==================================================================================================
    # public API
    # ==================================================================================================
    def load_synthetic(self) -> None:
        xs, ys = synthetic_client_data(self.cfg.model, self.client_ids, self.cfg.shard, self.cfg.seed, self.spec)
        self.load_client_data(xs, ys)

    def load_datasets(self, datasets: List[Dict[str, Any]], work_dir: str = "/tmp/olsb_data") -> None:
        """Real data with the reference's shard contract (engine/data.py): ``datasets`` = the task's ``target.data``
        entries in population order, each ``{"data_path", "data_split_type", "data_transfer_type", "num_clients"}``.
        A dataset without a path falls back to the synthetic generator for its clients."""
        from .data import fetch_dataset, load_client_shards
        image = self.cfg.model in ("cnn2", "resnet18")
        shape = (32, 32, 3) if image else tuple(self.spec.input_shape)
        xs_parts, ys_parts, off = [], [], 0
        for d in datasets:
            n = int(d["num_clients"])
            ids = self.client_ids[(self.client_ids >= off) & (self.client_ids < off + n)]
            if len(ids):
                root = fetch_dataset(d, os.path.join(work_dir, str(d.get("name", off))))
                if root:
                    x, y = load_client_shards(root, ids - off, self.cfg.shard, bool(d.get("data_split_type", False)), n, shape)
                    if image and x.dtype != np.uint8:
                        raise ValueError("image families take uint8 NHWC samples")
                    if not image:
                        x = x.astype(np.int64 if self.spec.input_dtype == torch.int64 else np.float32)
                else:
                    x, y = synthetic_client_data(self.cfg.model, ids, self.cfg.shard, self.cfg.seed, self.spec)
                xs_parts.append(x); ys_parts.append(y)
            off += n
        if off != self.cfg.num_clients:
            raise ValueError(f"datasets describe {off} clients, the engine simulates {self.cfg.num_clients}")
        self.load_client_data(np.concatenate(xs_parts), np.concatenate(ys_parts))

    def load_client_data(self, xs: np.ndarray, labels: np.ndarray) -> None:
        """xs: uint8 NHWC [C, shard, 32,32,3] for the image families, float/int arrays otherwise."""
        self.host_x = torch.from_numpy(np.ascontiguousarray(xs))
        self.host_y = torch.from_numpy(np.ascontiguousarray(labels))
        if self.device.type == "cuda":
            self.host_x = self.host_x.pin_memory()
            self.host_y = self.host_y.pin_memory()
        if self.backend == "native":
            self.ws.load_data(self.host_x.to(self.device, non_blocking=True), self.host_y.to(self.device))
        else:
            x = self.host_x.to(self.device)
            if x.dtype == torch.uint8:
                x = reference.normalize_u8(x).permute(0, 1, 4, 2, 3).contiguous()   # C,S,3,32,32
            self.x, self.y = x, self.host_y.to(self.device)

    def _issue_h2d(self, slot: int) -> None:
        """Enqueue the pinned-host -> device copy of one round's inputs on the copy stream (double-buffered)."""
        st = self._copy_stream
        # the staging slot may still be read by the ingest kernel of two rounds ago
        st.wait_event(self._ingested[slot])
        with torch.cuda.stream(st):
            self._stage_x[slot].copy_(self.host_x, non_blocking=True)
            self._stage_y[slot].copy_(self.host_y, non_blocking=True)
            self._copied[slot].record(st)

    def stream_inputs(self, prefetch_next: bool = True) -> int:
        """End-to-end mode: this round's client samples + labels travel from pinned host memory to the GPU every round
        and are ingested there (native: uint8 -> bf16 operand layout).  The copy runs on its own stream into one of
        two staging buffers, so the transfer of round r+1 overlaps the compute of round r.  Returns bytes H2D."""
        nbytes = self.host_x.numel() * self.host_x.element_size() + self.host_y.numel() * self.host_y.element_size()
        if self.backend != "native":
            x = self.host_x.to(self.device, non_blocking=True)
            if x.dtype == torch.uint8:
                x = reference.normalize_u8(x).permute(0, 1, 4, 2, 3).contiguous()
            self.x, self.y = x, self.host_y.to(self.device, non_blocking=True)
            return nbytes
        if self.cfg.model != "cnn2":
            # MLP: bf16 cast; ResNet-18: uint8 -> stem im2col operand; BERT-mini: int32 token ids
            self.ws.load_data(self.host_x.to(self.device, non_blocking=True), self.host_y.to(self.device, non_blocking=True))
            return nbytes
        if not hasattr(self, "_stage_x"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage_x = [torch.empty_like(self.host_x, device=self.device) for _ in range(2)]
            self._stage_y = [torch.empty_like(self.host_y, device=self.device) for _ in range(2)]
            self._copied = [torch.cuda.Event() for _ in range(2)]
            self._ingested = [torch.cuda.Event() for _ in range(2)]
            for e in self._ingested:
                e.record(torch.cuda.current_stream())
            self._slot, self._pending = 0, False
        cur = torch.cuda.current_stream()
        if not self._pending:
            self._issue_h2d(self._slot)
        slot = self._slot
        cur.wait_event(self._copied[slot])
        self.ws.labels.copy_(self._stage_y[slot], non_blocking=True)
        self.ws.ingest_slice(self._stage_x[slot].reshape(-1, 32, 32, 3), 0)
        self._ingested[slot].record(cur)
        self._slot ^= 1
        self._pending = prefetch_next
        if prefetch_next:
            self._issue_h2d(self._slot)
        return nbytes

    def run_round(self, round_idx: int, stream_inputs: bool = False) -> RoundResult:
        t0 = time.perf_counter()
        if self.faults:
            self.faults.at_round_start(round_idx, self.rank)
        if stream_inputs:
            h2d = self.stream_inputs()
        if self.uses_ws:
            self._native_round_body(round_idx)
            res = self._native_finish(round_idx)
        else:
            res = self._torch_round(round_idx)
        res.wall_s = time.perf_counter() - t0
        # reference semantics (task_manager.py:743-748): a round with more failed clients than dynamic_nums is a failure
        res.extras["within_tolerance"] = res.failed <= self.cfg.dynamic_nums
        if self._flow_cache.get("plan"):
            res.extras["flow_plan"] = {k: self._flow_cache["plan"].get(k, []) for k in ("timing", "amount")}
        if self.cfg.profile:
            res.extras["phases_ms"] = self.timer.collect(self.group, self.world)
            res.extras["phases_ms_min"] = dict(self.timer.last_min)
        if stream_inputs:
            res.extras["h2d_bytes"] = h2d
            res.extras["d2h_bytes"] = 4 * native_round_rec() if self.uses_ws else 64
        self.round_log.append(res)
        return res

    def enqueue_round(self, round_idx: int) -> None:
        """Device-only enqueue of a round (no host read-back) — used by the device-timed benchmark."""
        assert self.uses_ws
        self._native_round_body(round_idx)

    def global_model(self) -> torch.Tensor:
        return self.ws.get_global() if self.uses_ws else self.global_w

    def spec_flat_global(self) -> torch.Tensor:
        """The global model in the zoo's flat (torch state_dict order) format, whatever the backend's packing."""
        g = self.global_model()
        if self.backend != "native":
            return g                                      # torch / hybrid rows are already in the zoo's flat order
        if self.cfg.model == "cnn2":
            return self.spec.flatten(CNN2.unpack(g))
        return self.spec.flatten(self.spec.unflatten(self.ws.L.unpack(g.cpu())))

    def set_global_model(self, flat: torch.Tensor) -> None:
        if self.uses_ws:
            self.ws.set_global(flat)
        else:
            self.global_w = flat.to(self.device).clone()

    def evaluate(self) -> Dict[str, float]:
        """Mean loss / accuracy of the current global model over every local client's shard."""
        g = self.spec_flat_global()
        B = self.cfg.batch
        x = self.host_x.to(self.device)[:, :B]
        y = self.host_y.to(self.device)[:, :B]
        if x.dtype == torch.uint8:
            x = reference.normalize_u8(x).permute(0, 1, 4, 2, 3)
        from torch.func import vmap
        w = g.to(self.device)[None, :].expand(self.C, -1)
        logits = vmap(lambda f, xx: self.spec.apply(self.spec.unflatten(f), xx))(w, x.contiguous())
        loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), y.reshape(-1).long())
        acc = (logits.argmax(-1) == y).float().mean()
        out = torch.tensor([loss.item(), acc.item(), 1.0], dtype=torch.float64)
        if self.world > 1:
            t = out.to(self.device) if self.device.type == "cuda" else out
            dist.all_reduce(t, group=self.group)
            out = t.cpu()
        return {"loss": float(out[0] / out[2]), "acc": float(out[1] / out[2])}

    def kernel_launches(self) -> int:
        return self.ws.kernel_launches if self.uses_ws else 0

    def state_dict(self) -> Dict[str, Any]:
        sd = {"config": asdict(self.cfg), "global": self.global_model().detach().cpu(), "rounds_done": len(self.round_log)}
        if self.uses_ws:
            # the server-optimiser state is sharded by rank (ZeRO-1-like); checkpoints store the FULL vectors so a job
            # restarted on a different number of ranks (elastic shrink) can re-slice them
            sd["opt_m"], sd["opt_v"] = (self._gather_opt(t) for t in (self.ws.opt_m, self.ws.opt_v))
            sd["server_step"] = self.ws.server_step
        else:
            sd["server_state"] = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in self.server_state.items()}
        return sd

    def load_state_dict(self, sd: Dict[str, Any]) -> None:
        self.set_global_model(sd["global"])
        if self.uses_ws:
            for name, dst in (("opt_m", self.ws.opt_m), ("opt_v", self.ws.opt_v)):
                dst.copy_(take_opt_shard(sd[name], self.rank, self.ws.shard_len, dst.numel()))
            self.ws.server_step = sd["server_step"]
        else:
            self.server_state = {k: (v.to(self.device) if torch.is_tensor(v) else v) for k, v in sd.get("server_state", {}).items()}

    def _gather_opt(self, shard: torch.Tensor) -> torch.Tensor:
        """Full optimiser-state vector from the per-rank shards (rank r owns [r*shard_len, (r+1)*shard_len))."""
        if self.world == 1:
            return shard.detach().cpu().clone()
        parts = [torch.empty_like(shard) for _ in range(self.world)]
        dist.all_gather(parts, shard.contiguous(), group=self.group)
        return merge_opt_shards([p.cpu() for p in parts], self.P)

    def close(self) -> None:
        if self.uses_ws:
            self._graphs.clear()          # graphs reference workspace memory: drop them before the slabs are freed
            self.ws.close()
