#!/usr/bin/env python
"""Flagship benchmark: FedAvg, 2-layer CNN (CIFAR-10-shaped synthetic shards), 1024 clients, 5 local steps.

    python bench.py --gpus N --steps K --warmup W            # ours (hand-written sm_100a path) + a short library arm
    python bench.py --impl reference ...                      # the unmodified reference (not installable: see DESIGN.md)
    python bench.py --impl nccl_torch ...                     # the library baseline alone (grouped cuDNN convs / vmap, bf16,
                                                              # CUDA graphs, NCCL broadcast + all-reduce)
    python bench.py --config 3|4|5                            # the other BASELINE.json configs (ResNet-18, FedProx, BERT-mini)

A "step" is one federated round: global-model broadcast into the packed replicas, ``local_steps`` SGD steps on
every resident client (batch 32), masked sample-weighted FedAvg reduce fused with the server update.  The headline
``value`` is simulated-clients x local-steps per second for the whole job (all N GPUs, strong scaling: the clients are
sharded over the ranks); ``rounds_per_sec`` is reported next to it.  Timing: W warm-up rounds, then exactly K rounds
bracketed by barrier + cuda synchronize, CUDA events on the launching stream, max over ranks.

``vs_baseline``: the reference publishes no number and cannot be installed, so the denominator is the
library build of the same round loop (``baseline/nccl_torch.py``) measured in this very process, on the same GPUs, at
the same precision (``baseline`` block of the JSON line; ``--no-baseline`` skips it).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

METRIC = "fl_client_local_steps_per_sec"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "nccl_torch"])
    ap.add_argument("--clients", type=int, default=1024)
    ap.add_argument("--local-steps", type=int, default=5)
    ap.add_argument("--shard", type=int, default=64)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-baseline", action="store_true", help="ours arm: skip the in-process library arm (vs_baseline stays null)")
    ap.add_argument("--baseline-steps", type=int, default=4, help="timed rounds of the in-process library arm")
    ap.add_argument("--baseline-dtype", default="bf16", choices=["bf16", "tf32", "fp32"],
                    help="precision of the library arm: bf16 autocast (same as ours), TF32 or plain fp32")
    ap.add_argument("--phases", action="store_true",
                    help="after the timed regions, run 3 extra rounds with CUDA-event phase timers and report "
                         "broadcast / local / reduce times and the peer-memory bus bandwidth they imply")
    ap.add_argument("--backend", default=None, choices=[None, "native", "hybrid"],
                    help="ours arm: force the backend (default native; hybrid = library forward/backward + native FL kernels)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 FedAvg CNN-2 x1024 (flagship, default), 3 FedAvg ResNet-18 x512 with "
                         "online/offline + spike trace, 4 FedProx CNN-2 x4096 heterogeneous steps, 5 FedAvg BERT-mini x256")
    return ap.parse_args()


# BASELINE.json configs other than the flagship; --clients / --local-steps override the preset when given explicitly
PRESETS = {
    2: dict(model="cnn2", clients=1024, local_steps=5, shard=64, algorithm="fedavg"),
    3: dict(model="resnet18", clients=512, local_steps=5, shard=64, algorithm="fedavg", p_online=0.8,
            online_trace="0.6+0.4*math.sin(t/4)", drop_probability=0.05),
    4: dict(model="cnn2", clients=4096, local_steps=5, shard=64, algorithm="fedprox", mu=0.01, hetero_steps=True),
    5: dict(model="bert_mini", clients=256, local_steps=2, shard=64, algorithm="fedavg"),
}
MODEL_DESC = {"cnn2": "cnn2 (conv5x5x32-pool-conv5x5x64-pool-fc10, 94666 params)", "resnet18": "resnet18 (GroupNorm, CIFAR stem)",
              "bert_mini": "bert_mini (4 layers, hidden 256, seq 128)"}


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs (rank 0 only)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_arm(args) -> int:
    """The reference cannot be installed: /root/reference has neither setup.py nor pyproject.toml (pip: "Neither
    'setup.py' nor 'pyproject.toml' found"), its modules import a non-existent `ols` package, four imported modules
    are missing from the tree, it needs Ray/Pulsar/MySQL/S3/K8s services, and it contains no model or training code
    (DESIGN.md)."""
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"impl": "reference", "unavailable":
                          "pip install of /root/reference fails (no setup.py/pyproject.toml); tree imports missing "
                          "'ols' package + 4 absent modules; needs Ray/Pulsar/MySQL; contains no model/training code"}))
    return 0


class Ctx:
    """rank / device / sync + max-over-ranks helpers shared by both arms"""

    def __init__(self):
        from olearning_sim_b200.parallel.comm import init_distributed
        self.rank, self.world, self.local = init_distributed()
        self.dev = torch.device("cuda", self.local)

    def sync(self):
        import torch.distributed as dist
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            dist.barrier()
            torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v: float) -> float:
        import torch.distributed as dist
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def run_library(ctx: Ctx, args, model: str, pre: dict, K: int, W: int, e2e: bool, dtype: str, sampler=None) -> dict:
    """The library arm (baseline/nccl_torch.py): same round semantics, data generator, timing and e2e copies."""
    from baseline import nccl_torch
    from olearning_sim_b200.engine.fed_engine import synthetic_client_data
    from olearning_sim_b200.models.zoo import get_model
    from olearning_sim_b200.parallel.comm import shard_range
    import numpy as np
    kw = dict(num_clients=args.clients, local_steps=args.local_steps, shard=args.shard, lr=0.05, rank=ctx.rank, world=ctx.world,
              device=ctx.dev, dtype=dtype, seed=0)
    for k in ("mu", "hetero_steps", "p_online", "drop_probability", "algorithm", "online_trace"):
        if k in pre:
            kw[k] = pre[k]
    lib = nccl_torch.build(model, **kw)
    host_x = host_y = None
    if model == "cnn2":
        c0, c1 = shard_range(args.clients, ctx.world, ctx.rank)
        xs, ys = synthetic_client_data("cnn2", np.arange(c0, c1), args.shard, 0, get_model("cnn2"))
        host_x, host_y = torch.from_numpy(xs).pin_memory(), torch.from_numpy(ys).pin_memory()
        lib.load(host_x.to(ctx.dev), host_y.to(ctx.dev))
    else:
        lib.load_synthetic()
    for r in range(W):
        lib.round(r)
    ctx.sync()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(K):
        lib.round(W + r)
    e1.record()
    ctx.sync()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / K
    clocks = sampler.stop() if sampler else None
    client_steps = args.clients * args.local_steps
    out = {"value": client_steps / (ms / 1e3), "ms_per_step": ms, "steps": K, "warmup": W, "dtype": dtype, "clocks": clocks,
           "impl": "nccl_torch: grouped cuDNN convs + bmm, CUDA graph, NCCL broadcast/all_reduce" if model == "cnn2"
           else "nccl_torch: torch.func vmap(grad) + NCCL all_reduce",
           "chunk": getattr(lib, "chunk", None), "cuda_graph": bool(getattr(lib, "_graphs", None))}
    if e2e:
        t0 = time.perf_counter()
        h2d, last = 0, None
        for r in range(K):
            if model == "cnn2":
                lib.load(host_x.to(ctx.dev, non_blocking=True), host_y.to(ctx.dev, non_blocking=True))
                h2d = host_x.numel() + host_y.numel() * host_y.element_size()
                lib.round(W + r)                             # same round indices as the device-timed region
                last = lib.mean_loss()                       # D2H read of the round's loss
            else:
                res = lib.eng.run_round(W + r, stream_inputs=True)
                h2d, last = res.extras["h2d_bytes"], res.loss
        ctx.sync()
        wall = ctx.max_over_ranks(time.perf_counter() - t0)
        out["e2e"] = {"value": client_steps * K / wall, "unit": "client_steps/s", "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": 4 if model == "cnn2" else 64, "ms_per_step": wall / K * 1e3, "last_loss": last}
    del lib
    torch.cuda.empty_cache()
    return out


def main() -> int:
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if not torch.cuda.is_available():
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"metric": METRIC, "value": None, "error": "no CUDA device"}))
        return 1
    import torch.distributed as dist
    ctx = Ctx()
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    pre = dict(PRESETS[args.config])
    if args.config != 2:
        # presets carry their own sizes unless the flags were given explicitly
        if "--clients" not in sys.argv: args.clients = pre["clients"]
        if "--local-steps" not in sys.argv: args.local_steps = pre["local_steps"]
    model = pre.pop("model"); pre.pop("clients"); pre.pop("local_steps"); pre.pop("shard")
    K, W = args.steps, max(args.warmup, 3)
    client_steps = args.clients * args.local_steps
    cfg_desc = {"id": args.config, "model": MODEL_DESC[model], "algorithm": pre.get("algorithm", "fedavg"),
                "clients": args.clients, "local_steps": args.local_steps, "local_batch": 32,
                "samples_per_client": args.shard, "global_batch": args.clients * 32,
                "seq_len": 128 if model == "bert_mini" else None, "parallelism": f"clients sharded dp{world}",
                "l2": "inputs larger than L2 (per-round working set of GBs streams through HBM)"}

    if args.impl == "nccl_torch":
        sampler = ClockSampler(ctx.local) if rank == 0 else None
        lib = run_library(ctx, args, model, pre, K, W, not args.no_e2e, args.baseline_dtype, sampler)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": lib["value"], "unit": "client_steps/s", "n_gpus": world, "steps": K,
                              "warmup": W, "ms_per_step": lib["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                              "vs_baseline": 1.0, "dtype": lib["dtype"], "data": "synthetic", "impl": "nccl_torch",
                              "rounds_per_sec": 1e3 / lib["ms_per_step"], "config": {**cfg_desc, "backend": lib["impl"],
                                                                                    "chunk": lib["chunk"], "cuda_graph": lib["cuda_graph"]},
                              "clocks": lib["clocks"], "e2e": lib.get("e2e"), "gpu_launches": 0}))
        ctx.sync()
        if world > 1:
            dist.destroy_process_group()
        return 0

    from olearning_sim_b200.engine.fed_engine import NATIVE_FAMILIES, FedConfig, FedEngine
    # ours: hand-written kernels end to end; --backend hybrid = library forward/backward + the native FL kernels
    backend = args.backend or ("native" if model in NATIVE_FAMILIES else "hybrid")
    cfg = FedConfig(model=model, num_clients=args.clients, local_steps=args.local_steps, batch=32, shard=args.shard,
                    lr=0.05, backend=backend, device=str(dev), seed=0, amp_bf16=backend == "hybrid", **pre)
    eng = FedEngine(cfg, rank, world)
    eng.load_synthetic()
    sync = ctx.sync
    # ------------------------------------------------------------------ device-timed (kernel path only)
    for r in range(W):
        eng.run_round(r)
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    sync()
    if sampler:
        sampler.start()
    launches0 = eng.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(K):
        eng.enqueue_round(W + r)
    e1.record()
    sync()
    launches = eng.kernel_launches() - launches0
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    ms_per_round = ms_total / K
    value = client_steps / (ms_per_round / 1e3)

    # ------------------------------------------------------------------ end to end through the public API
    e2e = None
    if not args.no_e2e:
        for r in range(2):
            eng.run_round(W + K + r, stream_inputs=True)
        sync()
        t0 = time.perf_counter()
        h2d = d2h = 0
        for r in range(K):
            # the same round indices as the device-timed region: participation traces (config 3) are functions of the round
            res = eng.run_round(W + r, stream_inputs=True)
            h2d, d2h = res.extras["h2d_bytes"], res.extras["d2h_bytes"]
        sync()
        wall = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": client_steps * K / wall, "unit": "client_steps/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": wall / K * 1e3, "timing": "host wall clock, max over ranks",
               "last_loss": res.loss}

    phases = None
    if args.phases:
        eng.cfg.profile = True
        eng.timer.enabled = True
        acc, accmin = {}, {}
        eng.run_round(W + 2 * K + 3)       # discarded: the first collect() creates the NCCL communicator used for max/min
        sync()
        for r in range(3):
            res = eng.run_round(W + 2 * K + 4 + r)
            for k, v in res.extras["phases_ms"].items():
                acc[k] = acc.get(k, 0.0) + v / 3
            for k, v in res.extras["phases_ms_min"].items():
                accmin[k] = accmin.get(k, 0.0) + v / 3
        P4 = eng.P * 4
        far = (world - 1) / world
        # per rank: broadcast pulls the (world-1)/world remote shards of the model; reduce pulls `groups` partial
        # shards from every peer.  Bus bandwidth = bytes crossing NVLink into this GPU / phase time.
        groups = getattr(eng.ws, "groups", 1)
        # phase times: max over ranks, and min over ranks (= the work itself for phases that begin with a cross-rank
        # barrier: every rank but the slowest also waits there).  Local HBM traffic of the two fused paths per rank:
        # broadcast writes C_local replicas, reduce reads them and writes / pulls `groups` partial rows.
        c_local = eng.C
        phases = {"ms_max": acc, "ms_min": accmin,
                  "bcast_remote_bytes": int(P4 * far), "reduce_remote_bytes": int(groups * P4 / world * (world - 1)),
                  "bcast_hbm_bytes": int(P4 * (c_local + 1)), "reduce_hbm_bytes": int(P4 * (c_local + 2 * groups)),
                  "bcast_bus_GBps": P4 * far / max(accmin.get("broadcast", 0), 1e-9) / 1e6,
                  "reduce_bus_GBps": groups * P4 / world * (world - 1) / max(accmin.get("reduce_update", 0), 1e-9) / 1e6,
                  "bcast_hbm_GBps": P4 * (c_local + 1) / max(accmin.get("broadcast", 0), 1e-9) / 1e6,
                  "reduce_hbm_GBps": P4 * (c_local + 2 * groups) / max(accmin.get("reduce_update", 0), 1e-9) / 1e6}
    graph_error = getattr(eng, "graph_error", None)
    sync()
    comm_desc = None
    hp = getattr(getattr(eng, "ws", None), "heap", None)
    if hp is not None and world > 1:
        gb = hp.buffers.get("global")
        comm_desc = {"transport": "peer-memory kernels over NVLink (no NCCL on the model path)",
                     "nvls_multicast": bool(gb is not None and gb.mc_ptr), "nvls_error": hp.nvls_error}
    eng.close()
    del eng
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ library arm, same process / GPUs / precision
    base = None
    if not args.no_baseline:
        try:
            # a participation trace is a function of the round index: then both arms time the same rounds
            kb = K if pre.get("online_trace") else max(args.baseline_steps, 3)
            base = run_library(ctx, args, model, pre, kb, W, not args.no_e2e, args.baseline_dtype)
        except Exception as e:  # noqa: BLE001 - the headline must survive a baseline failure
            base = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "client_steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_round, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": (value / base["value"]) if base and base.get("value") else None,
            "dtype": "bf16" if backend == "native" else "bf16 autocast (library forward/backward)", "data": "synthetic",
            "impl": args.impl, "rounds_per_sec": 1e3 / ms_per_round,
            "config": {**cfg_desc, "backend": backend},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        }
        if base:
            out["baseline"] = {"what": "library build of the same round loop measured in this process (baseline/nccl_torch.py); the "
                                       "reference publishes no number and cannot be installed", **base}
            if base.get("e2e") and e2e:
                out["baseline"]["e2e_ratio"] = e2e["value"] / base["e2e"]["value"]
        if comm_desc:
            out["comm"] = comm_desc
        if graph_error:
            out["cuda_graph_error"] = graph_error
        if phases:
            out["phases"] = phases
        print(json.dumps(out))
    sync()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
