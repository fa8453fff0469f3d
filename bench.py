#!/usr/bin/env python
"""Flagship benchmark: FedAvg, 2-layer CNN (CIFAR-10-shaped synthetic shards), 1024 clients, 5 local steps.

    python bench.py --gpus N --steps K --warmup W            # ours (hand-written sm_100a path)
    python bench.py --impl reference ...                      # the unmodified reference (not installable: see DESIGN.md)
    python bench.py --impl nccl_torch ...                     # honest library baseline (torch.func + cuDNN + NCCL)

A "step" is one federated round: global-model broadcast into the packed replicas, ``local_steps`` SGD steps on
every resident client (batch 32), masked sample-weighted FedAvg reduce fused with the server update.  The headline
``value`` is simulated-clients x local-steps per second for the whole job (all N GPUs, strong scaling: the 1024
clients are sharded over the ranks); ``rounds_per_sec`` is reported next to it.  Timing: W warm-up rounds, then
exactly K rounds bracketed by barrier + cuda synchronize, CUDA events on the launching stream, max over ranks.
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
    ap.add_argument("--phases", action="store_true",
                    help="after the timed regions, run 3 extra rounds with CUDA-event phase timers and report "
                         "broadcast / local / reduce times and the peer-memory bus bandwidth they imply")
    ap.add_argument("--amp", action="store_true", help="hybrid backend: bf16 autocast for the library forward/backward (default for config 5; it is "
                         "slower than fp32 for the vmapped ResNet convolutions)")
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


def main() -> int:
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine
    from olearning_sim_b200.parallel.comm import init_distributed
    import torch.distributed as dist

    rank, world, local = init_distributed()
    if not torch.cuda.is_available():
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": None, "error": "no CUDA device"}))
        return 1
    dev = torch.device("cuda", local)
    pre = dict(PRESETS[args.config])
    if args.config != 2:
        # presets carry their own sizes unless the flags were given explicitly
        if "--clients" not in sys.argv: args.clients = pre["clients"]
        if "--local-steps" not in sys.argv: args.local_steps = pre["local_steps"]
    model = pre.pop("model"); pre.pop("clients"); pre.pop("local_steps"); pre.pop("shard")
    native_ok = model in ("cnn2", "mlp2", "mlp", "resnet18", "bert_mini")
    # ours: hand-written kernels end to end where they exist (cnn2 / mlp2), otherwise library forward/backward + the
    # native FL kernels (broadcast, masks, fused client SGD, peer-memory reduce + server optimiser): backend "hybrid"
    backend = ("native" if native_ok else "hybrid") if args.impl == "ours" else "torch"
    cfg = FedConfig(model=model, num_clients=args.clients, local_steps=args.local_steps, batch=32, shard=args.shard,
                    lr=0.05, backend=backend, device=str(dev), seed=0, amp_bf16=bool((args.amp or args.config == 5) and backend == "hybrid"),
                    **pre)
    eng = FedEngine(cfg, rank, world)
    eng.load_synthetic()

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    K, W = args.steps, max(args.warmup, 3)
    # ------------------------------------------------------------------ device-timed (kernel path only)
    for r in range(W):
        eng.run_round(r)
    sampler = ClockSampler(local) if rank == 0 else None
    sync()
    if sampler:
        sampler.start()
    launches0 = eng.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(K):
        if backend in ("native", "hybrid"):
            eng.enqueue_round(W + r)
        else:
            eng.run_round(W + r)
    e1.record()
    sync()
    launches = eng.kernel_launches() - launches0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clocks = sampler.stop() if sampler else None
    ms_per_round = ms_total / K
    client_steps = args.clients * args.local_steps
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
            res = eng.run_round(W + K + 2 + r, stream_inputs=True)
            h2d, d2h = res.extras["h2d_bytes"], res.extras["d2h_bytes"]
        sync()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall = float(tw.item())
        e2e = {"value": client_steps * K / wall, "unit": "client_steps/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": wall / K * 1e3, "timing": "host wall clock, max over ranks",
               "last_loss": res.loss}

    phases = None
    if args.phases and backend in ("native", "hybrid"):
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
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "client_steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_round, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if backend == "native" else ("bf16 autocast (library forward/backward)" if cfg.amp_bf16
                                                         else "fp32 (library forward/backward)"), "data": "synthetic",
            "impl": args.impl, "rounds_per_sec": 1e3 / ms_per_round,
            "config": {"id": args.config, "model": MODEL_DESC[model], "algorithm": cfg.algorithm, "backend": backend,
                       "clients": args.clients, "local_steps": args.local_steps, "local_batch": 32,
                       "samples_per_client": args.shard, "global_batch": args.clients * 32,
                       "seq_len": 128 if model == "bert_mini" else None,
                       "parallelism": f"clients sharded dp{world}",
                       "l2": "inputs larger than L2 (per-round working set of GBs streams through HBM)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        }
        if phases:
            out["phases"] = phases
        print(json.dumps(out))
    sync()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
