#!/usr/bin/env python
"""Multi-GPU correctness against an NCCL oracle (SURVEY §4 "distributed, GPU").

Every rank runs the SAME federated task twice on its GPU:
  * native backend  — hand-written local step, broadcast / reduce through the peer-memory kernels (no NCCL on the path)
  * torch backend   — fp32 torch.func local step, ``dist.all_reduce`` (NCCL) for the weighted delta sum, server step in torch
and rank 0 compares the global models after R rounds, the per-round masks / counters exactly and the loss.  A second
check isolates the communication kernels: identical replicas are given rank-dependent synthetic "updates" and the
kernel reduce + server step is compared with ``dist.all_reduce`` in fp32.

    torchrun --nproc-per-node N tools/mgpu_oracle.py            MODEL=cnn2|mlp2|resnet18|bert_mini  OPT=avg|momentum|adam
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine  # noqa: E402
from olearning_sim_b200.ops.layout import CNN2  # noqa: E402
from olearning_sim_b200.parallel.comm import init_distributed  # noqa: E402

rank, world, local = init_distributed()
MODEL, OPT = os.environ.get("MODEL", "cnn2"), os.environ.get("OPT", "avg")
CLIENTS = int(os.environ.get("CLIENTS", {"cnn2": 48, "mlp2": 48, "resnet18": 8, "bert_mini": 8}[MODEL]))
ROUNDS = int(os.environ.get("ROUNDS", 2))
kw = dict(model=MODEL, num_clients=CLIENTS, local_steps=2, shard=64, batch=32, lr=0.02, seed=7, p_online=0.9, drop_probability=0.1,
          hetero_steps=True, server_opt=OPT, server_lr=0.01 if OPT == "adam" else 1.0, device=f"cuda:{local}")
nat = FedEngine(FedConfig(backend="native", **kw), rank, world)
nat.load_synthetic()
ora = FedEngine(FedConfig(backend="torch", **kw), rank, world)
ora.load_synthetic()
g0 = nat.spec.init(11)
if MODEL == "cnn2":
    nat.set_global_model(CNN2.pack(nat.spec.unflatten(g0)))
elif MODEL in ("mlp2", "mlp"):
    nat.set_global_model(nat.ws.L.pack(g0))
else:
    nat.set_global_model(nat.ws.L.pack(g0))
ora.set_global_model(g0)
rows, ok = [], True
for r in range(ROUNDS):
    a, b = nat.run_round(r), ora.run_round(r)
    same = (a.participants, a.dropped, a.late, a.arrived, a.failed, a.client_steps) == (b.participants, b.dropped, b.late, b.arrived, b.failed, b.client_steps)
    rows.append({"round": r, "counters_equal": same, "loss_native": round(a.loss, 5), "loss_oracle": round(b.loss, 5)})
    ok = ok and same and abs(a.loss - b.loss) < 3e-2 * max(1.0, abs(b.loss))
ga = nat.spec_flat_global().cpu()[: nat.spec.num_params]
gb = ora.global_model().cpu()[: nat.spec.num_params]
d0 = g0[: nat.spec.num_params]
rel = float((ga - gb).norm() / (gb - d0).norm().clamp_min(1e-20))
# measured at 2 and 8 ranks (profiles/r2_mgpu_oracle_n*.jsonl): cnn2 3.4e-3, mlp2 1.1e-2, resnet18 3.3e-2, bert_mini 4.4e-3
tol = {"cnn2": 1.5e-2, "mlp2": 2.5e-2, "resnet18": 6e-2, "bert_mini": 1.5e-2}[MODEL] * (2 if OPT == "adam" else 1)
ok = ok and rel < tol

# ---- communication kernels alone vs NCCL all_reduce (fp32)
ws = nat.ws
torch.manual_seed(100 + rank)
ws.broadcast(None)
P = ws.P
upd = torch.randn(ws.C, P, device=nat.device) * 1e-2
base = ws.get_global().clone()
ws.master.copy_(base[None, :] + upd)
w = torch.rand(ws.C, device=nat.device) + 0.5
if ws.C > 1:
    w[0] = 0.0                                                  # one masked client per rank
ws.reset_round_stats()
ws.reduce_and_update(w, opt="avg", server_lr=1.0)
got = ws.get_global()
pack = torch.cat([(upd * w[:, None]).sum(0), w.sum()[None]]).double()
dist.all_reduce(pack) if world > 1 else None
want = base.double() + pack[:-1] / pack[-1]
rel_comm = float((got.double() - want).norm() / (want - base.double()).norm())
ok = ok and rel_comm < 1e-5
# ---- time of the fused reduce + server step (device events, max over ranks)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    ws.reduce_and_update(w, opt="avg", server_lr=1.0)
if world > 1:
    dist.barrier()
e0.record()
for _ in range(10):
    ws.reduce_and_update(w, opt="avg", server_lr=1.0)
e1.record()
torch.cuda.synchronize()
t_red = torch.tensor([e0.elapsed_time(e1) * 100.0], device=nat.device)      # us per call
if world > 1:
    dist.all_reduce(t_red, op=dist.ReduceOp.MAX)
gbuf = ws.heap.buffers["global"]
if world > 1:
    dist.barrier()
if rank == 0:
    print(json.dumps({"world": world, "model": MODEL, "server_opt": OPT, "clients": CLIENTS, "rounds": rows,
                      "model_delta_rel_err_vs_nccl_oracle": rel, "tol": tol, "reduce_kernel_rel_err_vs_nccl_allreduce": rel_comm, "nvls_multicast": bool(gbuf.mc_ptr),
                      "nvls_error": ws.heap.nvls_error, "reduce_update_us": round(float(t_red), 1), "params": int(P), "ok": bool(ok)}),
          flush=True)
nat.close()
if world > 1:
    dist.destroy_process_group()
