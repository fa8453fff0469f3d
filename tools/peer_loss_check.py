#!/usr/bin/env python
"""Rank loss mid-round (2+ ranks): the LAST rank stops taking part after round 0 (it stays alive but never launches round 1 —
what a hung or dead peer looks like to the others).  Every other rank must get ``PeerLostError`` out of ``run_round(1)``
within the spin budget, keep a working CUDA context, and still hold the round-0 global model bit for bit.

    torchrun --nproc-per-node 2 tools/peer_loss_check.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine  # noqa: E402
from olearning_sim_b200.ops import native  # noqa: E402
from olearning_sim_b200.parallel.comm import PeerLostError, init_distributed  # noqa: E402

rank, world, local = init_distributed()
native.require()
native.set_option("peer_spin_limit", 1 << 13)          # x256 polls: a second or two instead of minutes
eng = FedEngine(FedConfig(model="cnn2", num_clients=16 * world, local_steps=2, shard=64, backend="native", seed=3,
                          device=f"cuda:{local}"), rank, world)
eng.load_synthetic()
eng.run_round(0)
g0 = eng.ws.get_global().clone()
torch.cuda.synchronize()
dist.barrier()
res = {"rank": rank, "world": world}
if rank == world - 1:
    time.sleep(float(os.environ.get("LOST_SLEEP", "12")))       # the "lost" rank: alive, silent
    res["role"] = "lost"
else:
    t0 = time.time()
    try:
        eng.run_round(1)
        res.update(role="survivor", raised=False)
    except PeerLostError as e:
        res.update(role="survivor", raised=True, peer=e.peer, seconds=round(time.time() - t0, 2))
    x = torch.ones(1 << 20, device=f"cuda:{local}")
    res["context_alive"] = bool(float((x * 2).sum().item()) == float(2 << 20))
    res["model_untouched"] = bool(torch.equal(eng.ws.get_global(), g0))
    res["ok"] = bool(res.get("raised") and res["context_alive"] and res["model_untouched"] and res.get("peer") == world - 1)
print(json.dumps(res), flush=True)
os._exit(0)          # no collective teardown with a rank that left the protocol
