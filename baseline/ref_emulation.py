#!/usr/bin/env python
"""ref_emulation — "what olearning-sim would do": the reference's execution model without Ray / Pulsar / MySQL.

Faithful to the mechanisms, not the infrastructure (none of it is installable offline):

* contiguous client -> worker sharding, floor(N/R) each, the first N mod R workers one more
  (ols_core/taskMgr/run_task.py:85-105); workers run in parallel, clients of a worker strictly sequentially
  (ols_core/taskMgr/utils/utils_run_task.py:481-494)
* ONE fresh python subprocess per simulated client per round: ``python3 <operator>/<entry> --params f.json``
  (utils_run_task.py:496-514) that loads the global model FILE of the round
  ("{task_id}_{current_round}_result_model", utils_run_task.py:335-350), trains locally and publishes its update
* updates travel as messages through a single-consumer sorter -> shelf -> dispatcher path
  (ols_core/deviceflow/non_grpc/sorter.py:37-91, dispatcher.py:84-171) to an aggregator that writes the next model file

    python baseline/ref_emulation.py --clients 16 --workers 4 --rounds 2 --local-steps 5
prints one JSON line with rounds/s and client-steps/s.  CPU only, like the reference's image (torch==1.13.1+cpu).
"""
from __future__ import annotations

import argparse
import concurrent.futures
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CLIENT_SCRIPT = r'''
import json, sys, os, base64, io
sys.path.insert(0, {root!r})
import torch, numpy as np
from olearning_sim_b200.models.zoo import get_model
from olearning_sim_b200.engine.fed_engine import synthetic_client_data
p = json.load(open(sys.argv[sys.argv.index("--params") + 1]))
torch.set_num_threads(1)
spec = get_model("cnn2")
w = torch.load(p["model"]["path"])                      # the per-round global model file
xs, ys = synthetic_client_data("cnn2", np.array([p["split_index"]]), p["operator_params"]["shard"], 0, spec)
x = ((torch.from_numpy(xs[0]).float() / 255 - 0.5) * 2).permute(0, 3, 1, 2)
y = torch.from_numpy(ys[0]).long()
w0 = w.clone()
w.requires_grad_(True)
B = 32
for s in range(p["operator_params"]["local_steps"]):
    off = (s * B) % x.shape[0]
    loss = spec.loss(w, x[off:off + B], y[off:off + B])
    (g,) = torch.autograd.grad(loss, w)
    with torch.no_grad():
        w -= p["operator_params"]["lr"] * g
buf = io.BytesIO(); torch.save((w.detach() - w0), buf)
msg = {{"routing_key": p["deviceflow"]["routing_key"], "compute_resource": "logical_simulation",
       "message": json.dumps({{"client": p["split_index"], "n": int(x.shape[0]), "delta": base64.b64encode(buf.getvalue()).decode()}})}}
with open(os.path.join(p["deviceflow"]["inbound_dir"], f"msg_{{p['split_index']}}.json"), "w") as f:
    json.dump(msg, f)
'''


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--clients", type=int, default=16)
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--local-steps", type=int, default=5)
    ap.add_argument("--shard", type=int, default=64)
    a = ap.parse_args()
    import base64, io
    import torch
    from olearning_sim_b200.deviceflow.bus import InMemoryBus
    from olearning_sim_b200.deviceflow.rooms import parse_envelope
    from olearning_sim_b200.models.zoo import get_model
    from olearning_sim_b200.parallel.comm import shard_range
    spec = get_model("cnn2")
    work = tempfile.mkdtemp(prefix="ref_emul_")
    op = os.path.join(work, "training")
    os.makedirs(op)
    with open(os.path.join(op, "train.py"), "w") as f:
        f.write(CLIENT_SCRIPT.format(root=ROOT))
    task_id = "emul"
    g = spec.init(0)
    bus = InMemoryBus(virtual_time=True)

    def actor(worker: int, rnd: int, inbound: str, model_path: str):
        s, e = shard_range(a.clients, a.workers, worker)
        ok = 0
        for idx in range(s, e):                               # sequential inside an actor
            prm = {"task_id": task_id, "operator_name": "training", "round": rnd, "device_type": "High", "split_index": idx,
                   "model": {"path": model_path}, "operator_params": {"local_steps": a.local_steps, "lr": 0.05, "shard": a.shard},
                   "deviceflow": {"routing_key": f"{task_id}_training_{rnd}", "inbound_dir": inbound}}
            pf = os.path.join(work, f"params_{worker}_{idx}.json")
            json.dump(prm, open(pf, "w"))
            ok += subprocess.call([sys.executable, os.path.join(op, "train.py"), "--params", pf]) == 0
        return ok

    t0 = time.perf_counter()
    for rnd in range(a.rounds):
        model_path = os.path.join(work, f"{task_id}_{rnd}_result_model.pt")
        torch.save(g, model_path)                             # "download" = every client process re-reads this file
        inbound = os.path.join(work, f"inbound_{rnd}")
        os.makedirs(inbound)
        with concurrent.futures.ThreadPoolExecutor(a.workers) as pool:
            done = sum(pool.map(lambda w: actor(w, rnd, inbound, model_path), range(a.workers)))
        # single-consumer sorter -> shelf -> dispatcher -> aggregator
        prod = bus.create_producer("inbound")
        shelf = bus.subscribe("inbound", "sorter")
        for fn in sorted(os.listdir(inbound)):
            prod.send(open(os.path.join(inbound, fn), "rb").read())
        tot, wsum = torch.zeros_like(g), 0.0
        for _ in range(done):
            m = shelf.receive(0)
            env = parse_envelope(m.value())
            body = json.loads(env["message"])
            delta = torch.load(io.BytesIO(base64.b64decode(body["delta"])), weights_only=True)
            tot += body["n"] * delta
            wsum += body["n"]
            shelf.acknowledge(m)
        g = g + tot / max(wsum, 1.0)
    wall = time.perf_counter() - t0
    print(json.dumps({"impl": "ref_emulation", "clients": a.clients, "workers": a.workers, "rounds": a.rounds,
                      "local_steps": a.local_steps, "rounds_per_sec": a.rounds / wall,
                      "client_steps_per_sec": a.clients * a.local_steps * a.rounds / wall, "wall_s": wall,
                      "note": "CPU, one python subprocess per client per round, file model exchange, queue update path"}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
