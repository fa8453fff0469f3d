"""GPU probe: measured deviation of the native (bf16 tensor-core) backend from the fp32 torch backend for the engine-level
tests (per-tensor model delta and loss, three seeds per family) — the tolerances in tests/test_gpu_kernels.py are set from
its output (profiles/r2_gpu_tests.log)."""
import sys, os
sys.path.insert(0, "/root/repo")
import torch
from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine
from olearning_sim_b200.ops.layout import CNN2
from olearning_sim_b200.models.zoo import get_model
out = {}
for seed in (3, 4, 5):
    kw = dict(model="cnn2", num_clients=12, local_steps=2, shard=64, lr=0.05, p_online=0.8, drop_probability=0.2, hetero_steps=True, seed=seed)
    a = FedEngine(FedConfig(backend="native", **kw)); a.load_synthetic()
    b = FedEngine(FedConfig(backend="torch", **kw)); b.load_synthetic()
    g0 = CNN2.init(11); spec = get_model("cnn2")
    a.set_global_model(g0); b.set_global_model(spec.flatten(CNN2.unpack(g0)))
    ra, rb = a.run_round(0), b.run_round(0)
    ga = CNN2.unpack(a.global_model()); gb = spec.unflatten(b.global_model().cpu())
    rels = {k: float(((ga[k] - CNN2.unpack(g0)[k]) - (gb[k].cpu() - CNN2.unpack(g0)[k])).norm() / ((gb[k].cpu() - CNN2.unpack(g0)[k]).norm() + 1e-12)) for k in ga}
    print("cnn2 seed", seed, "loss", abs(ra.loss - rb.loss) / max(1, abs(rb.loss)), "max rel", max(rels.values()), rels)
    a.close()
for model, mk, n, extra in (("mlp2", {}, 10, {}), ("mlp", {"in_dim": 784, "hidden": [256, 128, 64]}, 6, {"algorithm": "fedprox", "mu": 0.05})):
    for seed in (5, 6, 7):
        kw = dict(model=model, num_clients=n, local_steps=2, shard=64, lr=0.1, p_online=0.8, hetero_steps=True, seed=seed, **extra)
        if mk: kw["model_kwargs"] = mk
        if model == "mlp2": kw["drop_probability"] = 0.2
        a = FedEngine(FedConfig(backend="native", use_cuda_graph=False, **kw)); a.load_synthetic()
        b = FedEngine(FedConfig(backend="torch", device="cuda", **kw)); b.load_synthetic()
        ra, rb = a.run_round(0), b.run_round(0)
        ga, gb = a.spec_flat_global().cpu(), b.spec_flat_global().cpu()
        g0 = a.spec.init(seed)
        print(model, "seed", seed, "loss", abs(ra.loss - rb.loss) / max(1, abs(rb.loss)), "rel", float(((ga - g0) - (gb - g0)).norm() / (gb - g0).norm()))
        a.close()
