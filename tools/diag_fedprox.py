"""GPU diagnostic: ResNet-18 FedProx with sampling, native vs torch backend over 3 rounds."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine
for lr in (0.01,):
    kw = dict(model="resnet18", num_clients=6, local_steps=2, shard=64, algorithm="fedprox", mu=0.01, p_online=0.7,
              drop_probability=0.1, lr=lr, seed=0)
    a = FedEngine(FedConfig(backend="native", wave=4, **kw)); a.load_synthetic()
    b = FedEngine(FedConfig(backend="torch", device="cuda", **kw)); b.load_synthetic()
    g0 = a.spec.init(0)
    a.set_global_model(a.ws.L.pack(g0)); b.set_global_model(g0)
    for r in range(3):
        ra, rb = a.run_round(r), b.run_round(r)
        da = a.spec_flat_global().cpu() - g0
        db = b.global_model().cpu() - g0
        print(lr, r, "native", ra.loss, ra.participants, ra.arrived, "torch", rb.loss, rb.participants, rb.arrived,
              "rel", float((da - db).norm() / db.norm()))
    a.close()
