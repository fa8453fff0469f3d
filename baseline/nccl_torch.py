"""`nccl_torch` — the library build of the FL round loop: cuDNN / cuBLAS through ``torch`` + NCCL collectives.

This is the number the hand-written path is compared with (the reference itself ships no training or
collective code).  It is meant to be the *strongest* formulation the libraries allow, at the product's precision:

* CNN-2 (configs 2 and 4): all clients of a chunk are folded into the channel dimension and run as ONE grouped
  convolution per layer (``groups = clients``, channels_last, bf16 autocast or TF32), the classifier as one ``bmm``;
  one backward pass yields every client's gradient (the per-client losses are independent); SGD / FedProx is a fused
  ``foreach`` update; the whole local step (forward + backward + update) is captured in a CUDA graph and replayed.
  The chunk size (clients per grouped conv) is auto-tuned once.
* other families: ``torch.func.vmap(grad)`` under bf16 autocast / TF32 (the generic torch backend of ``FedEngine``).
* communication: ``dist.broadcast`` of the global model, ``dist.all_reduce`` of the weighted delta sum (NCCL), server
  step in torch on every rank.

Same round semantics, data, masks and timing as the native arm (``bench.py``).
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from olearning_sim_b200.deviceflow import philox  # noqa: E402
from olearning_sim_b200.models.zoo import get_model  # noqa: E402
from olearning_sim_b200.ops import reference  # noqa: E402
from olearning_sim_b200.parallel.comm import shard_range  # noqa: E402


def set_precision(dtype: str) -> None:
    """bf16: autocast; tf32: fp32 tensors with TF32 tensor-core math; fp32: plain."""
    tf32 = dtype in ("bf16", "tf32")
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True


class Cnn2Library:
    """Grouped-conv CNN-2 population on one rank."""

    def __init__(self, num_clients: int, local_steps: int, shard: int, lr: float, rank: int, world: int, device,
                 dtype: str = "bf16", mu: float = 0.0, seed: int = 0, chunk: Optional[int] = None, use_graph: bool = True,
                 hetero_steps: bool = False, p_online: float = 1.0, drop_probability: float = 0.0):
        self.N, self.S, self.shard, self.lr, self.mu = num_clients, local_steps, shard, lr, mu
        self.rank, self.world, self.dev, self.dtype, self.seed = rank, world, torch.device(device), dtype, seed
        self.c0, self.c1 = shard_range(num_clients, world, rank)
        self.C = self.c1 - self.c0
        self.spec = get_model("cnn2")
        self.p_online, self.drop_p, self.hetero = p_online, drop_probability, hetero_steps
        set_precision(dtype)
        g0 = self.spec.init(seed).to(self.dev)
        self.global_flat = g0.clone()
        gd = self.spec.unflatten(g0)
        self.params: Dict[str, torch.Tensor] = {n: gd[n][None].repeat(self.C, *([1] * gd[n].dim())).contiguous().requires_grad_(True)
                                                for n, _ in self.spec.shapes}
        self.names = [n for n, _ in self.spec.shapes]
        self.active = torch.ones(local_steps, self.C, device=self.dev)
        self.weight = torch.zeros(self.C, device=self.dev)
        self.loss_sum = torch.zeros(self.C, device=self.dev)
        rng = np.random.default_rng([seed, 991])
        all_steps = rng.integers(1, local_steps + 1, size=num_clients).astype(np.int32) if hetero_steps else \
            np.full(num_clients, local_steps, np.int32)
        self.local_steps_np = all_steps[self.c0:self.c1]
        self.chunk = chunk
        self.use_graph = use_graph
        self._graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self.xg: Dict[int, torch.Tensor] = {}
        self.y: Optional[torch.Tensor] = None

    # ---- data ---------------------------------------------------------------------------------------
    def load(self, x_u8: torch.Tensor, labels: torch.Tensor) -> None:
        """x_u8 [C, shard, 32, 32, 3] uint8 (device), labels [C, shard]: fold clients into channels per batch offset."""
        x = reference.normalize_u8(x_u8).permute(0, 1, 4, 2, 3)                       # C, shard, 3, 32, 32
        y = labels.to(self.dev).long()
        if self.y is None or self.y.shape != y.shape:
            self.y = y
        else:
            self.y.copy_(y)                                                              # graph-captured address stays valid
        B = 32
        act_dtype = torch.bfloat16 if self.dtype == "bf16" else torch.float32
        for off in sorted({(s * B) % self.shard for s in range(self.S)}):
            xb = x[:, off:off + B].permute(1, 0, 2, 3, 4).reshape(B, self.C * 3, 32, 32)
            buf = self.xg.get(off)
            new = xb.to(act_dtype).contiguous(memory_format=torch.channels_last)
            if buf is None:
                self.xg[off] = new
            else:
                buf.copy_(new)                                                           # graph-captured address stays valid

    # ---- local step ---------------------------------------------------------------------------------
    def _forward_chunk(self, off: int, k0: int, k1: int) -> torch.Tensor:
        p, k, B = self.params, k1 - k0, 32
        x = self.xg[off][:, k0 * 3:k1 * 3]
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.dtype == "bf16"):
            h = F.conv2d(x, p["conv1.weight"][k0:k1].reshape(k * 32, 3, 5, 5), p["conv1.bias"][k0:k1].reshape(-1), padding=2, groups=k)
            h = F.max_pool2d(F.relu(h), 2)
            h = F.conv2d(h, p["conv2.weight"][k0:k1].reshape(k * 64, 32, 5, 5), p["conv2.bias"][k0:k1].reshape(-1), padding=2, groups=k)
            h = F.max_pool2d(F.relu(h), 2)                                              # B, k*64, 8, 8
            feat = h.reshape(B, k, 4096).transpose(0, 1)                                # k, B, 4096 (NCHW flatten per client)
            logits = torch.baddbmm(p["fc.bias"][k0:k1, None, :].to(feat.dtype), feat, p["fc.weight"][k0:k1].transpose(1, 2).to(feat.dtype))
        y = self.y[k0:k1, off:off + B]
        loss = F.cross_entropy(logits.float().reshape(k * B, -1), y.reshape(-1), reduction="none").reshape(k, B).mean(1)
        return loss

    def _step_eager(self, s: int) -> None:
        off = (s * 32) % self.shard
        ck = self.chunk or self.C
        plist = [self.params[n] for n in self.names]
        for p in plist:
            p.grad = None
        losses = []
        for k0 in range(0, self.C, ck):
            loss = self._forward_chunk(off, k0, min(k0 + ck, self.C))
            loss.sum().backward()
            losses.append(loss.detach())
        act = self.active[s]
        self.loss_sum += torch.cat(losses) * act
        with torch.no_grad():
            for n, p in self.params.items():
                g = p.grad
                if self.mu > 0:
                    g = g + self.mu * (p - self._gview[n])
                p.addcmul_(g, act.view(-1, *([1] * (p.dim() - 1))), value=-self.lr)

    def _autotune(self) -> None:
        if self.chunk is not None:
            return
        best, best_t = None, 1e30
        for ck in (32, 128, 512, self.C):
            if ck > self.C and ck != self.C:
                continue
            self.chunk = min(ck, self.C)
            try:
                self._step_eager(0); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); self._step_eager(0); self._step_eager(0); e1.record(); torch.cuda.synchronize()
                t = e0.elapsed_time(e1)
            except RuntimeError:
                torch.cuda.empty_cache()
                continue
            if t < best_t:
                best, best_t = self.chunk, t
        self.chunk = best or min(128, self.C)
        if self.world > 1:                              # same choice on every rank (max over ranks)
            t = torch.tensor([self.chunk], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            self.chunk = int(t.item())

    def local_steps(self) -> None:
        for s in range(self.S):
            if not self.use_graph:
                self._step_eager(s)
                continue
            g = self._graphs.get(s)
            if g is None:
                # warm-up on a side stream, then capture (whole-network capture: forward, backward, update)
                st = torch.cuda.Stream(device=self.dev)
                st.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(st):
                    self._step_eager(s)
                torch.cuda.current_stream().wait_stream(st)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                try:
                    with torch.cuda.graph(g):
                        self._step_eager(s)
                    self._graphs[s] = g
                except Exception:  # noqa: BLE001
                    self.use_graph = False
                    torch.cuda.synchronize()
                continue
            g.replay()

    # ---- round --------------------------------------------------------------------------------------
    def round(self, r: int) -> float:
        gd = self.spec.unflatten(self.global_flat)
        self._gview = {n: gd[n][None] for n in self.names}
        if self.world > 1:
            dist.broadcast(self.global_flat, src=0)
        sampled = self.p_online < 1.0
        m = philox.round_mask(self.C, self.c0, self.seed, r, np.full(self.C, self.p_online, np.float32) if sampled else None,
                              self.drop_p, None, 0, np.full(self.C, self.shard, np.int32),
                              self.local_steps_np if self.hetero else None, self.S)
        part = torch.from_numpy(m["participate"].astype(bool)).to(self.dev)
        self.active.copy_(torch.from_numpy(m["step_active"].astype(np.float32)).to(self.dev))
        self.weight.copy_(torch.from_numpy(m["agg_weight"]).to(self.dev))
        with torch.no_grad():
            for n, p in self.params.items():
                if sampled:
                    p.copy_(torch.where(part.view(-1, *([1] * (p.dim() - 1))), self._gview[n].expand_as(p), p))
                else:
                    p.copy_(self._gview[n].expand_as(p))
        self.loss_sum.zero_()
        if r == 0:
            self._autotune()
            with torch.no_grad():
                for n, p in self.params.items():
                    p.copy_(self._gview[n].expand_as(p))
            self.loss_sum.zero_()
        self.local_steps()
        with torch.no_grad():
            w = self.weight
            parts = [((p - self._gview[n]) * w.view(-1, *([1] * (p.dim() - 1)))).sum(0).reshape(-1) for n, p in self.params.items()]
            pack = torch.cat(parts + [w.sum()[None]])
            if self.world > 1:
                dist.all_reduce(pack)
            wsum = pack[-1]
            delta = pack[:-1] / torch.clamp(wsum, min=1e-12)
            self.global_flat[: delta.numel()] += torch.where(wsum > 0, delta, torch.zeros_like(delta))
        return float(0.0)

    def mean_loss(self) -> float:
        return float((self.loss_sum / self.S).mean().item())


class GenericLibrary:
    """Every other family: the generic torch.func backend of FedEngine (vmap of grad, NCCL all-reduce) at bf16 / TF32."""

    def __init__(self, model: str, num_clients: int, local_steps: int, shard: int, lr: float, rank: int, world: int, device,
                 dtype: str = "bf16", **kw):
        from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine
        set_precision(dtype)
        self.eng = FedEngine(FedConfig(model=model, num_clients=num_clients, local_steps=local_steps, shard=shard, lr=lr,
                                       backend="torch", device=str(device), amp_bf16=dtype == "bf16", **kw), rank, world)
        self.host_loaded = False

    def load_synthetic(self) -> None:
        self.eng.load_synthetic()

    def round(self, r: int) -> float:
        return self.eng.run_round(r).loss


def build(model: str, **kw):
    if model == "cnn2":
        allowed = {"num_clients", "local_steps", "shard", "lr", "rank", "world", "device", "dtype", "mu", "seed", "chunk",
                   "use_graph", "hetero_steps", "p_online", "drop_probability"}
        return Cnn2Library(**{k: v for k, v in kw.items() if k in allowed})
    kw.pop("chunk", None); kw.pop("use_graph", None)
    return GenericLibrary(model, **kw)
