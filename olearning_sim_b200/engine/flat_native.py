"""Model-agnostic GPU-resident FL state: packed fp32 replicas + the native communication / optimiser kernels.

This is what every model family gets even when its per-client forward / backward still runs through the library
(``torch.func``): the global-model broadcast is the fused peer-memory all-gather + replica fan-out
(``bcast_scatter_generic_kernel``), the local update is the fused SGD / FedProx kernel (``client_sgd_kernel``), the
participation / drop / straggler masks come from the Philox DeviceFlow kernel, and the reduce is the masked weighted
partial sum + owner-side pull over NVLink + server SGD / momentum / Adam (``fedavg_partial`` / ``fedavg_finish``).  No NCCL
and no per-round host arithmetic on any of those paths.  The reference does the same three steps with model files on
S3, one Python process per client and a Pulsar topic.
"""
from __future__ import annotations

from typing import Optional

import torch

from ..ops import native
from ..parallel.comm import SymmetricHeap
from .cnn2_native import OPT_CODES


class FlatWorkspace:
    def __init__(self, num_clients: int, num_params: int, device="cuda", heap: Optional[SymmetricHeap] = None):
        native.require()
        assert num_params % 4 == 0, "packed parameter rows are padded to a multiple of 4 floats"
        self.C, self.P, self.dev = num_clients, num_params, torch.device(device)
        self.heap = heap or SymmetricHeap(device=self.dev if self.dev.index is not None else None)
        self.rank, self.world = self.heap.rank, self.heap.world
        P, C, d = num_params, num_clients, self.dev
        f32, i32 = torch.float32, torch.int32
        self.shard_len = ((P // 4 + self.world - 1) // self.world) * 4
        self.master = torch.zeros(C, P, dtype=f32, device=d)
        self.global_local = torch.zeros(P, dtype=f32, device=d)
        self.loss_sum = torch.zeros(C, dtype=f32, device=d)
        self.nonfinite = torch.zeros(C, dtype=i32, device=d)
        self.groups = max(1, min(16, max(1, C // 8)))
        hp = self.heap
        hp.alloc("global", P * 4, multicast=True); hp.alloc("partial", self.groups * P * 4, multicast=True); hp.alloc("wsum", 64 * 4)
        hp.alloc("flags", 256); hp.alloc("round", 256)
        self.global_w = hp.tensor("global", f32, (P,))
        self.partial = hp.tensor("partial", f32, (self.groups, P))
        self.wsum = hp.tensor("wsum", f32, (self.groups,))
        self.round_rec = hp.tensor("round", f32, (native.ROUND_REC,))
        self.round_rec.zero_()
        self.round_out = torch.zeros(native.ROUND_REC + 8, dtype=f32, device=d)      # [ROUND_REC] = peer-lost code
        self.opt_m = torch.zeros(self.shard_len, dtype=f32, device=d)
        self.opt_v = torch.zeros(self.shard_len, dtype=f32, device=d)
        self.stats = torch.zeros(4, dtype=f32, device=d)
        self.arrive_ctr = torch.zeros(4, dtype=i32, device=d)                        # [2] = sticky peer-lost code
        self.epoch = self.server_step = self.kernel_launches = 0

    def set_global(self, row: torch.Tensor) -> None:
        self.global_w.copy_(row.to(self.dev))
        self.global_local.copy_(self.global_w)
        self.heap.barrier()

    def get_global(self) -> torch.Tensor:
        out = torch.empty(self.P, dtype=torch.float32, device=self.dev)
        for r in range(self.world):
            s0, s1 = r * self.shard_len, min((r + 1) * self.shard_len, self.P)
            if s0 < s1:
                out[s0:s1].copy_(self.heap.tensor("global", torch.float32, (self.P,), peer=r)[s0:s1])
        return out

    def broadcast(self, participate: Optional[torch.Tensor] = None) -> None:
        native.bcast_scatter_generic(self.heap.buffers["global"].peer_ptrs, self.shard_len, self.P, self.global_local,
                                     self.master, None, participate)
        self.kernel_launches += 1

    def sgd_step(self, grads: torch.Tensor, lr: float, mu: float = 0.0, active: Optional[torch.Tensor] = None) -> None:
        """w_c -= lr * (g_c + mu * (w_c - w_global)) for every active client, in one launch."""
        native.client_sgd(self.master, grads, self.global_local if mu > 0 else None, active, lr, mu)
        self.kernel_launches += 1

    def reduce_and_update(self, weight: torch.Tensor, opt: str = "avg", server_lr: float = 1.0, beta1: float = 0.9,
                          beta2: float = 0.99, eps: float = 1e-3, accounting=None) -> None:
        if accounting is not None:
            part, step_active, counters, nsteps = accounting[:4]
            class_id, nclass = (accounting[4], accounting[5]) if len(accounting) > 4 else (None, 0)
            native.round_stats(self.loss_sum, self.nonfinite, part, step_active, counters, self.C, step_active.shape[0],
                               1.0 / nsteps, self.round_rec, class_id, nclass)
        native.fedavg_partial(self.master, self.global_local, weight, self.nonfinite, self.groups, self.partial, self.wsum)
        self.epoch += 1
        self.server_step += 1
        hb = self.heap.buffers
        native.fedavg_finish(hb["partial"].peer_ptrs, hb["wsum"].peer_ptrs, hb["flags"].peer_ptrs if self.world > 1 else None,
                             self.rank, self.groups, self.P, self.shard_len, self.epoch, self.global_w, self.opt_m,
                             self.opt_v, OPT_CODES[opt], server_lr, beta1, beta2, eps, self.server_step, self.stats,
                             self.arrive_ctr, hb["round"].peer_ptrs if accounting is not None else None, self.round_out,
                             8 + 2 * (accounting[5] if accounting is not None and len(accounting) > 4 else 0),
                             partial_mc=hb["partial"].mc_ptr, global_mc=hb["global"].mc_ptr)
        self.kernel_launches += 2 + (accounting is not None)

    def reset_round_stats(self) -> None:
        self.loss_sum.zero_()
        self.nonfinite.zero_()

    def close(self) -> None:
        self.heap.close()
