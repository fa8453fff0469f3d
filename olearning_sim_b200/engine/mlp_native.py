"""GPU-resident state of an MLP client population (fully-connected families) on one rank.

Same structure as ``Cnn2Workspace`` but built from the generic per-client GEMM kernels (``csrc/linear_tc.cuh``):
a local step of a 2-layer MLP is six launches — fc1 forward (+bias+ReLU), fc2 forward (+bias), softmax-xent
(+fc2-bias SGD), fc2 data gradient (+ReLU gate, +fc1-bias SGD), fc2 weight gradient (+SGD), fc1 weight gradient (+SGD).
The packed parameter row is [W1 | b1 | W2 | b2] with every block starting on an 8-element boundary so the bf16 copy
of the row can be used directly as TMA operand storage.
"""
from __future__ import annotations

from typing import Optional

import torch

from ..ops import native
from ..parallel.comm import SymmetricHeap
from .cnn2_native import OPT_CODES

B = 32


def _al8(n: int) -> int:
    return (n + 7) // 8 * 8


class MlpLayout:
    def __init__(self, in_dim: int = 784, hidden: int = 200, classes: int = 10):
        assert in_dim % 8 == 0 and hidden % 8 == 0, "row strides must be multiples of 16 bytes for TMA"
        self.in_dim, self.hidden, self.classes = in_dim, hidden, classes
        self.W1 = 0
        self.B1 = _al8(self.W1 + hidden * in_dim)
        self.W2 = _al8(self.B1 + hidden)
        self.B2 = _al8(self.W2 + classes * hidden)
        self.P = _al8(self.B2 + classes)
        self.LD = 16                                     # padded class dimension of logits / dlogits

    def pack(self, flat_spec: torch.Tensor) -> torch.Tensor:
        """zoo.mlp2 flat order (fc1.weight, fc1.bias, fc2.weight, fc2.bias) -> aligned packed row."""
        out = torch.zeros(self.P)
        h, i, c = self.hidden, self.in_dim, self.classes
        o = 0
        for dst, n in ((self.W1, h * i), (self.B1, h), (self.W2, c * h), (self.B2, c)):
            out[dst:dst + n] = flat_spec[o:o + n]
            o += n
        return out

    def unpack(self, row: torch.Tensor) -> torch.Tensor:
        h, i, c = self.hidden, self.in_dim, self.classes
        return torch.cat([row[self.W1:self.W1 + h * i], row[self.B1:self.B1 + h], row[self.W2:self.W2 + c * h],
                          row[self.B2:self.B2 + c]])


class MlpWorkspace:
    def __init__(self, num_clients: int, shard: int, layout: Optional[MlpLayout] = None, device="cuda",
                 heap: Optional[SymmetricHeap] = None):
        native.require()
        self.L = layout or MlpLayout()
        L = self.L
        self.C, self.shard, self.dev = num_clients, shard, torch.device(device)
        self.heap = heap or SymmetricHeap(device=self.dev if self.dev.index is not None else None)
        self.rank, self.world = self.heap.rank, self.heap.world
        P = self.P = L.P
        self.shard_len = ((P // 4 + self.world - 1) // self.world) * 4
        C, d = num_clients, self.dev
        bf, f32, i32 = torch.bfloat16, torch.float32, torch.int32
        self.x = torch.zeros(C, shard, L.in_dim, dtype=bf, device=d)
        self.labels = torch.zeros(C, shard, dtype=i32, device=d)
        self.master = torch.zeros(C, P, dtype=f32, device=d)
        self.master_bf16 = torch.zeros(C, P, dtype=bf, device=d)
        self.global_local = torch.zeros(P, dtype=f32, device=d)
        self.h = torch.zeros(C, B, L.hidden, dtype=bf, device=d)
        self.dh = torch.zeros(C, B, L.hidden, dtype=bf, device=d)
        self.logits = torch.zeros(C, B, L.LD, dtype=bf, device=d)
        self.dlogits = torch.zeros(C, B, L.LD, dtype=bf, device=d)
        self.loss_sum = torch.zeros(C, dtype=f32, device=d)
        self.nonfinite = torch.zeros(C, dtype=i32, device=d)
        self.n_samples = torch.full((C,), shard, dtype=i32, device=d)
        self.groups = max(1, min(16, max(1, C // 8)))
        hp = self.heap
        hp.alloc("global", P * 4); hp.alloc("partial", self.groups * P * 4); hp.alloc("wsum", 64 * 4); hp.alloc("flags", 256)
        hp.alloc("round", 256)
        self.round_rec = hp.tensor("round", f32, (8,))
        self.round_rec.zero_()
        self.round_out = torch.zeros(8, dtype=f32, device=d)
        self.global_w = hp.tensor("global", f32, (P,))
        self.partial = hp.tensor("partial", f32, (self.groups, P))
        self.wsum = hp.tensor("wsum", f32, (self.groups,))
        self.opt_m = torch.zeros(self.shard_len, dtype=f32, device=d)
        self.opt_v = torch.zeros(self.shard_len, dtype=f32, device=d)
        self.stats = torch.zeros(4, dtype=f32, device=d)
        self.arrive_ctr = torch.zeros(2, dtype=i32, device=d)
        self.epoch = self.server_step = self.kernel_launches = 0

    # ---- data & model -------------------------------------------------------------------------------
    def load_data(self, x: torch.Tensor, labels: torch.Tensor) -> None:
        self.x.copy_(x.to(self.dev).reshape(self.C, self.shard, -1).to(torch.bfloat16))
        self.labels.copy_(labels.to(self.dev).to(torch.int32))

    def set_global(self, row: torch.Tensor) -> None:
        self.global_w.copy_(row.to(self.dev))
        self.global_local.copy_(self.global_w)
        self.heap.barrier()

    def get_global(self) -> torch.Tensor:
        out = torch.empty(self.P, dtype=torch.float32, device=self.dev)
        for r in range(self.world):
            s0, s1 = r * self.shard_len, min((r + 1) * self.shard_len, self.P)
            out[s0:s1].copy_(self.heap.tensor("global", torch.float32, (self.P,), peer=r)[s0:s1])
        return out

    # ---- round phases -------------------------------------------------------------------------------
    def broadcast(self, participate: Optional[torch.Tensor] = None) -> None:
        native.bcast_scatter_generic(self.heap.buffers["global"].peer_ptrs, self.shard_len, self.P, self.global_local,
                                     self.master, self.master_bf16, participate)
        self.kernel_launches += 1

    def _gemm(self, **kw) -> None:
        a = native.GemmArgs()
        for k, v in kw.items():
            setattr(a, k, v)
        a.C = self.C
        native.client_gemm(a)
        self.kernel_launches += 1

    def local_step(self, step_idx: int, lr: float, mu: float = 0.0, active: Optional[torch.Tensor] = None) -> None:
        L, P, C = self.L, self.P, self.C
        off = (step_idx * B) % self.shard
        act = native._vp(active)
        gw = self.global_local if mu > 0 else None
        mb, mf = self.master_bf16, self.master
        op = native.operand
        X = op(self.x, off * L.in_dim, self.shard * L.in_dim, L.in_dim, B, L.in_dim)
        W1 = op(mb, L.W1, P, L.in_dim, L.hidden, L.in_dim)
        W2 = op(mb, L.W2, P, L.hidden, L.classes, L.hidden)
        H = op(self.h, 0, B * L.hidden, L.hidden, B, L.hidden)
        DH = op(self.dh, 0, B * L.hidden, L.hidden, B, L.hidden)
        DL = op(self.dlogits, 0, B * L.LD, L.LD, B, L.LD)
        vp = native._vp
        # fc1 forward: H^T[f,s] = W1[f,:].X[s,:] + b1, ReLU
        self._gemm(variant=native.GEMM_FWD, M=L.hidden, N=B, K=L.in_dim, a=W1, b=X, active=act,
                   out_t=vp(self.h), out_stride=B * L.hidden, out_ld=L.hidden, bias=vp(mf, L.B1), bias_stride=P, relu=1)
        # fc2 forward: logits
        self._gemm(variant=native.GEMM_FWD, M=L.classes, N=B, K=L.hidden, a=W2, b=H, active=act,
                   out_t=vp(self.logits), out_stride=B * L.LD, out_ld=L.LD, bias=vp(mf, L.B2), bias_stride=P, relu=0)
        native.client_xent(C, B, L.classes, L.LD, active, self.logits, self.labels, off, self.dlogits, self.loss_sum,
                           self.nonfinite, b_master=mf, b_offset=L.B2, b_stride=P, b_global=gw, lr=lr, mu=mu)
        self.kernel_launches += 1
        # fc2 dgrad: dH^T[i,s] = sum_o W2[o,i] dL[s,o], gated by H>0; fc1-bias SGD on the row sums
        self._gemm(variant=native.GEMM_DGRAD, M=L.hidden, N=B, K=L.classes, a=W2, b=DL, active=act, lr=lr, mu=mu,
                   out_t=vp(self.dh), out_stride=B * L.hidden, out_ld=L.hidden, mask_t=vp(self.h),
                   b_master=vp(mf, L.B1), b_stride=P, b_global=vp(gw, L.B1) if gw is not None else None)
        # fc2 wgrad + SGD: dW2[o,i] = sum_s dL[s,o] H[s,i]
        self._gemm(variant=native.GEMM_WGRAD, M=L.classes, N=L.hidden, K=B, a=DL, b=H, active=act, lr=lr, mu=mu,
                   w_master=vp(mf, L.W2), w_stride=P, w_ld=L.hidden, w_global=vp(gw, L.W2) if gw is not None else None,
                   w_bf16=vp(mb, L.W2), wb_stride=P)
        # fc1 wgrad + SGD: dW1[f,i] = sum_s dH[s,f] X[s,i]
        self._gemm(variant=native.GEMM_WGRAD, M=L.hidden, N=L.in_dim, K=B, a=DH, b=X, active=act, lr=lr, mu=mu,
                   w_master=vp(mf, L.W1), w_stride=P, w_ld=L.in_dim, w_global=vp(gw, L.W1) if gw is not None else None,
                   w_bf16=vp(mb, L.W1), wb_stride=P)

    def reduce_and_update(self, weight: torch.Tensor, opt: str = "avg", server_lr: float = 1.0, beta1: float = 0.9,
                          beta2: float = 0.99, eps: float = 1e-3, accounting=None) -> None:
        if accounting is not None:
            # (participate, step_active, counters, local_steps): 8-float record, summed over ranks by the finish kernel
            part, step_active, counters, nsteps = accounting
            native.round_stats(self.loss_sum, self.nonfinite, part, step_active, counters, self.C, step_active.shape[0],
                               1.0 / nsteps, self.round_rec)
        native.fedavg_partial(self.master, self.global_local, weight, self.nonfinite, self.groups, self.partial, self.wsum)
        self.epoch += 1
        self.server_step += 1
        hb = self.heap.buffers
        native.fedavg_finish(hb["partial"].peer_ptrs, hb["wsum"].peer_ptrs, hb["flags"].peer_ptrs if self.world > 1 else None,
                             self.rank, self.groups, self.P, self.shard_len, self.epoch, self.global_w, self.opt_m,
                             self.opt_v, OPT_CODES[opt], server_lr, beta1, beta2, eps, self.server_step, self.stats,
                             self.arrive_ctr, hb["round"].peer_ptrs if accounting is not None else None, self.round_out)
        self.kernel_launches += 2 + (accounting is not None)

    def reset_round_stats(self) -> None:
        self.loss_sum.zero_()
        self.nonfinite.zero_()

    def close(self) -> None:
        self.heap.close()
