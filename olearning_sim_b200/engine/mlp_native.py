"""GPU-resident state of an MLP client population (fully-connected families) on one rank.

Same structure as ``Cnn2Workspace`` but built from the generic per-client GEMM kernels (``csrc/linear_tc.cuh``):
a local step of an L-layer MLP is 3L launches — L forward GEMMs (+bias, ReLU on hidden layers), softmax-xent (+last
bias SGD), and per layer from the top a data gradient (+ReLU gate, +bias SGD of the layer below) and a weight gradient
(+SGD); six for the 2-layer MLP of BASELINE config #1.
The packed parameter row is [W1 | b1 | W2 | b2 | ...] with every block starting on an 8-element boundary so the bf16 copy
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
    """Packed row of an MLP with any number of hidden layers: [W1 | b1 | W2 | b2 | ...], every block on an 8-element
    boundary.  ``dims`` = (in_dim, hidden_1, ..., hidden_k, classes); the 2-layer constructor arguments are kept."""

    def __init__(self, in_dim: int = 784, hidden=200, classes: int = 10):
        hid = [int(hidden)] if isinstance(hidden, (int, float)) else [int(h) for h in hidden]
        self.dims = [int(in_dim)] + hid + [int(classes)]
        assert all(d % 8 == 0 for d in self.dims[:-1]), "row strides must be multiples of 16 bytes for TMA"
        self.in_dim, self.hidden, self.classes = self.dims[0], hid[0], self.dims[-1]
        self.nlayers = len(self.dims) - 1
        self.W, self.B = [], []
        o = 0
        for l in range(1, len(self.dims)):
            self.W.append(o)
            o = _al8(o + self.dims[l] * self.dims[l - 1])
            self.B.append(o)
            o = _al8(o + self.dims[l])
        self.P = o
        self.LD = 16                                     # padded class dimension of logits / dlogits
        assert self.classes <= self.LD
        # 2-layer names used by older call sites
        self.W1, self.B1 = self.W[0], self.B[0]
        self.W2, self.B2 = (self.W[1], self.B[1]) if self.nlayers > 1 else (self.W[0], self.B[0])

    def _blocks(self):
        for l in range(self.nlayers):
            yield self.W[l], self.dims[l + 1] * self.dims[l]
            yield self.B[l], self.dims[l + 1]

    def pack(self, flat_spec: torch.Tensor) -> torch.Tensor:
        """zoo flat order (fc1.weight, fc1.bias, fc2.weight, ...) -> aligned packed row."""
        out = torch.zeros(self.P)
        o = 0
        for dst, n in self._blocks():
            out[dst:dst + n] = flat_spec[o:o + n]
            o += n
        return out

    def unpack(self, row: torch.Tensor) -> torch.Tensor:
        return torch.cat([row[dst:dst + n] for dst, n in self._blocks()])


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
        # hidden activations / their gradients, one pair per hidden layer
        self.acts = [torch.zeros(C, B, h, dtype=bf, device=d) for h in L.dims[1:-1]]
        self.dacts = [torch.zeros(C, B, h, dtype=bf, device=d) for h in L.dims[1:-1]]
        self.h, self.dh = self.acts[0], self.dacts[0]
        self.logits = torch.zeros(C, B, L.LD, dtype=bf, device=d)
        self.dlogits = torch.zeros(C, B, L.LD, dtype=bf, device=d)
        self.loss_sum = torch.zeros(C, dtype=f32, device=d)
        self.nonfinite = torch.zeros(C, dtype=i32, device=d)
        self.n_samples = torch.full((C,), shard, dtype=i32, device=d)
        self.groups = max(1, min(16, max(1, C // 8)))
        hp = self.heap
        hp.alloc("global", P * 4, multicast=True); hp.alloc("partial", self.groups * P * 4, multicast=True); hp.alloc("wsum", 64 * 4); hp.alloc("flags", 256)
        hp.alloc("round", 256)
        self.round_rec = hp.tensor("round", f32, (native.ROUND_REC,))
        self.round_rec.zero_()
        self.round_out = torch.zeros(native.ROUND_REC + 8, dtype=f32, device=d)      # [ROUND_REC] = peer-lost code
        self.global_w = hp.tensor("global", f32, (P,))
        self.partial = hp.tensor("partial", f32, (self.groups, P))
        self.wsum = hp.tensor("wsum", f32, (self.groups,))
        self.opt_m = torch.zeros(self.shard_len, dtype=f32, device=d)
        self.opt_v = torch.zeros(self.shard_len, dtype=f32, device=d)
        self.stats = torch.zeros(4, dtype=f32, device=d)
        self.arrive_ctr = torch.zeros(4, dtype=i32, device=d)                        # [2] = sticky peer-lost code
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
        """One SGD / FedProx step of every client: L forward GEMMs (+bias, ReLU), softmax-xent (+last bias SGD), then per
        layer from the top: data gradient (+ReLU gate, +bias SGD of the layer below) and weight gradient (+SGD).
        3L launches for an L-layer MLP."""
        L, P, C = self.L, self.P, self.C
        off = (step_idx * B) % self.shard
        act = native._vp(active)
        gw = self.global_local if mu > 0 else None
        mb, mf = self.master_bf16, self.master
        op, vp = native.operand, native._vp
        d, nl = L.dims, L.nlayers
        # operands: layer inputs a[0..nl-1], weights, output-side gradients dz[1..nl]
        a_in = [op(self.x, off * d[0], self.shard * d[0], d[0], B, d[0])] + \
               [op(t, 0, B * h, h, B, h) for t, h in zip(self.acts, d[1:-1])]
        w = [op(mb, L.W[l], P, d[l], d[l + 1], d[l]) for l in range(nl)]
        dz = [op(t, 0, B * h, h, B, h) for t, h in zip(self.dacts, d[1:-1])] + [op(self.dlogits, 0, B * L.LD, L.LD, B, L.LD)]
        for l in range(nl):                                  # forward
            last = l == nl - 1
            out = self.logits if last else self.acts[l]
            ld = L.LD if last else d[l + 1]
            self._gemm(variant=native.GEMM_FWD, M=d[l + 1], N=B, K=d[l], a=w[l], b=a_in[l], active=act,
                       out_t=vp(out), out_stride=B * ld, out_ld=ld, bias=vp(mf, L.B[l]), bias_stride=P, relu=0 if last else 1)
        native.client_xent(C, B, L.classes, L.LD, active, self.logits, self.labels, off, self.dlogits, self.loss_sum,
                           self.nonfinite, b_master=mf, b_offset=L.B[nl - 1], b_stride=P, b_global=gw, lr=lr, mu=mu)
        self.kernel_launches += 1
        for l in range(nl - 1, -1, -1):                      # backward, top layer first
            if l > 0:
                # dA_{l}^T[i,s] = sum_o W_l[o,i] dZ_l[s,o], gated by A_l > 0; bias SGD of layer l-1 on the row sums
                self._gemm(variant=native.GEMM_DGRAD, M=d[l], N=B, K=d[l + 1], a=w[l], b=dz[l], active=act, lr=lr, mu=mu,
                           out_t=vp(self.dacts[l - 1]), out_stride=B * d[l], out_ld=d[l], mask_t=vp(self.acts[l - 1]),
                           b_master=vp(mf, L.B[l - 1]), b_stride=P,
                           b_global=vp(gw, L.B[l - 1]) if gw is not None else None)
            # dW_l[o,i] = sum_s dZ_l[s,o] A_{l-1}[s,i]; fused SGD on the fp32 master + bf16 operand copy
            self._gemm(variant=native.GEMM_WGRAD, M=d[l + 1], N=d[l], K=B, a=dz[l], b=a_in[l], active=act, lr=lr, mu=mu,
                       w_master=vp(mf, L.W[l]), w_stride=P, w_ld=d[l], w_global=vp(gw, L.W[l]) if gw is not None else None,
                       w_bf16=vp(mb, L.W[l]), wb_stride=P)

    def reduce_and_update(self, weight: torch.Tensor, opt: str = "avg", server_lr: float = 1.0, beta1: float = 0.9,
                          beta2: float = 0.99, eps: float = 1e-3, accounting=None) -> None:
        if accounting is not None:
            # (participate, step_active, counters, local_steps): 8-float record, summed over ranks by the finish kernel
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
