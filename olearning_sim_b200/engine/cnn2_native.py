"""GPU-resident state of the CNN-2 client population on one rank (the B200-native `Actor`).

The reference gives every simulated phone a fresh python subprocess inside a Ray actor
(ols_core/taskMgr/utils/utils_run_task.py:146-147,481-514).  Here a rank holds *all* of its clients as rows of
packed tensors that never leave HBM during a task:

    data      x5   [C*shard, 32,32,16] bf16 (ingested once)      labels [C, shard] int32
    weights   master [C, P] fp32  +  bf16 tensor-core operand copies (w1_fwd, w2_fwd, w2_dgrad)
    acts      p1/code1, p2/code2, dy2, dp1 (one local batch per client)
    server    global [P] fp32 in the symmetric heap (each rank owns the optimiser state of one shard)

and a round is: ``broadcast`` (1 kernel) -> ``local_step`` x S (6 kernels each) -> ``reduce_and_update`` (2 kernels).
"""
from __future__ import annotations

from typing import Optional

import torch

from ..ops import native
from ..ops.layout import CNN2
from ..parallel.comm import SymmetricHeap

B = CNN2.BATCH
OPT_CODES = {"avg": 0, "sgd": 0, "fedavg": 0, "momentum": 1, "fedavgm": 1, "adam": 2, "fedadam": 2}


class Cnn2Workspace:
    def __init__(self, num_clients: int, shard: int, device="cuda", heap: Optional[SymmetricHeap] = None,
                 reduce_groups: int = 0, wave: Optional[int] = None, hbm_budget_bytes: Optional[int] = None):
        """``wave``: number of clients whose activations are resident at once.  Weights, optimiser-free masters and data
        shards of *all* clients stay in HBM (1.1 MB + data per client); the 4.6 MB of per-step activations are
        time-multiplexed over waves when the population does not fit (the reference scales clients by adding Ray
        actors; here the "long axis" is clients x parameters).  Default: everything resident if it fits
        in ``hbm_budget_bytes`` (80 % of free memory), otherwise the largest wave that does."""
        assert shard % B == 0, f"client shard size must be a multiple of the local batch ({B})"
        native.require()
        self.C, self.shard, self.dev = num_clients, shard, torch.device(device)
        self.heap = heap or SymmetricHeap(device=self.dev if self.dev.index is not None else None)
        self.rank, self.world = self.heap.rank, self.heap.world
        P = CNN2.P
        self.P = P
        self.shard_len = ((P // 4 + self.world - 1) // self.world) * 4
        sms = native.num_sms()
        if reduce_groups <= 0:
            # enough (param-chunk x client-group) CTAs to saturate HBM: P/4/128 = 186 blocks per group
            reduce_groups = max(1, min(16, (4 * sms) // 186, max(1, num_clients // 8)))
        self.groups = reduce_groups
        C = num_clients
        d = self.dev
        bf, u8, f32, i32 = torch.bfloat16, torch.uint8, torch.float32, torch.int32
        act_bytes = B * (16 * 16 * 32 * 3 + 8 * 8 * 64 * 3 + 16 * 16 * 64 * 2 + 16 * 16 * 32 * 2)      # per client
        fixed_bytes = P * 4 + (CNN2.W1_N + CNN2.W2_N) * 2 + shard * (32 * 32 * 16 * 2 + 4)
        if wave is None:
            if hbm_budget_bytes is None:
                free, _ = torch.cuda.mem_get_info(d)
                hbm_budget_bytes = int(free * 0.8)
            room = hbm_budget_bytes - C * fixed_bytes
            wave = C if room >= C * act_bytes else max(sms, (room // act_bytes) // sms * sms)
            if wave <= 0:
                raise MemoryError(f"{C} CNN-2 clients need {C * fixed_bytes / 2**30:.1f} GiB of weights+data alone")
        self.wave = W = max(1, min(int(wave), C))
        self.x5 = torch.empty(C * shard, 32, 32, 16, dtype=bf, device=d)
        self.labels = torch.zeros(C, shard, dtype=i32, device=d)
        self.master = torch.zeros(C, P, dtype=f32, device=d)
        self.global_local = torch.zeros(P, dtype=f32, device=d)
        self.w1_fwd = torch.zeros(C, CNN2.W1_N, dtype=bf, device=d)
        self.w2_fwd = torch.zeros(C, CNN2.W2_N, dtype=bf, device=d)
        self.w2_dgrad = None      # not needed: the data-gradient kernel reads w2_fwd as an MN-major operand
        self.p1 = torch.empty(W * B, 16, 16, 32, dtype=bf, device=d)
        self.code1 = torch.empty(W * B, 16, 16, 32, dtype=u8, device=d)
        self.p2 = torch.empty(W * B, 8, 8, 64, dtype=bf, device=d)
        self.code2 = torch.empty(W * B, 8, 8, 64, dtype=u8, device=d)
        self.dy2 = torch.empty(W * B, 16, 16, 64, dtype=bf, device=d)
        self.dp1 = torch.empty(W * B, 16, 16, 32, dtype=bf, device=d)
        self.loss_sum = torch.zeros(C, dtype=f32, device=d)
        self.nonfinite = torch.zeros(C, dtype=i32, device=d)
        self.n_samples = torch.full((C,), shard, dtype=i32, device=d)
        # symmetric (peer-visible) state
        h = self.heap
        h.alloc("global", P * 4, multicast=True)
        h.alloc("partial", self.groups * P * 4, multicast=True)
        h.alloc("wsum", max(self.groups, 64) * 4)
        h.alloc("flags", 256)
        h.alloc("round", 256)
        self.round_rec = h.tensor("round", f32, (native.ROUND_REC,))
        self.round_rec.zero_()
        self.round_out = torch.zeros(native.ROUND_REC + 8, dtype=f32, device=d)      # [ROUND_REC] = peer-lost code
        self.global_w = h.tensor("global", f32, (P,))
        self.partial = h.tensor("partial", f32, (self.groups, P))
        self.wsum = h.tensor("wsum", f32, (self.groups,))
        self.opt_m = torch.zeros(self.shard_len, dtype=f32, device=d)
        self.opt_v = torch.zeros(self.shard_len, dtype=f32, device=d)
        self.stats = torch.zeros(4, dtype=f32, device=d)
        self.arrive_ctr = torch.zeros(4, dtype=i32, device=d)                        # [2] = sticky peer-lost code
        self.epoch = 0
        self.server_step = 0
        self.kernel_launches = 0

    # ---- data & model ----------------------------------------------------------------------------
    def load_data(self, x_u8: torch.Tensor, labels: torch.Tensor, n_samples: Optional[torch.Tensor] = None) -> None:
        """x_u8 uint8 [C, shard, 32,32,3] on the device (or pinned host; copied), labels int32 [C, shard]."""
        x = x_u8.to(self.dev, non_blocking=True).reshape(self.C * self.shard, 32, 32, 3).contiguous()
        native.ingest_u8(x, self.x5)
        self.labels.copy_(labels.to(self.dev).to(torch.int32))
        if n_samples is not None:
            self.n_samples.copy_(n_samples.to(self.dev).to(torch.int32))
        self.kernel_launches += 1

    def ingest_slice(self, x_u8_dev: torch.Tensor, first_sample: int) -> None:
        """Ingest a contiguous block of samples (streaming H2D path of the data feeder)."""
        n = x_u8_dev.shape[0]
        native.ingest_u8(x_u8_dev.reshape(n, 32, 32, 3), self.x5[first_sample:first_sample + n])
        self.kernel_launches += 1

    def set_global(self, flat: torch.Tensor) -> None:
        self.global_w.copy_(flat.to(self.dev))
        self.global_local.copy_(self.global_w)
        self.heap.barrier()

    def get_global(self) -> torch.Tensor:
        """Assemble the authoritative global model (every rank owns one shard) on this rank."""
        out = torch.empty(self.P, dtype=torch.float32, device=self.dev)
        for r in range(self.world):
            s0 = r * self.shard_len
            s1 = min(s0 + self.shard_len, self.P)
            peer = self.heap.tensor("global", torch.float32, (self.P,), peer=r)
            out[s0:s1].copy_(peer[s0:s1])
        return out

    # ---- round phases ----------------------------------------------------------------------------
    def broadcast(self, participate: Optional[torch.Tensor] = None) -> None:
        native.bcast_scatter_cnn2(self.heap.buffers["global"].peer_ptrs, self.rank, self.shard_len, self.global_local,
                                  self.master, participate, self.w1_fwd, self.w2_fwd, self.w2_dgrad)
        self.kernel_launches += 1

    def local_step(self, step_idx: int, lr: float, mu: float = 0.0, active: Optional[torch.Tensor] = None,
                   nsplit: int = 0) -> None:
        off = (step_idx * B) % self.shard
        for w0 in range(0, self.C, self.wave):                       # waves share the activation buffers
            w1 = min(w0 + self.wave, self.C)
            a = native.Cnn2StepArgs()
            x5 = self.x5[w0 * self.shard:w1 * self.shard]
            a.x5 = x5.data_ptr(); a.x_total = x5.shape[0]; a.x_n_stride = self.shard; a.x_n_off = off
            a.labels = self.labels[w0:w1].data_ptr(); a.label_stride = self.shard; a.label_off = off
            a.master = self.master[w0:w1].data_ptr(); a.global_w = self.global_local.data_ptr(); a.P = self.P; a.C = w1 - w0
            a.active = active[w0:w1].data_ptr() if active is not None else None
            a.w1_fwd = self.w1_fwd[w0:w1].data_ptr(); a.w2_fwd = self.w2_fwd[w0:w1].data_ptr(); a.w2_dgrad = None
            a.p1 = self.p1.data_ptr(); a.code1 = self.code1.data_ptr(); a.p2 = self.p2.data_ptr()
            a.code2 = self.code2.data_ptr(); a.dy2 = self.dy2.data_ptr(); a.dp1 = self.dp1.data_ptr()
            a.loss_sum = self.loss_sum[w0:w1].data_ptr(); a.nonfinite = self.nonfinite[w0:w1].data_ptr()
            a.lr = lr; a.mu = mu; a.use_prox = 1 if mu > 0 else 0; a.nsplit = nsplit
            native.cnn2_local_step(a)
            self.kernel_launches += native.KERNELS_PER_LOCAL_STEP

    def reduce_and_update(self, weight: torch.Tensor, opt: str = "avg", server_lr: float = 1.0, beta1: float = 0.9,
                          beta2: float = 0.99, eps: float = 1e-3, accounting=None) -> None:
        """FedAvg reduce (masked, weighted) + server optimiser; publishes the new global shard to all peers."""
        if accounting is not None:
            # (participate, step_active, counters, local_steps): 8-float record, summed over ranks by the finish kernel
            part, step_active, counters, nsteps = accounting[:4]
            class_id, nclass = (accounting[4], accounting[5]) if len(accounting) > 4 else (None, 0)
            native.round_stats(self.loss_sum, self.nonfinite, part, step_active, counters, self.C, step_active.shape[0],
                               1.0 / nsteps, self.round_rec, class_id, nclass)
        native.fedavg_partial(self.master, self.global_local, weight, self.nonfinite, self.groups, self.partial,
                              self.wsum)
        self.epoch += 1
        self.server_step += 1
        hb = self.heap.buffers
        native.fedavg_finish(hb["partial"].peer_ptrs, hb["wsum"].peer_ptrs,
                             hb["flags"].peer_ptrs if self.world > 1 else None, self.rank, self.groups, self.P,
                             self.shard_len, self.epoch, self.global_w, self.opt_m, self.opt_v,
                             OPT_CODES[opt], server_lr, beta1, beta2, eps, self.server_step, self.stats,
                             self.arrive_ctr, hb["round"].peer_ptrs if accounting is not None else None, self.round_out,
                             8 + 2 * (accounting[5] if accounting is not None and len(accounting) > 4 else 0),
                             partial_mc=hb["partial"].mc_ptr, global_mc=hb["global"].mc_ptr)
        self.kernel_launches += 2 + (accounting is not None)

    def reset_round_stats(self) -> None:
        self.loss_sum.zero_()
        self.nonfinite.zero_()

    def close(self) -> None:
        self.heap.close()
