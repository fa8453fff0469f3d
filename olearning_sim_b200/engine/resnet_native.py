"""GPU-resident state + native local step of a ResNet-18 (GroupNorm, CIFAR stem) client population on one rank.

BASELINE config 3.  Every convolution of every client is an implicit GEMM on tcgen05 (``csrc/bgemm_tc.cuh``): the
forward pass writes the raw convolution output and accumulates the GroupNorm statistics in its epilogue, the data
gradient reads the *forward* weight copy transposed (MN-major operand, no second copy), the weight gradient keeps
three (tap, channel-block) accumulators in TMEM per tile and applies SGD / FedProx to the fp32 master row and its bf16
operand copy in the epilogue — clients never materialise weight gradients.  GroupNorm forward / backward, the pooled
classifier head and the SGD step on the normalisation / head parameters are the bandwidth kernels of
``csrc/nn_kernels.cuh``.  Replaces the per-phone training subprocess of the reference
(ols_core/taskMgr/utils/utils_run_task.py:481-514) for this family.

Packed parameter row (fp32 master, one row per client):
    [ conv weights as [Cout][kh*kw][Cin] (stem: [64][32], K = (kh*3+kw)*3+c padded to 32) | GN gamma/beta | fc.weight | fc.bias ]
The leading conv region has a bf16 operand copy (``wb``), refreshed by the broadcast and by every weight-gradient
epilogue; the trailing "small" region is updated by one fused SGD launch per step from the fp32 gradient side buffer.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from ..models.zoo import ModelSpec
from ..ops import bgemm as bgm
from ..ops import native
from ..parallel.comm import SymmetricHeap
from .flat_native import FlatWorkspace

B = 32           # samples per local step


def _al(n: int, a: int = 8) -> int:
    return (n + a - 1) // a * a


# ReLU-backward gates in the data-gradient epilogues (GroupNorm backward then reads no mask); OLSB_RELU_FUSE=0: gate inside
# the GroupNorm backward passes
FUSE_RELU_GATE = __import__("os").environ.get("OLSB_RELU_FUSE", "1") != "0"

class ResNetLayout:
    """Offsets of every parameter of the zoo's ``resnet18`` spec inside the packed row + pack / unpack."""

    def __init__(self, spec: ModelSpec, groups: int = 2):
        self.spec, self.groups = spec, groups
        self.conv: Dict[str, Tuple[int, int, int, int, int]] = {}      # name -> (offset, cout, cin_eff, ksize, ktot)
        self.small: Dict[str, Tuple[int, int]] = {}                    # name -> (offset in row, numel)
        o = 0
        for name, shp in spec.shapes:
            if name.endswith(".weight") and len(shp) == 4:
                cout, cin, kh, _ = shp
                ktot = 32 if name == "stem.weight" else kh * kh * cin
                self.conv[name] = (o, cout, cin, kh, ktot)
                o = _al(o + cout * ktot)
        self.P_bf16 = self.small_off = o
        for name, shp in spec.shapes:
            if not (name.endswith(".weight") and len(shp) == 4):
                n = math.prod(shp)
                self.small[name] = (o, n)
                o = _al(o + n)
        self.P = _al(o, 8)
        self.P_small = self.P - self.small_off
        self.width = spec.shapes[0][1][0]
        self.num_classes = dict(spec.shapes)["fc.weight"][0]
        # block structure: (name, stride, downsample, cin, cout)
        self.blocks: List[Tuple[str, int, bool, int, int]] = []
        cin = self.width
        for stage, (mult, stride) in enumerate([(1, 1), (2, 2), (4, 2), (8, 2)]):
            cout = self.width * mult
            for blk in range(2):
                s = stride if blk == 0 else 1
                self.blocks.append((f"l{stage}.{blk}", s, s != 1 or cin != cout, cin, cout))
                cin = cout
        self.feat = cin

    def pack(self, flat_spec: torch.Tensor) -> torch.Tensor:
        p = self.spec.unflatten(flat_spec.float().cpu())
        row = torch.zeros(self.P)
        for name, (off, cout, cin, k, ktot) in self.conv.items():
            w = p[name].permute(0, 2, 3, 1).reshape(cout, k * k * cin)          # [cout][kh][kw][cin]
            if name == "stem.weight":
                w = torch.nn.functional.pad(w, (0, ktot - w.shape[1]))
            row[off:off + cout * ktot] = w.reshape(-1)
        for name, (off, n) in self.small.items():
            row[off:off + n] = p[name].reshape(-1)
        return row

    def unpack(self, row: torch.Tensor) -> torch.Tensor:
        row = row.float().cpu()
        out = {}
        for name, shp in self.spec.shapes:
            if name in self.conv:
                off, cout, cin, k, ktot = self.conv[name]
                w = row[off:off + cout * ktot].reshape(cout, ktot)[:, :k * k * cin].reshape(cout, k, k, cin)
                out[name] = w.permute(0, 3, 1, 2).contiguous()
            else:
                off, n = self.small[name]
                out[name] = row[off:off + n].reshape(shp)
        return self.spec.flatten(out)


class ResNetWorkspace(FlatWorkspace):
    """Packed replicas + activations of ``num_clients`` ResNet-18 clients; ``wave`` clients train at a time."""

    def __init__(self, num_clients: int, shard: int, spec: ModelSpec, device="cuda", heap: Optional[SymmetricHeap] = None,
                 wave: Optional[int] = None, groups: int = 2):
        self.L = ResNetLayout(spec, groups)
        super().__init__(num_clients, self.L.P, device=device, heap=heap)
        L, d, C = self.L, self.dev, num_clients
        bf, f32 = torch.bfloat16, torch.float32
        self.shard = shard
        self.wb = torch.zeros(C, L.P_bf16, dtype=bf, device=d)
        self.n_samples = torch.full((C,), shard, dtype=torch.int32, device=d)
        self.xcol = torch.zeros(C * shard, 32, 32, 32, dtype=bf, device=d)          # stem im2col of every sample
        self.labels = torch.zeros(C, shard, dtype=torch.int32, device=d)
        W = self.width = L.width
        # activation bytes per client: see _alloc_wave; size the wave from free memory unless told
        per_client = self._act_elems() * 2 + L.P_small * 4
        if wave is None:
            free, _ = torch.cuda.mem_get_info(d)
            wave = max(1, min(C, int(free * 0.8) // max(per_client, 1)))
        self.wave = max(1, min(wave, C))
        self._alloc_wave(self.wave)

    # ---- buffers ------------------------------------------------------------------------------------
    def _stage_dims(self):
        """(resolution, channels) of the four stages."""
        W = self.L.width
        return [(32, W), (16, 2 * W), (8, 4 * W), (4, 8 * W)]

    def _act_elems(self) -> int:
        W = self.L.width
        n = 2 * B * 32 * 32 * W                                   # stem raw + act
        res, ch = 32, W
        for (name, s, down, cin, cout) in self.L.blocks:
            res = res // s
            n += (4 + (1 if down else 0)) * B * res * res * cout   # raw1, y1, raw2, hout (+rawd)
        n += 3 * B * 32 * 32 * W                                   # gradient scratch X, Y, Z
        return n

    def _alloc_wave(self, wv: int) -> None:
        d, bf, f32 = self.dev, torch.bfloat16, torch.float32
        L, W = self.L, self.L.width

        def act(res, ch):
            return torch.empty(wv * B, res, res, ch, dtype=bf, device=d)
        self.raw0, self.a0 = act(32, W), act(32, W)
        self.blk: List[Dict[str, torch.Tensor]] = []
        res = 32
        ngn = 1
        for (name, s, down, cin, cout) in L.blocks:
            res //= s
            t = {"raw1": act(res, cout), "y1": act(res, cout), "raw2": act(res, cout), "hout": act(res, cout)}
            if down:
                t["rawd"] = act(res, cout)
            ngn += 3 if down else 2
            self.blk.append(t)
        self.gx, self.gy, self.gz = act(32, W), act(32, W), act(32, W)
        self.stats = torch.zeros(ngn, wv * B, L.groups, 2, dtype=f32, device=d)
        self.sb = torch.zeros(wv * B, L.groups, 2, dtype=f32, device=d)
        self.gsmall = torch.zeros(wv, L.P_small, dtype=f32, device=d)

    # ---- data & model -------------------------------------------------------------------------------
    def load_data(self, x_u8: torch.Tensor, labels: torch.Tensor) -> None:
        """x_u8: [C, shard, 32, 32, 3] uint8 on this device."""
        before = bgm.launches
        bgm.stem_im2col(x_u8.reshape(-1, 32, 32, 3).contiguous(), self.xcol)
        self.labels.copy_(labels.to(self.dev).to(torch.int32))
        self.kernel_launches += bgm.launches - before

    def broadcast(self, participate: Optional[torch.Tensor] = None) -> None:
        native.bcast_scatter_generic(self.heap.buffers["global"].peer_ptrs, self.shard_len, self.P, self.global_local,
                                     self.master, self.wb, participate, self.L.P_bf16, self.L.P_bf16)
        self.kernel_launches += 1

    # ---- one local SGD / FedProx step of every client ----------------------------------------------
    def local_step(self, step_idx: int, lr: float, mu: float = 0.0, active: Optional[torch.Tensor] = None) -> None:
        before = bgm.launches
        for w0 in range(0, self.C, self.wave):
            w1 = min(w0 + self.wave, self.C)
            self._step_wave(w0, w1, step_idx, lr, mu, None if active is None else active[w0:w1])
        self.kernel_launches += bgm.launches - before

    def _gn_apply(self, n, act, li, raw, gname, out, res_h, ch, relu, raw2=None, li2=None, gname2=None, res=None):
        L = self.L
        p = bgm.GnApplyParams()
        p.C, p.imgs, p.HW, p.Ch, p.groups = n, B, res_h * res_h, ch, L.groups
        p.active = bgm._ptr(act)
        p.raw, p.stats = bgm._ptr(raw), bgm._ptr(self.stats[li])
        p.gamma_off, p.beta_off = L.small[gname + ".gamma"][0], L.small[gname + ".beta"][0]
        if raw2 is not None:
            p.raw2, p.stats2 = bgm._ptr(raw2), bgm._ptr(self.stats[li2])
            p.gamma2_off, p.beta2_off = L.small[gname2 + ".gamma"][0], L.small[gname2 + ".beta"][0]
        p.res = bgm._ptr(res)
        p.master, p.P, p.relu, p.out = bgm._ptr(self._m), self.P, 1 if relu else 0, bgm._ptr(out)
        bgm.gn_apply(p)

    def _gn_bwd(self, n, act, li, dy, mask, raw, gname, dx, res_h, ch, dres=None):
        L = self.L
        p = bgm.GnBwdParams()
        p.C, p.imgs, p.HW, p.Ch, p.groups = n, B, res_h * res_h, ch, L.groups
        p.active = bgm._ptr(act)
        p.dy, p.mask, p.raw, p.stats = bgm._ptr(dy), bgm._ptr(mask), bgm._ptr(raw), bgm._ptr(self.stats[li])
        p.master, p.P, p.gamma_off = bgm._ptr(self._m), self.P, L.small[gname + ".gamma"][0]
        p.gsmall, p.gs_stride = bgm._ptr(self.gsmall), L.P_small
        p.g_gamma_off = L.small[gname + ".gamma"][0] - L.small_off
        p.g_beta_off = L.small[gname + ".beta"][0] - L.small_off
        p.sb, p.dx, p.dres = bgm._ptr(self.sb), bgm._ptr(dx), bgm._ptr(dres)
        bgm.gn_bwd(p)

    def _step_wave(self, w0: int, w1: int, step_idx: int, lr: float, mu: float, act: Optional[torch.Tensor]) -> None:
        L, n = self.L, w1 - w0
        W, G = L.width, L.groups
        self._m = self.master[w0:w1]
        m, wb, Pb, P = self._m, self.wb[w0:w1], L.P_bf16, self.P
        gw = self.global_local if mu > 0 else None
        off = (step_idx * B) % self.shard
        bgm.memset(self.stats); bgm.memset(self.gsmall)
        xcol = self.xcol[w0 * self.shard:w1 * self.shard]

        def cf(x, name, H, cin, cout, k, s, raw, li, **kw):
            bgm.conv_fwd(n, B, H, cin, cout, k, s, x, wb, L.conv[name][0], Pb, raw, self.stats[li], G, act, **kw)

        def cd(dy, name, H, cin, cout, k, s, dx, addend=None, relu_mask=None, mask_even=True):
            bgm.conv_dgrad(n, B, H, cin, cout, k, s, dy, wb, L.conv[name][0], Pb, dx, addend, act, relu_mask=relu_mask,
                           mask_even=mask_even)

        def cw(dy, x, name, H, cin, cout, k, s, **kw):
            bgm.conv_wgrad(n, B, H, cin, cout, k, s, dy, x, m, wb, L.conv[name][0], P, Pb, lr, mu, gw, act, **kw)

        # ------------------------------------------------------------------ forward
        cf(xcol, "stem.weight", 32, 32, W, 1, 1, self.raw0, 0, x_cimgs=self.shard, x_img0=off)
        self._gn_apply(n, act, 0, self.raw0, "stem.gn", self.a0, 32, W, True)
        h, res, li = self.a0, 32, 1
        gn_idx = []
        for bi, (name, s, down, cin, cout) in enumerate(L.blocks):
            t = self.blk[bi]
            ro = res // s
            l1, l2, ld = li, li + 1, li + 2
            cf(h, f"{name}.c1.weight", res, cin, cout, 3, s, t["raw1"], l1)
            self._gn_apply(n, act, l1, t["raw1"], f"{name}.g1", t["y1"], ro, cout, True)
            cf(t["y1"], f"{name}.c2.weight", ro, cout, cout, 3, 1, t["raw2"], l2)
            if down:
                cf(h, f"{name}.ds.weight", res, cin, cout, 1, s, t["rawd"], ld)
                self._gn_apply(n, act, l2, t["raw2"], f"{name}.g2", t["hout"], ro, cout, True, raw2=t["rawd"], li2=ld,
                               gname2=f"{name}.dg")
            else:
                self._gn_apply(n, act, l2, t["raw2"], f"{name}.g2", t["hout"], ro, cout, True, res=h)
            gn_idx.append((l1, l2, ld))
            li += 3 if down else 2
            h, res = t["hout"], ro
        # ------------------------------------------------------------------ head (loss, fc update, d hout)
        X, Y, Z = self.gx, self.gy, self.gz
        hp = bgm.PoolHeadParams()
        hp.C, hp.imgs, hp.HW, hp.Ch, hp.classes = n, B, res * res, L.feat, L.num_classes
        hp.active, hp.h, hp.dh = bgm._ptr(act), bgm._ptr(h), bgm._ptr(X)
        hp.labels, hp.label_stride, hp.label_off = bgm._ptr(self.labels[w0:w1]), self.shard, off
        hp.master, hp.P = bgm._ptr(m), P
        hp.w_off, hp.b_off = L.small["fc.weight"][0], L.small["fc.bias"][0]
        hp.global_w = bgm._ptr(gw) if gw is not None else None
        hp.lr, hp.mu = lr, mu
        hp.loss_sum, hp.nonfinite = bgm._ptr(self.loss_sum[w0:w1]), bgm._ptr(self.nonfinite[w0:w1])
        hp.apply_update = 1
        bgm.pool_head(hp)
        # ------------------------------------------------------------------ backward
        # `cur` holds d hout of the block.  Every data-gradient epilogue gates its output with the ReLU of the activation it
        # belongs to (relu_mask), so the GroupNorm backward passes read (dy, raw) only; the head's gradient is the one
        # ungated input (gated inside the first GroupNorm backward, which then also writes the gated copy for the shortcut).
        fuse = FUSE_RELU_GATE
        cur, spare, gated = X, Z, False
        for bi in range(len(L.blocks) - 1, -1, -1):
            name, s, down, cin, cout = L.blocks[bi]
            t = self.blk[bi]
            l1, l2, ld = gn_idx[bi]
            ro = res
            rin = ro * s
            hin = self.blk[bi - 1]["hout"] if bi > 0 else self.a0
            # gate in the epilogue only where the data gradient is compute-bound (>= 128 channels): at 64 channels it is
            # close to HBM-bound and a third operand stream costs what the GroupNorm passes save (measured)
            f2 = fuse and cout >= 128          # d y1  gated by conv2's data gradient
            f1 = fuse and cin >= 128           # d hin gated by conv1's (and the shortcut's) data gradient
            if gated:      # R = gated d hout (also the gradient of the shortcut branch), T = free buffer
                self._gn_bwd(n, act, l2, cur, None, t["raw2"], f"{name}.g2", Y, ro, cout)
                R, T = cur, spare
            else:
                self._gn_bwd(n, act, l2, cur, t["hout"], t["raw2"], f"{name}.g2", Y, ro, cout, dres=spare)
                R, T = spare, cur
            cd(Y, f"{name}.c2.weight", ro, cout, cout, 3, 1, T, relu_mask=t["y1"] if f2 else None)      # T = d y1
            cw(Y, t["y1"], f"{name}.c2.weight", ro, cout, cout, 3, 1)
            self._gn_bwd(n, act, l1, T, None if f2 else t["y1"], t["raw1"], f"{name}.g1", Y, ro, cout)   # Y = d raw1
            if down:
                # conv path into all four pixel parities; the (even, even) class stays ungated until the shortcut is added
                cd(Y, f"{name}.c1.weight", rin, cin, cout, 3, s, T, relu_mask=hin if f1 else None, mask_even=False)
                cw(Y, hin, f"{name}.c1.weight", rin, cin, cout, 3, s)
                self._gn_bwd(n, act, ld, R, None, t["rawd"], f"{name}.dg", Y, ro, cout)     # Y = d rawd
                cd(Y, f"{name}.ds.weight", rin, cin, cout, 1, s, T, addend=T, relu_mask=hin if f1 else None)   # += shortcut
                cw(Y, hin, f"{name}.ds.weight", rin, cin, cout, 1, s)
            else:
                cd(Y, f"{name}.c1.weight", rin, cin, cout, 3, 1, T, addend=R, relu_mask=hin if f1 else None)   # + identity
                cw(Y, hin, f"{name}.c1.weight", rin, cin, cout, 3, 1)
            cur, spare, gated = T, R, f1
            res = rin
        self._gn_bwd(n, act, 0, cur, None if gated else self.a0, self.raw0, "stem.gn", Y, 32, W)
        cw(Y, xcol, "stem.weight", 32, 32, W, 1, 1, x_cimgs=self.shard, x_img0=off)
        # ------------------------------------------------------------------ GN parameters: one fused SGD launch
        gn_n = L.small["fc.weight"][0] - L.small_off        # fc was updated by the head kernel
        bgm.slice_sgd(m, P, L.small_off, gn_n, self.gsmall, L.P_small, gw, act, n, lr, mu)
