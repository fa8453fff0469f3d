"""GPU-resident state + native local step of a BERT-mini (4 layers, hidden 256, 4 heads, seq 128) client population.

BASELINE config 5.  Every matrix product of every client — QKV / output / feed-forward projections, their data and
weight gradients, and the attention products Q K^T, P V and their five backward counterparts — is one launch of the
batched tcgen05 GEMM family (``csrc/bgemm_tc.cuh``) over all clients of the rank, with the element-wise work fused
into the epilogues: bias + residual add, GELU (saving the pre-activation), GELU', softmax over the 128-wide score row a
thread owns in TMEM, softmax backward, and SGD / FedProx on the fp32 master + bf16 operand copy for the weight
gradients.  LayerNorm forward / backward (which also yields the bias gradients of the preceding projection), the
embedding gather and its sparse scatter-SGD, and the classifier head are the bandwidth kernels of
``csrc/nn_kernels.cuh``.  Replaces the per-phone training subprocess of the reference
(ols_core/taskMgr/utils/utils_run_task.py:481-514) for this family.

Packed row:  [ per layer: Wqkv | Wo | Wff1 | Wff2  (bf16 operand copy) | LN / bias vectors | pos.emb | cls | tok.emb ]
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

B = 32


def _al(n: int, a: int = 8) -> int:
    return (n + a - 1) // a * a


class BertLayout:
    def __init__(self, spec: ModelSpec):
        self.spec = spec
        shp = dict(spec.shapes)
        self.vocab, self.H = shp["tok.emb"]
        self.S = shp["pos.emb"][0]
        self.layers = sum(1 for n, _ in spec.shapes if n.endswith(".qkv.weight"))
        self.FF = shp["h0.ff1.weight"][0]
        self.classes = shp["cls.weight"][0]
        assert self.H == 256 and self.S == 128, "the native path is built for hidden 256 / seq 128 (BERT-mini)"
        self.heads = self.H // 64
        self.off: Dict[str, int] = {}
        o = 0
        big = [n for n, s in spec.shapes if n.endswith(".weight") and n.startswith("h")]
        for n in big:
            self.off[n] = o
            o = _al(o + math.prod(shp[n]))
        self.P_bf16 = self.small_off = o
        tail = ("pos.emb", "cls.weight", "cls.bias", "tok.emb")
        for n, s in spec.shapes:
            if n in big or n in tail:
                continue
            self.off[n] = o
            o = _al(o + math.prod(s))
        for n in tail:
            self.off[n] = o
            o = _al(o + math.prod(shp[n]))
        self.P = _al(o, 8)
        self.sgd_len = self.off["cls.weight"] - self.small_off        # norms, biases, pos.emb: one fused SGD launch
        self.P_small = self.sgd_len

    def _items(self):
        shp = dict(self.spec.shapes)
        for n, _ in self.spec.shapes:
            yield n, self.off[n], math.prod(shp[n]), shp[n]

    def pack(self, flat_spec: torch.Tensor) -> torch.Tensor:
        p = self.spec.unflatten(flat_spec.float().cpu())
        row = torch.zeros(self.P)
        for n, o, k, _ in self._items():
            row[o:o + k] = p[n].reshape(-1)
        return row

    def unpack(self, row: torch.Tensor) -> torch.Tensor:
        row = row.float().cpu()
        return self.spec.flatten({n: row[o:o + k].reshape(s) for n, o, k, s in self._items()})


class BertWorkspace(FlatWorkspace):
    def __init__(self, num_clients: int, shard: int, spec: ModelSpec, device="cuda", heap: Optional[SymmetricHeap] = None,
                 wave: Optional[int] = None):
        self.L = BertLayout(spec)
        super().__init__(num_clients, self.L.P, device=device, heap=heap)
        L, d, C = self.L, self.dev, num_clients
        self.shard = shard
        self.wb = torch.zeros(C, L.P_bf16, dtype=torch.bfloat16, device=d)
        self.n_samples = torch.full((C,), shard, dtype=torch.int32, device=d)
        self.ids = torch.zeros(C, shard * L.S, dtype=torch.int32, device=d)
        self.labels = torch.zeros(C, shard, dtype=torch.int32, device=d)
        T = B * L.S
        per_client = (L.layers * (3 * L.H + 2 * L.FF + 6 * L.H + B * L.heads * L.S * L.S // T) + 10 * L.H + L.FF + 3 * L.H) * T * 2
        if wave is None:
            free, _ = torch.cuda.mem_get_info(d)
            wave = max(1, min(C, int(free * 0.8) // max(per_client, 1)))
        self.wave = max(1, min(wave, C))
        self._alloc_wave(self.wave)

    def _alloc_wave(self, n: int) -> None:
        L, d = self.L, self.dev
        bf, f32 = torch.bfloat16, torch.float32
        T, H, FF, S = B * L.S, L.H, L.FF, L.S

        def buf(cols):
            return torch.empty(n, T, cols, dtype=bf, device=d)
        self.xhat0, self.h0, self.rstd0 = buf(H), buf(H), torch.empty(n, T, dtype=f32, device=d)
        self.lay: List[Dict[str, torch.Tensor]] = []
        for _ in range(L.layers):
            self.lay.append({"qkv": buf(3 * H), "probs": torch.empty(n * B * L.heads, S, S, dtype=bf, device=d), "att": buf(H),
                             "xhat1": buf(H), "h1": buf(H), "rstd1": torch.empty(n, T, dtype=f32, device=d),
                             "pre_ff": buf(FF), "a1": buf(FF), "xhat2": buf(H), "hout": buf(H),
                             "rstd2": torch.empty(n, T, dtype=f32, device=d)})
        self.pre = buf(H)
        self.dh, self.dp2, self.dp1, self.datt = buf(H), buf(H), buf(H), buf(H)
        self.dff, self.dqkv = buf(FF), buf(3 * H)
        self.ds = torch.empty(n * B * L.heads, S, S, dtype=bf, device=d)
        self.gsmall = torch.zeros(n, L.P_small, dtype=f32, device=d)

    # ---- data & model -------------------------------------------------------------------------------
    def load_data(self, ids: torch.Tensor, labels: torch.Tensor) -> None:
        """ids: integer token ids [C, shard, S]."""
        self.ids.copy_(ids.to(self.dev).reshape(self.C, -1).to(torch.int32))
        self.labels.copy_(labels.to(self.dev).to(torch.int32))

    def broadcast(self, participate: Optional[torch.Tensor] = None) -> None:
        native.bcast_scatter_generic(self.heap.buffers["global"].peer_ptrs, self.shard_len, self.P, self.global_local,
                                     self.master, self.wb, participate, self.L.P_bf16, self.L.P_bf16)
        self.kernel_launches += 1

    def local_step(self, step_idx: int, lr: float, mu: float = 0.0, active: Optional[torch.Tensor] = None) -> None:
        before = bgm.launches
        for w0 in range(0, self.C, self.wave):
            w1 = min(w0 + self.wave, self.C)
            self._step_wave(w0, w1, step_idx, lr, mu, None if active is None else active[w0:w1])
        self.kernel_launches += bgm.launches - before

    # ---- helpers -------------------------------------------------------------------------------------
    def _ln(self, n, act, m, pre, gname, xhat, y, rstd, ids=None, ids_off=0):
        L = self.L
        p = bgm.LnParams()
        p.C, p.T, p.active = n, B * L.S, bgm._ptr(act)
        p.pre = bgm._ptr(pre)
        if ids is not None:
            p.ids, p.ids_stride, p.ids_off, p.S = bgm._ptr(ids), self.shard * L.S, ids_off, L.S
            p.tok_off, p.pos_off = L.off["tok.emb"], L.off["pos.emb"]
        p.master, p.P = bgm._ptr(m), self.P
        p.gamma_off, p.beta_off = L.off[gname + ".gamma"], L.off[gname + ".beta"]
        p.xhat, p.y, p.rstd = bgm._ptr(xhat), bgm._ptr(y), bgm._ptr(rstd)
        bgm.ln_fwd(p)

    def _ln_bwd(self, n, act, m, dy, gname, xhat, rstd, dx, bias_name=None):
        L = self.L
        p = bgm.LnParams()
        p.C, p.T, p.active = n, B * L.S, bgm._ptr(act)
        p.master, p.P, p.gamma_off = bgm._ptr(m), self.P, L.off[gname + ".gamma"]
        p.xhat, p.rstd, p.dy, p.dx = bgm._ptr(xhat), bgm._ptr(rstd), bgm._ptr(dy), bgm._ptr(dx)
        p.gsmall, p.gs_stride = bgm._ptr(self.gsmall), L.P_small
        p.g_gamma_off, p.g_beta_off = L.off[gname + ".gamma"] - L.small_off, L.off[gname + ".beta"] - L.small_off
        p.g_bias_off = (L.off[bias_name] - L.small_off) if bias_name else -1
        bgm.ln_bwd(p)

    def _colsum(self, n, act, x, N, name):
        L = self.L
        p = bgm.ColsumParams()
        p.C, p.T, p.N, p.chunk, p.active = n, B * L.S, N, 256, bgm._ptr(act)
        p.x, p.gsmall, p.gs_stride, p.g_off = bgm._ptr(x), bgm._ptr(self.gsmall), L.P_small, L.off[name] - L.small_off
        bgm.colsum(p)

    def _step_wave(self, w0: int, w1: int, step_idx: int, lr: float, mu: float, act: Optional[torch.Tensor]) -> None:
        L, n = self.L, w1 - w0
        H, FF, S, T, heads = L.H, L.FF, L.S, B * L.S, L.heads
        m, wb, Pb, P = self.master[w0:w1], self.wb[w0:w1], L.P_bf16, self.P
        gw = self.global_local if mu > 0 else None
        off = (step_idx * B) % self.shard
        ids = self.ids[w0:w1]
        bgm.memset(self.gsmall)
        CB, adiv = n * B, B * heads

        def fwd(x, K, N, wname, out, bname, **kw):
            bgm.linear_fwd(n, T, K, N, x, wb, L.off[wname], Pb, out, m, L.off[bname], P, active=act, **kw)

        def dgrad(dy, N, K, wname, dx, **kw):
            bgm.linear_dgrad(n, T, N, K, dy, wb, L.off[wname], Pb, dx, active=act, **kw)

        def wgrad(dy, x, N, K, wname):
            bgm.linear_wgrad(n, T, N, K, dy, x, m, wb, L.off[wname], P, Pb, lr, mu, gw, act)

        # ------------------------------------------------------------------ forward
        self._ln(n, act, m, None, "emb.ln", self.xhat0, self.h0, self.rstd0, ids=ids, ids_off=off * S)
        h = self.h0
        for l in range(L.layers):
            t, nm = self.lay[l], f"h{l}"
            fwd(h, H, 3 * H, f"{nm}.qkv.weight", t["qkv"], f"{nm}.qkv.bias")
            bgm.attn_scores(CB, S, heads, t["qkv"], H, t["probs"], act, adiv)
            bgm.attn_pv(CB, S, heads, t["probs"], t["qkv"], H, t["att"], act, adiv)
            fwd(t["att"], H, H, f"{nm}.o.weight", self.pre, f"{nm}.o.bias", addend=h)
            self._ln(n, act, m, self.pre, f"{nm}.ln1", t["xhat1"], t["h1"], t["rstd1"])
            fwd(t["h1"], H, FF, f"{nm}.ff1.weight", t["a1"], f"{nm}.ff1.bias", act=bgm.ACT_GELU, aux_out=t["pre_ff"])
            fwd(t["a1"], FF, H, f"{nm}.ff2.weight", self.pre, f"{nm}.ff2.bias", addend=t["h1"])
            self._ln(n, act, m, self.pre, f"{nm}.ln2", t["xhat2"], t["hout"], t["rstd2"])
            h = t["hout"]
        # ------------------------------------------------------------------ head: loss, cls update, d h_last
        cp = bgm.ClsHeadParams()
        cp.C, cp.imgs, cp.S, cp.classes = n, B, S, L.classes
        cp.active, cp.y, cp.dy = bgm._ptr(act), bgm._ptr(h), bgm._ptr(self.dh)
        cp.labels, cp.label_stride, cp.label_off = bgm._ptr(self.labels[w0:w1]), self.shard, off
        cp.master, cp.P, cp.w_off, cp.b_off = bgm._ptr(m), P, L.off["cls.weight"], L.off["cls.bias"]
        cp.global_w = bgm._ptr(gw) if gw is not None else None
        cp.lr, cp.mu = lr, mu
        cp.loss_sum, cp.nonfinite = bgm._ptr(self.loss_sum[w0:w1]), bgm._ptr(self.nonfinite[w0:w1])
        cp.apply_update = 1
        bgm.cls_head(cp)
        # ------------------------------------------------------------------ backward
        for l in range(L.layers - 1, -1, -1):
            t, nm = self.lay[l], f"h{l}"
            hin = self.lay[l - 1]["hout"] if l > 0 else self.h0
            self._ln_bwd(n, act, m, self.dh, f"{nm}.ln2", t["xhat2"], t["rstd2"], self.dp2, bias_name=f"{nm}.ff2.bias")
            dgrad(self.dp2, H, FF, f"{nm}.ff2.weight", self.dff, gelu_pre=t["pre_ff"])
            wgrad(self.dp2, t["a1"], H, FF, f"{nm}.ff2.weight")
            self._colsum(n, act, self.dff, FF, f"{nm}.ff1.bias")
            dgrad(self.dff, FF, H, f"{nm}.ff1.weight", self.dh, addend=self.dp2)           # d h1 (residual + FF path)
            wgrad(self.dff, t["h1"], FF, H, f"{nm}.ff1.weight")
            self._ln_bwd(n, act, m, self.dh, f"{nm}.ln1", t["xhat1"], t["rstd1"], self.dp1, bias_name=f"{nm}.o.bias")
            dgrad(self.dp1, H, H, f"{nm}.o.weight", self.datt)
            wgrad(self.dp1, t["att"], H, H, f"{nm}.o.weight")
            bgm.attn_bwd(CB, S, heads, H, t["qkv"], t["probs"], self.datt, self.ds, self.dqkv, act, adiv)
            self._colsum(n, act, self.dqkv, 3 * H, f"{nm}.qkv.bias")
            dgrad(self.dqkv, 3 * H, H, f"{nm}.qkv.weight", self.dh, addend=self.dp1)       # d h_in (residual + attention)
            wgrad(self.dqkv, hin, 3 * H, H, f"{nm}.qkv.weight")
        self._ln_bwd(n, act, m, self.dh, "emb.ln", self.xhat0, self.rstd0, self.dp1)
        ep = bgm.EmbBwdParams()
        ep.C, ep.T, ep.S, ep.active, ep.dpre = n, T, S, bgm._ptr(act), bgm._ptr(self.dp1)
        ep.ids, ep.ids_stride, ep.ids_off = bgm._ptr(ids), self.shard * S, off * S
        ep.master, ep.P, ep.tok_off, ep.lr = bgm._ptr(m), P, L.off["tok.emb"], lr
        ep.gsmall, ep.gs_stride, ep.g_pos_off = bgm._ptr(self.gsmall), L.P_small, L.off["pos.emb"] - L.small_off
        bgm.emb_bwd(ep)
        bgm.slice_sgd(m, P, L.small_off, L.sgd_len, self.gsmall, L.P_small, gw, act, n, lr, mu)
        if mu > 0:       # dense proximal pull on the sparsely-updated token table
            bgm.slice_sgd(m, P, L.off["tok.emb"], L.vocab * H, None, 0, gw, act, n, lr, mu)
