"""ctypes binding + operation builders of the batched tcgen05 GEMM / implicit-conv family (``csrc/bgemm_tc.cuh``) and
the neural-net support kernels (``csrc/nn_kernels.cuh``).

Every builder fills a ``Bgemm`` block (two TMA tensor-map descriptions + the kernel's ``Params``) for one layer
operation over *all* clients of a rank and launches it on the current torch stream.  The blocks mirror the C structs
field by field; ``_check_mirror()`` compares sizes / offsets with the library at first use so a drift fails loudly.

Operation             GEMM view  (per client / per (client, sample, head))
--------------------  ---------------------------------------------------------------------------------------------
linear_fwd            Y[T, N]   = X[T, K] W[N, K]^T (+bias, ReLU/GELU)             A_K  x B_K    E_STORE
linear_dgrad          dX[T, K]  = dY[T, N] W[N, K]  (+addend, GELU')               A_K  x B_MN   E_STORE
linear_wgrad          W[N, K]  -= lr * dY[T, N]^T X[T, K]                            A_MN x B_MN   E_SGD
attn_scores / ...     softmax(Q K^T / sqrt(d)), P V, and the five backward products  (see ``engine/bert_native.py``)
conv_fwd              raw[pix, Cout] = window(X)[pix, tap*Cin] W[Cout, tap*Cin]^T  + GroupNorm statistics
conv_dgrad            dX[pix, Cin]   = window(dY)[pix, tap*Cout] W^T               (stride 2: four parity classes)
conv_wgrad            W[Cout, tap*Cin] -= lr * dY[pix, Cout]^T window(X)[pix, tap*Cin]
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import torch

from . import native

A_K, A_MN, A_CONV, A_HALO, A_HALOW, A_CONVX = 0, 1, 2, 3, 4, 5
B_K, B_MN, B_CONVX, B_WT = 0, 1, 2, 3
E_STORE, E_GNSTAT, E_SOFTMAX, E_DSOFT, E_SGD, E_SGD_T = 0, 1, 2, 3, 4, 5
ACT_NONE, ACT_RELU, ACT_GELU, ACT_GELU_BWD, ACT_RELU_BWD = 0, 1, 2, 3, 4

I9 = C.c_int * 9


class Tmap(C.Structure):
    _fields_ = [("base", C.c_void_p), ("rank", C.c_int), ("dims", C.c_ulonglong * 5), ("strides", C.c_ulonglong * 4),
                ("box", C.c_uint * 5), ("estr", C.c_uint * 5)]


class Params(C.Structure):
    _fields_ = [
        ("batches", C.c_int), ("M", C.c_int), ("N", C.c_int), ("m_tiles", C.c_int), ("n_tiles", C.c_int), ("k_chunks", C.c_int),
        ("active", C.c_void_p), ("act_div", C.c_int),
        ("a_bdiv", C.c_int), ("b_bdiv", C.c_int), ("pf_ahead", C.c_int), ("direct_store", C.c_int),
        ("tw_log", C.c_int), ("th_log", C.c_int), ("tiles_x", C.c_int), ("tiles_y", C.c_int), ("imgs", C.c_int),
        ("src_stride", C.c_int), ("src_cimgs", C.c_int), ("src_img0", C.c_int), ("cchunks", C.c_int), ("ntaps", C.c_int), ("ktap", C.c_int),
        ("tap_dy", I9), ("tap_dx", I9), ("tap_w", I9),
        ("kb_h", C.c_int), ("kb_n", C.c_int), ("kb_cpi", C.c_int), ("nblocks", C.c_int), ("cin", C.c_int),
        ("out", C.c_void_p), ("o_bs1", C.c_longlong), ("o_bs0", C.c_longlong), ("o_bdiv", C.c_int), ("ld", C.c_int),
        ("out_conv", C.c_int), ("os", C.c_int), ("py", C.c_int), ("px", C.c_int), ("OW", C.c_int), ("OH", C.c_int),
        ("bias", C.c_void_p), ("bias_bs", C.c_longlong), ("bias_div", C.c_int),
        ("addend", C.c_void_p),
        ("act", C.c_int), ("aux_in", C.c_void_p), ("aux_out", C.c_void_p),
        ("scale", C.c_float),
        ("gn_stats", C.c_void_p), ("gn_groups", C.c_int), ("gn_cpg", C.c_int),
        ("w_master", C.c_void_p), ("w_bs", C.c_longlong), ("w_ld", C.c_int), ("w_global", C.c_void_p),
        ("w_bf16", C.c_void_p), ("wb_bs", C.c_longlong),
        ("grad_out", C.c_void_p), ("g_bs", C.c_longlong), ("lr", C.c_float), ("mu", C.c_float),
    ]


class Bgemm(C.Structure):
    _fields_ = [("am", C.c_int), ("bm", C.c_int), ("n_tile", C.c_int), ("epi", C.c_int), ("a", Tmap), ("b", Tmap), ("p", Params)]


class GnApplyParams(C.Structure):
    _fields_ = [("C", C.c_int), ("imgs", C.c_int), ("HW", C.c_int), ("Ch", C.c_int), ("groups", C.c_int), ("active", C.c_void_p),
                ("raw", C.c_void_p), ("stats", C.c_void_p), ("gamma_off", C.c_int), ("beta_off", C.c_int),
                ("raw2", C.c_void_p), ("stats2", C.c_void_p), ("gamma2_off", C.c_int), ("beta2_off", C.c_int),
                ("res", C.c_void_p), ("master", C.c_void_p), ("P", C.c_longlong), ("relu", C.c_int), ("out", C.c_void_p)]


class GnBwdParams(C.Structure):
    _fields_ = [("C", C.c_int), ("imgs", C.c_int), ("HW", C.c_int), ("Ch", C.c_int), ("groups", C.c_int), ("active", C.c_void_p),
                ("dy", C.c_void_p), ("mask", C.c_void_p), ("raw", C.c_void_p), ("stats", C.c_void_p),
                ("master", C.c_void_p), ("P", C.c_longlong), ("gamma_off", C.c_int),
                ("gsmall", C.c_void_p), ("gs_stride", C.c_longlong), ("g_gamma_off", C.c_int), ("g_beta_off", C.c_int),
                ("sb", C.c_void_p), ("dx", C.c_void_p), ("dres", C.c_void_p)]


class PoolHeadParams(C.Structure):
    _fields_ = [("C", C.c_int), ("imgs", C.c_int), ("HW", C.c_int), ("Ch", C.c_int), ("classes", C.c_int), ("active", C.c_void_p),
                ("h", C.c_void_p), ("dh", C.c_void_p), ("labels", C.c_void_p), ("label_stride", C.c_int), ("label_off", C.c_int),
                ("master", C.c_void_p), ("P", C.c_longlong), ("w_off", C.c_int), ("b_off", C.c_int), ("global_w", C.c_void_p),
                ("lr", C.c_float), ("mu", C.c_float), ("loss_sum", C.c_void_p), ("nonfinite", C.c_void_p),
                ("logits_out", C.c_void_p), ("apply_update", C.c_int)]


class LnParams(C.Structure):
    _fields_ = [("C", C.c_int), ("T", C.c_int), ("active", C.c_void_p), ("pre", C.c_void_p),
                ("ids", C.c_void_p), ("ids_stride", C.c_int), ("ids_off", C.c_int), ("S", C.c_int),
                ("tok_off", C.c_int), ("pos_off", C.c_int),
                ("master", C.c_void_p), ("P", C.c_longlong), ("gamma_off", C.c_int), ("beta_off", C.c_int),
                ("xhat", C.c_void_p), ("y", C.c_void_p), ("rstd", C.c_void_p),
                ("dy", C.c_void_p), ("dx", C.c_void_p),
                ("gsmall", C.c_void_p), ("gs_stride", C.c_longlong), ("g_gamma_off", C.c_int), ("g_beta_off", C.c_int),
                ("g_bias_off", C.c_int)]


class EmbBwdParams(C.Structure):
    _fields_ = [("C", C.c_int), ("T", C.c_int), ("S", C.c_int), ("active", C.c_void_p), ("dpre", C.c_void_p),
                ("ids", C.c_void_p), ("ids_stride", C.c_int), ("ids_off", C.c_int),
                ("master", C.c_void_p), ("P", C.c_longlong), ("tok_off", C.c_int), ("lr", C.c_float),
                ("gsmall", C.c_void_p), ("gs_stride", C.c_longlong), ("g_pos_off", C.c_int)]


class ClsHeadParams(C.Structure):
    _fields_ = [("C", C.c_int), ("imgs", C.c_int), ("S", C.c_int), ("classes", C.c_int), ("active", C.c_void_p),
                ("y", C.c_void_p), ("dy", C.c_void_p), ("labels", C.c_void_p), ("label_stride", C.c_int), ("label_off", C.c_int),
                ("master", C.c_void_p), ("P", C.c_longlong), ("w_off", C.c_int), ("b_off", C.c_int), ("global_w", C.c_void_p),
                ("lr", C.c_float), ("mu", C.c_float), ("loss_sum", C.c_void_p), ("nonfinite", C.c_void_p),
                ("logits_out", C.c_void_p), ("apply_update", C.c_int)]


class ColsumParams(C.Structure):
    _fields_ = [("C", C.c_int), ("T", C.c_int), ("N", C.c_int), ("chunk", C.c_int), ("active", C.c_void_p), ("x", C.c_void_p),
                ("gsmall", C.c_void_p), ("gs_stride", C.c_longlong), ("g_off", C.c_int)]


_checked = False
launches = 0          # kernels launched through this module (the engine adds this to its launch counter)


def _lib():
    global _checked
    l = native.require()
    if not _checked:
        l.olsb_nn_last_error.restype = C.c_char_p
        out = (C.c_int * 16)()
        n = l.olsb_nn_sizes(out, 16)
        want = [C.sizeof(Tmap), C.sizeof(Params), C.sizeof(Bgemm), Bgemm.p.offset, Params.out.offset, Params.w_master.offset,
                Params.lr.offset, C.sizeof(GnApplyParams), C.sizeof(GnBwdParams), C.sizeof(PoolHeadParams), C.sizeof(LnParams),
                C.sizeof(EmbBwdParams), C.sizeof(ClsHeadParams), C.sizeof(ColsumParams)]
        got = [int(out[i]) for i in range(n)]
        if got != want:
            raise native.NativeError(f"ctypes mirror of the bgemm parameter blocks is out of sync: C {got} vs python {want}")
        _checked = True
    return l


def _ck(rc: int) -> None:
    if rc != 0:
        raise native.NativeError(f"native nn call failed rc={rc}: {_lib().olsb_nn_last_error().decode(errors='replace')}")


def _ptr(t: Optional[torch.Tensor], elem_off: int = 0):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr() + t.element_size() * int(elem_off))


def tmap(t: torch.Tensor, elem_off: int, dims: Sequence[int], strides: Sequence[int], box: Sequence[int],
         estr: Optional[Sequence[int]] = None) -> Tmap:
    """bf16 tensor map: ``dims`` innermost first, ``strides`` (elements) of dims[1:], ``box`` per dim."""
    assert t.dtype == torch.bfloat16 and t.is_cuda
    m = Tmap()
    base = t.data_ptr() + 2 * int(elem_off)
    assert base % 16 == 0 and all((int(s) * 2) % 16 == 0 for s in strides), "TMA needs 16-byte aligned bases and strides"
    m.base, m.rank = C.c_void_p(base), len(dims)
    for i, d in enumerate(dims):
        m.dims[i] = int(d); m.box[i] = int(box[i]); m.estr[i] = int(estr[i]) if estr else 1
    for i, s in enumerate(strides):
        m.strides[i] = int(s)
    return m


PF_AHEAD = int(__import__("os").environ.get("OLSB_PF_AHEAD", "0"))       # L2 prefetch distance (pipeline chunks) of the TMA producer


DIRECT_STORE = __import__("os").environ.get("OLSB_DIRECT_STORE", "0")   # "0" staged through smem (measured best everywhere), "1" direct


def launch(g: Bgemm) -> None:
    global launches
    g.p.pf_ahead = PF_AHEAD
    g.p.direct_store = int(DIRECT_STORE) if DIRECT_STORE in ("0", "1") else (1 if g.p.out_conv else 0)
    _ck(_lib().olsb_bgemm(C.byref(g), native._stream()))
    launches += 1


def _count(n: int = 1) -> None:
    global launches
    launches += n


# ======================================================================================================
# fully-connected layers (tokens on the M axis)
# ======================================================================================================
def linear_fwd(C_: int, T: int, K: int, N: int, x: torch.Tensor, wb: torch.Tensor, w_off: int, wb_stride: int,
               out: torch.Tensor, bias_master: Optional[torch.Tensor] = None, bias_off: int = 0, P: int = 0,
               act: int = ACT_NONE, aux_out: Optional[torch.Tensor] = None, addend: Optional[torch.Tensor] = None,
               active: Optional[torch.Tensor] = None, x_off: int = 0, x_cstride: Optional[int] = None,
               x_ld: Optional[int] = None) -> None:
    """out[c][t][n] = act(x[c][t][:] . W_c[n][:] + bias_c[n] (+ addend)); W_c = wb[c*wb_stride + w_off] as [N][K] bf16."""
    g = Bgemm()
    g.am, g.bm, g.n_tile, g.epi = A_K, B_K, 256, E_STORE
    x_ld = x_ld or K
    g.a = tmap(x, x_off, (K, T, C_), (x_ld, x_cstride or T * x_ld), (64, 128, 1))
    g.b = tmap(wb, w_off, (K, N, C_), (K, wb_stride), (64, 256, 1))
    p = g.p
    p.batches, p.M, p.N = C_, T, N
    p.m_tiles, p.n_tiles, p.k_chunks = (T + 127) // 128, (N + 255) // 256, (K + 63) // 64
    p.active, p.act_div = _ptr(active), 1
    p.out, p.o_bs1, p.o_bs0, p.o_bdiv, p.ld = _ptr(out), T * N, 0, 1, N
    if bias_master is not None:
        p.bias, p.bias_bs, p.bias_div = _ptr(bias_master, bias_off), P, 1
    p.addend, p.act, p.aux_out, p.scale = _ptr(addend), act, _ptr(aux_out), 1.0
    launch(g)


def linear_dgrad(C_: int, T: int, N: int, K: int, dy: torch.Tensor, wb: torch.Tensor, w_off: int, wb_stride: int,
                 dx: torch.Tensor, addend: Optional[torch.Tensor] = None, gelu_pre: Optional[torch.Tensor] = None,
                 active: Optional[torch.Tensor] = None) -> None:
    """dx[c][t][k] = sum_n dy[c][t][n] W_c[n][k]  (+ addend)  (* gelu'(gelu_pre))."""
    g = Bgemm()
    g.am, g.bm, g.n_tile, g.epi = A_K, B_MN, 256, E_STORE
    g.a = tmap(dy, 0, (N, T, C_), (N, T * N), (64, 128, 1))
    g.b = tmap(wb, w_off, (K, N, C_), (K, wb_stride), (64, 64, 1))
    p = g.p
    p.batches, p.M, p.N = C_, T, K
    p.m_tiles, p.n_tiles, p.k_chunks = (T + 127) // 128, (K + 255) // 256, (N + 63) // 64
    p.active, p.act_div = _ptr(active), 1
    p.out, p.o_bs1, p.o_bs0, p.o_bdiv, p.ld = _ptr(dx), T * K, 0, 1, K
    p.addend, p.scale = _ptr(addend), 1.0
    if gelu_pre is not None:
        p.act, p.aux_in = ACT_GELU_BWD, _ptr(gelu_pre)
    launch(g)


def linear_wgrad(C_: int, T: int, N: int, K: int, dy: torch.Tensor, x: torch.Tensor, master: torch.Tensor, wb: torch.Tensor,
                 w_off: int, P: int, wb_stride: int, lr: float, mu: float = 0.0, global_w: Optional[torch.Tensor] = None,
                 active: Optional[torch.Tensor] = None, grad_out: Optional[torch.Tensor] = None, x_off: int = 0,
                 x_cstride: Optional[int] = None, x_ld: Optional[int] = None) -> None:
    """W_c[n][k] -= lr * (sum_t dy[c][t][n] x[c][t][k] + mu (W_c - W_global)); fp32 master + bf16 copy."""
    g = Bgemm()
    g.am, g.bm, g.n_tile, g.epi = A_MN, B_MN, 256, E_SGD
    x_ld = x_ld or K
    g.a = tmap(dy, 0, (N, T, C_), (N, T * N), (64, 64, 1))
    g.b = tmap(x, x_off, (K, T, C_), (x_ld, x_cstride or T * x_ld), (64, 64, 1))
    p = g.p
    p.batches, p.M, p.N = C_, N, K
    p.m_tiles, p.n_tiles, p.k_chunks = (N + 127) // 128, (K + 255) // 256, (T + 63) // 64
    p.active, p.act_div = _ptr(active), 1
    p.w_master, p.w_bs, p.w_ld = _ptr(master, w_off), P, K
    p.w_bf16, p.wb_bs = _ptr(wb, w_off), wb_stride
    p.w_global = _ptr(global_w, w_off) if (global_w is not None and mu > 0) else None
    p.lr, p.mu = lr, mu
    if grad_out is not None:
        p.grad_out, p.g_bs = _ptr(grad_out), N * K
    launch(g)


# ======================================================================================================
# attention (seq 128, head dim 64): batch index = (client*imgs + sample) * heads + head
# ======================================================================================================
def _head_map(t: torch.Tensor, col_off: int, ld: int, S: int, heads: int, CB: int, box_rows: int) -> Tmap:
    """[CB][S][ld] tensor, head h occupies columns col_off + 64 h .. : 4-D map (d, s, head, cb)."""
    return tmap(t, col_off, (64, S, heads, CB), (ld, 64, S * ld), (64, box_rows, 1, 1))


def attn_scores(CB: int, S: int, heads: int, qkv: torch.Tensor, hidden: int, probs: torch.Tensor, active, act_div: int) -> None:
    """probs[b][s][s'] = softmax_s'(Q K^T / sqrt(64))."""
    assert S == 128
    g = Bgemm()
    g.am, g.bm, g.n_tile, g.epi = A_K, B_K, 128, E_SOFTMAX
    ld = 3 * hidden
    g.a = _head_map(qkv, 0, ld, S, heads, CB, 128)
    g.b = _head_map(qkv, hidden, ld, S, heads, CB, 128)
    p = g.p
    p.batches, p.M, p.N, p.m_tiles, p.n_tiles, p.k_chunks = CB * heads, S, S, 1, 1, 1
    p.a_bdiv = p.b_bdiv = heads
    p.active, p.act_div = _ptr(active), act_div
    p.out, p.o_bs1, p.o_bs0, p.o_bdiv, p.ld = _ptr(probs), S * S, 0, 1, S
    p.scale = 1.0 / math.sqrt(64.0)
    launch(g)


def attn_pv(CB, S, heads, probs, qkv, hidden, att, active, act_div) -> None:
    """att[cb][s][h*64+d] = sum_s' probs[b][s][s'] V[s'][d]."""
    g = Bgemm()
    g.am, g.bm, g.n_tile, g.epi = A_K, B_MN, 64, E_STORE
    g.a = tmap(probs, 0, (S, S, CB * heads), (S, S * S), (64, 128, 1))
    g.b = _head_map(qkv, 2 * hidden, 3 * hidden, S, heads, CB, 64)
    p = g.p
    p.batches, p.M, p.N, p.m_tiles, p.n_tiles, p.k_chunks = CB * heads, S, 64, 1, 1, S // 64
    p.b_bdiv = heads
    p.active, p.act_div = _ptr(active), act_div
    p.out, p.o_bs1, p.o_bs0, p.o_bdiv, p.ld = _ptr(att), S * hidden, 64, heads, hidden
    p.scale = 1.0
    launch(g)


def attn_bwd(CB, S, heads, hidden, qkv, probs, datt, ds, dqkv, active, act_div) -> None:
    """Five products: dV = P^T dO, dS = softmax'(dO V^T), dQ = dS K, dK = dS^T Q  (dqkv written in qkv layout)."""
    ld3 = 3 * hidden
    nb = CB * heads

    def common(g, N, kch):
        p = g.p
        p.batches, p.M, p.N, p.m_tiles, p.n_tiles, p.k_chunks = nb, S, N, 1, 1, kch
        p.active, p.act_div, p.scale = _ptr(active), act_div, 1.0
        return p
    # dV[s'][d] = sum_s P[s][s'] dO[s][d]
    g = Bgemm(); g.am, g.bm, g.n_tile, g.epi = A_MN, B_MN, 64, E_STORE
    g.a = tmap(probs, 0, (S, S, nb), (S, S * S), (64, 64, 1))
    g.b = _head_map(datt, 0, hidden, S, heads, CB, 64)
    p = common(g, 64, S // 64); p.b_bdiv = heads
    p.out, p.o_bs1, p.o_bs0, p.o_bdiv, p.ld = _ptr(dqkv, 2 * hidden), S * ld3, 64, heads, ld3
    launch(g)
    # dS = P o (dP - rowsum(dP o P)) / sqrt(d),  dP = dO V^T
    g = Bgemm(); g.am, g.bm, g.n_tile, g.epi = A_K, B_K, 128, E_DSOFT
    g.a = _head_map(datt, 0, hidden, S, heads, CB, 128)
    g.b = _head_map(qkv, 2 * hidden, ld3, S, heads, CB, 128)
    p = common(g, S, 1); p.a_bdiv = p.b_bdiv = heads
    p.out, p.o_bs1, p.o_bs0, p.o_bdiv, p.ld = _ptr(ds), S * S, 0, 1, S
    p.aux_in, p.scale = _ptr(probs), 1.0 / math.sqrt(64.0)
    launch(g)
    # dQ[s][d] = sum_s' dS[s][s'] K[s'][d]
    g = Bgemm(); g.am, g.bm, g.n_tile, g.epi = A_K, B_MN, 64, E_STORE
    g.a = tmap(ds, 0, (S, S, nb), (S, S * S), (64, 128, 1))
    g.b = _head_map(qkv, hidden, ld3, S, heads, CB, 64)
    p = common(g, 64, S // 64); p.b_bdiv = heads
    p.out, p.o_bs1, p.o_bs0, p.o_bdiv, p.ld = _ptr(dqkv, 0), S * ld3, 64, heads, ld3
    launch(g)
    # dK[s'][d] = sum_s dS[s][s'] Q[s][d]
    g = Bgemm(); g.am, g.bm, g.n_tile, g.epi = A_MN, B_MN, 64, E_STORE
    g.a = tmap(ds, 0, (S, S, nb), (S, S * S), (64, 64, 1))
    g.b = _head_map(qkv, 0, ld3, S, heads, CB, 64)
    p = common(g, 64, S // 64); p.b_bdiv = heads
    p.out, p.o_bs1, p.o_bs0, p.o_bdiv, p.ld = _ptr(dqkv, hidden), S * ld3, 64, heads, ld3
    launch(g)


# ======================================================================================================
# convolutions (NHWC bf16, `imgs` samples per client, weights [Cout][taps][Cin] bf16 in the operand copy)
# ======================================================================================================
def _tile_geom(OW: int, OH: int) -> Tuple[int, int, int]:
    """(log2 tile width, log2 tile height, images per 128-pixel tile) on an OW x OH grid."""
    assert OW == OH and OW in (4, 8, 16, 32), "square power-of-two feature maps"
    if OW >= 32:
        return 5, 2, 1
    if OW == 16:
        return 4, 3, 1
    if OW == 8:
        return 3, 3, 2
    return 2, 2, 8


def _n_tile(n: int) -> int:
    return 256 if n >= 256 else (128 if n >= 128 else 64)


HALO = __import__("os").environ.get("OLSB_HALO", "1") != "0"      # halo-box form of the 3x3 / 64 -> 64 / 32x32 convolutions


HALO_WGRAD = __import__("os").environ.get("OLSB_HALO_WGRAD", "1") != "0"
WGRAD_T = __import__("os").environ.get("OLSB_WGRAD_T", "1") != "0"      # weight gradients with rows = (tap, cin), columns = cout


def _halo_ok(H: int, Cin: int, Cout: int, ksize: int, stride: int) -> bool:
    return HALO and ksize == 3 and stride == 1 and Cin == 64 and Cout == 64 and H % 16 == 0


def conv_fwd(C_: int, imgs: int, H: int, Cin: int, Cout: int, ksize: int, stride: int, x: torch.Tensor, wb: torch.Tensor,
             w_off: int, wb_stride: int, raw: torch.Tensor, stats: torch.Tensor, groups: int, active=None,
             x_cimgs: int = 0, x_img0: int = 0) -> None:
    """raw[c,b,oy,ox,:] = conv(x)[...], plus GroupNorm sums in ``stats`` ([C*imgs][groups][2], pre-zeroed)."""
    OH = H // stride
    pad = ksize // 2
    twl, thl, tn = _tile_geom(OH, OH)
    TW, TH = 1 << twl, 1 << thl
    nt = _n_tile(Cout)
    g = Bgemm()
    halo = _halo_ok(H, Cin, Cout, ksize, stride)
    g.am, g.bm, g.n_tile, g.epi = A_HALO if halo else A_CONV, B_K, nt, E_GNSTAT
    nimg = x.numel() // (H * H * Cin)
    if halo:            # 16 x 8-pixel tiles of one image, one (16+2) x (8+2) box each
        twl, thl, tn, TW, TH = 3, 4, 1, 8, 16
        g.a = tmap(x, 0, (Cin, H, H, nimg), (Cin, H * Cin, H * H * Cin), (64, TW + 2, TH + 2, 1))
    else:
        g.a = tmap(x, 0, (Cin, H, H, nimg), (Cin, H * Cin, H * H * Cin), (64, TW * stride, TH * stride, tn), (1, stride, stride, 1))
    ktot = ksize * ksize * Cin
    g.b = tmap(wb, w_off, (ktot, Cout, C_), (ktot, wb_stride), (64, nt, 1))
    p = g.p
    p.batches, p.M, p.N = C_, imgs * OH * OH, Cout
    p.m_tiles, p.n_tiles = p.M // 128, Cout // nt
    p.cchunks = (Cin + 63) // 64
    p.ntaps, p.ktap = ksize * ksize, Cin
    p.k_chunks = p.ntaps * p.cchunks
    for kh in range(ksize):
        for kw in range(ksize):
            t = kh * ksize + kw
            p.tap_dy[t], p.tap_dx[t], p.tap_w[t] = kh - pad, kw - pad, t
    p.tw_log, p.th_log, p.tiles_x, p.tiles_y, p.imgs, p.src_stride = twl, thl, OH // TW, OH // TH, imgs, stride
    p.src_cimgs, p.src_img0 = x_cimgs or imgs, x_img0
    p.active, p.act_div = _ptr(active), 1
    p.out, p.ld, p.out_conv, p.os, p.py, p.px, p.OW, p.OH = _ptr(raw), Cout, 1, 1, 0, 0, OH, OH
    p.o_bdiv, p.scale = 1, 1.0
    p.gn_stats, p.gn_groups, p.gn_cpg = _ptr(stats), groups, Cout // groups
    launch(g)


def conv_dgrad(C_: int, imgs: int, H: int, Cin: int, Cout: int, ksize: int, stride: int, dy: torch.Tensor, wb: torch.Tensor,
               w_off: int, wb_stride: int, dx: torch.Tensor, addend: Optional[torch.Tensor] = None, active=None,
               relu_mask: Optional[torch.Tensor] = None, mask_even: bool = True) -> None:
    """dx[c,b,y,x,:Cin] = sum over taps of dy (.) W^T.  ``H`` is the INPUT (dx) resolution; stride 2 runs one launch per
    output-pixel parity class (each class is a stride-1 correlation of dy with a subset of the taps).
    ``relu_mask``: the forward activation this gradient belongs to (same shape as dx); the epilogue zeroes the gradient
    where it is not positive (ReLU backward), after the addend.  ``mask_even=False`` leaves the (even, even) pixel class
    of a stride-2 layer ungated — a following 1x1 stride-2 shortcut gradient adds to exactly those pixels and gates them."""
    OH = H // stride
    pad = ksize // 2
    twl, thl, tn = _tile_geom(OH, OH)
    TW, TH = 1 << twl, 1 << thl
    nt = _n_tile(Cin)
    ktaps = ksize * ksize
    classes = [(0, 0)] if stride == 1 else ([(0, 0)] if ksize == 1 else [(0, 0), (0, 1), (1, 0), (1, 1)])
    for (py, px) in classes:
        g = Bgemm()
        halo = _halo_ok(H, Cin, Cout, ksize, stride)
        g.am, g.bm, g.n_tile, g.epi = A_HALO if halo else A_CONV, B_WT, nt, E_STORE
        if halo:
            twl, thl, tn, TW, TH = 3, 4, 1, 8, 16
            g.a = tmap(dy, 0, (Cout, OH, OH, C_ * imgs), (Cout, OH * Cout, OH * OH * Cout), (64, TW + 2, TH + 2, 1))
        else:
            g.a = tmap(dy, 0, (Cout, OH, OH, C_ * imgs), (Cout, OH * Cout, OH * OH * Cout), (64, TW, TH, tn))
        g.b = tmap(wb, w_off, (Cin, ktaps, Cout, C_), (Cin, ktaps * Cin, wb_stride), (64, 1, 64, 1))
        p = g.p
        taps = []
        for kh in range(ksize):
            for kw in range(ksize):
                if stride == 1:
                    taps.append((pad - kh, pad - kw, kh * ksize + kw))
                else:
                    ny, nx = py + pad - kh, px + pad - kw
                    if ny % 2 == 0 and nx % 2 == 0:
                        taps.append((ny // 2, nx // 2, kh * ksize + kw))
        p.batches, p.M, p.N = C_, imgs * OH * OH, Cin
        p.m_tiles, p.n_tiles = p.M // 128, (Cin + nt - 1) // nt
        p.cchunks, p.ntaps, p.ktap = Cout // 64, len(taps), Cout
        p.k_chunks = p.ntaps * p.cchunks
        for t, (dyy, dxx, tw_) in enumerate(taps):
            p.tap_dy[t], p.tap_dx[t], p.tap_w[t] = dyy, dxx, tw_
        p.tw_log, p.th_log, p.tiles_x, p.tiles_y, p.imgs, p.src_stride = twl, thl, OH // TW, OH // TH, imgs, 1
        p.active, p.act_div = _ptr(active), 1
        p.out, p.ld, p.out_conv, p.os, p.py, p.px, p.OW, p.OH = _ptr(dx), Cin, 1, stride, py, px, H, H
        p.o_bdiv, p.scale = 1, 1.0
        p.addend = _ptr(addend)
        if relu_mask is not None and (mask_even or (py, px) != (0, 0)):
            p.act, p.aux_in = ACT_RELU_BWD, _ptr(relu_mask)
        launch(g)


def conv_wgrad(C_: int, imgs: int, H: int, Cin: int, Cout: int, ksize: int, stride: int, dy: torch.Tensor, x: torch.Tensor,
               master: torch.Tensor, wb: torch.Tensor, w_off: int, P: int, wb_stride: int, lr: float, mu: float = 0.0,
               global_w=None, active=None, grad_out=None, x_cimgs: int = 0, x_img0: int = 0) -> None:
    """W_c[cout][tap][cin] -= lr * (sum_pix dy[pix][cout] window(x)[pix][tap][cin] + mu (W - Wg))."""
    OH = H // stride
    pad = ksize // 2
    ktaps = ksize * ksize
    ktot = ktaps * Cin
    if _halo_ok(H, Cin, Cout, ksize, stride) and HALO_WGRAD:
        # halo form: work item = (client, set of tap pairs), accumulating over every 16 x 8-pixel tile of the client
        g = Bgemm()
        g.am, g.bm, g.n_tile, g.epi = A_HALOW, B_MN, 192, E_SGD_T
        nimg = x.numel() // (H * H * Cin)
        g.a = tmap(x, 0, (Cin, H, H, nimg), (Cin, H * Cin, H * H * Cin), (64, 10, 18, 1))
        g.b = tmap(dy, 0, (Cout, OH, OH, C_ * imgs), (Cout, OH * Cout, OH * OH * Cout), (64, 8, 16, 1))
        p = g.p
        p.batches, p.M, p.N, p.m_tiles, p.n_tiles = C_, ktot, Cout, 1, 2
        p.tw_log, p.th_log, p.tiles_x, p.tiles_y, p.imgs, p.src_stride = 3, 4, H // 8, H // 16, imgs, 1
        p.src_cimgs, p.src_img0 = x_cimgs or imgs, x_img0
        p.active, p.act_div = _ptr(active), 1
        p.w_master, p.w_bs, p.w_ld = _ptr(master, w_off), P, ktot
        p.w_bf16, p.wb_bs = _ptr(wb, w_off), wb_stride
        p.w_global = _ptr(global_w, w_off) if (global_w is not None and mu > 0) else None
        p.lr, p.mu, p.o_bdiv = lr, mu, 1
        if grad_out is not None:
            p.grad_out, p.g_bs = _ptr(grad_out), Cout * ktot
        launch(g)
        return
    kb_h, kb_n = (64 // OH, 1) if OH * OH >= 64 else (OH, 64 // (OH * OH))
    cpi = max(OH * OH // 64, 1)
    g = Bgemm()
    # measured (64 clients): 4x4x512 393 vs 600 us, 8x8x256 160 vs 212 us; with 128 output channels the 128-column tile
    # needs more operand bytes per MAC than the 192-column form below (206 vs 174 us) and stays on it
    transposed = WGRAD_T and Cout % 256 == 0 and Cin % 64 == 0
    npix = imgs * OH * OH
    nimg = x.numel() // (H * H * Cin)
    t_dy = tmap(dy, 0, (Cout, npix, C_), (Cout, npix * Cout), (64, 64, 1))
    t_x = tmap(x, 0, (Cin, H, H, nimg), (Cin, H * Cin, H * H * Cin), (64, OH * stride, kb_h * stride, kb_n), (1, stride, stride, 1))
    p = g.p
    p.cchunks = (Cin + 63) // 64
    p.nblocks, p.cin, p.ntaps = ktaps * p.cchunks, Cin, ktaps
    if transposed:
        # D[(tap, cin)][cout]: the update walks W[cout][tap][cin] with the lanes of a warp on contiguous addresses
        nt = 256
        g.am, g.bm, g.n_tile, g.epi = A_CONVX, B_MN, nt, E_SGD_T
        g.a, g.b = t_x, t_dy
        p.batches, p.M, p.N = C_, p.nblocks * 64, Cout
        p.m_tiles, p.n_tiles, p.k_chunks = (p.nblocks + 1) // 2, Cout // nt, npix // 64
    else:
        g.am, g.bm, g.n_tile, g.epi = A_MN, B_CONVX, 192, E_SGD
        g.a, g.b = t_dy, t_x
        p.batches, p.M, p.N = C_, Cout, ktot
        p.m_tiles, p.n_tiles, p.k_chunks = (Cout + 127) // 128, (p.nblocks + 2) // 3, npix // 64
    for kh in range(ksize):
        for kw in range(ksize):
            t = kh * ksize + kw
            p.tap_dy[t], p.tap_dx[t], p.tap_w[t] = kh - pad, kw - pad, t
    p.kb_h, p.kb_n, p.kb_cpi, p.imgs, p.src_stride = kb_h, kb_n, cpi, imgs, stride
    p.src_cimgs, p.src_img0 = x_cimgs or imgs, x_img0
    p.active, p.act_div = _ptr(active), 1
    p.w_master, p.w_bs, p.w_ld = _ptr(master, w_off), P, ktot
    p.w_bf16, p.wb_bs = _ptr(wb, w_off), wb_stride
    p.w_global = _ptr(global_w, w_off) if (global_w is not None and mu > 0) else None
    p.lr, p.mu, p.o_bdiv = lr, mu, 1
    if grad_out is not None:
        p.grad_out, p.g_bs = _ptr(grad_out), Cout * ktot
    launch(g)


# ======================================================================================================
# support kernels
# ======================================================================================================
def stem_im2col(src_u8: torch.Tensor, dst: torch.Tensor) -> None:
    n = src_u8.shape[0]
    assert src_u8.dtype == torch.uint8 and tuple(dst.shape) == (n, 32, 32, 32) and dst.dtype == torch.bfloat16
    _ck(_lib().olsb_stem_im2col(_ptr(src_u8), _ptr(dst), C.c_longlong(n), native._stream()))
    _count()


def gn_apply(p: GnApplyParams) -> None:
    _ck(_lib().olsb_gn_apply(C.byref(p), native._stream())); _count()


# one-kernel GroupNorm backward (gradient parked in smem).  Measured slower than the two-pass form on the ResNet-18 step
# (134 vs 123 us per layer at 64 clients: the 128 KB images of stage 1 leave one 512-thread CTA per SM, too few bytes in
# flight), so it is opt-in
GN_FUSED = __import__("os").environ.get("OLSB_GN_FUSED", "0") == "1"
_gn_mode_set = False


def gn_bwd(p: GnBwdParams) -> None:
    global _gn_mode_set
    if not _gn_mode_set:
        _lib().olsb_gn_set_fused(C.c_int(1 if GN_FUSED else 0))
        _gn_mode_set = True
    _ck(_lib().olsb_gn_bwd(C.byref(p), native._stream())); _count(1 if GN_FUSED else 2)


def pool_head(p: PoolHeadParams) -> None:
    _ck(_lib().olsb_pool_head(C.byref(p), native._stream())); _count()


def ln_fwd(p: LnParams) -> None:
    _ck(_lib().olsb_ln_fwd(C.byref(p), native._stream())); _count()


def ln_bwd(p: LnParams) -> None:
    _ck(_lib().olsb_ln_bwd(C.byref(p), native._stream())); _count()


def emb_bwd(p: EmbBwdParams) -> None:
    _ck(_lib().olsb_emb_bwd(C.byref(p), native._stream())); _count()


def cls_head(p: ClsHeadParams) -> None:
    _ck(_lib().olsb_cls_head(C.byref(p), native._stream())); _count()


def colsum(p: ColsumParams) -> None:
    _ck(_lib().olsb_colsum(C.byref(p), native._stream())); _count()


def slice_sgd(master: torch.Tensor, P: int, off: int, n: int, g: Optional[torch.Tensor], gs_stride: int,
              global_w: Optional[torch.Tensor], active, C_: int, lr: float, mu: float) -> None:
    _ck(_lib().olsb_slice_sgd(_ptr(master), C.c_longlong(P), C.c_int(off), C.c_int(n), _ptr(g), C.c_longlong(gs_stride),
                              _ptr(global_w) if (global_w is not None and mu > 0) else None, _ptr(active), C.c_int(C_),
                              C.c_float(lr), C.c_float(mu), native._stream()))
    _count()


def memset(t: torch.Tensor, value: int = 0) -> None:
    _ck(_lib().olsb_memset_async(_ptr(t), C.c_int(value), C.c_ulonglong(t.numel() * t.element_size()), native._stream()))
