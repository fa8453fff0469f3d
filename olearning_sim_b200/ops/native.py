"""ctypes binding of the sm_100a kernel library (``csrc/api.cu``).

Every function takes torch CUDA tensors, passes raw device pointers and the *current* torch CUDA stream, and
raises ``NativeError`` on a non-zero return code.  There is deliberately **no** eager fallback here: on a GPU box
a missing library is an error (``require()``), so a silent PyTorch path can never masquerade as the product.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import build as _build
from .layout import CNN2

_lib: Optional[C.CDLL] = None


class NativeError(RuntimeError):
    pass


def available() -> bool:
    return torch.cuda.is_available() and (_build.LIB.exists() or _build._nvcc() is not None)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = _build.build()
        _lib = C.CDLL(str(path))
        _lib.olsb_last_error.restype = C.c_char_p
        _lib.olsb_philox_host.restype = None
    return _lib


def require() -> C.CDLL:
    if not torch.cuda.is_available():
        raise NativeError("CUDA device required for the native sm_100a path")
    l = lib()
    _ck(l.olsb_init())
    return l


def _ck(rc: int) -> None:
    if rc != 0:
        msg = lib().olsb_last_error().decode(errors="replace")
        raise NativeError(f"native call failed rc={rc}: {msg}")


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "native ops need contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr_array(ptrs: Sequence[int]):
    arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])
    return arr


def num_sms() -> int:
    require()
    return lib().olsb_num_sms()


# --------------------------------------------------------------------------------------------------
def ingest_u8(src_u8: torch.Tensor, dst_x5: torch.Tensor) -> None:
    """uint8 [N,32,32,3] -> bf16 [N,32,32,16] (normalised, kw-unrolled)."""
    n = src_u8.shape[0]
    assert src_u8.dtype == torch.uint8 and dst_x5.dtype == torch.bfloat16
    assert tuple(dst_x5.shape) == (n, 32, 32, 16)
    _ck(require().olsb_ingest_u8(_p(src_u8), _p(dst_x5), C.c_longlong(n), _stream()))


def conv1_fwd(x5, x_n_stride, x_n_off, w1_fwd, master, num_clients, active, p1, code1, nsplit=0):
    _ck(require().olsb_conv1_fwd(_p(x5), C.c_longlong(x5.shape[0]), C.c_int(x_n_stride), C.c_int(x_n_off), _p(w1_fwd),
                                 _p(master), C.c_longlong(master.shape[1]), C.c_int(num_clients), _p(active), _p(p1),
                                 _p(code1), C.c_int(nsplit), _stream()))


def conv2_fwd(p1, w2_fwd, master, num_clients, active, p2, code2, nsplit=0):
    _ck(require().olsb_conv2_fwd(_p(p1), _p(w2_fwd), _p(master), C.c_longlong(master.shape[1]), C.c_int(num_clients),
                                 _p(active), _p(p2), _p(code2), C.c_int(nsplit), _stream()))


def conv2_dgrad(dy2, w2_dgrad, master, global_w, num_clients, active, p1, dp1, lr, mu, dbias_out=None, nsplit=0):
    _ck(require().olsb_conv2_dgrad(_p(dy2), _p(w2_dgrad), _p(master), _p(global_w), C.c_longlong(master.shape[1]),
                                   C.c_int(num_clients), _p(active), _p(p1), _p(dp1), C.c_float(lr),
                                   C.c_float(mu), _p(dbias_out), C.c_int(nsplit), _stream()))


def conv2_wgrad(p1, dy2, master, global_w, num_clients, active, w2_fwd, w2_dgrad, lr, mu, grad_out=None):
    _ck(require().olsb_conv2_wgrad(_p(p1), _p(dy2), _p(master), _p(global_w), C.c_longlong(master.shape[1]),
                                   C.c_int(num_clients), _p(active), _p(w2_fwd), _p(w2_dgrad), C.c_float(lr),
                                   C.c_float(mu), _p(grad_out), _stream()))


def conv1_wgrad(x5, x_n_stride, x_n_off, dp1, code1, master, global_w, num_clients, active, w1_fwd, lr, mu,
                grad_out=None):
    _ck(require().olsb_conv1_wgrad(_p(x5), C.c_longlong(x5.shape[0]), C.c_int(x_n_stride), C.c_int(x_n_off), _p(dp1),
                                   _p(code1), _p(master), _p(global_w), C.c_longlong(master.shape[1]), C.c_int(num_clients),
                                   _p(active), _p(w1_fwd), C.c_float(lr), C.c_float(mu), _p(grad_out), _stream()))


def head(num_clients, active, p2, code2, labels, label_off, master, global_w, lr, mu, dy2, loss_sum, nonfinite,
         logits_out=None, apply_update=True, dwfc_out=None, db2_out=None):
    assert labels.dtype == torch.int32
    _ck(require().olsb_head(C.c_int(num_clients), _p(active), _p(p2), _p(code2), _p(labels),
                            C.c_int(labels.shape[1]), C.c_int(label_off), _p(master), _p(global_w),
                            C.c_longlong(master.shape[1]), C.c_float(lr), C.c_float(mu), _p(dy2), _p(loss_sum),
                            _p(nonfinite), _p(logits_out), C.c_int(1 if apply_update else 0), _p(dwfc_out),
                            _p(db2_out), _stream()))


class Cnn2StepArgs(C.Structure):
    _fields_ = [
        ("x5", C.c_void_p), ("x_total", C.c_longlong), ("x_n_stride", C.c_int), ("x_n_off", C.c_int),
        ("labels", C.c_void_p), ("label_stride", C.c_int), ("label_off", C.c_int),
        ("master", C.c_void_p), ("global_w", C.c_void_p), ("P", C.c_longlong), ("C", C.c_int),
        ("active", C.c_void_p),
        ("w1_fwd", C.c_void_p), ("w2_fwd", C.c_void_p), ("w2_dgrad", C.c_void_p),
        ("p1", C.c_void_p), ("code1", C.c_void_p), ("p2", C.c_void_p), ("code2", C.c_void_p),
        ("dy2", C.c_void_p), ("dp1", C.c_void_p),
        ("loss_sum", C.c_void_p), ("nonfinite", C.c_void_p),
        ("lr", C.c_float), ("mu", C.c_float), ("use_prox", C.c_int), ("nsplit", C.c_int),
    ]


def cnn2_local_step(args: Cnn2StepArgs) -> None:
    _ck(require().olsb_cnn2_local_step(C.byref(args), _stream()))


KERNELS_PER_LOCAL_STEP = 6


# --------------------------------------------------------------------------------------------------
def bcast_scatter_cnn2(global_peer_ptrs: Sequence[int], rank: int, shard_len: int, global_local, master,
                       participate, w1_fwd, w2_fwd, w2_dgrad) -> None:
    n = len(global_peer_ptrs)
    _ck(require().olsb_bcast_scatter_cnn2(_ptr_array(global_peer_ptrs), C.c_int(n), C.c_int(rank),
                                          C.c_int(shard_len), _p(global_local), _p(master),
                                          C.c_int(master.shape[0]), _p(participate), _p(w1_fwd), _p(w2_fwd),
                                          _p(w2_dgrad), _stream()))


def fedavg_partial(master, global_w, weight, nonfinite, groups: int, partial, wsum) -> None:
    c, p = master.shape
    _ck(require().olsb_fedavg_partial(_p(master), _p(global_w), _p(weight), _p(nonfinite), C.c_int(c), C.c_int(p),
                                      C.c_int(groups), _p(partial), _p(wsum), _stream()))


def fedavg_finish(partial_peer_ptrs, wsum_peer_ptrs, flag_peer_ptrs, rank: int, groups: int, P: int, shard_len: int,
                  epoch: int, global_w, opt_m, opt_v, opt: int, server_lr: float, beta1: float, beta2: float,
                  eps: float, step: int, stats, arrive_ctr, round_peer_ptrs=None, round_out=None, round_len: int = 8,
                  partial_mc: int = 0, global_mc: int = 0) -> None:
    """``partial_mc`` / ``global_mc``: NVSwitch multicast addresses of the partial-sum and global-model slabs
    (parallel/nvls.py); non-zero selects the multimem.ld_reduce / multimem.st path of the kernel."""
    n = len(partial_peer_ptrs)
    flags = _ptr_array(flag_peer_ptrs) if flag_peer_ptrs is not None else None
    rnd = _ptr_array(round_peer_ptrs) if round_peer_ptrs is not None else None
    _ck(require().olsb_fedavg_finish(_ptr_array(partial_peer_ptrs), _ptr_array(wsum_peer_ptrs), flags, C.c_int(n),
                                     C.c_int(rank), C.c_int(groups), C.c_int(P), C.c_int(shard_len),
                                     C.c_uint(epoch), _p(global_w), _p(opt_m), _p(opt_v), C.c_int(opt),
                                     C.c_float(server_lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                                     C.c_int(step), _p(stats), _p(arrive_ctr), rnd,
                                     _p(round_out) if rnd is not None else None, C.c_int(round_len),
                                     C.c_void_p(partial_mc or None), C.c_void_p(global_mc or None), _stream()))


def set_option(name: str, value: int) -> int:
    """Runtime kernel-variant switch (``wgrad2_mode``: 0 five-box, 1 one wide box, 2 wide box + swizzle base offset)."""
    old = require().olsb_set_option(name.encode(), C.c_int(value))
    if old < 0:
        raise KeyError(name)
    return int(old)


ROUND_REC = 56        # floats per rank record: 8 job-wide + 2 per (dataset, device class), up to 24 classes


def round_stats(loss_sum, nonfinite, participate, step_active, counters, num_clients: int, steps: int, inv_steps: float,
                out, class_id=None, nclass: int = 0) -> None:
    """Per-rank round record: [loss, finite participants, counters[0..3], non-finite clients, client-steps] followed by
    [failed_k, participants_k] for every device class k (``class_id``: int32 class index per local client)."""
    _ck(require().olsb_round_stats(_p(loss_sum), _p(nonfinite), _p(participate), _p(step_active), _p(counters),
                                   C.c_int(num_clients), C.c_int(steps), C.c_float(inv_steps), _p(out), _p(class_id),
                                   C.c_int(nclass), _stream()))


def deviceflow_mask(num_clients, client_base, seed, round_idx, p_online, drop_probability, arrive_tick, max_tick,
                    n_samples, local_steps, max_steps, participate, step_active, agg_weight, counters) -> None:
    _ck(require().olsb_deviceflow_mask(C.c_int(num_clients), C.c_int(client_base), C.c_ulonglong(seed),
                                       C.c_uint(round_idx), _p(p_online), C.c_float(drop_probability),
                                       _p(arrive_tick), C.c_int(max_tick), _p(n_samples), _p(local_steps),
                                       C.c_int(max_steps), _p(participate), _p(step_active), _p(agg_weight),
                                       _p(counters), _stream()))


def flow_ticks(num_clients_total: int, c0: int, num_local: int, seed: int, round_idx: int, p_online_all, scale: float, scratch: dict,
               cum, nslots: int, dropmap, tick) -> None:
    """arrive_tick of this rank's clients for a flow-dispatch schedule, ranked on the device (4 launches)."""
    _ck(require().olsb_flow_ticks(C.c_int(num_clients_total), C.c_int(c0), C.c_int(num_local), C.c_ulonglong(seed), C.c_uint(round_idx),
                                  _p(p_online_all), C.c_float(scale), _p(scratch["keys"]), _p(scratch["hist"]), _p(scratch["start"]),
                                  _p(scratch["cursor"]), _p(scratch["items"]), _p(cum), C.c_int(nslots), _p(dropmap), _p(tick),
                                  _stream()))


def rate_curve_area(y: torch.Tensor, area: torch.Tensor) -> None:
    """y: float64 [bins, 101] rate-function samples; area: float64 [bins] positive-trapezoid integrals."""
    assert y.dtype == torch.float64 and area.dtype == torch.float64 and y.shape[-1] == 101
    _ck(require().olsb_rate_curve_area(_p(y), _p(area), C.c_int(area.numel()), _stream()))


def client_sgd(w, g, wg, active, lr, mu=0.0, momentum=0.0, mom=None) -> None:
    c, p = w.shape
    assert p % 4 == 0
    _ck(require().olsb_client_sgd(_p(w), _p(g), _p(wg), _p(active), C.c_int(c), C.c_int(p), C.c_float(lr),
                                  C.c_float(mu), C.c_float(momentum), _p(mom), _stream()))


def philox_host(c0, c1, c2, c3, k0, k1):
    out = (C.c_uint * 4)()
    lib().olsb_philox_host(C.c_uint(c0), C.c_uint(c1), C.c_uint(c2), C.c_uint(c3), C.c_uint(k0), C.c_uint(k1), out)
    return tuple(int(v) for v in out)


# --------------------------------------------------------------------------------------------------
# generic per-client GEMM family (fully-connected layers) — csrc/linear_tc.cuh
class Operand(C.Structure):
    _fields_ = [("base", C.c_void_p), ("client_stride", C.c_longlong), ("row_stride", C.c_longlong),
                ("rows", C.c_int), ("inner", C.c_int)]


class GemmArgs(C.Structure):
    _fields_ = [("variant", C.c_int), ("C", C.c_int), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("a", Operand), ("b", Operand), ("active", C.c_void_p), ("lr", C.c_float), ("mu", C.c_float),
                ("out_t", C.c_void_p), ("out_stride", C.c_longlong), ("out_ld", C.c_int),
                ("bias", C.c_void_p), ("bias_stride", C.c_longlong), ("relu", C.c_int),
                ("mask_t", C.c_void_p), ("b_master", C.c_void_p), ("b_stride", C.c_longlong), ("b_global", C.c_void_p),
                ("db_out", C.c_void_p),
                ("w_master", C.c_void_p), ("w_stride", C.c_longlong), ("w_ld", C.c_int), ("w_global", C.c_void_p),
                ("w_bf16", C.c_void_p), ("wb_stride", C.c_longlong), ("grad_out", C.c_void_p), ("g_stride", C.c_longlong)]


GEMM_FWD, GEMM_DGRAD, GEMM_WGRAD = 0, 1, 2


def operand(t: torch.Tensor, elem_offset: int, client_stride: int, row_stride: int, rows: int, inner: int) -> Operand:
    """bf16 operand view [C][rows][inner] inside tensor `t` (strides in elements).  TMA needs 16-byte aligned bases
    and strides."""
    assert t.dtype == torch.bfloat16 and t.is_cuda
    base = t.data_ptr() + 2 * elem_offset
    assert base % 16 == 0 and (client_stride * 2) % 16 == 0 and (row_stride * 2) % 16 == 0, "TMA alignment"
    return Operand(C.c_void_p(base), client_stride, row_stride, rows, inner)


def _vp(t: Optional[torch.Tensor], elem_offset: int = 0):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr() + t.element_size() * elem_offset)


def client_gemm(args: GemmArgs) -> None:
    _ck(require().olsb_client_gemm(C.byref(args), _stream()))


def client_xent(num_clients, batch, classes, ld, active, logits, labels, label_off, dlogits, loss_sum, nonfinite,
                b_master=None, b_offset=0, b_stride=0, b_global=None, lr=0.0, mu=0.0, db_out=None) -> None:
    _ck(require().olsb_client_xent(C.c_int(num_clients), C.c_int(batch), C.c_int(classes), C.c_int(ld), _p(active),
                                   _p(logits), _p(labels), C.c_int(labels.shape[1]), C.c_int(label_off), _p(dlogits),
                                   _p(loss_sum), _p(nonfinite), _vp(b_master, b_offset), C.c_longlong(b_stride),
                                   _vp(b_global, b_offset) if b_global is not None else None, C.c_float(lr),
                                   C.c_float(mu), _p(db_out), _stream()))


def bcast_scatter_generic(global_peer_ptrs: Sequence[int], shard_len: int, P: int, global_local, master, master_bf16,
                          participate, p_bf16: int = 0, bf16_stride: int = 0) -> None:
    """``p_bf16`` / ``bf16_stride``: the bf16 operand copy holds only the leading ``p_bf16`` elements of every row
    (the tensor-core operands), rows ``bf16_stride`` apart; 0 = the whole row."""
    n = len(global_peer_ptrs)
    _ck(require().olsb_bcast_scatter_generic(_ptr_array(global_peer_ptrs), C.c_int(n), C.c_int(shard_len), C.c_int(P),
                                             _p(global_local), _p(master), _p(master_bf16), C.c_int(master.shape[0]),
                                             _p(participate), C.c_int(p_bf16), C.c_longlong(bf16_stride), _stream()))


assert CNN2.P % 4 == 0
