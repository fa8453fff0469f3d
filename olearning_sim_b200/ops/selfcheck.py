"""Numerics self-check of every CUDA kernel against the plain-PyTorch fp32 oracle.

``python -m olearning_sim_b200.ops.selfcheck [--clients N] [--json out.json]`` runs all checks, prints one line per
check and never stops at the first failure (a GPU call is expensive; we want the whole picture per run).  The same
functions back ``tests/test_gpu_kernels.py``.
"""
from __future__ import annotations

import argparse
import json
import sys
import traceback
from typing import Callable, Dict, List

import torch
import torch.nn.functional as F

from . import native, reference
from .layout import CNN2

B = CNN2.BATCH


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    denom = b.norm().item()
    return (a - b).norm().item() / (denom if denom > 0 else 1.0)


def _rand_master(c: int, seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    rows = [CNN2.init(seed + i) for i in range(c)]
    m = torch.stack(rows) * scale
    return m.cuda()


def _operand_copies(master: torch.Tensor):
    c = master.shape[0]
    L = CNN2
    w1f = master[:, L.W1:L.W1 + L.W1_N].to(torch.bfloat16).contiguous()
    w2 = master[:, L.W2:L.W2 + L.W2_N]
    w2f = w2.to(torch.bfloat16).contiguous()
    w2d = w2.reshape(c, 64, 800).transpose(1, 2).to(torch.bfloat16).contiguous()   # [C][tap*32+cin][cout]
    return w1f, w2f, w2d


def _conv_ref(x_nhwc: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    """per-client conv, x [C,B,H,W,Cin] fp32, w [C,Cout,Cin,5,5] -> [C,B,H,W,Cout]"""
    c, b, h, ww, cin = x_nhwc.shape
    cout = w.shape[1]
    xg = x_nhwc.permute(1, 0, 4, 2, 3).reshape(b, c * cin, h, ww)
    y = F.conv2d(xg, w.reshape(c * cout, cin, 5, 5), None if bias is None else bias.reshape(-1), padding=2, groups=c)
    return y.reshape(b, c, cout, h, ww).permute(1, 0, 3, 4, 2).contiguous()


def _pool_ref(a: torch.Tensor) -> torch.Tensor:
    c, b, h, w, ch = a.shape
    return a.reshape(c, b, h // 2, 2, w // 2, 2, ch).amax(dim=(3, 5))


def _unpool_with_code(g: torch.Tensor, code: torch.Tensor) -> torch.Tensor:
    """g, code [N,h,w,ch] -> [N,2h,2w,ch] placing g at the coded position."""
    n, h, w, ch = g.shape
    out = torch.zeros(n, h, 2, w, 2, ch, device=g.device, dtype=g.dtype)
    for pos in range(4):
        out[:, :, pos >> 1, :, pos & 1, :] = torch.where(code == pos, g, torch.zeros_like(g))
    return out.reshape(n, 2 * h, 2 * w, ch)


# --------------------------------------------------------------------------------------------------
X5_DEFAULT = 2          # library default of the ``x5_rows`` switch (conv1 input layout)


def check_ingest(c: int) -> Dict:
    n = c * 4
    x = torch.randint(0, 256, (n, 32, 32, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty(n, 32, 32, 16, dtype=torch.bfloat16, device="cuda")
    native.ingest_u8(x, out)
    ref = reference.ingest_ref(x, preswizzle=X5_DEFAULT).to(torch.bfloat16)
    return {"err": rel_err(out.float(), ref.float()), "tol": 1e-4}


def _make_x5(c: int, shard: int, seed: int = 1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randint(0, 256, (c * shard, 32, 32, 3), dtype=torch.uint8, device="cuda", generator=g)
    x5 = torch.empty(c * shard, 32, 32, 16, dtype=torch.bfloat16, device="cuda")
    native.ingest_u8(x, x5)
    return x, x5


def check_conv1_fwd(c: int) -> Dict:
    shard, off = 2 * B, B // 2
    xu8, x5 = _make_x5(c, shard)
    master = _rand_master(c)
    w1f, _, _ = _operand_copies(master)
    p1 = torch.zeros(c * B, 16, 16, 32, dtype=torch.bfloat16, device="cuda")
    code1 = torch.full((c * B, 16, 16, 32), 9, dtype=torch.uint8, device="cuda")
    native.conv1_fwd(x5, shard, off, w1f, master, c, None, p1, code1)
    torch.cuda.synchronize()
    x = reference.normalize_u8(xu8).reshape(c, shard, 32, 32, 3)[:, off:off + B].to(torch.bfloat16).float()
    v = reference.cnn2_views(master)
    a1 = F.relu(_conv_ref(x, v["w1"].to(torch.bfloat16).float(), v["b1"]))
    ref = _pool_ref(a1).reshape(c * B, 16, 16, 32)
    # code consistency: un-pooling the pooled value with the code must reproduce a1 at the argmax position
    a1f = a1.reshape(c * B, 16, 2, 16, 2, 32)
    pos_h = (code1 >> 1).long()
    pos_w = (code1 & 1).long()
    gathered = torch.gather(torch.gather(a1f, 2, pos_h[:, :, None, :, None, :].expand(-1, -1, 1, -1, 2, -1)), 4,
                            pos_w[:, :, None, :, None, :].expand(-1, -1, 1, -1, 1, -1)).reshape(c * B, 16, 16, 32)
    return {"err": rel_err(p1.float(), ref), "tol": 1e-2, "code_err": rel_err(gathered, ref),
            "code_max": int(code1.max().item())}


def check_conv2_fwd(c: int) -> Dict:
    master = _rand_master(c)
    _, w2f, _ = _operand_copies(master)
    p1 = (torch.randn(c * B, 16, 16, 32, device="cuda").clamp_min(0)).to(torch.bfloat16)
    p2 = torch.zeros(c * B, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    code2 = torch.full((c * B, 8, 8, 64), 9, dtype=torch.uint8, device="cuda")
    native.conv2_fwd(p1, w2f, master, c, None, p2, code2)
    torch.cuda.synchronize()
    v = reference.cnn2_views(master)
    a2 = F.relu(_conv_ref(p1.float().reshape(c, B, 16, 16, 32), v["w2"].to(torch.bfloat16).float(), v["b2"]))
    ref = _pool_ref(a2).reshape(c * B, 8, 8, 64)
    return {"err": rel_err(p2.float(), ref), "tol": 1e-2, "code_max": int(code2.max().item())}


def check_conv2_fwd_shift(c: int) -> Dict:
    """conv2 forward as shifted-view MMAs on 16x8 tiles (``fwd2_shift`` switch) vs the folded kernel: numerics (values and
    argmax codes against each other and the fp32 oracle) and time"""
    master = _rand_master(c)
    _, w2f, _ = _operand_copies(master)
    p1 = (torch.randn(c * B, 16, 16, 32, device="cuda").clamp_min(0)).to(torch.bfloat16)
    v = reference.cnn2_views(master)
    a2 = F.relu(_conv_ref(p1.float().reshape(c, B, 16, 16, 32), v["w2"].to(torch.bfloat16).float(), v["b2"]))
    ref = _pool_ref(a2).reshape(c * B, 8, 8, 64)
    res: Dict = {"tol": 1e-2, "err": 0.0}
    outs = {}
    try:
        for mode in (0, 1):
            native.set_option("fwd2_shift", mode)
            p2 = torch.zeros(c * B, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
            code2 = torch.full((c * B, 8, 8, 64), 9, dtype=torch.uint8, device="cuda")
            native.conv2_fwd(p1, w2f, master, c, None, p2, code2)
            torch.cuda.synchronize()
            outs[mode] = (p2, code2)
            res[f"shift{mode}_err"] = rel_err(p2.float(), ref)
            res[f"shift{mode}_code_max"] = int(code2.max().item())
            res[f"shift{mode}_us"] = round(_time_us(lambda: native.conv2_fwd(p1, w2f, master, c, None, p2, code2)), 1)
            res["err"] = max(res["err"], res[f"shift{mode}_err"])
        res["code_mismatch"] = float((outs[0][1] != outs[1][1]).float().mean().item())
        res["value_mismatch"] = rel_err(outs[1][0].float(), outs[0][0].float())
    finally:
        native.set_option("fwd2_shift", -1)       # default: automatic by client count
    return res


def check_conv1_window(c: int) -> Dict:
    """conv1 forward with pooling done by the accumulator layout on the column-parity-plane input (``x5_rows`` = 2) next
    to the row layout (= 1): ingest layout, forward (values + argmax codes), weight gradient, and the time of both"""
    res: Dict = {"tol": 1e-2, "err": 0.0}
    try:
        for mode in (1, 2):
            native.set_option("x5_rows", mode)
            n = 8
            xi = torch.randint(0, 256, (n, 32, 32, 3), dtype=torch.uint8, device="cuda")
            out = torch.empty(n, 32, 32, 16, dtype=torch.bfloat16, device="cuda")
            native.ingest_u8(xi, out)
            res[f"m{mode}_ingest"] = rel_err(out.float(), reference.ingest_ref(xi, preswizzle=mode).to(torch.bfloat16).float())
            f = check_conv1_fwd(c)
            w = check_conv1_wgrad(c)
            res[f"m{mode}_fwd"], res[f"m{mode}_code"], res[f"m{mode}_wgrad"] = f["err"], f["code_err"], w["err"]
            res["err"] = max(res["err"], res[f"m{mode}_ingest"] * 100, f["err"], f["code_err"], w["err"])
            # timing
            shard, off = 2 * B, B // 2
            _, x5 = _make_x5(c, shard)
            master = _rand_master(c)
            w1f, _, _ = _operand_copies(master)
            p1 = torch.zeros(c * B, 16, 16, 32, dtype=torch.bfloat16, device="cuda")
            code1 = torch.zeros(c * B, 16, 16, 32, dtype=torch.uint8, device="cuda")
            dp1 = (torch.randn(c * B, 16, 16, 32, device="cuda") * 0.1).to(torch.bfloat16)
            gout = torch.zeros(c, CNN2.W1_N, device="cuda")
            res[f"m{mode}_fwd_us"] = round(_time_us(lambda: native.conv1_fwd(x5, shard, off, w1f, master, c, None, p1, code1)), 1)
            res[f"m{mode}_wgrad_us"] = round(_time_us(lambda: native.conv1_wgrad(x5, shard, off, dp1, code1, master, None, c,
                                                                                None, w1f, 0.1, 0.0, gout)), 1)
    finally:
        native.set_option("x5_rows", X5_DEFAULT)
    return res


def check_zero_round(c: int) -> Dict:
    """engine-level edge case: a round in which no client is online (availability trace = 0) must leave the global model
    untouched and finite on the native path, then training resumes; also a round where every update is dropped"""
    from ..engine.fed_engine import FedConfig, FedEngine
    res: Dict = {"tol": 0.0, "err": 0.0}
    e = FedEngine(FedConfig(model="cnn2", num_clients=max(c, 2), local_steps=2, shard=64, backend="native", seed=2,
                            p_online=1.0, online_trace="0 if t == 1 else 1"))
    e.load_synthetic()
    r0 = e.run_round(0)
    g0 = e.global_model().clone()
    r1 = e.run_round(1)
    g1 = e.global_model().clone()
    r2 = e.run_round(2)
    res.update(part=[r0.participants, r1.participants, r2.participants], loss=[r0.loss, r1.loss, r2.loss],
               unchanged=float((g1 - g0).abs().max()), moved=float((e.global_model() - g1).abs().max()),
               finite=bool(torch.isfinite(e.global_model()).all()))
    ok = r1.participants == 0 and res["unchanged"] == 0.0 and res["moved"] > 0 and res["finite"] and r1.client_steps == 0
    e.close()
    d = FedEngine(FedConfig(model="cnn2", num_clients=max(c, 2), local_steps=1, shard=64, backend="native", seed=2,
                            drop_probability=1.0))
    d.load_synthetic()
    gd = d.global_model().clone()
    rd = d.run_round(0)
    res.update(dropped=rd.dropped, arrived=rd.arrived, drop_unchanged=float((d.global_model() - gd).abs().max()))
    ok = ok and rd.arrived == 0 and rd.dropped == rd.participants and res["drop_unchanged"] == 0.0
    d.close()
    res["err"] = 0.0 if ok else 1.0
    return res


def check_head(c: int) -> Dict:
    master = _rand_master(c)
    p2 = torch.randn(c * B, 8, 8, 64, device="cuda").clamp_min(0).to(torch.bfloat16)
    code2 = torch.randint(0, 4, (c * B, 8, 8, 64), dtype=torch.uint8, device="cuda")
    labels = torch.randint(0, 10, (c, 2 * B), dtype=torch.int32, device="cuda")
    off = B // 2
    dy2 = torch.zeros(c * B, 16, 16, 64, dtype=torch.bfloat16, device="cuda")
    loss = torch.zeros(c, device="cuda")
    nonf = torch.zeros(c, dtype=torch.int32, device="cuda")
    logits = torch.zeros(c, B, 10, device="cuda")
    dwfc = torch.zeros(c, CNN2.WFC_N + 16, device="cuda")
    db2 = torch.zeros(c, 64, device="cuda")
    m2 = master.clone()
    native.head(c, None, p2, code2, labels, off, m2, None, 0.1, 0.0, dy2, loss, nonf, logits, False, dwfc, db2)
    torch.cuda.synchronize()
    v = reference.cnn2_views(master)
    feat = p2.float().reshape(c, B, 4096).requires_grad_(True)
    wfc = v["wfc"].clone().requires_grad_(True)
    bfc = v["bfc"].clone().requires_grad_(True)
    lg = torch.bmm(feat, wfc.transpose(1, 2)) + bfc[:, None, :]
    y = labels[:, off:off + B].long()
    l = F.cross_entropy(lg.reshape(c * B, 10), y.reshape(-1), reduction="none").reshape(c, B).mean(1)
    l.sum().backward()
    dfeat = torch.where(feat > 0, feat.grad, torch.zeros_like(feat.grad)).reshape(c * B, 8, 8, 64)
    dy2_ref = _unpool_with_code(dfeat, code2)
    res = {
        "logits": rel_err(logits, lg.detach()), "loss": rel_err(loss, l.detach()),
        "dwfc": rel_err(dwfc[:, :CNN2.WFC_N], wfc.grad.reshape(c, -1)),
        "dbfc": rel_err(dwfc[:, CNN2.WFC_N:CNN2.WFC_N + 10], bfc.grad),
        "dy2": rel_err(dy2.float(), dy2_ref), "db2": rel_err(db2, dfeat.reshape(c, -1, 64).sum(1)),
        "unchanged": float((m2 - master).abs().max().item()),
    }
    res["err"] = max(res["logits"], res["loss"], res["dwfc"], res["dbfc"], res["dy2"], res["db2"])
    res["tol"] = 1e-2
    # update mode: w_fc' = w_fc - lr*g
    m3 = master.clone()
    native.head(c, None, p2, code2, labels, off, m3, None, 0.1, 0.0, dy2, loss, nonf, None, True, None, None)
    torch.cuda.synchronize()
    exp = master[:, CNN2.WFC:CNN2.WFC + CNN2.WFC_N] - 0.1 * wfc.grad.reshape(c, -1)
    res["sgd"] = rel_err(m3[:, CNN2.WFC:CNN2.WFC + CNN2.WFC_N], exp)
    res["err"] = max(res["err"], res["sgd"])
    return res


def check_conv2_dgrad(c: int) -> Dict:
    master = _rand_master(c)
    _, w2f, _ = _operand_copies(master)
    dy2 = (torch.randn(c * B, 16, 16, 64, device="cuda") * 0.1).to(torch.bfloat16)
    p1 = torch.randn(c * B, 16, 16, 32, device="cuda").clamp_min(0).to(torch.bfloat16)
    dp1 = torch.zeros(c * B, 16, 16, 32, dtype=torch.bfloat16, device="cuda")
    db1 = torch.zeros(c, 32, device="cuda")
    native.conv2_dgrad(dy2, w2f, master, None, c, None, p1, dp1, 0.1, 0.0, db1)      # B operand = forward weight copy
    torch.cuda.synchronize()
    v = reference.cnn2_views(master)
    x = p1.float().reshape(c, B, 16, 16, 32).clone().requires_grad_(True)
    y = _conv_ref(x, v["w2"].to(torch.bfloat16).float(), None)
    y.backward(dy2.float().reshape(c, B, 16, 16, 64))
    ref = torch.where(x > 0, x.grad, torch.zeros_like(x.grad)).reshape(c * B, 16, 16, 32)
    return {"err": max(rel_err(dp1.float(), ref), rel_err(db1, ref.reshape(c, -1, 32).sum(1))), "tol": 1.5e-2,
            "dp1": rel_err(dp1.float(), ref), "db1": rel_err(db1, ref.reshape(c, -1, 32).sum(1))}


def check_conv2_dgrad_pair(c: int) -> Dict:
    """conv2 data gradient on CTA pairs (cta_group::2, ``dgrad_pair`` switch) vs the single-CTA kernel: numerics, time"""
    master = _rand_master(c)
    _, w2f, _ = _operand_copies(master)
    dy2 = (torch.randn(c * B, 16, 16, 64, device="cuda") * 0.1).to(torch.bfloat16)
    p1 = torch.randn(c * B, 16, 16, 32, device="cuda").clamp_min(0).to(torch.bfloat16)
    v = reference.cnn2_views(master)
    x = p1.float().reshape(c, B, 16, 16, 32).clone().requires_grad_(True)
    y = _conv_ref(x, v["w2"].to(torch.bfloat16).float(), None)
    y.backward(dy2.float().reshape(c, B, 16, 16, 64))
    ref = torch.where(x > 0, x.grad, torch.zeros_like(x.grad)).reshape(c * B, 16, 16, 32)
    res: Dict = {"tol": 1.5e-2, "err": 0.0}
    try:
        for mode in (0, 1):
            native.set_option("dgrad_pair", mode)
            dp1 = torch.zeros(c * B, 16, 16, 32, dtype=torch.bfloat16, device="cuda")
            db1 = torch.zeros(c, 32, device="cuda")
            native.conv2_dgrad(dy2, w2f, master, None, c, None, p1, dp1, 0.1, 0.0, db1)
            torch.cuda.synchronize()
            res[f"pair{mode}_dp1"] = rel_err(dp1.float(), ref)
            res[f"pair{mode}_db1"] = rel_err(db1, ref.reshape(c, -1, 32).sum(1))
            res[f"pair{mode}_us"] = round(_time_us(lambda: native.conv2_dgrad(dy2, w2f, master, None, c, None, p1, dp1,
                                                                           0.1, 0.0, db1)), 1)
            res["err"] = max(res["err"], res[f"pair{mode}_dp1"], res[f"pair{mode}_db1"])
    finally:
        native.set_option("dgrad_pair", -1)       # default: automatic by client count
    return res


def check_conv2_wgrad(c: int) -> Dict:
    master = _rand_master(c)
    w1f, w2f, w2d = _operand_copies(master)
    dy2 = (torch.randn(c * B, 16, 16, 64, device="cuda") * 0.1).to(torch.bfloat16)
    p1 = torch.randn(c * B, 16, 16, 32, device="cuda").clamp_min(0).to(torch.bfloat16)
    gout = torch.zeros(c, CNN2.W2_N, device="cuda")
    native.conv2_wgrad(p1, dy2, master, None, c, None, w2f, w2d, 0.1, 0.0, gout)
    torch.cuda.synchronize()
    w = reference.cnn2_views(master)["w2"].clone().requires_grad_(True)
    y = _conv_ref(p1.float().reshape(c, B, 16, 16, 32), w, None)
    y.backward(dy2.float().reshape(c, B, 16, 16, 64))
    gref = w.grad.permute(0, 1, 3, 4, 2).reshape(c, -1)       # cout, kh, kw, cin
    res = {"err": rel_err(gout, gref), "tol": 1e-2}
    # fused-update mode
    m2 = master.clone()
    native.conv2_wgrad(p1, dy2, m2, None, c, None, w2f, w2d, 0.1, 0.0, None)
    torch.cuda.synchronize()
    exp = master[:, CNN2.W2:CNN2.W2 + CNN2.W2_N] - 0.1 * gref
    res["sgd"] = rel_err(m2[:, CNN2.W2:CNN2.W2 + CNN2.W2_N], exp)
    res["w2f"] = rel_err(w2f.float(), exp.to(torch.bfloat16).float())
    res["w2d"] = rel_err(w2d.float(), exp.reshape(c, 64, 800).transpose(1, 2).to(torch.bfloat16).float())
    res["err"] = max(res["err"], res["sgd"], res["w2f"], res["w2d"])
    return res


def check_conv2_wgrad_wide(c: int) -> Dict:
    """conv2 weight gradient with one 12x20 halo box per chunk (modes 1/2 of the ``wgrad2_mode`` switch): error per
    horizontal tap (the taps are 64-byte operand shifts inside a swizzle pattern) and time against the 5-box kernel"""
    master = _rand_master(c)
    _, w2f, w2d = _operand_copies(master)
    dy2 = (torch.randn(c * B, 16, 16, 64, device="cuda") * 0.1).to(torch.bfloat16)
    p1 = torch.randn(c * B, 16, 16, 32, device="cuda").clamp_min(0).to(torch.bfloat16)
    w = reference.cnn2_views(master)["w2"].clone().requires_grad_(True)
    y = _conv_ref(p1.float().reshape(c, B, 16, 16, 32), w, None)
    y.backward(dy2.float().reshape(c, B, 16, 16, 64))
    gref = w.grad.permute(0, 1, 3, 4, 2).reshape(c, 64, 5, 5, 32)       # cout, kh, kw, cin
    res: Dict = {"tol": 1e-2, "err": 0.0}
    try:
        for mode, ws in ((0, 0), (0, 1), (1, 0), (1, 1)):
            native.set_option("wgrad2_mode", mode)
            native.set_option("wgrad2_ws", ws)
            gout = torch.zeros(c, CNN2.W2_N, device="cuda")
            native.conv2_wgrad(p1, dy2, master, None, c, None, w2f, w2d, 0.1, 0.0, gout)
            torch.cuda.synchronize()
            g = gout.reshape(c, 64, 5, 5, 32)
            tag = f"box{mode}_ws{ws}"
            res[f"{tag}_kw"] = [round(rel_err(g[:, :, :, k], gref[:, :, :, k]), 5) for k in range(5)]
            res[f"{tag}_kh"] = [round(rel_err(g[:, :, k], gref[:, :, k]), 5) for k in range(5)]
            res[f"{tag}_us"] = round(_time_us(lambda: native.conv2_wgrad(p1, dy2, master, None, c, None, w2f, w2d,
                                                                         0.1, 0.0, gout)), 1)
    finally:
        native.set_option("wgrad2_mode", 1)      # defaults: one wide box per chunk + weight-stationary MMAs
        native.set_option("wgrad2_ws", 1)
    return res


def _time_us(fn, iters: int = 5) -> float:
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / iters


def check_conv1_wgrad(c: int) -> Dict:
    shard, off = 2 * B, B
    xu8, x5 = _make_x5(c, shard, seed=5)
    master = _rand_master(c)
    w1f, _, _ = _operand_copies(master)
    dp1 = (torch.randn(c * B, 16, 16, 32, device="cuda") * 0.1).to(torch.bfloat16)
    code1 = torch.randint(0, 4, (c * B, 16, 16, 32), dtype=torch.uint8, device="cuda")
    gout = torch.zeros(c, CNN2.W1_N, device="cuda")
    native.conv1_wgrad(x5, shard, off, dp1, code1, master, None, c, None, w1f, 0.1, 0.0, gout)
    torch.cuda.synchronize()
    dy1 = _unpool_with_code(dp1.float(), code1)
    x = reference.normalize_u8(xu8).reshape(c, shard, 32, 32, 3)[:, off:off + B].to(torch.bfloat16).float()
    w = reference.cnn2_views(master)["w1"].clone().requires_grad_(True)
    y = _conv_ref(x, w, None)
    y.backward(dy1.reshape(c, B, 32, 32, 32))
    gref = torch.zeros(c, 32, 5, 16, device="cuda")
    gref[..., :15] = w.grad.permute(0, 1, 3, 4, 2).reshape(c, 32, 5, 15)
    return {"err": rel_err(gout, gref.reshape(c, -1)), "tol": 1e-2}


def check_local_step(c: int) -> Dict:
    """whole fused step (6 kernels) vs autograd SGD step on the same bf16-rounded operands"""
    from ..engine.cnn2_native import Cnn2Workspace
    shard = 2 * B
    xu8, _ = _make_x5(c, shard, seed=11)
    labels = torch.randint(0, 10, (c, shard), dtype=torch.int32, device="cuda")
    ws = Cnn2Workspace(c, shard, device="cuda")
    ws.load_data(xu8.reshape(c, shard, 32, 32, 3), labels)
    g0 = CNN2.init(123).cuda()
    ws.set_global(g0)
    ws.broadcast()
    lr = 0.05
    ws.local_step(0, lr=lr)
    torch.cuda.synchronize()
    x = reference.normalize_u8(xu8).reshape(c, shard, 32, 32, 3)[:, :B]
    flat0 = g0[None, :].repeat(c, 1)
    new, loss = reference.cnn2_local_step(flat0, x, labels[:, :B], lr, emulate_bf16=True)
    d_k = ws.master - flat0
    d_r = new - flat0
    res = {"loss": rel_err(ws.loss_sum, loss), "delta": rel_err(d_k, d_r), "tol": 6e-2}
    res["err"] = max(res["loss"], res["delta"])
    return res


def check_reduce(c: int) -> Dict:
    P = CNN2.P
    master = torch.randn(c, P, device="cuda")
    g = torch.randn(P, device="cuda")
    w = torch.rand(c, device="cuda")
    w[::3] = 0
    nonf = torch.zeros(c, dtype=torch.int32, device="cuda")
    if c > 1:
        nonf[1] = 1
    groups = 4
    partial = torch.zeros(groups, P, device="cuda")
    wsum = torch.zeros(groups, device="cuda")
    native.fedavg_partial(master, g, w, nonf, groups, partial, wsum)
    weff = torch.where(nonf > 0, torch.zeros_like(w), w)
    dref, ws_ref = reference.fedavg_delta(master, g, weff)
    res = {"partial": rel_err(partial.sum(0), dref), "wsum": abs(wsum.sum().item() - ws_ref)}
    for opt in (0, 1, 2):
        gw = g.clone()
        m = torch.zeros(P, device="cuda")
        v = torch.zeros(P, device="cuda")
        stats = torch.zeros(4, device="cuda")
        ctr = torch.zeros(2, dtype=torch.int32, device="cuda")
        state: dict = {}
        ref = g.clone()
        for step in (1, 2):
            native.fedavg_finish([partial.data_ptr()], [wsum.data_ptr()], None, 0, groups, P, P, step, gw, m, v, opt,
                                 0.5, 0.9, 0.99, 1e-3, step, stats, ctr)
            ref = reference.server_update(ref, dref, ws_ref, opt, state, 0.5, 0.9, 0.99, 1e-3)
        torch.cuda.synchronize()
        res[f"opt{opt}"] = rel_err(gw, ref)
    res["err"] = max(res["partial"], res["opt0"], res["opt1"], res["opt2"])
    res["tol"] = 1e-4
    return res


def check_bcast(c: int) -> Dict:
    g = CNN2.init(7).cuda()
    P = CNN2.P
    master = torch.zeros(c, P, device="cuda")
    gl = torch.zeros(P, device="cuda")
    w1f = torch.zeros(c, CNN2.W1_N, dtype=torch.bfloat16, device="cuda")
    w2f = torch.zeros(c, CNN2.W2_N, dtype=torch.bfloat16, device="cuda")
    w2d = torch.zeros(c, CNN2.W2_N, dtype=torch.bfloat16, device="cuda")
    native.bcast_scatter_cnn2([g.data_ptr()], 0, P, gl, master, None, w1f, w2f, w2d)
    torch.cuda.synchronize()
    e1, e2, e3 = _operand_copies(g[None, :].repeat(c, 1))
    errs = [rel_err(master, g[None, :].repeat(c, 1)), rel_err(gl, g), rel_err(w1f.float(), e1.float()),
            rel_err(w2f.float(), e2.float()), rel_err(w2d.float().reshape(c, -1), e3.float().reshape(c, -1))]
    return {"err": max(errs), "tol": 1e-7, "parts": errs}


def check_mask(c: int) -> Dict:
    from ..deviceflow import philox
    n = max(c, 257)
    p_on = torch.rand(n, device="cuda")
    ns = torch.randint(1, 100, (n,), dtype=torch.int32, device="cuda")
    steps = torch.randint(1, 6, (n,), dtype=torch.int32, device="cuda")
    part = torch.zeros(n, dtype=torch.uint8, device="cuda")
    act = torch.zeros(5, n, dtype=torch.uint8, device="cuda")
    aw = torch.zeros(n, device="cuda")
    cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    native.deviceflow_mask(n, 1000, 0x1234567890ABCDEF, 3, p_on, 0.25, None, 0, ns, steps, 5, part, act, aw, cnt)
    torch.cuda.synchronize()
    ref = philox.round_mask(n, 1000, 0x1234567890ABCDEF, 3, p_on.cpu().numpy(), 0.25, None, 0, ns.cpu().numpy(),
                            steps.cpu().numpy(), 5)
    import numpy as np
    ok = (np.array_equal(part.cpu().numpy(), ref["participate"]) and np.array_equal(act.cpu().numpy(), ref["step_active"])
          and np.allclose(aw.cpu().numpy(), ref["agg_weight"]) and np.array_equal(cnt.cpu().numpy(), ref["counters"]))
    return {"err": 0.0 if ok else 1.0, "tol": 0.5, "counters": cnt.cpu().tolist()}


# ---- generic per-client GEMM family (linear_tc.cuh) ------------------------------------------------------------
def _mlp_ws(c: int, shard: int = 64):
    from ..engine.mlp_native import MlpLayout, MlpWorkspace
    from ..models.zoo import get_model
    ws = MlpWorkspace(c, shard, MlpLayout(), device="cuda")
    spec = get_model("mlp2")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(c, shard, 784, generator=g)
    y = torch.randint(0, 10, (c, shard), generator=g, dtype=torch.int32)
    ws.load_data(x, y)
    rows = torch.stack([ws.L.pack(spec.init(100 + i)) for i in range(c)]).cuda()
    ws.master.copy_(rows)
    ws.master_bf16.copy_(rows.to(torch.bfloat16))
    ws.global_local.copy_(rows[0])
    return ws, spec


def check_gemm(c: int) -> Dict:
    """the three GEMM variants against torch on the MLP-2 shapes (K tails, M=10 row tile, N=784 with a partial tile)"""
    ws, _ = _mlp_ws(c)
    L, P = ws.L, ws.P
    op, vp = native.operand, native._vp
    off = B
    Xf = ws.x[:, off:off + B].float()
    W1 = ws.master_bf16[:, L.W1:L.W1 + 200 * 784].float().reshape(c, 200, 784)
    W2 = ws.master_bf16[:, L.W2:L.W2 + 2000].float().reshape(c, 10, 200)
    b1 = ws.master[:, L.B1:L.B1 + 200]
    res = {}
    # forward (K-major x K-major, bias + ReLU, transposed store), K = 784 = 12*64 + 16
    a = native.GemmArgs(variant=native.GEMM_FWD, C=c, M=200, N=B, K=784, a=op(ws.master_bf16, L.W1, P, 784, 200, 784),
                        b=op(ws.x, off * 784, ws.shard * 784, 784, B, 784), out_t=vp(ws.h), out_stride=B * 200, out_ld=200,
                        bias=vp(ws.master, L.B1), bias_stride=P, relu=1)
    native.client_gemm(a)
    href = torch.relu(torch.bmm(Xf, W1.transpose(1, 2)) + b1[:, None, :])
    res["fwd"] = rel_err(ws.h.float(), href)
    # dgrad (A MN-major, K = 10), ReLU gate + raw bias gradient
    dl = (torch.randn(c, B, 16, device="cuda") * 0.1)
    dl[..., 10:] = 0
    ws.dlogits.copy_(dl.to(torch.bfloat16))
    dbo = torch.zeros(c, 200, device="cuda")
    a = native.GemmArgs(variant=native.GEMM_DGRAD, C=c, M=200, N=B, K=10, a=op(ws.master_bf16, L.W2, P, 200, 10, 200),
                        b=op(ws.dlogits, 0, B * 16, 16, B, 16), out_t=vp(ws.dh), out_stride=B * 200, out_ld=200,
                        mask_t=vp(ws.h), db_out=vp(dbo))
    native.client_gemm(a)
    dlf = ws.dlogits.float()[..., :10]
    dh_ref = torch.bmm(dlf, W2) * (ws.h.float() > 0)
    res["dgrad"] = rel_err(ws.dh.float(), dh_ref)
    res["db"] = rel_err(dbo, dh_ref.sum(1))
    # wgrad (both MN-major, K = batch), N = 784 -> 4 N tiles, the last one 16 wide
    gout = torch.zeros(c, 200 * 784, device="cuda")
    a = native.GemmArgs(variant=native.GEMM_WGRAD, C=c, M=200, N=784, K=B, a=op(ws.dh, 0, B * 200, 200, B, 200),
                        b=op(ws.x, off * 784, ws.shard * 784, 784, B, 784), w_ld=784, grad_out=vp(gout), g_stride=200 * 784)
    native.client_gemm(a)
    gref = torch.bmm(ws.dh.float().transpose(1, 2), Xf).reshape(c, -1)
    res["wgrad"] = rel_err(gout, gref)
    # wgrad with fused SGD on the small layer (M = 10)
    m0 = ws.master.clone()
    a = native.GemmArgs(variant=native.GEMM_WGRAD, C=c, M=10, N=200, K=B, a=op(ws.dlogits, 0, B * 16, 16, B, 16),
                        b=op(ws.h, 0, B * 200, 200, B, 200), lr=0.5, w_master=vp(ws.master, L.W2), w_stride=P, w_ld=200,
                        w_bf16=vp(ws.master_bf16, L.W2), wb_stride=P)
    native.client_gemm(a)
    torch.cuda.synchronize()
    exp = m0[:, L.W2:L.W2 + 2000] - 0.5 * torch.bmm(dlf.transpose(1, 2), ws.h.float()).reshape(c, -1)
    res["sgd"] = rel_err(ws.master[:, L.W2:L.W2 + 2000], exp)
    res["sgd_bf16"] = rel_err(ws.master_bf16[:, L.W2:L.W2 + 2000].float(), exp.to(torch.bfloat16).float())
    untouched = torch.cat([ws.master[:, :L.W2], ws.master[:, L.W2 + 2000:]], 1) - torch.cat([m0[:, :L.W2], m0[:, L.W2 + 2000:]], 1)
    res["untouched"] = float(untouched.abs().max())
    res["err"] = max(res["fwd"], res["dgrad"], res["db"], res["wgrad"], res["sgd"], res["sgd_bf16"], res["untouched"])
    res["tol"] = 1e-2
    ws.close()
    return res


def check_mlp_step(c: int) -> Dict:
    """six-launch MLP local step vs autograd SGD on the same bf16-rounded operands"""
    ws, spec = _mlp_ws(c)
    L = ws.L
    g0 = ws.L.pack(spec.init(7)).cuda()
    ws.set_global(g0)
    ws.broadcast()
    lr = 0.1
    ws.local_step(0, lr=lr)
    torch.cuda.synchronize()
    x = ws.x[:, :B].float()
    y = ws.labels[:, :B].long()
    w = g0[None, :].repeat(c, 1).clone().requires_grad_(True)
    wq = w.detach().to(torch.bfloat16).float() + (w - w.detach())   # bf16-rounded operand values, identity gradient
    W1 = wq[:, L.W1:L.W1 + 200 * 784].reshape(c, 200, 784)
    W2 = wq[:, L.W2:L.W2 + 2000].reshape(c, 10, 200)
    h = torch.relu(torch.bmm(x, W1.transpose(1, 2)) + w[:, None, L.B1:L.B1 + 200])
    hq = h.detach().to(torch.bfloat16).float() + (h - h.detach())
    lg = torch.bmm(hq, W2.transpose(1, 2)) + w[:, None, L.B2:L.B2 + 10]
    loss = F.cross_entropy(lg.reshape(c * B, 10), y.reshape(-1), reduction="none").reshape(c, B).mean(1)
    (g,) = torch.autograd.grad(loss.sum(), w)
    new = g0[None, :] - lr * g
    d_k, d_r = ws.master - g0[None, :], new - g0[None, :]
    res = {"loss": rel_err(ws.loss_sum, loss.detach()), "delta": rel_err(d_k, d_r), "tol": 5e-2}
    for name, o, n in (("W1", L.W1, 200 * 784), ("b1", L.B1, 200), ("W2", L.W2, 2000), ("b2", L.B2, 10)):
        res[f"d_{name}"] = rel_err(d_k[:, o:o + n], d_r[:, o:o + n])
        res[f"n_{name}"] = float(d_r[:, o:o + n].norm())
    res["err"] = max(res["loss"], res["delta"])
    ws.close()
    return res


CHECKS: Dict[str, Callable[[int], Dict]] = {
    "ingest": check_ingest, "conv1_fwd": check_conv1_fwd, "conv2_fwd": check_conv2_fwd, "head": check_head,
    "conv2_dgrad": check_conv2_dgrad, "conv2_wgrad": check_conv2_wgrad, "conv1_wgrad": check_conv1_wgrad,
    "bcast": check_bcast, "reduce": check_reduce, "mask": check_mask, "local_step": check_local_step,
    "gemm": check_gemm, "mlp_step": check_mlp_step,
}


# kernel variants under evaluation: run only when named explicitly with --only
EXPERIMENTAL: Dict[str, Callable[[int], Dict]] = {"conv2_wgrad_wide": check_conv2_wgrad_wide,
                                                    "conv2_dgrad_pair": check_conv2_dgrad_pair,
                                                    "conv2_fwd_shift": check_conv2_fwd_shift,
                                                    "conv1_window": check_conv1_window,
                                                    "zero_round": check_zero_round}


def run_all(clients: List[int], only: List[str] | None = None) -> List[Dict]:
    out = []
    for name, fn in {**CHECKS, **EXPERIMENTAL}.items():
        if (only and name not in only) or (not only and name in EXPERIMENTAL):
            continue
        for c in clients:
            rec: Dict = {"check": name, "clients": c}
            try:
                torch.manual_seed(1234 + c)
                r = fn(c)
                rec.update(r)
                rec["ok"] = bool(r["err"] <= r["tol"])
            except Exception as e:  # noqa: BLE001 - report and continue
                rec["ok"] = False
                rec["exception"] = f"{type(e).__name__}: {e}"
                rec["trace"] = traceback.format_exc(limit=4)
                try:
                    torch.cuda.synchronize()
                except Exception as e2:  # sticky CUDA error: nothing else can run
                    rec["fatal"] = str(e2)
                    out.append(rec)
                    print(json.dumps(rec), flush=True)
                    return out
            out.append(rec)
            print(json.dumps({k: v for k, v in rec.items() if k != "trace"}), flush=True)
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--clients", type=int, nargs="*", default=[3, 160])
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--json", default=None)
    a = ap.parse_args(argv)
    res = run_all(a.clients, a.only)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)
    bad = [r for r in res if not r["ok"]]
    print(f"selfcheck: {len(res) - len(bad)}/{len(res)} ok")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
