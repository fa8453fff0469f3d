#!/usr/bin/env python
"""Per-operation micro-benchmark of the batched tcgen05 GEMM / implicit-conv family at the shapes of the ResNet-18 and
BERT-mini local steps: CUDA-event time (steady state, inputs far larger than L2), algorithmic FLOPs and bytes, achieved
TFLOP/s / GB/s and the fraction of the roofline built from MEASURED_PEAKS.json.

    python tools/op_bench.py [--clients 64] [--reps 10] [--json out.json]
    ncu --set full ... python tools/op_bench.py --once          # every op exactly once after warm-up (profiler mode)
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from olearning_sim_b200.ops import bgemm as bgm  # noqa: E402
from olearning_sim_b200.ops import native  # noqa: E402

DEV = "cuda"


def rb(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(torch.bfloat16)


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, 1400.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clients", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--nowarm", action="store_true", help="profiler mode: every op exactly once, no warm-up")
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    native.require()
    C, B = a.clients, 32
    hbm, tf, src = peaks()
    ops = []

    def add(name, fn, flops, bytes_):
        if a.only and a.only not in name:
            return
        ops.append((name, fn, flops, bytes_))

    # ------------------------------------------------------------------ BERT-mini shapes (T = 4096 tokens / client)
    T, H, FF, S, heads = B * 128, 256, 1024, 128, 4
    P = H * 3 * H + 3 * H + 64
    wb = rb(C, FF * H + 64, scale=0.05)
    master = torch.zeros(C, FF * H + 64 + FF, device=DEV)
    x = rb(C, T, H)
    for (nm, K, N) in (("qkv", H, 3 * H), ("ff1", H, FF), ("ff2", FF, H)):
        out = torch.empty(C, T, N, dtype=torch.bfloat16, device=DEV)
        xin = x if K == H else rb(C, T, K)
        pre = torch.empty_like(out) if nm == "ff1" else None
        add(f"linear_fwd_{nm}" + ("_gelu" if pre is not None else ""),
            lambda xin=xin, K=K, N=N, out=out, pre=pre: bgm.linear_fwd(C, T, K, N, xin, wb, 0, wb.shape[1], out, master, FF * H + 64, master.shape[1],
                                                                      act=bgm.ACT_GELU if pre is not None else bgm.ACT_NONE, aux_out=pre),
            2 * C * T * K * N, C * (T * K + N * K + T * N * (2 if pre is not None else 1)) * 2)
        dy = rb(C, T, N, scale=0.1)
        dx = torch.empty(C, T, K, dtype=torch.bfloat16, device=DEV)
        gp = rb(C, T, K) if nm == "ff2" else None
        add(f"linear_dgrad_{nm}" + ("_gelubwd" if gp is not None else ""),
            lambda dy=dy, N=N, K=K, dx=dx, gp=gp: bgm.linear_dgrad(C, T, N, K, dy, wb, 0, wb.shape[1], dx, gelu_pre=gp),
            2 * C * T * K * N, C * (T * N + N * K + T * K * (2 if gp is not None else 1)) * 2)
        add(f"linear_wgrad_{nm}",
            lambda dy=dy, xin=xin, N=N, K=K: bgm.linear_wgrad(C, T, N, K, dy, xin, master, wb, 0, master.shape[1], wb.shape[1], 1e-4),
            2 * C * T * K * N, C * ((T * N + T * K) * 2 + N * K * 10))
    CB = C * B
    qkv = rb(CB, S, 3 * H, scale=0.5)
    probs = torch.empty(CB * heads, S, S, dtype=torch.bfloat16, device=DEV)
    att = torch.empty(CB, S, H, dtype=torch.bfloat16, device=DEV)
    add("attn_scores", lambda: bgm.attn_scores(CB, S, heads, qkv, H, probs, None, 1), 2 * CB * heads * S * S * 64,
        (CB * S * 2 * H + CB * heads * S * S) * 2)
    add("attn_pv", lambda: bgm.attn_pv(CB, S, heads, probs, qkv, H, att, None, 1), 2 * CB * heads * S * S * 64,
        (CB * S * 2 * H + CB * heads * S * S) * 2)
    datt, ds, dqkv = rb(CB, S, H, scale=0.1), torch.empty_like(probs), torch.empty_like(qkv)
    add("attn_bwd(4 launches)", lambda: bgm.attn_bwd(CB, S, heads, H, qkv, probs, datt, ds, dqkv, None, 1), 8 * CB * heads * S * S * 64,
        (CB * S * 8 * H + CB * heads * S * S * 5) * 2)
    # ------------------------------------------------------------------ ResNet-18 shapes
    for (Hh, Cin, Cout, k, s) in ((32, 64, 64, 3, 1), (16, 128, 128, 3, 1), (8, 256, 256, 3, 1), (4, 512, 512, 3, 1), (32, 64, 128, 3, 2)):
        OH = Hh // s
        ktot = k * k * Cin
        Pb = Cout * ktot
        wbc = rb(C, Pb, scale=1 / math.sqrt(ktot))
        mc = wbc.float()
        xx = rb(C * B, Hh, Hh, Cin)
        raw = torch.empty(C * B, OH, OH, Cout, dtype=torch.bfloat16, device=DEV)
        stats = torch.zeros(C * B, 2, 2, device=DEV)
        fl = 2 * C * B * OH * OH * Cout * ktot
        tag = f"{Hh}x{Hh}_{Cin}-{Cout}_s{s}"
        add(f"conv_fwd_{tag}", lambda Hh=Hh, Cin=Cin, Cout=Cout, k=k, s=s, xx=xx, wbc=wbc, Pb=Pb, raw=raw, stats=stats:
            bgm.conv_fwd(C, B, Hh, Cin, Cout, k, s, xx, wbc, 0, Pb, raw, stats, 2), fl, (xx.numel() + raw.numel() + C * Pb) * 2)
        dy = rb(C * B, OH, OH, Cout, scale=0.1)
        dx = torch.empty_like(xx)
        add(f"conv_dgrad_{tag}", lambda Hh=Hh, Cin=Cin, Cout=Cout, k=k, s=s, dy=dy, wbc=wbc, Pb=Pb, dx=dx:
            bgm.conv_dgrad(C, B, Hh, Cin, Cout, k, s, dy, wbc, 0, Pb, dx), fl, (xx.numel() + raw.numel() + C * Pb) * 2)
        add(f"conv_wgrad_{tag}", lambda Hh=Hh, Cin=Cin, Cout=Cout, k=k, s=s, dy=dy, xx=xx, mc=mc, wbc=wbc, Pb=Pb:
            bgm.conv_wgrad(C, B, Hh, Cin, Cout, k, s, dy, xx, mc, wbc, 0, Pb, Pb, 1e-4), fl, (xx.numel() + raw.numel()) * 2 + C * Pb * 10)
    # ------------------------------------------------------------------ run
    rows = []
    for name, fn, flops, nbytes in ops:
        if a.nowarm:
            fn(); torch.cuda.synchronize()
            continue
        fn(); fn()
        torch.cuda.synchronize()
        if a.once:
            fn(); torch.cuda.synchronize()
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.reps
        t_roof = max(flops / (tf * 1e12), nbytes / (hbm * 1e9)) * 1e6
        rows.append({"op": name, "us": round(us, 1), "tflops": round(flops / us / 1e6, 1), "gbs": round(nbytes / us / 1e3, 1),
                     "roofline_us": round(t_roof, 1), "frac": round(t_roof / us, 3),
                     "bound": "flops" if flops / (tf * 1e12) > nbytes / (hbm * 1e9) else "hbm"})
        print(f"{name:34s} {us:9.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {nbytes / us / 1e3:7.1f} GB/s  roofline {t_roof:8.1f} us "
              f"({rows[-1]['bound']})  frac {t_roof / us:5.2f}")
    if a.json:
        json.dump({"clients": C, "peaks": {"hbm_gbs": hbm, "bf16_tflops": tf, "source": src}, "rows": rows}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
