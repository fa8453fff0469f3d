"""GPU numerics of the batched tcgen05 GEMM / implicit-conv family and the ResNet-18 / BERT-mini native steps.

Every kernel is compared with a plain PyTorch fp32 computation of the same op on the same bf16-rounded inputs
(SURVEY §4 "unit, kernel").  The engine-level tests run a full federated round of the native backend against the
generic torch.func backend on identical seeds / data.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-20))


@pytest.fixture(scope="module")
def bgm():
    from olearning_sim_b200.ops import bgemm, native
    native.require()
    return bgemm


def rb(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(torch.bfloat16)


# ---------------------------------------------------------------------------------------------------------
# fully-connected layers
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,K,N", [(256, 256, 768), (384, 1024, 256), (128, 256, 256)])
def test_linear_fwd_dgrad_wgrad(bgm, T, K, N):
    torch.manual_seed(0)
    C = 3
    P = N * K + 64 + N
    wb = rb(C, N * K + 64, scale=0.05)
    master = torch.zeros(C, P, device=DEV)
    master[:, :N * K + 64] = wb.float()
    master[:, N * K + 64:] = torch.randn(C, N, device=DEV) * 0.1
    x = rb(C, T, K)
    W = wb[:, 8:8 + N * K].float().reshape(C, N, K)
    bias = master[:, N * K + 64:]
    out = torch.empty(C, T, N, dtype=torch.bfloat16, device=DEV)
    addend = rb(C, T, N)
    bgm.linear_fwd(C, T, K, N, x, wb, 8, wb.shape[1], out, master, N * K + 64, P, addend=addend)
    ref = torch.einsum("ctk,cnk->ctn", x.float(), W) + bias[:, None, :] + addend.float()
    assert rel(out, ref) < 6e-3
    # GELU + saved pre-activation
    pre = torch.empty_like(out)
    bgm.linear_fwd(C, T, K, N, x, wb, 8, wb.shape[1], out, master, N * K + 64, P, act=bgm.ACT_GELU, aux_out=pre)
    ref_pre = torch.einsum("ctk,cnk->ctn", x.float(), W) + bias[:, None, :]
    assert rel(pre, ref_pre) < 6e-3 and rel(out, F.gelu(ref_pre)) < 8e-3
    # data gradient (+ addend, + GELU')
    dy = rb(C, T, N, scale=0.1)
    dx = torch.empty(C, T, K, dtype=torch.bfloat16, device=DEV)
    add2 = rb(C, T, K, scale=0.1)
    bgm.linear_dgrad(C, T, N, K, dy, wb, 8, wb.shape[1], dx, addend=add2)
    ref = torch.einsum("ctn,cnk->ctk", dy.float(), W) + add2.float()
    assert rel(dx, ref) < 6e-3
    gp = rb(C, T, K)
    bgm.linear_dgrad(C, T, N, K, dy, wb, 8, wb.shape[1], dx, gelu_pre=gp)
    g = gp.float().requires_grad_(True)
    F.gelu(g).backward(torch.einsum("ctn,cnk->ctk", dy.float(), W))
    assert rel(dx, g.grad) < 8e-3
    # weight gradient: raw, then fused SGD + FedProx
    gout = torch.zeros(C, N, K, device=DEV)
    bgm.linear_wgrad(C, T, N, K, dy, x, master, wb, 8, P, wb.shape[1], 0.1, grad_out=gout)
    gref = torch.einsum("ctn,ctk->cnk", dy.float(), x.float())
    assert rel(gout, gref) < 3e-3
    gw = torch.randn(P, device=DEV) * 0.05
    m0 = master.clone()
    act = torch.tensor([1, 0, 1], dtype=torch.uint8, device=DEV)
    bgm.linear_wgrad(C, T, N, K, dy, x, master, wb, 8, P, wb.shape[1], 0.1, mu=0.3, global_w=gw, active=act)
    w0 = m0[:, 8:8 + N * K].reshape(C, N, K)
    want = w0 - 0.1 * (gref + 0.3 * (w0 - gw[8:8 + N * K].reshape(N, K)))
    got = master[:, 8:8 + N * K].reshape(C, N, K)
    assert rel(got[0] - w0[0], want[0] - w0[0]) < 4e-3 and rel(got[2] - w0[2], want[2] - w0[2]) < 4e-3
    assert torch.equal(got[1], w0[1])                                   # inactive client untouched
    assert rel(wb[:, 8:8 + N * K].float().reshape(C, N, K)[0], got[0]) < 5e-3
    assert torch.equal(master[:, :8], m0[:, :8]) and torch.equal(master[:, 8 + N * K:], m0[:, 8 + N * K:])


# ---------------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------------
def test_attention_fwd_bwd(bgm):
    torch.manual_seed(1)
    C, Bs, S, heads, H = 2, 3, 128, 4, 256
    CB = C * Bs
    qkv = rb(CB, S, 3 * H, scale=0.8)
    probs = torch.empty(CB * heads, S, S, dtype=torch.bfloat16, device=DEV)
    att = torch.empty(CB, S, H, dtype=torch.bfloat16, device=DEV)
    act = torch.ones(C, dtype=torch.uint8, device=DEV)
    bgm.attn_scores(CB, S, heads, qkv, H, probs, act, Bs * heads)
    bgm.attn_pv(CB, S, heads, probs, qkv, H, att, act, Bs * heads)
    q32 = qkv.float().requires_grad_(True)
    q, k, v = [q32[..., i * H:(i + 1) * H].reshape(CB, S, heads, 64).permute(0, 2, 1, 3) for i in range(3)]
    pr = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    o = (pr @ v).permute(0, 2, 1, 3).reshape(CB, S, H)
    assert rel(probs.reshape(CB, heads, S, S), pr) < 1e-2
    assert rel(att, o) < 1e-2
    datt = rb(CB, S, H, scale=0.1)
    o.backward(datt.float())
    ds = torch.empty_like(probs)
    dqkv = torch.empty_like(qkv)
    bgm.attn_bwd(CB, S, heads, H, qkv, probs, datt, ds, dqkv, act, Bs * heads)
    for i, nm in enumerate("qkv"):
        assert rel(dqkv[..., i * H:(i + 1) * H], q32.grad[..., i * H:(i + 1) * H]) < 2.5e-2, nm


# ---------------------------------------------------------------------------------------------------------
# convolutions: every (resolution, channels, kernel, stride) the ResNet-18 family uses
# ---------------------------------------------------------------------------------------------------------
CONVS = [(32, 32, 64, 1, 1), (32, 64, 64, 3, 1), (32, 64, 128, 3, 2), (32, 64, 128, 1, 2), (16, 128, 128, 3, 1),
         (16, 128, 256, 3, 2), (16, 128, 256, 1, 2), (8, 256, 256, 3, 1), (8, 256, 512, 3, 2), (8, 256, 512, 1, 2),
         (4, 512, 512, 3, 1)]


@pytest.mark.parametrize("H,Cin,Cout,k,s", CONVS)
def test_conv_fwd_dgrad_wgrad(bgm, H, Cin, Cout, k, s):
    torch.manual_seed(H * 7 + Cin + k + s)
    C, imgs, G = 2, 32, 2
    OH = H // s
    ktot = k * k * Cin
    Pb = Cout * ktot + 16
    wb = rb(C, Pb, scale=1.0 / math.sqrt(ktot))
    x = rb(C * imgs, H, H, Cin)
    raw = torch.empty(C * imgs, OH, OH, Cout, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(C * imgs, G, 2, device=DEV)
    bgm.conv_fwd(C, imgs, H, Cin, Cout, k, s, x, wb, 8, Pb, raw, stats, G)
    Wt = wb[:, 8:8 + Cout * ktot].float().reshape(C, Cout, k, k, Cin).permute(0, 1, 4, 2, 3).contiguous()
    xs = x.float().reshape(C, imgs, H, H, Cin).permute(0, 1, 4, 2, 3).contiguous().requires_grad_(True)
    Wt.requires_grad_(True)
    ys = torch.stack([F.conv2d(xs[c], Wt[c], stride=s, padding=k // 2) for c in range(C)])          # C,imgs,Cout,OH,OH
    ref = ys.permute(0, 1, 3, 4, 2).reshape(C * imgs, OH, OH, Cout)
    assert rel(raw, ref) < 6e-3
    rs = ref.reshape(C * imgs, OH * OH, G, Cout // G)
    st = torch.stack([rs.sum((1, 3)), (rs * rs).sum((1, 3))], -1)
    assert rel(stats, st) < 5e-3
    if Cin >= 64:
        dy = rb(C * imgs, OH, OH, Cout, scale=0.1)
        ys.backward(dy.float().reshape(C, imgs, OH, OH, Cout).permute(0, 1, 4, 2, 3))
        dx = torch.full((C * imgs, H, H, Cin), 7.0, dtype=torch.bfloat16, device=DEV)
        add = rb(C * imgs, H, H, Cin, scale=0.1)
        if k == 1 and s == 2:
            dx.copy_(add)                                   # 1x1 stride-2 accumulates into an existing gradient
            bgm.conv_dgrad(C, imgs, H, Cin, Cout, k, s, dy, wb, 8, Pb, dx, addend=dx)
        else:
            bgm.conv_dgrad(C, imgs, H, Cin, Cout, k, s, dy, wb, 8, Pb, dx, addend=add if s == 1 else None)
        want = xs.grad.permute(0, 1, 3, 4, 2).reshape(C * imgs, H, H, Cin)
        if s == 1 or k == 1:
            want = want + add.float()
        assert rel(dx, want) < 8e-3
        if k == 3:
            # ReLU-backward gate fused into the epilogue (after the addend); stride 2: every pixel parity class gated
            dxm = torch.full_like(dx, 7.0)
            bgm.conv_dgrad(C, imgs, H, Cin, Cout, k, s, dy, wb, 8, Pb, dxm, addend=add if s == 1 else None, relu_mask=x)
            assert rel(dxm, want * (x.float() > 0)) < 8e-3
    else:
        dy = rb(C * imgs, OH, OH, Cout, scale=0.1)
        ys.backward(dy.float().reshape(C, imgs, OH, OH, Cout).permute(0, 1, 4, 2, 3))
    gout = torch.zeros(C, Cout, ktot, device=DEV)
    P = Pb + 8
    master = torch.zeros(C, P, device=DEV)
    master[:, :Pb] = wb.float()
    bgm.conv_wgrad(C, imgs, H, Cin, Cout, k, s, dy, x, master, wb, 8, P, Pb, 0.1, grad_out=gout)
    gref = Wt.grad.permute(0, 1, 3, 4, 2).reshape(C, Cout, ktot)
    assert rel(gout, gref) < 4e-3
    m0 = master.clone()
    bgm.conv_wgrad(C, imgs, H, Cin, Cout, k, s, dy, x, master, wb, 8, P, Pb, 0.1)
    got = master[:, 8:8 + Cout * ktot].reshape(C, Cout, ktot) - m0[:, 8:8 + Cout * ktot].reshape(C, Cout, ktot)
    assert rel(got, -0.1 * gref) < 4e-3
    assert torch.equal(master[:, :8], m0[:, :8]) and torch.equal(master[:, 8 + Cout * ktot:], m0[:, 8 + Cout * ktot:])


# ---------------------------------------------------------------------------------------------------------
# GroupNorm / LayerNorm / heads
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("res,Ch", [(32, 64), (8, 256), (4, 512)])
def test_groupnorm_fwd_bwd(bgm, res, Ch):
    torch.manual_seed(res + Ch)
    C, imgs, G = 2, 32, 2
    P = 4 * Ch + 8
    master = torch.zeros(C, P, device=DEV)
    master[:, :Ch] = 1 + 0.2 * torch.randn(C, Ch, device=DEV)
    master[:, Ch:2 * Ch] = 0.2 * torch.randn(C, Ch, device=DEV)
    raw = rb(C * imgs, res, res, Ch, scale=2.0)
    resid = rb(C * imgs, res, res, Ch)
    r32 = raw.float().reshape(C * imgs, res * res, G, Ch // G)
    stats = torch.stack([r32.sum((1, 3)), (r32 * r32).sum((1, 3))], -1).contiguous()
    out = torch.empty_like(raw)
    p = bgm.GnApplyParams()
    p.C, p.imgs, p.HW, p.Ch, p.groups = C, imgs, res * res, Ch, G
    p.raw, p.stats, p.gamma_off, p.beta_off = bgm._ptr(raw), bgm._ptr(stats), 0, Ch
    p.res, p.master, p.P, p.relu, p.out = bgm._ptr(resid), bgm._ptr(master), P, 1, bgm._ptr(out)
    bgm.gn_apply(p)
    x = raw.float().reshape(C, imgs, res, res, Ch).permute(0, 1, 4, 2, 3).contiguous().requires_grad_(True)
    ga = master[:, :Ch].clone().requires_grad_(True)
    be = master[:, Ch:2 * Ch].clone().requires_grad_(True)
    y = torch.stack([F.group_norm(x[c], G, ga[c], be[c]) for c in range(C)])
    yo = F.relu(y + resid.float().reshape(C, imgs, res, res, Ch).permute(0, 1, 4, 2, 3))
    ref = yo.permute(0, 1, 3, 4, 2).reshape(C * imgs, res, res, Ch)
    assert rel(out, ref) < 6e-3
    dy = rb(C * imgs, res, res, Ch, scale=0.1)
    yo.backward(dy.float().reshape(C, imgs, res, res, Ch).permute(0, 1, 4, 2, 3))
    gsm = torch.zeros(C, 2 * Ch, device=DEV)
    sb = torch.zeros(C * imgs, G, 2, device=DEV)
    dx, dres = torch.empty_like(raw), torch.empty_like(raw)
    q = bgm.GnBwdParams()
    q.C, q.imgs, q.HW, q.Ch, q.groups = C, imgs, res * res, Ch, G
    q.dy, q.mask, q.raw, q.stats = bgm._ptr(dy), bgm._ptr(out), bgm._ptr(raw), bgm._ptr(stats)
    q.master, q.P, q.gamma_off = bgm._ptr(master), P, 0
    q.gsmall, q.gs_stride, q.g_gamma_off, q.g_beta_off = bgm._ptr(gsm), 2 * Ch, 0, Ch
    q.sb, q.dx, q.dres = bgm._ptr(sb), bgm._ptr(dx), bgm._ptr(dres)
    bgm.gn_bwd(q)
    want = x.grad.permute(0, 1, 3, 4, 2).reshape(C * imgs, res, res, Ch)
    assert rel(dx, want) < 1.2e-2
    assert rel(gsm[:, :Ch], ga.grad) < 1e-2 and rel(gsm[:, Ch:], be.grad) < 1e-2
    assert rel(dres, dy.float() * (ref > 0)) < 1e-6


def test_layernorm_colsum_slice_sgd(bgm):
    torch.manual_seed(5)
    C, T, H = 2, 512, 256
    P = 4 * H
    master = torch.zeros(C, P, device=DEV)
    master[:, :H] = 1 + 0.2 * torch.randn(C, H, device=DEV)
    master[:, H:2 * H] = 0.2 * torch.randn(C, H, device=DEV)
    pre = rb(C, T, H, scale=1.5)
    xhat, y, rstd = torch.empty_like(pre), torch.empty_like(pre), torch.empty(C, T, device=DEV)
    p = bgm.LnParams()
    p.C, p.T, p.pre, p.master, p.P, p.gamma_off, p.beta_off = C, T, bgm._ptr(pre), bgm._ptr(master), P, 0, H
    p.xhat, p.y, p.rstd = bgm._ptr(xhat), bgm._ptr(y), bgm._ptr(rstd)
    bgm.ln_fwd(p)
    x = pre.float().requires_grad_(True)
    ga = master[:, :H].clone().requires_grad_(True)
    be = master[:, H:2 * H].clone().requires_grad_(True)
    ref = torch.stack([F.layer_norm(x[c], (H,), ga[c], be[c]) for c in range(C)])
    assert rel(y, ref) < 6e-3
    dy = rb(C, T, H, scale=0.1)
    ref.backward(dy.float())
    gsm = torch.zeros(C, 3 * H, device=DEV)
    dx = torch.empty_like(pre)
    p.dy, p.dx = bgm._ptr(dy), bgm._ptr(dx)
    p.gsmall, p.gs_stride, p.g_gamma_off, p.g_beta_off, p.g_bias_off = bgm._ptr(gsm), 3 * H, 0, H, 2 * H
    bgm.ln_bwd(p)
    assert rel(dx, x.grad) < 1.2e-2
    assert rel(gsm[:, :H], ga.grad) < 1e-2 and rel(gsm[:, H:2 * H], be.grad) < 1e-2
    assert rel(gsm[:, 2 * H:], dx.float().sum(1)) < 1e-2
    # column sums
    cs = torch.zeros(C, 768 + 8, device=DEV)
    z = rb(C, T, 768)
    q = bgm.ColsumParams()
    q.C, q.T, q.N, q.chunk, q.x, q.gsmall, q.gs_stride, q.g_off = C, T, 768, 256, bgm._ptr(z), bgm._ptr(cs), 776, 8
    bgm.colsum(q)
    assert rel(cs[:, 8:], z.float().sum(1)) < 1e-3 and float(cs[:, :8].abs().max()) == 0
    # slice SGD with FedProx
    g = torch.randn(C, 100, device=DEV)
    gw = torch.randn(P, device=DEV)
    m0 = master.clone()
    act = torch.tensor([0, 1], dtype=torch.uint8, device=DEV)
    bgm.slice_sgd(master, P, 16, 100, g, 100, gw, act, C, 0.1, 0.5)
    want = m0[1, 16:116] - 0.1 * (g[1] + 0.5 * (m0[1, 16:116] - gw[16:116]))
    assert rel(master[1, 16:116], want) < 1e-6 and torch.equal(master[0], m0[0])


# ---------------------------------------------------------------------------------------------------------
# whole local steps / rounds: native backend vs the torch.func backend
# ---------------------------------------------------------------------------------------------------------
def _engine_pair(model, clients, steps, **kw):
    from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine
    base = dict(model=model, num_clients=clients, local_steps=steps, shard=64, lr=0.05, seed=3, **kw)
    a = FedEngine(FedConfig(backend="native", **base)); a.load_synthetic()
    b = FedEngine(FedConfig(backend="torch", device="cuda", **base)); b.load_synthetic()
    g0 = a.spec.init(11)
    a.set_global_model(a.ws.L.pack(g0))
    b.set_global_model(g0)
    return a, b, g0


def _compare_round(a, b, g0, tol):
    ra, rb_ = a.run_round(0), b.run_round(0)
    assert (ra.participants, ra.arrived, ra.client_steps) == (rb_.participants, rb_.arrived, rb_.client_steps)
    assert abs(ra.loss - rb_.loss) < 2e-2 * max(1.0, abs(rb_.loss)), (ra.loss, rb_.loss)
    ga, gb = a.spec.unflatten(a.spec_flat_global().cpu()), a.spec.unflatten(b.global_model().cpu())
    g0d = a.spec.unflatten(g0)
    worst = {}
    for k in ga:
        da, db = ga[k] - g0d[k], gb[k] - g0d[k]
        worst[k] = float((da - db).norm() / (db.norm() + 1e-12))
    bad = {k: v for k, v in worst.items() if v > tol}
    total = rel(a.spec_flat_global().cpu() - g0, b.global_model().cpu() - g0)
    print("round delta rel err total", total, "worst", sorted(worst.items(), key=lambda kv: -kv[1])[:5])
    assert total < tol and not bad, (total, bad)


def test_resnet18_native_round_vs_torch():
    """20 convolution layers with bf16 activations: the deviation from the fp32 library grows towards the stem.  The
    bound is therefore relative to what the LIBRARY shows at the same precision (bf16 autocast vs fp32)."""
    from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine
    a, b, g0 = _engine_pair("resnet18", 4, 2, hetero_steps=True)
    assert a.backend == "native"
    c = FedEngine(FedConfig(backend="torch", device="cuda", amp_bf16=True, model="resnet18", num_clients=4, local_steps=2,
                            shard=64, lr=0.05, seed=3, hetero_steps=True))
    c.load_synthetic(); c.set_global_model(g0)
    ra, rb_, rc = a.run_round(0), b.run_round(0), c.run_round(0)
    assert (ra.participants, ra.arrived, ra.client_steps) == (rb_.participants, rb_.arrived, rb_.client_steps)
    assert abs(ra.loss - rb_.loss) < 2e-2 * max(1.0, abs(rb_.loss)), (ra.loss, rb_.loss)
    da = a.spec_flat_global().cpu() - g0
    db, dc = b.global_model().cpu() - g0, c.global_model().cpu() - g0
    e_native, e_lib = rel(da, db), rel(dc, db)
    print("resnet18 round delta vs fp32 library: native", e_native, "bf16-autocast library", e_lib)
    assert e_native < 4e-2 and e_native < 1.5 * e_lib + 5e-3, (e_native, e_lib)
    ga, gb = a.spec.unflatten(a.spec_flat_global().cpu()), a.spec.unflatten(b.global_model().cpu())
    g0d = a.spec.unflatten(g0)
    for k in ga:
        r = rel(ga[k] - g0d[k], gb[k] - g0d[k])
        assert r < 0.12, (k, r)
    a.close()


def test_bert_mini_native_round_vs_torch():
    a, b, g0 = _engine_pair("bert_mini", 3, 2)
    assert a.backend == "native"
    _compare_round(a, b, g0, 6e-2)
    a.close()


def test_resnet18_fedprox_sampled_waves_track_the_torch_backend():
    """FedProx + client sampling + drops + two activation waves, three rounds: the native loss follows the fp32 library"""
    from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine
    kw = dict(model="resnet18", num_clients=6, local_steps=2, shard=64, algorithm="fedprox", mu=0.01, p_online=0.7,
              drop_probability=0.1, lr=0.01, seed=0)
    a = FedEngine(FedConfig(backend="native", wave=4, **kw)); a.load_synthetic()
    b = FedEngine(FedConfig(backend="torch", device="cuda", **kw)); b.load_synthetic()
    g0 = a.spec.init(0)
    a.set_global_model(a.ws.L.pack(g0)); b.set_global_model(g0)
    for r in range(3):
        ra, rb_ = a.run_round(r), b.run_round(r)
        assert (ra.participants, ra.dropped, ra.arrived, ra.failed) == (rb_.participants, rb_.dropped, rb_.arrived, 0)
        assert abs(ra.loss - rb_.loss) < 4e-2 * max(1.0, abs(rb_.loss)), (r, ra.loss, rb_.loss)
    assert rel(a.spec_flat_global().cpu() - g0, b.global_model().cpu() - g0) < 8e-2
    a.close()


# ---------------------------------------------------------------------------------------------------------
# DeviceFlow on the device: rate-curve integration (K8) and flow-dispatch ranking vs their host twins
# ---------------------------------------------------------------------------------------------------------
def test_rate_curve_kernel_is_bit_identical_to_the_host_integrator():
    import numpy as np
    from olearning_sim_b200.deviceflow.strategy import Strategy
    from olearning_sim_b200.ops import native
    for iv, dom, fn in (([0, 40], [0.0, 6.283], "math.sin(t)+0.3"), ([5, 17], [-1.0, 2.0], "t*t-0.5"), ([0, 3], [0, 1], "2*t")):
        y = Strategy.interval_samples(iv, dom, fn)
        yt = torch.from_numpy(np.ascontiguousarray(y)).to(DEV)
        out = torch.empty(y.shape[0], dtype=torch.float64, device=DEV)
        native.rate_curve_area(yt, out)
        _, want = Strategy.interval_areas(iv, dom, fn)
        assert out.cpu().tolist() == want                              # exact: same operations in the same order


@pytest.mark.parametrize("drop", [None, {"drop_probability": [0.3, 0.0]}, {"drop_amounts": [5, 2]}])
def test_flow_ticks_on_device_match_the_host_ranking(drop):
    import json as _json
    from olearning_sim_b200.engine.fed_engine import FedConfig, FedEngine
    si = {"use": True, "time_type": "relative", "intervals": [[0, 6], [8, 12]],
          "dispatch_rules": {"domains": [[0, 3.14], [0, 1]], "functions": ["math.sin(t)+0.2", "2*t"]}}
    if drop:
        si["drop_simulation"] = drop
    st = {"real_time_dispatch": {"use_strategy": False}, "flow_dispatch": {"use_strategy": True, "total_dispatch_amount": 150, "specific_interval": si}}
    eng = FedEngine(FedConfig(model="cnn2", num_clients=257, local_steps=1, shard=64, backend="native", p_online=0.85,
                              online_trace="0.7+0.3*math.cos(t)", deviceflow_strategy=_json.dumps(st), seed=5))
    eng.load_synthetic()
    for r in range(3):
        scale, _, ticks, max_tick = eng._round_flow(r)                  # host twin: philox + stable argsort + schedule_to_ticks
        assert eng._device_flow(r, scale) and eng._max_tick == max_tick
        torch.cuda.synchronize()
        assert eng.tick_dev.cpu().tolist() == ticks.tolist(), r
    # and the whole round agrees with the host path end to end
    a = eng.run_round(5)
    eng.device_flow_ticks = False
    eng2 = FedEngine(FedConfig(model="cnn2", num_clients=257, local_steps=1, shard=64, backend="native", p_online=0.85,
                               online_trace="0.7+0.3*math.cos(t)", deviceflow_strategy=_json.dumps(st), seed=5))
    eng2.device_flow_ticks = False
    eng2.load_synthetic()
    b = eng2.run_round(5)
    assert (a.participants, a.arrived, a.late, a.dropped) == (b.participants, b.arrived, b.late, b.dropped) and a.late > 0
    eng.close(); eng2.close()
