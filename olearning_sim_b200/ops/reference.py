"""Plain-PyTorch fp32 oracles for the CUDA kernels and the CPU execution path of the engine.

Everything here works on the *packed* parameter layout (``ops/layout.py``) and is batched over clients with
grouped convolutions / ``bmm`` so the CPU path (``gloo``, config #1) and the numerics tests share one
implementation.  Nothing in this file is on the GPU product path.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .layout import CNN2


def normalize_u8(x_u8: torch.Tensor) -> torch.Tensor:
    """uint8 NHWC -> float NHWC in [-1, 1] (same arithmetic as ingest_u8_kernel)."""
    return (x_u8.float() * (1.0 / 255.0) - 0.5) * 2.0


def ingest_ref(x_u8: torch.Tensor, preswizzle=False) -> torch.Tensor:
    """uint8 [N,32,32,3] -> float [N,32,32,16] kw-unrolled conv1 operand.  ``preswizzle``: the device layouts of
    ingest_u8_kernel — 1/True: pixels with (w & 4) store their two 8-channel halves swapped; 2: each row additionally
    split into its two column-parity planes ([row][w & 1][w >> 1]), halves swapped where (plane index & 4)."""
    x = normalize_u8(x_u8)                        # N,H,W,3
    n = x.shape[0]
    xp = F.pad(x, (0, 0, 2, 2))                   # pad W by 2 on both sides
    out = torch.zeros(n, 32, 32, 16, dtype=torch.float32, device=x.device)
    for kw in range(5):
        out[..., kw * 3:kw * 3 + 3] = xp[:, :, kw:kw + 32, :]
    if int(preswizzle) == 2:
        planes = torch.stack([out[:, :, 0::2], out[:, :, 1::2]], dim=2)           # N, H, 2, 16, 16ch
        sw = (torch.arange(16, device=x.device) & 4).bool()
        planes[:, :, :, sw] = torch.cat([planes[:, :, :, sw, 8:], planes[:, :, :, sw, :8]], dim=-1)
        return planes.reshape(n, 32, 32, 16)
    if preswizzle:
        sw = (torch.arange(32, device=x.device) & 4).bool()
        out[:, :, sw] = torch.cat([out[:, :, sw, 8:], out[:, :, sw, :8]], dim=-1)
    return out


def _bf16(t: torch.Tensor, emulate: bool) -> torch.Tensor:
    return t.to(torch.bfloat16).float() if emulate else t


def cnn2_views(flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Views of packed rows [C,P] as torch conv weights (per client)."""
    c = flat.shape[0]
    L = CNN2
    w1 = flat[:, L.W1:L.W1 + L.W1_N].reshape(c, 32, 5, 16)[..., :15].reshape(c, 32, 5, 5, 3).permute(0, 1, 4, 2, 3)
    w2 = flat[:, L.W2:L.W2 + L.W2_N].reshape(c, 64, 5, 5, 32).permute(0, 1, 4, 2, 3)
    wfc = flat[:, L.WFC:L.WFC + L.WFC_N].reshape(c, 10, 4096)
    return {"w1": w1, "b1": flat[:, L.B1:L.B1 + 32], "w2": w2, "b2": flat[:, L.B2:L.B2 + 64], "wfc": wfc,
            "bfc": flat[:, L.BFC:L.BFC + 10]}


def cnn2_forward(flat: torch.Tensor, x_nhwc: torch.Tensor, emulate_bf16: bool = False):
    """Batched-over-clients forward.  flat [C,P], x [C,B,32,32,3] (already normalised floats).
    Returns logits [C,B,10] and intermediates (NHWC)."""
    c, b = x_nhwc.shape[:2]
    v = cnn2_views(flat)
    x = _bf16(x_nhwc, emulate_bf16).permute(0, 1, 4, 2, 3)          # C,B,3,32,32
    # grouped conv: fold clients into channel groups
    xg = x.permute(1, 0, 2, 3, 4).reshape(b, c * 3, 32, 32)
    w1 = _bf16(v["w1"], emulate_bf16).reshape(c * 32, 3, 5, 5)
    a1 = F.conv2d(xg, w1, v["b1"].reshape(-1), padding=2, groups=c)  # B, C*32, 32,32
    a1 = F.relu(a1)
    p1 = F.max_pool2d(a1, 2)
    p1q = _bf16(p1, emulate_bf16)
    w2 = _bf16(v["w2"], emulate_bf16).reshape(c * 64, 32, 5, 5)
    a2 = F.relu(F.conv2d(p1q, w2, v["b2"].reshape(-1), padding=2, groups=c))
    p2 = F.max_pool2d(a2, 2)                                          # B, C*64, 8, 8
    p2q = _bf16(p2, emulate_bf16)
    feat = p2q.reshape(b, c, 64, 8, 8).permute(1, 0, 3, 4, 2).reshape(c, b, 4096)   # NHWC flatten
    logits = torch.bmm(feat, v["wfc"].transpose(1, 2)) + v["bfc"][:, None, :]
    inter = {"p1": p1.reshape(b, c, 32, 16, 16).permute(1, 0, 3, 4, 2), "p2": p2.reshape(b, c, 64, 8, 8).permute(1, 0, 3, 4, 2)}
    return logits, inter


def cnn2_loss(flat: torch.Tensor, x_nhwc: torch.Tensor, labels: torch.Tensor, emulate_bf16: bool = False):
    """Mean cross-entropy per client -> [C]."""
    logits, _ = cnn2_forward(flat, x_nhwc, emulate_bf16)
    c, b = labels.shape
    loss = F.cross_entropy(logits.reshape(c * b, 10), labels.reshape(-1).long(), reduction="none").reshape(c, b)
    return loss.mean(dim=1), logits


def cnn2_local_step(flat: torch.Tensor, x_nhwc: torch.Tensor, labels: torch.Tensor, lr: float, mu: float = 0.0,
                    global_w: Optional[torch.Tensor] = None, active: Optional[torch.Tensor] = None,
                    emulate_bf16: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """One SGD/FedProx step for every client.  Returns (new_flat, loss[C])."""
    w = flat.detach().clone().requires_grad_(True)
    loss, _ = cnn2_loss(w, x_nhwc, labels, emulate_bf16)
    (g,) = torch.autograd.grad(loss.sum(), w)
    g = g * CNN2.pad_mask().to(g.device)
    if mu and global_w is not None:
        g = g + mu * (flat - global_w[None, :]) * CNN2.pad_mask().to(g.device)
    new = flat - lr * g
    if active is not None:
        a = active.bool()[:, None]
        new = torch.where(a, new, flat)
    return new.detach(), loss.detach()


# --------------------------------------------------------------------------------------------------
# generic FL arithmetic (model-agnostic, used by the CPU engine and as oracle for K1/K2)
# --------------------------------------------------------------------------------------------------
def fedavg_delta(master: torch.Tensor, global_w: torch.Tensor, weight: torch.Tensor) -> Tuple[torch.Tensor, float]:
    """sum_c a_c (w_c - w_g), sum_c a_c"""
    # rows with zero weight are skipped outright (a poisoned replica is NaN, and 0 * NaN would poison the sum)
    delta = torch.where(weight[:, None] > 0, master - global_w[None, :], torch.zeros_like(master))
    d = (weight[:, None] * delta).sum(dim=0)
    return d, float(weight.sum())


def server_update(global_w: torch.Tensor, delta_sum: torch.Tensor, wsum: float, opt: str, state: dict,
                  server_lr: float = 1.0, beta1: float = 0.9, beta2: float = 0.99, eps: float = 1e-3) -> torch.Tensor:
    """FedAvg / FedAvgM / FedAdam server step on the averaged client delta (same formulas as fedavg_finish_kernel)."""
    if wsum <= 0:
        return global_w
    d = delta_sum / wsum
    if opt in ("avg", "sgd", 0):
        return global_w + server_lr * d
    if opt in ("momentum", 1):
        m = state.get("m", torch.zeros_like(d))
        m = beta1 * m + d
        state["m"] = m
        return global_w + server_lr * m
    if opt in ("adam", 2):
        step = state.get("step", 0) + 1
        m = beta1 * state.get("m", torch.zeros_like(d)) + (1 - beta1) * d
        v = beta2 * state.get("v", torch.zeros_like(d)) + (1 - beta2) * d * d
        state.update(m=m, v=v, step=step)
        mh = m / (1 - beta1 ** step)
        vh = v / (1 - beta2 ** step)
        return global_w + server_lr * mh / (vh.sqrt() + eps)
    raise ValueError(f"unknown server optimiser {opt!r}")
