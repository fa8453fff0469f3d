"""Packed parameter layouts shared between python and the CUDA kernels (``csrc/fl_kernels.cuh``).

A simulated client is one row of a packed ``[C, P]`` fp32 tensor.  The order *inside* a row is chosen for the
tensor-core kernels (tap-major, channels-last), not for ``torch.nn``; ``pack``/``unpack`` convert to and from
ordinary ``state_dict`` tensors so checkpoints stay interoperable.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch


@dataclass(frozen=True)
class _Cnn2Layout:
    W1: int = 0          # [32][5][16]  cout, kh, kw*3+c (15 used + 1 pad)
    B1: int = 2560       # [32]
    W2: int = 2592       # [64][25][32] cout, kh*5+kw, cin
    B2: int = 53792      # [64]
    WFC: int = 53856     # [10][4096]   class, (ph*8+pw)*64+ch
    BFC: int = 94816     # [10] (+6 pad)
    P: int = 94832
    W1_N: int = 2560
    W2_N: int = 51200
    WFC_N: int = 40960
    BATCH: int = 32
    NUM_CLASSES: int = 10
    REAL_PARAMS: int = 32 * 75 + 32 + 64 * 800 + 64 + 10 * 4096 + 10

    # ---- conversions ---------------------------------------------------------------------------
    def pack(self, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
        """state_dict {conv1.weight [32,3,5,5], conv1.bias, conv2.weight [64,32,5,5], conv2.bias,
        fc.weight [10,4096] (NCHW flatten), fc.bias} -> packed fp32 [P]."""
        out = torch.zeros(self.P, dtype=torch.float32)
        w1 = sd["conv1.weight"].float()                      # [cout, c, kh, kw]
        w1p = torch.zeros(32, 5, 16)
        w1p[:, :, :15] = w1.permute(0, 2, 3, 1).reshape(32, 5, 15)   # (cout, kh, kw, c) -> kw*3+c
        out[self.W1:self.W1 + self.W1_N] = w1p.reshape(-1)
        out[self.B1:self.B1 + 32] = sd["conv1.bias"].float()
        w2 = sd["conv2.weight"].float()                      # [cout, cin, kh, kw]
        out[self.W2:self.W2 + self.W2_N] = w2.permute(0, 2, 3, 1).reshape(64, 25, 32).reshape(-1)
        out[self.B2:self.B2 + 64] = sd["conv2.bias"].float()
        wfc = sd["fc.weight"].float().reshape(10, 64, 8, 8)  # NCHW flatten: ch, ph, pw
        out[self.WFC:self.WFC + self.WFC_N] = wfc.permute(0, 2, 3, 1).reshape(10, 4096).reshape(-1)
        out[self.BFC:self.BFC + 10] = sd["fc.bias"].float()
        return out

    def unpack(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        flat = flat.detach().float().cpu()
        w1p = flat[self.W1:self.W1 + self.W1_N].reshape(32, 5, 16)[:, :, :15].reshape(32, 5, 5, 3)
        w2p = flat[self.W2:self.W2 + self.W2_N].reshape(64, 5, 5, 32)
        wfc = flat[self.WFC:self.WFC + self.WFC_N].reshape(10, 8, 8, 64)
        return {
            "conv1.weight": w1p.permute(0, 3, 1, 2).contiguous(),
            "conv1.bias": flat[self.B1:self.B1 + 32].clone(),
            "conv2.weight": w2p.permute(0, 3, 1, 2).contiguous(),
            "conv2.bias": flat[self.B2:self.B2 + 64].clone(),
            "fc.weight": wfc.permute(0, 3, 1, 2).reshape(10, 4096).contiguous(),
            "fc.bias": flat[self.BFC:self.BFC + 10].clone(),
        }

    def pad_mask(self) -> torch.Tensor:
        """1 for real parameters, 0 for layout padding (must stay zero)."""
        m = torch.ones(self.P)
        w1 = torch.ones(32, 5, 16)
        w1[:, :, 15] = 0
        m[self.W1:self.W1 + self.W1_N] = w1.reshape(-1)
        m[self.BFC + 10:self.P] = 0
        return m

    def init(self, seed: int = 0) -> torch.Tensor:
        """Kaiming-uniform random init (torch.nn defaults) in packed form."""
        g = torch.Generator().manual_seed(seed)

        def u(shape, fan_in):
            bound = 1.0 / (fan_in ** 0.5)
            return (torch.rand(shape, generator=g) * 2 - 1) * bound

        sd = {
            "conv1.weight": u((32, 3, 5, 5), 75), "conv1.bias": u((32,), 75),
            "conv2.weight": u((64, 32, 5, 5), 800), "conv2.bias": u((64,), 800),
            "fc.weight": u((10, 4096), 4096), "fc.bias": u((10,), 4096),
        }
        return self.pack(sd)


CNN2 = _Cnn2Layout()
