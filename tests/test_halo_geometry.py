"""CPU model of the address algebra of the halo-box convolution kernels (ops/csrc/bgemm_tc.cuh, A_HALO / A_HALOW).

The kernels never materialise an im2col matrix: one (16+2) x (8+2)-pixel box per tile sits in shared memory as 180 rows of
128 bytes (64 channels), and every operand the tensor core reads is a *view* of that one buffer described by a start row and
two strides.  This test rebuilds those views with plain index arithmetic — exactly the numbers the kernel puts into its
descriptors — and checks that the resulting GEMMs are the convolution / its weight gradient:

* forward / data gradient: tile row r = (ty, tx) = (r >> 3, r & 7) reads box row  start + (r >> 3) * PITCH + (r & 7)  with
  start = (dy + 1) * PITCH + (dx + 1) for tap (dy, dx): sixteen 8-row groups at a group stride of PITCH = 10 rows;
* weight gradient: M = 128 = two taps x 64 input channels; the second tap of pair mt is reached through the leading-
  dimension offset lbo[mt] (rows), K index k = (ty, tx) walks the same 8-row groups, K16 step = two image rows.
"""
import numpy as np
import torch
import torch.nn.functional as F

PITCH, BOX_H, BOX_W, TH, TW = 10, 18, 10, 16, 8


def load_box(x, ty0, tx0):
    """TMA box load with out-of-bounds zero fill: rows = box pixels (row-major, PITCH per image row), columns = channels"""
    H, W, C = x.shape
    box = np.zeros((BOX_H * BOX_W, C), dtype=x.dtype)
    for by in range(BOX_H):
        for bx in range(BOX_W):
            y, xx = ty0 - 1 + by, tx0 - 1 + bx
            if 0 <= y < H and 0 <= xx < W:
                box[by * BOX_W + bx] = x[y, xx]
    return box


def a_view(box, start_row):
    """K-major A operand: 128 rows = sixteen 8-row groups, group stride PITCH rows, starting at ``start_row``"""
    rows = [start_row + (r >> 3) * PITCH + (r & 7) for r in range(128)]
    return box[rows]


def test_forward_and_data_gradient_are_shifted_views_of_one_box():
    rng = np.random.default_rng(0)
    H = W = 32
    x = rng.standard_normal((H, W, 64)).astype(np.float32)
    w = rng.standard_normal((64, 3, 3, 64)).astype(np.float32)            # [cout][kh][kw][cin], the packed layout
    ref = F.conv2d(torch.from_numpy(x).permute(2, 0, 1)[None], torch.from_numpy(w).permute(0, 3, 1, 2), padding=1)[0].permute(1, 2, 0).numpy()
    out = np.zeros((H, W, 64), dtype=np.float32)
    for ty0 in range(0, H, TH):
        for tx0 in range(0, W, TW):
            box = load_box(x, ty0, tx0)
            acc = np.zeros((128, 64), dtype=np.float32)
            for kh in range(3):
                for kw in range(3):
                    dy, dx = kh - 1, kw - 1
                    start = (dy + 1) * PITCH + (dx + 1)                       # aoff / 128 in the kernel
                    acc += a_view(box, start) @ w[:, kh, kw, :].T            # B = [cout][cin] block, K-major
            for r in range(128):
                out[ty0 + (r >> 3), tx0 + (r & 7)] = acc[r]
    assert np.allclose(out, ref, atol=2e-3)
    # data gradient = the same kernel on dY with flipped taps and the weight block read MN-major ([cout][cin] as K x N)
    dy_t = rng.standard_normal((H, W, 64)).astype(np.float32)
    xs = torch.from_numpy(x).permute(2, 0, 1)[None].requires_grad_(True)
    F.conv2d(xs, torch.from_numpy(w).permute(0, 3, 1, 2), padding=1).backward(torch.from_numpy(dy_t).permute(2, 0, 1)[None])
    ref_dx = xs.grad[0].permute(1, 2, 0).numpy()
    dx_out = np.zeros_like(ref_dx)
    for ty0 in range(0, H, TH):
        for tx0 in range(0, W, TW):
            box = load_box(dy_t, ty0, tx0)
            acc = np.zeros((128, 64), dtype=np.float32)
            for kh in range(3):
                for kw in range(3):
                    dyy, dxx = 1 - kh, 1 - kw                                 # conv_dgrad: taps.append((pad - kh, pad - kw, ...))
                    acc += a_view(box, (dyy + 1) * PITCH + (dxx + 1)) @ w[:, kh, kw, :]
            for r in range(128):
                dx_out[ty0 + (r >> 3), tx0 + (r & 7)] = acc[r]
    assert np.allclose(dx_out, ref_dx, atol=2e-3)


def test_weight_gradient_tap_pairs_through_the_leading_dimension_offset():
    rng = np.random.default_rng(1)
    H = W = 32
    x = rng.standard_normal((H, W, 64)).astype(np.float32)
    dy = rng.standard_normal((H, W, 64)).astype(np.float32)
    wz = torch.zeros(64, 64, 3, 3, requires_grad=True)
    F.conv2d(torch.from_numpy(x).permute(2, 0, 1)[None], wz, padding=1).backward(torch.from_numpy(dy).permute(2, 0, 1)[None])
    ref = wz.grad.permute(0, 2, 3, 1).numpy()                                # [cout][kh][kw][cin]
    # the kernel's five M tiles: start row o0 and leading-dimension offset (rows) to the second tap of the pair
    desc = []
    for mt in range(5):
        t0, t1 = 2 * mt, min(2 * mt + 1, 8)
        o0, o1 = (t0 // 3) * PITCH + t0 % 3, (t1 // 3) * PITCH + t1 % 3
        desc.append((o0, o1 - o0 if o1 > o0 else 1))
    assert desc == [(0, 1), (2, 8), (11, 1), (20, 1), (22, 1)]                # +128 B, +1024 B, ... in bytes / 128
    acc = np.zeros((5, 128, 64), dtype=np.float32)                           # [M tile][(tap slot, cin)][cout]
    for ty0 in range(0, H, TH):
        for tx0 in range(0, W, TW):
            box = load_box(x, ty0, tx0)
            b = np.stack([dy[ty0 + (k >> 3), tx0 + (k & 7)] for k in range(128)])          # dY tile, K x N
            for mt, (o0, lbo) in enumerate(desc):
                for ks in range(8):                                          # K16 = two image rows of the tile
                    krows = [o0 + ks * 2 * PITCH + (k >> 3) * PITCH + (k & 7) for k in range(16)]
                    a0 = box[krows]                                          # K x 64 (MN-major block 0: first tap)
                    a1 = box[[min(r + lbo, BOX_H * BOX_W - 1) for r in krows]]               # block 1: + leading-dim offset
                    a = np.concatenate([a0, a1], axis=1)                     # K x 128
                    acc[mt] += a.T @ b[ks * 16:(ks + 1) * 16]
    got = np.zeros_like(ref)
    for mt in range(5):
        for slot in range(2):
            tap = 2 * mt + slot
            if tap < 9:
                got[:, tap // 3, tap % 3, :] = acc[mt, slot * 64:(slot + 1) * 64, :].T
    assert np.allclose(got, ref, atol=5e-2, rtol=1e-4)
