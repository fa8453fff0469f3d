"""Counter-based RNG (Philox4x32-10) — the CPU twin of ``deviceflow_mask_kernel``.

The reference draws ``random.random()`` per message inside the dispatcher thread
(ols_core/deviceflow/non_grpc/dispatcher.py:103-109, strategy.py:290-292), which makes a trace irreproducible.
Here every (seed, round, client) triple owns a fixed random stream, so the CPU simulation, the CUDA kernel and
any number of ranks agree bit-for-bit on who participates, who is dropped and who is late.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox.  Counters/keys are uint32 arrays (broadcastable); returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32).copy()
    c1 = np.asarray(c1, dtype=np.uint32).copy()
    c2 = np.asarray(c2, dtype=np.uint32).copy()
    c3 = np.asarray(c3, dtype=np.uint32).copy()
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n2 = hi0 ^ c3 ^ k1
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def uniform24(r: np.ndarray) -> np.ndarray:
    """uint32 -> float32 in [0,1) using the top 24 bits (exactly what the kernel does)."""
    return ((r >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def round_uniforms(num_clients: int, client_base: int, seed: int, round_idx: int, stream: int = 0):
    ids = (np.arange(num_clients, dtype=np.uint64) + np.uint64(client_base)).astype(np.uint32)
    r = philox4x32_10(np.uint32(round_idx), ids, np.uint32(stream), np.uint32(0), seed & 0xFFFFFFFF,
                      (seed >> 32) & 0xFFFFFFFF)
    return [uniform24(x) for x in r]


def round_mask(num_clients: int, client_base: int, seed: int, round_idx: int, p_online: Optional[np.ndarray],
               drop_probability: float, arrive_tick: Optional[np.ndarray], max_tick: int, n_samples: np.ndarray,
               local_steps: Optional[np.ndarray], max_steps: int) -> Dict[str, np.ndarray]:
    """Per-round participation / drop / lateness, identical to the CUDA kernel."""
    u = round_uniforms(num_clients, client_base, seed, round_idx)
    pon = np.ones(num_clients, np.float32) if p_online is None else np.asarray(p_online, np.float32)
    part = u[0] < pon
    dp = np.float32(drop_probability)
    if dp >= 1.0:
        dropped = part.copy()
    elif dp > 0.0:
        dropped = part & ~(u[1] > dp)
    else:
        dropped = np.zeros(num_clients, bool)
    late = np.zeros(num_clients, bool)
    if arrive_tick is not None:
        t = np.asarray(arrive_tick)
        late = part & ~dropped & ((t < 0) | (t > max_tick))
    arrived = part & ~dropped & ~late
    steps = np.full(num_clients, max_steps) if local_steps is None else np.minimum(np.asarray(local_steps), max_steps)
    step_active = (np.arange(max_steps)[:, None] < steps[None, :]) & part[None, :]
    return {
        "participate": part.astype(np.uint8),
        "step_active": step_active.astype(np.uint8),
        "agg_weight": np.where(arrived, np.asarray(n_samples, np.float32), np.float32(0)).astype(np.float32),
        "counters": np.array([part.sum(), dropped.sum(), late.sum(), arrived.sum()], np.int32),
        "dropped": dropped, "late": late, "arrived": arrived,
    }
