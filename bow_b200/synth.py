"""Host-side (numpy) mirror of the device generators in csrc/generate.cu — bench / test utility.

u(seed, col, i) = splitmix64(seed + 0x9E3779B97F4A7C15 * (i + (col << 40)))   

Any sub-range of the synthetic data sets can be regenerated here bit-for-bit, which is how parity
is checked on configurations that do not fit host memory and how the CPU reference arm of bench.py
gets the same inputs without touching the GPU.
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
T0_DEFAULT = 1_700_000_000_000_000_000
STEP_DEFAULT = 1_000_000_000


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + _GOLD
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def synth_u(seed: int, col: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        i = idx.astype(np.uint64) + (np.uint64(col) << np.uint64(40))
        return splitmix64(np.uint64(seed) + _GOLD * i)


def regular_time(row0: int, n: int, t0: int = T0_DEFAULT, step: int = STEP_DEFAULT) -> np.ndarray:
    return t0 + (row0 + np.arange(n, dtype=np.int64)) * step


def values(seed: int, col: int, row0: int, n: int, is_int: bool = False, null_mod: int = 0):
    """value column `col` (1-based, as on the device) -> (values, valid mask | None)"""
    idx = (row0 + np.arange(n, dtype=np.int64)).astype(np.uint64)
    u = synth_u(seed, col, idx)
    if is_int:
        v = (u & np.uint64(0xFFFFF)).astype(np.int64)
    else:
        v = (u >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    mask = None
    if null_mod:
        mask = (synth_u(seed + 1, col, idx) % np.uint64(null_mod)) != 0
    return v, mask


def regular_frame(row0: int, n: int, ncols: int = 1, seed: int = 42, t0: int = T0_DEFAULT, step: int = STEP_DEFAULT,
                  null_mask: int = 0, int_mask: int = 0, null_mod: int = 10):
    cols = [(regular_time(row0, n, t0, step), None)]
    for c in range(ncols):
        nulls = (null_mask >> c) & 1
        cols.append(values(seed, c + 1, row0, n, bool((int_mask >> c) & 1), null_mod if nulls else 0))
    return cols
