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


# ---- BURSTY (BASELINE configs[3]), mirror of gen_bursty_counts / gen_time_bursty --------------------------------
def bursty_counts(seed: int, nw: int) -> np.ndarray:
    """rows per window of the pattern, windows 0 .. nw-1"""
    u = synth_u(seed, 0, np.arange(nw, dtype=np.uint64))
    r = u % np.uint64(100)
    small = np.uint64(1) + (u >> np.uint64(8)) % np.uint64(100)
    big = (np.uint64(1000) + (u >> np.uint64(8)) % np.uint64(1000)) << ((u >> np.uint64(32)) % np.uint64(10))
    c = np.where(r < 50, np.uint64(0), np.where(r < 95, small, big))
    return c.astype(np.int64)


def bursty_windows_needed(rows_end: int) -> int:
    return rows_end // 2000 + 4096


def bursty_offsets(seed: int, rows_end: int) -> np.ndarray:
    """off[k] = global index of the first row of window k (exclusive scan of the counts; off[nw] = rows covered)"""
    c = bursty_counts(seed, bursty_windows_needed(rows_end))
    off = np.zeros(len(c) + 1, dtype=np.int64)
    np.cumsum(c, out=off[1:])
    assert off[-1] >= rows_end, "pattern too short"
    return off


def bursty_time(seed: int, row0: int, n: int, t0: int, interval: int, off: np.ndarray = None) -> np.ndarray:
    if off is None:
        off = bursty_offsets(seed, row0 + n)
    gi = row0 + np.arange(n, dtype=np.int64)
    k = np.searchsorted(off, gi, side="right") - 1
    c = off[k + 1] - off[k]
    j = gi - off[k]
    shifted = ((synth_u(seed, 0, k.astype(np.uint64)) >> np.uint64(40)) & np.uint64(1)).astype(bool)
    dt = np.where(shifted, ((2 * j + 1) * interval) // (2 * c), (j * interval) // c)
    return t0 + k * interval + dt


def bursty_frame(row0: int, n: int, ncols: int = 1, seed: int = 42, t0: int = T0_DEFAULT, interval: int = STEP_DEFAULT,
                 null_mask: int = 0, int_mask: int = 0, null_mod: int = 10, off: np.ndarray = None):
    cols = [(bursty_time(seed, row0, n, t0, interval, off), None)]
    for c in range(ncols):
        nulls = (null_mask >> c) & 1
        cols.append(values(seed, c + 1, row0, n, bool((int_mask >> c) & 1), null_mod if nulls else 0))
    return cols


def bursty_lower_bound(off: np.ndarray, seed: int, t0: int, interval: int, n_rows: int):
    """lower bound callable (first global row with time >= x) for bow_b200.partition.plan"""
    def lb(x: int) -> int:
        if x <= t0:
            return 0
        k = (x - t0) // interval
        if k >= len(off) - 1:
            return n_rows
        if (x - t0) % interval == 0:
            return int(min(n_rows, off[k]))
        lo, c = int(off[k]), int(off[k + 1] - off[k])
        if c == 0:
            return int(min(n_rows, lo))
        tw = bursty_time(seed, lo, c, t0, interval, off)
        return int(min(n_rows, lo + np.searchsorted(tw, x, side="left")))
    return lb
