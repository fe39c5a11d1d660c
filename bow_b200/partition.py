"""Range partitioning of the interval-rolling path across the GPUs of one box.

Windows are independent units, so the path shards with NO data-path collective: the window index
range [0, W) is cut into G contiguous pieces (cuts are multiples of 64 windows so that per-shard
validity bitmaps concatenate byte-aligned), every cut is mapped to a row by a lower-bound search on
the sorted time column, and each shard additionally receives a halo: the rows up to and including
the first row at or after its last window's end (the single row an inclusive window may borrow,
reference rolling/rolling.go:201-218).  Every shard starts exactly on a window start, so a plain
IntervalRolling on the shard reproduces the global window lattice (rolling.go:96-99), and per-shard
outputs are simply concatenated.

Pure host integer logic; the data never passes through this module.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List


@dataclass(frozen=True)
class Shard:
    rank: int
    k_lo: int      # first window owned
    k_hi: int      # one past the last window owned
    row_lo: int    # first row owned (= lower bound of S_{k_lo}; 0 for the first shard)
    row_hi: int    # one past the last row owned (= lower bound of S_{k_hi}; n for the last shard)
    halo_hi: int   # rows [row_hi, halo_hi) are shipped too (inclusive row of the last window)

    @property
    def num_windows(self) -> int:
        return self.k_hi - self.k_lo

    @property
    def num_rows(self) -> int:
        return self.row_hi - self.row_lo


def go_div(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def normalise_offset(interval: int, offset: int) -> int:
    """enforceIntervalAndOffset, rolling/rolling.go:114-128"""
    if interval <= 0:
        raise ValueError("strictly positive interval required")
    if offset >= interval or offset <= -interval:
        offset = offset - go_div(offset, interval) * interval
    if offset < 0:
        offset += interval
    return offset


def first_window_start(t_first: int, interval: int, offset: int) -> int:
    """rolling/rolling.go:96-99 (offset already normalised)"""
    s0 = go_div(t_first, interval) * interval + offset
    if s0 > t_first:
        s0 -= interval
    return s0


def num_windows(t_first: int, t_last: int, interval: int, offset: int) -> int:
    """countWindows, rolling/rolling.go:143-154"""
    s0 = first_window_start(t_first, interval, offset)
    return 0 if s0 > t_last else (t_last - s0) // interval + 1


def plan(n_rows: int, t_first: int, t_last: int, interval: int, offset: int, n_shards: int,
         lower_bound: Callable[[int], int], align: int = 64) -> List[Shard]:
    """Cuts [0, W) into n_shards window ranges and maps them to rows.

    lower_bound(x) -> first row index whose time is >= x (n_rows if none); it is evaluated
    2*(n_shards-1) + 1 times at most (on the device for resident data, arithmetically for the
    synthetic generators).
    """
    offset = normalise_offset(interval, offset)
    if n_rows == 0:
        return [Shard(g, 0, 0, 0, 0, 0) for g in range(n_shards)]
    s0 = first_window_start(t_first, interval, offset)
    W = num_windows(t_first, t_last, interval, offset)
    cuts = [0]
    for g in range(1, n_shards):
        k = (g * W + n_shards // 2) // n_shards
        k = (k // align) * align
        cuts.append(max(k, cuts[-1]))
    cuts.append(W)
    rows = [0] + [lower_bound(s0 + k * interval) for k in cuts[1:-1]] + [n_rows]
    shards = []
    for g in range(n_shards):
        row_hi = rows[g + 1]
        halo_hi = row_hi
        if g + 1 < n_shards and row_hi < n_rows and cuts[g + 1] > cuts[g]:
            halo_hi = row_hi + 1   # first row at or after the end of the shard's last window
        shards.append(Shard(g, cuts[g], cuts[g + 1], rows[g], row_hi, halo_hi))
    return shards


def regular_lower_bound(t0: int, step: int, n_rows: int) -> Callable[[int], int]:
    """lower bound for t[i] = t0 + i*step (the REGULAR synthetic generator)"""
    def lb(x: int) -> int:
        if x <= t0:
            return 0
        return min(n_rows, -((t0 - x) // step))
    return lb
