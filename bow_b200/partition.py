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
    frame_lo: int = -1   # first row shipped: rows [frame_lo, row_lo) are the LEFT halo (Interpolate only)
    extra_windows: int = 0   # Interpolate only: 1 if the start row of window k_hi is produced here as well
    plain: bool = False  # the whole frame, to be run as an ordinary (unsharded) rolling: see `plan`

    @property
    def num_windows(self) -> int:
        return self.k_hi - self.k_lo

    @property
    def num_rows(self) -> int:
        return self.row_hi - self.row_lo

    @property
    def first_row(self) -> int:
        return self.row_lo if self.frame_lo < 0 else self.frame_lo

    @property
    def lead_rows(self) -> int:
        return self.row_lo - self.first_row


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
    if t_first < s0:
        # Negative timestamps with an offset: Go's truncating division can leave the first window start AFTER the
        # first row (rolling.go:96-99; t_first = -15, interval 10, offset 7 -> s0 = -13).  Those leading rows belong to
        # window 0 iff that window holds a row of its own (rolling.go:194-211) - a rule of the unsharded iterator that a
        # shard (whose rows before its first window start are a halo) does not apply.  The frame is not cut: shard 0
        # takes all of it and runs as an ordinary rolling, the other shards are empty.
        return [Shard(0, 0, W, 0, n_rows, n_rows, plain=True)] + [Shard(g, W, W, n_rows, n_rows, n_rows) for g in range(1, n_shards)]
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


def plan_interpolate(n_rows: int, t_first: int, t_last: int, interval: int, offset: int, n_shards: int,
                     lower_bound: Callable[[int], int], prev_valid_row: Callable[[int], int],
                     next_valid_row: Callable[[int], int], align: int = 64) -> List[Shard]:
    """Shards for Rolling.Interpolate (and the Aggregate that follows it on the interpolated frame).

    On top of `plan`, shard g ships
      * a LEFT halo back to `prev_valid_row(row_lo)`: the smallest, over the interpolated columns, of the last
        row before row_lo holding a valid value (0 if a column has none) - what interpolation.Linear /
        StepPrevious look up for the shard's first windows (reference linear.go:14-18, stepprevious.go:12-20;
        for the first shard this role is played by Options.PrevRow);
      * one extra WINDOW on the right (the one-window halo of the north star): the start row of window k_hi,
        real or synthetic, is the inclusive row of the shard's last window when an inclusive aggregation
        (IntegralTrapezoid / WeightedAverageLinear) follows, so it is interpolated here too;
      * a RIGHT halo up to `next_valid_row(r)`: one past the largest, over the interpolated columns, of the
        first row at or after r holding a valid value (n_rows if none), r = first row of the last window
        interpolated here (linear.go:20-27 looks beyond the window).
    """
    base = plan(n_rows, t_first, t_last, interval, offset, n_shards, lower_bound, align)
    if n_rows == 0 or base[0].plain:
        return base
    off = normalise_offset(interval, offset)
    s0 = first_window_start(t_first, interval, off)
    W = num_windows(t_first, t_last, interval, off)
    out = []
    for sh in base:
        if sh.num_windows == 0:
            out.append(Shard(sh.rank, sh.k_lo, sh.k_hi, sh.row_lo, sh.row_hi, sh.row_hi, sh.row_lo, 0))
            continue
        extra = 1 if sh.k_hi < W else 0
        k_last = sh.k_hi - 1 + extra                       # last window interpolated on this shard
        end_rows = lower_bound(s0 + (k_last + 1) * interval) if k_last + 1 < W else n_rows
        halo_hi = min(n_rows, max(end_rows + 1, next_valid_row(lower_bound(s0 + k_last * interval))))
        frame_lo = 0 if sh.row_lo == 0 else max(0, min(sh.row_lo - 1, prev_valid_row(sh.row_lo)))
        out.append(Shard(sh.rank, sh.k_lo, sh.k_hi, sh.row_lo, sh.row_hi, halo_hi, frame_lo, extra))
    return out
