"""Multi-GPU driver of the interval-rolling path: one process per GPU, range partitioning, no
collective on the data path.

Every rank owns a contiguous range of WINDOWS of the global window lattice (bow_b200.partition) and
holds only the rows of those windows plus a one-row halo.  It runs the ordinary single-GPU kernels on
its shard through `bowgpu_rolling_create_shard`, which pins the lattice (first window start, number of
windows) instead of deriving it from the shard's first and last row.  Per-shard outputs are W_shard
long and are simply concatenated; `gather_outputs` collects them on one rank with torch.distributed
(gloo or nccl) AFTER the hot path — it moves W-length results, never rows.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import partition as P

NpCol = Tuple[np.ndarray, Optional[np.ndarray]]
Outputs = List[Tuple[np.ndarray, np.ndarray]]          # per spec: (values, valid mask)
Executor = Callable[[Sequence[NpCol], int, int, int, int, bool, Sequence[tuple]], Outputs]


def gpu_executor(cols: Sequence[NpCol], time_col: int, interval: int, s0: int, num_windows: int, inclusive: bool,
                 specs: Sequence[tuple]) -> Outputs:
    """Runs one shard on this process' GPU through the C ABI."""
    from . import native as N
    from .runtime import default_ctx
    fr = N.Frame.from_numpy(default_ctx(), cols)
    if num_windows < 0:   # a `plain` shard (partition.plan): the whole frame as an ordinary rolling
        r = N.Rolling(fr, time_col, interval, offset=s0 % interval, inclusive=inclusive)
    else:
        r = N.Rolling(fr, time_col, interval, inclusive=inclusive, shard=(s0, num_windows))
    try:
        return r.aggregate(specs)
    finally:
        r.close()
        fr.close()


def slice_cols(cols: Sequence[NpCol], lo: int, hi: int) -> List[NpCol]:
    return [(v[lo:hi], None if m is None else m[lo:hi]) for v, m in cols]


def plan_for_columns(time: np.ndarray, interval: int, offset: int, n_shards: int) -> Tuple[List[P.Shard], int]:
    """Shards of a host-resident sorted time column -> (shards, global first window start)."""
    n = len(time)
    if n == 0:
        return P.plan(0, 0, 0, interval, offset, n_shards, lambda x: 0), 0
    off = P.normalise_offset(interval, offset)
    s0 = P.first_window_start(int(time[0]), interval, off)
    shards = P.plan(n, int(time[0]), int(time[-1]), interval, offset, n_shards,
                    lambda x: int(np.searchsorted(time, x, side="left")))
    return shards, s0


def aggregate_shard(cols: Sequence[NpCol], shard: P.Shard, time_col: int, interval: int, s0_global: int,
                    inclusive: bool, specs: Sequence[tuple], executor: Executor = gpu_executor) -> Outputs:
    """Aggregates the windows owned by `shard`.  `cols` are the GLOBAL columns (tests) or already the shard's
    rows [row_lo, halo_hi) when `len(cols[0][0]) == halo_hi - row_lo`."""
    nloc = shard.halo_hi - shard.row_lo
    local = cols if len(cols[0][0]) == nloc else slice_cols(cols, shard.row_lo, shard.halo_hi)
    if shard.plain:       # (num_windows < 0 tells the executor to run an ordinary rolling with offset s0 mod interval)
        return executor(local, time_col, interval, s0_global, -1, inclusive, specs)
    return executor(local, time_col, interval, s0_global + shard.k_lo * interval, shard.num_windows, inclusive, specs)


def concat_outputs(per_shard: Sequence[Outputs]) -> Outputs:
    """Per-shard results are disjoint window ranges in rank order: the global result is their concatenation."""
    nspec = len(per_shard[0])
    return [(np.concatenate([o[j][0] for o in per_shard]), np.concatenate([o[j][1] for o in per_shard]))
            for j in range(nspec)]


def gather_outputs(local: Outputs, dst: int = 0) -> Optional[Outputs]:
    """Collects every rank's outputs on rank `dst` (None elsewhere).  Works with any initialised
    torch.distributed backend; without one (single process) it returns `local`."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    gathered: List[Optional[Outputs]] = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst)
    if rank != dst:
        return None
    return concat_outputs(gathered)


# ---- Rolling.Interpolate across shards ------------------------------------------------------------------------
InterpExecutor = Callable[[Sequence[NpCol], int, int, int, int, bool, Sequence, Optional[Sequence[NpCol]]], List[NpCol]]


def halo_searches(cols: Sequence[NpCol], interp_cols: Sequence[int]):
    """prev_valid_row / next_valid_row callables of `partition.plan_interpolate` for host-resident columns."""
    n = len(cols[0][0])

    def prev_valid_row(row: int) -> int:
        lo = row
        for j in interp_cols:
            m = cols[j][1]
            if m is None:
                lo = min(lo, max(row - 1, 0))
                continue
            idx = np.flatnonzero(m[:row])
            lo = min(lo, int(idx[-1]) if len(idx) else 0)
        return lo

    def next_valid_row(row: int) -> int:
        hi = row
        for j in interp_cols:
            m = cols[j][1]
            if m is None:
                hi = max(hi, min(row + 1, n))
                continue
            idx = np.flatnonzero(m[row:])
            hi = max(hi, row + int(idx[0]) + 1 if len(idx) else n)
        return hi

    return prev_valid_row, next_valid_row


def plan_interpolate_for_columns(cols: Sequence[NpCol], time_col: int, interval: int, offset: int, n_shards: int,
                                 ops: Sequence) -> Tuple[List[P.Shard], int]:
    time = cols[time_col][0]
    n = len(time)
    if n == 0:
        return P.plan(0, 0, 0, interval, offset, n_shards, lambda x: 0), 0
    need = [j for j, o in enumerate(ops) if str(o) in ("Linear", "StepPrevious", "StepNext", "1", "2", "4")]
    pv, nv = halo_searches(cols, need)
    off = P.normalise_offset(interval, offset)
    s0 = P.first_window_start(int(time[0]), interval, off)
    shards = P.plan_interpolate(n, int(time[0]), int(time[-1]), interval, offset, n_shards,
                                lambda x: int(np.searchsorted(time, x, side="left")), pv, nv)
    return shards, s0


def gpu_interpolate_executor(cols: Sequence[NpCol], time_col: int, interval: int, s0: int, num_windows: int,
                             inclusive: bool, ops: Sequence, prev_row: Optional[Sequence[NpCol]]) -> List[NpCol]:
    """Interpolates windows [0, num_windows) of the lattice starting at s0 on this process' GPU; rows of `cols`
    before s0 are the left halo."""
    from . import native as N
    from .runtime import default_ctx
    fr = N.Frame.from_numpy(default_ctx(), cols)
    if num_windows < 0:   # a `plain` shard
        r = N.Rolling(fr, time_col, interval, offset=s0 % interval, inclusive=inclusive, prev_row=prev_row)
    else:
        r = N.Rolling(fr, time_col, interval, inclusive=inclusive, prev_row=prev_row, shard=(s0, num_windows))
    try:
        out = r.interpolate(ops)
        try:
            return out.download()
        finally:
            out.close()
    finally:
        r.close()
        fr.close()


def interpolate_shard(cols: Sequence[NpCol], shard: P.Shard, time_col: int, interval: int, s0_global: int,
                      inclusive: bool, ops: Sequence, prev_row: Optional[Sequence[NpCol]] = None,
                      executor: InterpExecutor = gpu_interpolate_executor) -> Tuple[List[NpCol], int]:
    """-> (interpolated rows of windows [k_lo, k_hi + extra_windows), number of those rows that belong to the
    shard's OWN windows).  The global interpolated frame is the concatenation of every shard's first `own` rows;
    the remaining rows (start row of the next shard's first window onwards) only serve a following inclusive
    Aggregate on this shard."""
    ncols = len(cols)
    if shard.num_windows == 0:
        return [(np.zeros(0, dtype=v.dtype), np.zeros(0, dtype=bool)) for v, _ in cols], 0
    nloc = shard.halo_hi - shard.first_row
    local = cols if len(cols[0][0]) == nloc else slice_cols(cols, shard.first_row, shard.halo_hi)
    if shard.plain:
        out = executor(local, time_col, interval, s0_global, -1, inclusive, ops, prev_row)
        return out, len(out[time_col][0])
    s0 = s0_global + shard.k_lo * interval
    out = executor(local, time_col, interval, s0, shard.num_windows + shard.extra_windows, inclusive, ops, prev_row)
    assert len(out) == ncols
    t_out = out[time_col][0]
    own = int(np.searchsorted(t_out, s0 + shard.num_windows * interval, side="left")) if shard.extra_windows else len(t_out)
    return out, own


def concat_frames(per_shard: Sequence[Tuple[List[NpCol], int]]) -> List[NpCol]:
    ncols = len(per_shard[0][0])
    return [(np.concatenate([o[j][0][:own] for o, own in per_shard]),
             np.concatenate([o[j][1][:own] for o, own in per_shard])) for j in range(ncols)]


def gpu_interpolate_aggregate_executor(cols: Sequence[NpCol], time_col: int, interval: int, s0: int, num_windows: int,
                                       ops: Sequence, specs: Sequence[tuple],
                                       prev_row: Optional[Sequence[NpCol]]) -> Outputs:
    """Fused Interpolate -> Aggregate of the windows [0, num_windows) of the lattice starting at s0 on this process'
    GPU (bowgpu_rolling_interpolate_aggregate): rows before s0 are the left halo, rows after the last window the right
    halo; the interpolated frame is never materialised."""
    from . import native as N
    from .runtime import default_ctx
    fr = N.Frame.from_numpy(default_ctx(), cols)
    if num_windows < 0:   # a `plain` shard
        r = N.Rolling(fr, time_col, interval, offset=s0 % interval, prev_row=prev_row)
    else:
        r = N.Rolling(fr, time_col, interval, prev_row=prev_row, shard=(s0, num_windows))
    try:
        return r.interpolate_aggregate(ops, specs)
    finally:
        r.close()
        fr.close()


def interpolate_aggregate_shard(cols: Sequence[NpCol], shard: P.Shard, time_col: int, interval: int, s0_global: int,
                                ops: Sequence, specs: Sequence[tuple], prev_row: Optional[Sequence[NpCol]] = None,
                                executor=gpu_interpolate_aggregate_executor) -> Outputs:
    """Rolling.Interpolate(ops).Aggregate(specs) of the windows owned by `shard` (a `plan_interpolate` shard: left halo,
    one extra window and right halo are shipped).  Per-shard outputs concatenate to the global result."""
    nloc = shard.halo_hi - shard.first_row
    local = cols if len(cols[0][0]) == nloc else slice_cols(cols, shard.first_row, shard.halo_hi)
    if shard.plain:
        return executor(local, time_col, interval, s0_global, -1, ops, specs, prev_row)
    return executor(local, time_col, interval, s0_global + shard.k_lo * interval, shard.num_windows, ops, specs, prev_row)
