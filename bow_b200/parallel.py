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
