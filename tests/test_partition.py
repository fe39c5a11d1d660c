"""Host-side sharding logic of the multi-GPU path (bow_b200.partition / bow_b200.parallel) on CPU:
planner invariants, and a world_size-2 gloo run in which every rank aggregates its shard and rank 0
concatenates — with the ORACLE injected as the per-shard executor (there is no GPU here; the executor is a
test double, the planner / halo / lattice pinning / gather logic is the code under test)."""
import os
import socket

import numpy as np
import pytest

from bow_b200 import parallel as PP
from bow_b200 import partition as P
from oracle import refc as R
from tests import helpers as H

SPECS = [("WindowStart", 0), ("Count", 1), ("Sum", 1), ("ArithmeticMean", 1), ("Min", 1), ("Max", 1), ("First", 1),
         ("Last", 1), ("IntegralStep", 1), ("IntegralTrapezoid", 1), ("WeightedAverageLinear", 1)]


def oracle_executor(cols, time_col, interval, s0, num_windows, inclusive, specs):
    """Shard executor backed by the oracle: runs it on the shard's rows and re-indexes its windows onto the
    pinned lattice (s0, num_windows); windows without rows get the empty-window defaults."""
    if num_windows < 0:      # a `plain` shard (rows before the first window start): the whole frame, unsharded
        return R.RefRolling(R.Frame(cols), time_col, interval, offset=s0 % interval, inclusive=inclusive).aggregate(specs)
    outs = []
    n = len(cols[0][0])
    ref_out, k_off, wl = None, 0, 0
    if n:
        t = cols[time_col][0]
        assert int(t[0]) >= s0
        ref = R.RefRolling(R.Frame(cols), time_col, interval, offset=s0 % interval, inclusive=inclusive)
        assert (ref.first_window_start - s0) % interval == 0
        k_off = (ref.first_window_start - s0) // interval
        ref_out = ref.aggregate(specs)
        wl = ref.num_windows
    for j, sp in enumerate(specs):
        op = sp[0]
        dt = np.int64 if op in ("WindowStart", "Count") or (op in ("First", "Last") and cols[sp[1]][0].dtype == np.int64) \
            else np.float64
        v = np.zeros(num_windows, dtype=dt)
        m = np.zeros(num_windows, dtype=bool)
        if op == "WindowStart":
            v[:] = s0 + np.arange(num_windows, dtype=np.int64) * interval
            m[:] = True
        elif op in ("Count", "Sum"):
            m[:] = True
        if ref_out is not None:
            take = max(0, min(wl, num_windows - k_off))
            v[k_off:k_off + take] = ref_out[j][0][:take]
            m[k_off:k_off + take] = ref_out[j][1][:take]
        outs.append((v, m))
    return outs


def check_plan(shards, n, W):
    assert shards[0].k_lo == 0 and shards[0].row_lo == 0
    assert shards[-1].k_hi == W and shards[-1].row_hi == n
    for a, b in zip(shards, shards[1:]):
        assert a.k_hi == b.k_lo and a.row_hi == b.row_lo          # disjoint, contiguous
        assert a.k_hi % 64 == 0 or a.k_hi in (0, W)               # bitmaps concatenate byte aligned
        assert a.row_hi <= a.halo_hi <= a.row_hi + 1


@pytest.mark.parametrize("kind", ["regular", "dense", "sparse", "bursty"])
@pytest.mark.parametrize("g", [1, 2, 3, 8])
def test_plan_and_concat_matches_unsharded(kind, g):
    rng = np.random.default_rng(H.seed_of((kind, g)))
    for n, interval in ((0, 5), (1, 5), (50, 3), (5000, 7), (20000, 2), (20000, 400)):
        t = H.random_times(rng, n, kind)
        if n:
            t = t - int(t[0]) + 1000      # non-negative times (negative ones need the single-shard early-row path)
        v = H.random_values(rng, n, np.float64, 0.2)
        cols = [(t, None), v]
        offset = int(rng.integers(-interval, interval))
        inclusive = bool(rng.integers(0, 2))
        shards, s0 = PP.plan_for_columns(t, interval, offset, g)
        ref = R.RefRolling(R.Frame(cols), 0, interval, offset=offset, inclusive=inclusive)
        W = ref.num_windows
        check_plan(shards, n, W)
        assert sum(s.num_windows for s in shards) == W
        if n == 0:
            continue
        assert s0 == ref.first_window_start
        per = [PP.aggregate_shard(cols, s, 0, interval, s0, inclusive, SPECS, executor=oracle_executor) for s in shards]
        got = PP.concat_outputs(per)
        want = ref.aggregate(SPECS)
        for sp, (gv, gm), (wv, wm) in zip(SPECS, got, want):
            assert np.array_equal(gm, wm), (kind, g, n, interval, sp)
            assert np.array_equal(gv[gm].view(np.int64), wv[wm].view(np.int64)), (kind, g, n, interval, sp)


def test_regular_lower_bound():
    lb = P.regular_lower_bound(1000, 10, 50)
    t = 1000 + np.arange(50) * 10
    for x in (0, 999, 1000, 1001, 1005, 1010, 1489, 1490, 1491, 99999):
        assert lb(x) == int(np.searchsorted(t, x, side="left"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)          # every rank regenerates the same global series
        n, interval, offset = 30000, 11, 4
        t = (np.cumsum(rng.integers(0, 6, size=n)) + 100).astype(np.int64)
        v = (rng.normal(size=n), rng.random(n) > 0.15)
        cols = [(t, None), v]
        shards, s0 = PP.plan_for_columns(t, interval, offset, world)
        local = PP.slice_cols(cols, shards[rank].row_lo, shards[rank].halo_hi)     # a rank only holds its rows
        out = PP.aggregate_shard(local, shards[rank], 0, interval, s0, True, SPECS, executor=oracle_executor)
        full = PP.gather_outputs(out, dst=0)
        if rank == 0:
            want = R.RefRolling(R.Frame(cols), 0, interval, offset=offset, inclusive=True).aggregate(SPECS)
            ok = all(np.array_equal(gm, wm) and np.array_equal(gv[gm].view(np.int64), wv[wm].view(np.int64))
                     for (gv, gm), (wv, wm) in zip(full, want))
            ret.put(("ok" if ok else "mismatch", len(full[0][0]), len(want[0][0])))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    status, wg, ww = ret.get(timeout=10)
    assert status == "ok" and wg == ww


# ---- Rolling.Interpolate across shards -------------------------------------------------------------------------
OPS = ["WindowStart", "Linear", "StepPrevious"]
OPS_NEXT = ["WindowStart", "StepNext", "Linear"]


def oracle_interp_executor(cols, time_col, interval, s0, num_windows, inclusive, ops, prev_row):
    """Shard executor backed by the oracle.  The oracle has no pinned-lattice mode, so it interpolates the
    shard's rows INCLUDING the left halo (same lattice: offset = s0 mod interval) and the rows of the windows
    [s0, s0 + num_windows*interval) are cut out by time afterwards; windows of the lattice that lie after the
    last local row are empty and produce their start rows from the halo, like the kernel does."""
    t = cols[time_col][0]
    end = s0 + num_windows * interval
    # one sentinel row at the end of the lattice makes the oracle emit the trailing empty windows; it is cut away
    ext = []
    for j, (v, m) in enumerate(cols):
        sv = np.array([end if j == time_col else 0], dtype=v.dtype)
        sm = np.array([j == time_col], dtype=bool)
        mm = np.ones(len(v), dtype=bool) if m is None else m
        ext.append((np.concatenate([v, sv]), np.concatenate([mm, sm])))
    need_sentinel = len(t) == 0 or int(t[-1]) < end
    use = ext if need_sentinel else [(v, np.ones(len(v), dtype=bool) if m is None else m) for v, m in cols]
    pr = R.Frame(prev_row) if prev_row is not None else None
    ref = R.RefRolling(R.Frame(use), time_col, interval, offset=s0 % interval, inclusive=inclusive, prev_row=pr)
    out = ref.interpolate(ops)
    to = out[time_col][0]
    lo, hi = int(np.searchsorted(to, s0, side="left")), int(np.searchsorted(to, end, side="left"))
    return [(v[lo:hi], m[lo:hi]) for v, m in out]


@pytest.mark.parametrize("OPS", [OPS, OPS_NEXT], ids=["prev", "next"])
@pytest.mark.parametrize("kind", ["regular", "sparse", "bursty"])
@pytest.mark.parametrize("g", [1, 2, 5])
def test_sharded_interpolate_matches_unsharded(kind, g, OPS):
    rng = np.random.default_rng(H.seed_of((kind, g, "i")))
    for n, interval in ((1, 5), (40, 3), (6000, 7), (20000, 300)):
        t = H.random_times(rng, n, kind)
        t = t - int(t[0]) + 1000
        cols = [(t, None), H.random_values(rng, n, np.float64, 0.5), H.random_values(rng, n, np.int64, 0.9)]
        offset = int(rng.integers(-interval, interval))
        prev = [(np.array([990], dtype=np.int64), None), (np.array([1.5]), None), (np.array([7], dtype=np.int64), None)]
        shards, s0 = PP.plan_interpolate_for_columns(cols, 0, interval, offset, g, OPS)
        ref = R.RefRolling(R.Frame(cols), 0, interval, offset=offset, prev_row=R.Frame(prev))
        want = ref.interpolate(OPS)
        per = [PP.interpolate_shard(cols, s, 0, interval, s0, False, OPS, prev, executor=oracle_interp_executor)
               for s in shards]
        got = PP.concat_frames(per)
        for j in range(3):
            assert np.array_equal(got[j][1], want[j][1]), (kind, g, n, interval, j)
            assert np.array_equal(got[j][0][got[j][1]].view(np.int64), want[j][0][want[j][1]].view(np.int64)), \
                (kind, g, n, interval, j)
        # the rows a shard keeps beyond its own windows start with the next shard's first row
        for (o, own), nxt in zip(per, per[1:]):
            if len(o[0][0]) > own and nxt[1] > 0:
                assert o[0][0][own] == nxt[0][0][0][0]


@pytest.mark.parametrize("g", [2, 3, 8])
def test_rows_before_the_first_window_start_are_not_cut(g):
    """negative timestamps with an offset: Go's truncating division puts the first window start AFTER the first rows
    (t_first = -15, interval 10, offset 7 -> s0 = -13; rolling.go:96-99).  Window 0 keeps those rows iff it holds a row of
    its own (rolling.go:194-211): only the unsharded iterator knows, so `plan` hands the whole frame to shard 0."""
    for t_list, interval, offset in (([-15, -14, -12, -3, 4, 8, 25, 31], 10, 7), ([-15, -1, 0, 9, 13], 10, 7),
                                     ([-29, -28, 40, 41, 99], 10, 3)):
        t = np.array(t_list, dtype=np.int64)
        v = (np.arange(len(t), dtype=np.float64) + 0.5, None)
        cols = [(t, None), v]
        off = P.normalise_offset(interval, offset)
        assert int(t[0]) < P.first_window_start(int(t[0]), interval, off)
        shards, s0 = PP.plan_for_columns(t, interval, offset, g)
        assert shards[0].plain and shards[0].row_hi == len(t) and all(s.num_windows == 0 for s in shards[1:])
        per = [PP.aggregate_shard(cols, s, 0, interval, s0, False, SPECS, executor=oracle_executor) for s in shards if s.num_windows]
        got = PP.concat_outputs(per)
        want = R.RefRolling(R.Frame(cols), 0, interval, offset=offset).aggregate(SPECS)
        for sp, (gv, gm), (wv, wm) in zip(SPECS, got, want):
            assert np.array_equal(gm, wm) and np.array_equal(gv[gm].view(np.int64), wv[wm].view(np.int64)), sp
