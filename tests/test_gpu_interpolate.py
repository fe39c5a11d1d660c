"""Parity of Rolling.Interpolate on the GPU (through the C ABI) against the oracle: row placement,
values and validity of every output row.  Bit-exact (WindowStart / StepPrevious / None / copied rows;
Linear evaluates the reference's five float64 operations in the same order without FMA, so it is
bit-exact too; the stated tolerance for Linear is 1e-12 relative)."""
import numpy as np
import pytest

from oracle import literal as L
from oracle import refc as R
from tests import helpers as H
from tests.golden import reference_vectors as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from bow_b200 import native as N
    c = N.Ctx(0)
    yield c
    c.close()


def bits(a):
    a = np.asarray(a)
    return a.view(np.int64) if a.dtype == np.float64 else a


def assert_frames_equal(got, want, what=""):
    assert len(got) == len(want)
    for j, ((gv, gm), (wv, wm)) in enumerate(zip(got, want)):
        assert len(gv) == len(wv), (what, j, len(gv), len(wv))
        assert gv.dtype == wv.dtype, (what, j)
        assert np.array_equal(gm, wm), f"{what} col {j}: validity differs at {np.flatnonzero(gm != wm)[:10]}"
        bad = np.flatnonzero((bits(gv) != bits(wv)) & gm)
        if gv.dtype == np.float64:
            bad = bad[~(np.isnan(gv[bad]) & np.isnan(wv[bad]))]
        assert bad.size == 0, f"{what} col {j}: rows {bad[:10]} got {gv[bad[:10]]} want {wv[bad[:10]]}"


@pytest.mark.parametrize("name,kind,rows,offset,expected,cite", G.INTERPOLATIONS, ids=[c[0] for c in G.INTERPOLATIONS])
def test_golden_interpolations(ctx, name, kind, rows, offset, expected, cite):
    from bow_b200 import native as N
    cols = H.np_cols_from_lists([[r[0] for r in rows], [r[1] for r in rows]], [L.INT64, L.FLOAT64])
    fr = N.Frame.from_numpy(ctx, cols)
    r = N.Rolling(fr, 0, 2, offset=offset)
    out = H.lists_from_np(r.interpolate(["WindowStart", "None_" if kind == "None" else kind]).download())
    H.assert_cols_equal(out, [[r[0] for r in expected], [r[1] for r in expected]], cite)


@pytest.mark.parametrize("name,times,offset,expected", G.INTERP_WINDOWSTART, ids=[c[0] for c in G.INTERP_WINDOWSTART])
def test_golden_interp_windowstart(ctx, name, times, offset, expected):
    from bow_b200 import native as N
    fr = N.Frame.from_numpy(ctx, H.np_cols_from_lists([times], [L.INT64]))
    r = N.Rolling(fr, 0, 2, offset=offset)
    H.assert_cols_equal(H.lists_from_np(r.interpolate(["WindowStart"]).download()), [expected])


ICASES = [(k, n) for k in ("regular", "dense", "sparse", "bursty") for n in (1, 2, 31, 33, 300, 2047, 2049, 4100, 30011)]


@pytest.mark.parametrize("kind,n", ICASES, ids=[f"{k}-{n}" for k, n in ICASES])
def test_random_interpolate_vs_oracle(ctx, kind, n):
    from bow_b200 import native as N
    rng = np.random.default_rng(H.seed_of((kind, n, "interp")))
    for trial in range(5):
        t = H.random_times(rng, n, kind)
        a = H.random_values(rng, n, np.float64, [0.0, 0.3, 0.1, 0.9, 0.5][trial])
        b = H.random_values(rng, n, np.int64, [0.2, 0.0, 0.6, 0.0, 0.97][trial])
        c = H.random_values(rng, n, np.float64, 0.4)
        cols = [a, (t, None), b, c]           # interval column in the middle
        ops = ["Linear", "WindowStart", ["StepPrevious", "Linear", "StepNext"][trial % 3], ["None_", "StepPrevious", "StepNext"][(trial + 1) % 3]]
        interval = int(rng.choice([1, 2, 5, 10, 60, 1000]))
        offset = int(rng.integers(-2 * interval, 2 * interval))
        inclusive = trial == 3
        prev = None
        if trial in (1, 4):
            prev = [(np.array([1.5]), np.array([trial == 1])), (np.array([int(t[0]) - 3], dtype=np.int64), None),
                    (np.array([7], dtype=np.int64), None), (np.array([-2.0]), np.array([True]))]
        so = int(rng.integers(0, 70)) if trial == 2 else 0
        fr = N.Frame.from_numpy(ctx, cols, offset=so)
        r = N.Rolling(fr, 1, interval, offset=offset, inclusive=inclusive, prev_row=prev)
        got = r.interpolate(ops).download()
        ref = R.RefRolling(R.Frame(cols), 1, interval, offset=offset, inclusive=inclusive,
                           prev_row=R.Frame(prev) if prev else None)
        want = ref.interpolate(ops)
        assert_frames_equal(got, want, f"{kind} n={n} I={interval} off={offset} inc={inclusive} trial={trial}")


def test_interpolate_then_aggregate(ctx):
    """the reference's usual chain: Interpolate(WindowStart, Linear) then WeightedAverageLinear / IntegralTrapezoid"""
    from bow_b200 import native as N
    rng = np.random.default_rng(11)
    n = 40000
    t = np.cumsum(rng.integers(1, 9, size=n)).astype(np.int64) + 1000
    v = H.random_values(rng, n, np.float64, 0.1)
    w = H.random_values(rng, n, np.int64, 0.0)
    cols = [(t, None), v, w]
    fr = N.Frame.from_numpy(ctx, cols)
    r = N.Rolling(fr, 0, 250, offset=70)
    fi = r.interpolate(["WindowStart", "Linear", "Linear"])
    r2 = N.Rolling(fi, 0, 250, offset=70)
    specs = [("WindowStart", 0), ("WeightedAverageLinear", 1), ("IntegralTrapezoid", 1), ("WeightedAverageStep", 2),
             ("Count", 1), ("ArithmeticMean", 2)]
    got = r2.aggregate(specs)
    ref = R.RefRolling(R.Frame(cols), 0, 250, offset=70)
    icols = ref.interpolate(["WindowStart", "Linear", "Linear"])
    icols = [(vv, None if mm.all() else mm) for vv, mm in icols]
    assert_frames_equal(fi.download(), [(vv, np.ones(len(vv), bool) if mm is None else mm) for vv, mm in icols], "interp")
    want = R.RefRolling(R.Frame(icols), 0, 250, offset=70).aggregate(specs)
    scales = {c: H.term_scales(icols, c, 250, offset=70) for c in (1, 2)}
    for j, sp in enumerate(specs):
        (gv, gm), (wv, wm) = got[j], want[j]
        assert np.array_equal(gm, wm), sp
        if gv.dtype == np.int64:
            assert np.array_equal(gv, wv), sp
        else:
            H.assert_in_tolerance_class(gv, wv, scales[sp[1]][sp[0]], str(sp))


def test_interpolate_errors_and_empty(ctx):
    from bow_b200 import native as N
    t = np.array([1, 2, 3], dtype=np.int64)
    v = np.array([1.0, 2.0, 3.0])
    fr = N.Frame.from_numpy(ctx, [(t, None), (v, None)])
    r = N.Rolling(fr, 0, 2)
    with pytest.raises(N.BowGpuError, match="ETYPE"):
        r.interpolate(["WindowStart", "WindowStart"])
    with pytest.raises(N.BowGpuError, match="EINVAL"):
        r.interpolate(["WindowStart"])
    fe = N.Frame.from_numpy(ctx, [(np.zeros(0, dtype=np.int64), None), (np.zeros(0), None)])
    out = N.Rolling(fe, 0, 2).interpolate(["WindowStart", "Linear"])
    assert out.num_rows == 0 and out.num_cols == 2


@pytest.mark.parametrize("g", [2, 3, 7])
def test_sharded_interpolate_matches_oracle(ctx, g):
    """range-partitioned Interpolate (left halo = rows before the shard's first window, right halo = one window +
    next valid rows): the concatenation of the shards' own rows must equal the unsharded oracle, and the chained
    per-shard Aggregate with inclusive aggregations must equal the unsharded chain"""
    from bow_b200 import native as N
    from bow_b200 import parallel as PP
    rng = np.random.default_rng(100 + g)
    ops = ["WindowStart", "Linear", "StepPrevious", "Linear"]
    specs = [("WindowStart", 0), ("WeightedAverageLinear", 1), ("IntegralTrapezoid", 3), ("Last", 2), ("Count", 1),
             ("IntegralStep", 1)]
    for kind, n, interval in (("regular", 30000, 40), ("sparse", 9000, 11), ("bursty", 50000, 700), ("regular", 300, 5)):
        t = H.random_times(rng, n, kind)
        t = t - int(t[0]) + 5000
        cols = [(t, None), H.random_values(rng, n, np.float64, 0.4), H.random_values(rng, n, np.int64, 0.93),
                H.random_values(rng, n, np.int64, 0.0)]
        offset = int(rng.integers(-interval, interval))
        prev = [(np.array([4990], dtype=np.int64), None), (np.array([0.25]), None), (np.array([3], dtype=np.int64), None),
                (np.array([-8], dtype=np.int64), None)]
        shards, s0 = PP.plan_interpolate_for_columns(cols, 0, interval, offset, g, ops)
        ref = R.RefRolling(R.Frame(cols), 0, interval, offset=offset, prev_row=R.Frame(prev))
        want = ref.interpolate(ops)
        per = [PP.interpolate_shard(cols, s, 0, interval, s0, False, ops, prev) for s in shards]
        got = PP.concat_frames(per)
        assert_frames_equal(got, want, f"sharded interp {kind} g={g}")
        # chained on the device, shard by shard
        icols = [(vv, None if mm.all() else mm) for vv, mm in want]
        want_agg = R.RefRolling(R.Frame(icols), 0, interval, offset=offset).aggregate(specs)
        outs = []
        for s in shards:
            if s.num_windows == 0:
                continue
            local = PP.slice_cols(cols, s.first_row, s.halo_hi)
            fr = N.Frame.from_numpy(ctx, local)
            s0g = s0 + s.k_lo * interval
            r = N.Rolling(fr, 0, interval, prev_row=prev, shard=(s0g, s.num_windows + s.extra_windows))
            fi = r.interpolate(ops)
            r2 = N.Rolling(fi, 0, interval, shard=(s0g, s.num_windows))
            outs.append(r2.aggregate(specs))
            for o in (r2, fi, r, fr):
                o.close()
        got_agg = PP.concat_outputs(outs)
        # the fused per-shard chain (no materialised frame) must agree with the materialising one and the oracle
        fused = PP.concat_outputs([PP.interpolate_aggregate_shard(cols, s, 0, interval, s0, ops, specs, prev)
                                   for s in shards if s.num_windows])
        for sp, (fv, fm), (gv, gm) in zip(specs, fused, got_agg):
            assert np.array_equal(fm, gm), (kind, g, sp, "fused validity")
            if gv.dtype == np.int64 or sp[0] in ("Last", "First", "Min", "Max"):
                assert np.array_equal(bits(fv[fm]), bits(gv[gm])), (kind, g, sp, "fused values")
            else:
                sc = np.maximum(np.abs(gv[gm]), 1.0)
                assert np.all(np.abs(fv[fm] - gv[gm]) <= 1e-12 * sc * max(1, interval)), (kind, g, sp, "fused values")
        for j, sp in enumerate(specs):
            (gv, gm), (wv, wm) = got_agg[j], want_agg[j]
            assert np.array_equal(gm, wm), (kind, g, sp)
            if gv.dtype == np.int64:
                assert np.array_equal(gv[gm], wv[wm]), (kind, g, sp)
            else:
                scale = np.maximum(np.abs(wv[wm]), 1.0)
                assert np.all(np.abs(gv[gm] - wv[wm]) <= 1e-12 * scale * max(1, interval)), (kind, g, sp)


ALL_AGGS = ["Count", "Sum", "ArithmeticMean", "Min", "Max", "First", "Last", "IntegralStep", "IntegralTrapezoid",
            "WeightedAverageStep", "WeightedAverageLinear"]


def chain_oracle(cols, interval, offset, ops, specs, prev=None):
    ref = R.RefRolling(R.Frame(cols), 0, interval, offset=offset, prev_row=R.Frame(prev) if prev else None)
    icols = ref.interpolate(ops)
    icols = [(vv, None if mm.all() else mm) for vv, mm in icols]
    return R.RefRolling(R.Frame(icols), 0, interval, offset=offset).aggregate(specs)


def assert_chain(got, want, specs, what, interval):
    for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
        assert gv.dtype == wv.dtype, (what, sp)
        assert np.array_equal(gm, wm), f"{what} {sp}: validity differs at windows {np.flatnonzero(gm != wm)[:8]}"
        a, b = gv[gm], wv[wm]
        if sp[0] in ("Sum", "ArithmeticMean", "IntegralStep", "IntegralTrapezoid", "WeightedAverageStep", "WeightedAverageLinear"):
            fin = np.isfinite(b)
            scale = 1e3 * (interval if "Integral" in sp[0] else 1.0)
            bad = np.flatnonzero(np.abs(a[fin] - b[fin]) > 1e-12 * np.maximum(np.abs(b[fin]), scale) * 64)
            assert bad.size == 0, f"{what} {sp}: {a[fin][bad[:3]]} vs {b[fin][bad[:3]]}"
            assert np.array_equal(np.isnan(a), np.isnan(b)), (what, sp)
        else:
            same = bits(a) == bits(b)
            if a.dtype == np.float64:
                same |= np.isnan(a) & np.isnan(b)
            assert same.all(), f"{what} {sp}: windows {np.flatnonzero(gm)[~same][:5]} got {a[~same][:3]} want {b[~same][:3]}"


@pytest.mark.parametrize("kind", ["regular", "sparse", "bursty", "dense"])
@pytest.mark.parametrize("n", [1, 40, 3000, 8192, 8193, 33000, 70000])
def test_fused_interpolate_aggregate_vs_oracle(ctx, kind, n):
    """bowgpu_rolling_interpolate_aggregate (no materialised frame) == oracle Interpolate -> Aggregate: synthetic start
    rows of non-empty AND empty windows, inclusive rows taken from the next window's start row, thread / tile edges"""
    from bow_b200 import native as N
    rng = np.random.default_rng(H.seed_of("fused", kind, n))
    for trial in range(2):
        interval = int(rng.choice([3, 11, 40, 700, 5000]))
        offset = int(rng.integers(-interval, interval))
        t = H.random_times(rng, n, kind)
        t = t - int(t[0]) + 1000
        cols = [(t, None), H.random_values(rng, n, np.float64, float(rng.choice([0.0, 0.3, 0.9])), specials=n < 100),
                H.random_values(rng, n, np.int64, float(rng.choice([0.0, 0.5]))),
                H.random_values(rng, n, np.float64, 0.2)]
        ops = ["WindowStart", str(rng.choice(["Linear", "StepPrevious", "None_", "StepNext"])),
               str(rng.choice(["Linear", "StepPrevious", "StepNext"])), "Linear"]
        prev = [(np.array([995], dtype=np.int64), None), (np.array([0.5]), None), (np.array([4], dtype=np.int64), None),
                (np.array([-1.0]), None)] if trial else None
        specs = [("WindowStart", 0), ("Count", 0)] + [(a, c) for c in (1, 2, 3) for a in ALL_AGGS]
        fr = N.Frame.from_numpy(ctx, cols)
        r = N.Rolling(fr, 0, interval, offset=offset, prev_row=prev)
        got = r.interpolate_aggregate(ops, specs)
        want = chain_oracle(cols, interval, offset, ops, specs, prev)
        assert_chain(got, want, specs, f"fused {kind} n={n} I={interval} off={offset} ops={ops}", interval)
        r.close()
        fr.close()


def test_fused_falls_back_when_starts_are_inexact(ctx):
    """ns timestamps beyond 2^53: a first row 123 ns after S_k passes the reference's float64 "has start" test
    (interpolation.go:121-128) without sitting on S_k; the fused path must hand over to the materialising one"""
    from bow_b200 import native as N
    n, interval = 20000, 60_000_000_000
    t = 1_700_000_000_000_000_000 + np.arange(n, dtype=np.int64) * 1_000_000_000 + 123
    rng = np.random.default_rng(9)
    cols = [(t, None), H.random_values(rng, n, np.float64, 0.1)]
    ops = ["WindowStart", "Linear"]
    specs = [("WindowStart", 0), ("Count", 1), ("WeightedAverageLinear", 1), ("Last", 1)]
    fr = N.Frame.from_numpy(ctx, cols)
    r = N.Rolling(fr, 0, interval)
    got = r.interpolate_aggregate(ops, specs)
    want = chain_oracle(cols, interval, 0, ops, specs)
    assert_chain(got, want, specs, "inexact starts", interval)
    # and with Options.Inclusive (duplicated inclusive rows in the interpolated frame) -> materialising chain too
    r2 = N.Rolling(fr, 0, interval, inclusive=True)
    got2 = r2.interpolate_aggregate(ops, specs)
    ref = R.RefRolling(R.Frame(cols), 0, interval, inclusive=True)
    icols = [(vv, None if mm.all() else mm) for vv, mm in ref.interpolate(ops)]
    want2 = R.RefRolling(R.Frame(icols), 0, interval, inclusive=True).aggregate(specs)
    assert_chain(got2, want2, specs, "inclusive", interval)


# ---- long null runs: the prev-valid / next-valid lookups climb a summary pyramid instead of walking the bitmap ----------
@pytest.mark.parametrize("pattern", ["ends", "middle", "none", "sparse"])
def test_interpolate_long_null_runs_vs_oracle(ctx, pattern):
    from bow_b200 import native as N
    n, interval = 300_000, 100
    rng = np.random.default_rng(H.seed_of("nullruns", pattern))
    t = np.arange(n, dtype=np.int64) * 3 + 7          # no row sits on a window start: every window gets a start row
    m = np.zeros(n, dtype=bool)
    if pattern == "ends":
        m[[0, n - 1]] = True
    elif pattern == "middle":
        m[n // 2 - 1] = m[n // 2 + 40_000] = True
    elif pattern == "sparse":
        m[rng.integers(0, n, size=12)] = True
    a = (rng.standard_normal(n), m)
    b = (rng.integers(-99, 99, size=n).astype(np.int64), m.copy())
    c = (rng.standard_normal(n), np.roll(m, 12345))
    cols = [(t, None), a, b, c]
    ops = ["WindowStart", "StepPrevious", "Linear", "StepNext"]
    fr = N.Frame.from_numpy(ctx, cols)
    r = N.Rolling(fr, 0, interval)
    out = r.interpolate(ops)
    got = out.download()
    out.close()
    want = R.RefRolling(R.Frame(cols), 0, interval).interpolate(ops)
    for j in range(4):
        assert np.array_equal(got[j][1], want[j][1]), (pattern, j)
        assert np.array_equal(got[j][0][got[j][1]].view(np.int64), want[j][0][want[j][1]].view(np.int64)), (pattern, j)
    r.close()
    fr.close()


def test_interpolate_all_null_column_is_not_quadratic(ctx):
    """5e7 rows, one valid value at each end, 5e5 windows: walking the bitmap word by word would read ~4e11 words"""
    import os
    import time
    from bow_b200 import native as N
    n = int(5e7 * float(os.environ.get("BOW_TEST_SCALE", "1")))
    interval = 100
    vals = np.zeros(n)
    vals[0], vals[-1] = 0.25, 9.5
    m = np.zeros(n, dtype=bool)
    m[0] = m[-1] = True
    time_col = np.arange(n, dtype=np.int64) * 2 + 7      # odd timestamps: no row sits on a window start
    fr = N.Frame.from_numpy(ctx, [(time_col, None), (vals, m), (vals, m.copy()), (vals, m.copy())])
    r = N.Rolling(fr, 0, interval)
    ctx.synchronize()
    t0 = time.perf_counter()
    out = r.interpolate(["WindowStart", "StepPrevious", "StepNext", "Linear"])
    ctx.synchronize()
    dt = time.perf_counter() - t0
    assert dt < 5.0, f"interpolate took {dt:.1f} s"
    W = r.num_windows
    n_out = out.num_rows
    assert n_out == n + W               # every window gets a start row
    got = out.download(n_out - 200, 200)
    out.close()
    r.close()
    fr.close()
    # the last window's start row: StepPrevious = the first value, StepNext = the last one, Linear in between
    ts, (sp, spm), (sn, snm), (li, lim) = got[0][0], got[1], got[2], got[3]
    k = np.flatnonzero(ts % interval == 0)[-1]
    assert spm[k] and sp[k] == 0.25 and snm[k] and sn[k] == 9.5 and lim[k] and 0.25 < li[k] < 9.5


@pytest.mark.parametrize("vtype", [L.INT64, L.FLOAT64])
def test_stepnext_is_pinned_on_the_reference_fillnext_goldens(ctx, vtype):
    """the CUDA StepNext against the reference's own FillNext golden table (G.STEPNEXT_VS_FILLNEXT), materialising and
    fused entry points"""
    from bow_b200 import native as N
    A = G.STEPNEXT_VS_FILLNEXT
    expected = [c for c in G.FILL_CASES if c[0] == A["expected_case"]][0][3]
    conv = (lambda x: x) if vtype == L.INT64 else (lambda x: None if x is None else float(x))
    cols = [[conv(r[c]) for r in G.FILL_ROWS] for c in range(5)]
    npcols = H.np_cols_from_lists([list(A["times"])] + cols, [L.INT64] + [vtype] * 5)
    fr = N.Frame.from_numpy(ctx, npcols)
    r = N.Rolling(fr, 0, A["interval"], offset=A["offset"])
    out = H.lists_from_np(r.interpolate(["WindowStart"] + ["StepNext"] * 5).download())
    syn = [i for i, t in enumerate(out[0]) if (t - A["offset"]) % A["interval"] == 0]
    assert len(syn) == len(A["times"])
    for k, i in enumerate(syn):
        H.assert_cols_equal([[out[c + 1][i]] for c in range(5)], [[conv(x)] for x in expected[k]], A["cite"])
    # fused: First of every window of the interpolated frame is its synthetic row
    got = r.interpolate_aggregate(["WindowStart"] + ["StepNext"] * 5, [("WindowStart", 0)] + [("First", c) for c in range(1, 6)])
    for c in range(5):
        gv, gm = got[c + 1]
        vals = [None if not ok else (int(x) if vtype == L.INT64 else float(x)) for x, ok in zip(gv.tolist(), gm.tolist())]
        assert vals == [conv(row[c]) for row in expected], (c, A["cite"])
    r.close()
    fr.close()


def test_has_start_row_goes_through_float64_above_2_53(ctx):
    """ns-epoch timestamps (> 2^53, one float64 ulp = 256 ns): the reference decides whether a window has a start row
    through float64 (interpolation.go:121-128), so a first row within +-128 ns of S_k counts as the start row.  The
    materialising Interpolate must reproduce that, and the fused call must notice it cannot (ST_INEXACT_START) and give
    the same results through the materialising chain."""
    from bow_b200 import native as N
    rng = np.random.default_rng(77)
    n, sec = 6000, 1_000_000_000
    interval = 60 * sec
    t0 = 1_700_000_000_000_000_000 // interval * interval       # a window start
    t = t0 + np.arange(n, dtype=np.int64) * sec
    jitter = rng.integers(-300, 300, size=n)                      # some land within 128 ns of a start, some do not
    jitter[rng.random(n) < 0.3] = 0
    t = np.sort(t + jitter)
    assert int(t[0]) > 2 ** 53
    a = H.random_values(rng, n, np.float64, 0.2)
    b = H.random_values(rng, n, np.int64, 0.1)
    cols = [(t, None), a, b]
    ops = ["WindowStart", "Linear", "StepPrevious"]
    fr = N.Frame.from_numpy(ctx, cols)
    r = N.Rolling(fr, 0, interval)
    out = r.interpolate(ops)
    got = out.download()
    out.close()
    ref = R.RefRolling(R.Frame(cols), 0, interval)
    want = ref.interpolate(ops)
    n_exact = int(np.sum(t % interval == 0))
    starts = ((t + interval // 2) // interval) * interval
    n_near = int(np.sum(np.abs(t - starts) <= 128)) - n_exact
    assert n_near > 5, "the test data must hold rows that match a window start only through float64"
    assert_frames_equal(got, want, "ns-epoch interpolate")
    specs = [("WindowStart", 0), ("Count", 1), ("First", 2), ("Last", 1), ("Max", 2)]
    fused = r.interpolate_aggregate(ops, specs)
    icols = [(vv, None if mm.all() else mm) for vv, mm in want]
    chain = R.RefRolling(R.Frame(icols), 0, interval).aggregate(specs)
    for sp, (gv, gm), (wv, wm) in zip(specs, fused, chain):
        assert np.array_equal(gm, wm), sp
        assert np.array_equal(bits(gv)[gm], bits(wv)[wm]), sp
    r.close()
    fr.close()


@pytest.mark.parametrize("kind,null_p", [("regular", 0.1), ("bursty", 0.6), ("sparse", 0.97)])
def test_pipelined_host_interpolate_aggregate(ctx, kind, null_p):
    """bowgpu_interpolate_aggregate_host: host columns in, host results out, every chunk of the window range with its
    left halo / extra window / right halo, over a device list.  Must equal Interpolate followed by Aggregate
    of the oracle — including the windows at chunk cuts and columns that are null for long stretches."""
    import torch
    from bow_b200 import native as N
    rng = np.random.default_rng(H.seed_of("hostinterp", kind))
    n = 500_000
    t = H.random_times(rng, n, kind)
    t = t - int(t[0]) + 12
    a = H.random_values(rng, n, np.float64, null_p)
    b = H.random_values(rng, n, np.int64, null_p / 2)
    c = H.random_values(rng, n, np.float64, 0.0)
    cols = [(t, None), a, b, c]
    interval = 41
    ops = ["WindowStart", "Linear", "StepPrevious", "StepNext"]
    specs = [("WindowStart", 0), ("Count", 1), ("First", 2), ("Last", 3), ("WeightedAverageLinear", 1), ("IntegralTrapezoid", 3),
             ("Min", 1), ("IntegralStep", 2)]
    ref = R.RefRolling(R.Frame(cols), 0, interval, offset=7)
    icols = [(vv, None if mm.all() else mm) for vv, mm in ref.interpolate(ops)]
    want = R.RefRolling(R.Frame(icols), 0, interval, offset=7).aggregate(specs)
    scales = {j: H.term_scales(icols, j, interval, offset=7, inclusive=True) for j in (1, 2, 3)}
    ngpu = torch.cuda.device_count()
    for devices in (None, [0, 0], list(range(ngpu))):
        got = N.interpolate_aggregate_host(ctx, cols, 0, interval, ops, specs, offset=7, devices=devices, chunk_rows=35_000)
        for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
            assert gv.dtype == wv.dtype and np.array_equal(gm, wm), (kind, devices, sp, np.flatnonzero(gm != wm)[:5])
            if sp[0] in H.TOL_OPS:
                H.assert_in_tolerance_class(gv, wv, scales[sp[1]][sp[0]], f"{kind} {devices} {sp}")
            else:
                assert np.array_equal(bits(gv)[gm], bits(wv)[wm]), (kind, devices, sp)
