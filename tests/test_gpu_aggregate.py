"""Parity of the CUDA path (through the C ABI) against the oracle: window bounds and the basic
aggregation family.  Bit-exact for bounds / Count / Min / Max / First / Last / WindowStart and for
validity bitmaps; float64 Sum / ArithmeticMean within 1e-12 * max(|ref|, sum|terms|) (reduction
order differs from the reference's left-to-right loop)."""
import numpy as np
import pytest

from oracle import literal as L
from oracle import refc as R
from tests import helpers as H
from tests.golden import reference_vectors as G

pytestmark = pytest.mark.gpu

EXACT = {"WindowStart", "Count", "Min", "Max", "First", "Last"}
TOL = 1e-12
BASIC = ["WindowStart", "Count", "Sum", "ArithmeticMean", "Min", "Max", "First", "Last"]


@pytest.fixture(scope="module")
def ctx():
    from bow_b200 import native as N
    c = N.Ctx(0)
    yield c
    c.close()


def bits(a):
    return np.asarray(a).view(np.int64) if np.asarray(a).dtype == np.float64 else np.asarray(a)


def compare_aggs(names, got, want, abs_sums=None, what="", int_input=False):
    """int_input: the aggregated column is int64 (|v| < 2^53 / n in these tests): every partial sum is an exactly
    representable integer, so Sum is order independent and must be BIT-EXACT (north star: "int64 Sum bit-exact")"""
    assert len(got) == len(want)
    for j, name in enumerate(names):
        (gv, gm), (wv, wm) = got[j], want[j]
        assert gv.dtype == wv.dtype, (what, name, gv.dtype, wv.dtype)
        assert np.array_equal(gm, wm), f"{what} {name}: validity differs at {np.flatnonzero(gm != wm)[:10]}"
        if name in EXACT or (int_input and name == "Sum"):
            bad = np.flatnonzero(bits(gv) != bits(wv))
            # NaN payloads are not compared (see tests/helpers.same_value)
            if gv.dtype == np.float64:
                bad = bad[~(np.isnan(gv[bad]) & np.isnan(wv[bad]))]
            assert bad.size == 0, f"{what} {name}: {bad[:10]} got {gv[bad[:10]]} want {wv[bad[:10]]}"
        else:
            H.assert_in_tolerance_class(gv, wv, abs_sums[name] if abs_sums is not None else 0.0, f"{what} {name}")
        # null slots hold value 0 (bowbuffer.go:22-40)
        assert not np.any(bits(gv)[~gm]), f"{what} {name}: non-zero value in a null slot"


def run_both(ctx, cols, interval, specs, offset=0, inclusive=False, slice_offset=0):
    from bow_b200 import native as N
    fr = N.Frame.from_numpy(ctx, cols, offset=slice_offset)
    r = N.Rolling(fr, 0, interval, offset=offset, inclusive=inclusive)
    got = r.aggregate(specs)
    ref = R.RefRolling(R.Frame(cols), 0, interval, offset=offset, inclusive=inclusive)
    assert r.num_windows == ref.num_windows
    assert r.first_window_start == ref.first_window_start or len(cols[0][0]) == 0
    want = ref.aggregate(specs)
    # sum |term| per window for the tolerance class of the float reductions
    abs_sums = H.term_scales(cols, specs[-1][1], interval, offset=offset, inclusive=inclusive)
    r.close()
    fr.close()
    return got, want, abs_sums


@pytest.mark.parametrize("agg,fixture,factor,vtype,expected,cite", G.AGGREGATIONS,
                         ids=[f"{c[0]}-{c[1]}-{c[2]}" for c in G.AGGREGATIONS])
def test_golden_aggregations(ctx, agg, fixture, factor, vtype, expected, cite):
    from bow_b200 import native as N
    rows = G.FIXTURES[fixture]
    cols = H.np_cols_from_lists([[r[0] for r in rows], [r[1] for r in rows]], [L.INT64, L.FLOAT64])
    fr = N.Frame.from_numpy(ctx, cols)
    r = N.Rolling(fr, 0, 10)
    out = H.lists_from_np(r.aggregate([("WindowStart", 0), (agg, 1, [factor] if factor is not None else [])]))
    H.assert_cols_equal(out, [[r[0] for r in expected], [r[1] for r in expected]], cite)


@pytest.mark.parametrize("name,opts,expected", G.ITERATE, ids=[c[0] for c in G.ITERATE])
def test_golden_bounds(ctx, name, opts, expected):
    from bow_b200 import native as N
    cols = H.np_cols_from_lists(G.ITERATE_COLS, [L.INT64, L.FLOAT64])
    fr = N.Frame.from_numpy(ctx, cols)
    r = N.Rolling(fr, 0, G.ITERATE_INTERVAL, **opts)
    first, inc = r.bounds()
    s0 = r.first_window_start
    got = []
    for k in range(r.num_windows):
        lo, hi = int(first[k]), int(first[k + 1]) + int(inc[k])
        got.append((k, s0 + k * G.ITERATE_INTERVAL, s0 + (k + 1) * G.ITERATE_INTERVAL, lo,
                    G.ITERATE_COLS[0][lo:hi], G.ITERATE_COLS[1][lo:hi]))
    assert got == expected


CASES = []
for kind in ("regular", "dense", "sparse", "bursty"):
    for n in (1, 2, 16, 17, 18, 300, 2175, 2176, 2177, 2178, 4352, 4353, 10000, 70001):
        CASES.append((kind, n))


@pytest.mark.parametrize("kind,n", CASES, ids=[f"{k}-{n}" for k, n in CASES])
def test_random_vs_oracle(ctx, kind, n):
    rng = np.random.default_rng(H.seed_of((kind, n)))
    for trial in range(4):
        t = H.random_times(rng, n, kind)
        dtype = np.int64 if trial == 1 else np.float64
        null_p = [0.0, 0.3, 0.1, 0.9][trial]
        v = H.random_values(rng, n, dtype, null_p, specials=(trial == 2))
        interval = int(rng.choice([1, 2, 5, 10, 60, 1000, 100000]))
        offset = int(rng.integers(-2 * interval, 2 * interval))
        inclusive = bool(rng.integers(0, 2))
        cols = [(t, None), v]
        specs = [("WindowStart", 0)] + [(a, 1) for a in BASIC[1:]]
        got, want, abs_sums = run_both(ctx, cols, interval, specs, offset=offset, inclusive=inclusive,
                                       slice_offset=int(rng.integers(0, 70)) if trial == 3 else 0)
        compare_aggs(BASIC, got, want, abs_sums, what=f"{kind} n={n} I={interval} off={offset} inc={inclusive}",
                     int_input=dtype == np.int64)


INTEGRALS = ["IntegralStep", "IntegralTrapezoid", "WeightedAverageStep", "WeightedAverageLinear"]
ICASES = [(k, n) for k in ("regular", "dense", "sparse", "bursty") for n in (1, 2, 17, 18, 300, 2176, 2177, 4353, 30011)]


@pytest.mark.parametrize("kind,n", ICASES, ids=[f"{k}-{n}" for k, n in ICASES])
def test_random_integrals_vs_oracle(ctx, kind, n):
    """float64 integrals / weighted averages: |gpu - ref| <= 1e-12 * max(|ref|, integral of |v|)"""
    from bow_b200 import native as N
    rng = np.random.default_rng(H.seed_of((kind, n, "i")))
    for trial in range(4):
        t = H.random_times(rng, n, kind)
        dtype = np.int64 if trial == 1 else np.float64
        v = H.random_values(rng, n, dtype, [0.0, 0.3, 0.1, 0.9][trial])
        interval = int(rng.choice([1, 2, 5, 10, 60, 1000, 100000]))
        offset = int(rng.integers(-2 * interval, 2 * interval))
        inclusive = bool(rng.integers(0, 2))
        cols = [(t, None), v]
        specs = [("WindowStart", 0)] + [(a, 1) for a in INTEGRALS] + [("IntegralStep", 1, [0.5]), ("Count", 1)]
        fr = N.Frame.from_numpy(ctx, cols, offset=int(rng.integers(0, 70)) if trial == 3 else 0)
        r = N.Rolling(fr, 0, interval, offset=offset, inclusive=inclusive)
        got = r.aggregate(specs)
        want = R.RefRolling(R.Frame(cols), 0, interval, offset=offset, inclusive=inclusive).aggregate(specs)
        scales = H.term_scales(cols, 1, interval, offset=offset, inclusive=inclusive)
        what = f"{kind} n={n} I={interval} off={offset} inc={inclusive} trial={trial}"
        for j, sp in enumerate(specs):
            (gv, gm), (wv, wm) = got[j], want[j]
            assert np.array_equal(gm, wm), f"{what} {sp}: validity differs at {np.flatnonzero(gm != wm)[:10]}"
            if sp[0] in ("WindowStart", "Count"):
                assert np.array_equal(gv, wv), (what, sp)
                continue
            fac = float(np.prod(np.abs(sp[2]))) if len(sp) > 2 else 1.0
            H.assert_in_tolerance_class(gv, wv, scales[sp[0]], f"{what} {sp}", factor=fac)
            assert not np.any(bits(gv)[~gm]), f"{what} {sp}: non-zero value in a null slot"


@pytest.mark.parametrize("kind", ["regular", "dense", "sparse", "bursty"])
def test_random_bounds_vs_oracle(ctx, kind):
    from bow_b200 import native as N
    rng = np.random.default_rng(7)
    for n in (1, 5, 100, 2176, 2177, 30000):
        for trial in range(3):
            t = H.random_times(rng, n, kind)
            interval = int(rng.choice([1, 3, 10, 500]))
            offset = int(rng.integers(-interval, interval))
            inclusive = bool(trial % 2)
            cols = [(t, None)]
            fr = N.Frame.from_numpy(ctx, cols)
            r = N.Rolling(fr, 0, interval, offset=offset, inclusive=inclusive)
            first, inc = r.bounds()
            w = R.RefRolling(R.Frame(cols), 0, interval, offset=offset, inclusive=inclusive).windows()
            W = r.num_windows
            assert W == len(w["lo"])
            early, kept = r.early_rows()
            assert np.array_equal(first[:W], w["first_index"]), (kind, n, interval, offset)
            hi = first[1:] + inc
            lo = first[:W].copy()
            if early and not kept and W > 0:      # rows before s0 are dropped with an empty window 0
                hi[0] = lo[0]
            nonempty = w["hi"] > w["lo"]
            assert np.array_equal(hi[nonempty], w["hi"][nonempty]), (kind, n, interval, offset, inclusive)
            assert np.array_equal(hi[~nonempty], lo[~nonempty])
            assert np.array_equal(inc, w["is_inclusive"])
            assert first[W] == n


def test_multi_column_and_duplicates(ctx):
    rng = np.random.default_rng(3)
    n = 20000
    t = H.random_times(rng, n, "regular")
    a = H.random_values(rng, n, np.float64, 0.2)
    b = H.random_values(rng, n, np.int64, 0.0)
    cols = [(t, None), a, b]
    specs = [("Sum", 1), ("WindowStart", 0), ("Sum", 1), ("Count", 2), ("Count", 2, [2.5]), ("Max", 2), ("Last", 2),
             ("Count", 0), ("First", 0), ("ArithmeticMean", 1, [0.1]), ("Min", 1, [3.0, -1.0])]
    names = [s[0] for s in specs]
    from bow_b200 import native as N
    fr = N.Frame.from_numpy(ctx, cols)
    r = N.Rolling(fr, 0, 37, offset=5)
    got = r.aggregate(specs)
    want = R.RefRolling(R.Frame(cols), 0, 37, offset=5).aggregate(specs)
    scales = H.term_scales(cols, 1, 37, offset=5)
    for j, name in enumerate(names):
        (gv, gm), (wv, wm) = got[j], want[j]
        assert np.array_equal(gm, wm), name
        if name in ("Sum", "ArithmeticMean"):   # (of the float column; Min carries factors applied to an exact value: exact)
            fac = float(np.prod(np.abs(specs[j][2]))) if len(specs[j]) > 2 else 1.0
            H.assert_in_tolerance_class(gv, wv, scales[name], f"{j} {name}", factor=fac)
        else:
            assert np.array_equal(bits(gv), bits(wv)), (j, name)


def test_errors(ctx):
    from bow_b200 import native as N
    t = np.array([3, 2, 1, 5], dtype=np.int64)
    v = np.array([1.0, 2.0, 3.0, 4.0])
    fr = N.Frame.from_numpy(ctx, [(t, None), (v, None)])
    with pytest.raises(N.BowGpuError, match="EINVAL"):
        N.Rolling(fr, 0, 0)
    with pytest.raises(N.BowGpuError, match="ETYPE"):
        N.Rolling(fr, 1, 10)
    r = N.Rolling(fr, 0, 10)
    with pytest.raises(N.BowGpuError, match="ENOINTERVALCOL"):
        r.aggregate([("Sum", 1)])
    with pytest.raises(N.BowGpuError, match="EUNSORTED"):
        r.aggregate([("WindowStart", 0), ("Sum", 1)])
    with pytest.raises(N.BowGpuError, match="EUNSORTED"):
        r.bounds()
    # the sticky flag is cleared: a sorted frame works afterwards
    fr2 = N.Frame.from_numpy(ctx, [(np.sort(t), None), (v, None)])
    out = N.Rolling(fr2, 0, 10).aggregate([("WindowStart", 0), ("Sum", 1)])
    assert out[1][0][0] == 10.0
    tn = (np.array([1, 2, 3], dtype=np.int64), np.array([True, False, True]))
    fr3 = N.Frame.from_numpy(ctx, [tn])
    with pytest.raises(N.BowGpuError, match="ENULLTIME"):
        N.Rolling(fr3, 0, 10)
    # empty frame
    fr4 = N.Frame.from_numpy(ctx, [(np.zeros(0, dtype=np.int64), None), (np.zeros(0), None)])
    r4 = N.Rolling(fr4, 0, 10)
    assert r4.num_windows == 0
    assert r4.aggregate([("WindowStart", 0), ("Sum", 1)])[0][0].size == 0


@pytest.mark.parametrize("g", [2, 5])
def test_sharded_matches_oracle(ctx, g):
    """range-partitioned execution (bowgpu_rolling_create_shard): every shard runs on this GPU, the
    concatenation must equal the unsharded oracle"""
    from bow_b200 import parallel as PP
    from bow_b200 import runtime
    runtime.set_default_ctx(ctx)
    rng = np.random.default_rng(21 + g)
    specs = [("WindowStart", 0)] + [(a, 1) for a in BASIC[1:]] + [(a, 1) for a in INTEGRALS]
    for kind, n, interval in (("regular", 50000, 37), ("bursty", 40000, 500), ("sparse", 9000, 3)):
        t = H.random_times(rng, n, kind)
        t = t - int(t[0]) + 7
        v = H.random_values(rng, n, np.float64, 0.25)
        cols = [(t, None), v]
        inclusive = bool(rng.integers(0, 2))
        shards, s0 = PP.plan_for_columns(t, interval, 3, g)
        per = [PP.aggregate_shard(cols, s, 0, interval, s0, inclusive, specs) for s in shards]
        got = PP.concat_outputs(per)
        want = R.RefRolling(R.Frame(cols), 0, interval, offset=3, inclusive=inclusive).aggregate(specs)
        scales = H.term_scales(cols, 1, interval, offset=3, inclusive=inclusive)
        for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
            assert np.array_equal(gm, wm), (kind, sp)
            if sp[0] in EXACT:
                assert np.array_equal(bits(gv), bits(wv)), (kind, sp)
            else:
                H.assert_in_tolerance_class(gv, wv, scales[sp[0]], f"{kind} {sp}")
    # rows before the first window start (negative timestamps with an offset): the plan hands everything to shard 0,
    # which runs as an ordinary rolling (partition.plan)
    t = np.array([-15, -14, -12, -3, 4, 8, 25, 31, 32, 47], dtype=np.int64)
    cols = [(t, None), (np.arange(len(t)) * 1.5 - 4.0, None)]
    shards, s0 = PP.plan_for_columns(t, 10, 7, g)
    assert shards[0].plain
    per = [PP.aggregate_shard(cols, s, 0, 10, s0, False, specs) for s in shards if s.num_windows]
    got = PP.concat_outputs(per)
    want = R.RefRolling(R.Frame(cols), 0, 10, offset=7).aggregate(specs)
    scales = H.term_scales(cols, 1, 10, offset=7)
    for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
        assert np.array_equal(gm, wm), ("early rows", sp)
        if sp[0] in EXACT:
            assert np.array_equal(bits(gv), bits(wv)), ("early rows", sp)
        else:
            H.assert_in_tolerance_class(gv, wv, scales[sp[0]], f"early rows {sp}")
    runtime.set_default_ctx(None)


# ---- aggregation.Aggregate over the whole Bow (rolling/aggregation/whole.go) ---------------------------------------
WHOLE_ALL = ["WindowStart", "Count", "Sum", "ArithmeticMean", "Min", "Max", "First", "Last", "IntegralStep",
             "IntegralTrapezoid", "WeightedAverageStep", "WeightedAverageLinear"]


@pytest.mark.parametrize("name,rows,aggs,expected,cite", [c for c in G.WHOLE_CASES if not isinstance(c[3], str)],
                         ids=[c[0] for c in G.WHOLE_CASES if not isinstance(c[3], str)])
def test_whole_golden(ctx, name, rows, aggs, expected, cite):
    from bow_b200 import native as N
    t = np.array([r[0] for r in rows], dtype=np.int64)
    v = np.array([0.0 if r[1] is None else r[1] for r in rows], dtype=np.float64)
    m = np.array([r[1] is not None for r in rows], dtype=bool)
    fr = N.Frame.from_numpy(ctx, [(t, None), (v, m)])
    got = fr.aggregate_whole(0, [(ctor, 0 if col == "time" else 1) for ctor, col, _ in aggs])
    for (gv, gm), want in zip(got, expected["cols"]):
        assert [x if ok else None for x, ok in zip(gv.tolist(), gm.tolist())] == want, cite
    fr.close()


@pytest.mark.parametrize("kind", ["regular", "dense", "sparse", "bursty"])
def test_whole_random_vs_oracle(ctx, kind):
    """one window over the whole frame, all aggregations on a float64 and an int64 column (+ Factor), sizes around
    the tile edges; bit-exact except sums / means / integrals (1e-12 relative to the sum of |terms|)"""
    from bow_b200 import native as N
    rng = np.random.default_rng(H.seed_of(("whole", kind)))
    for n in (0, 1, 2, 17, 300, 8191, 8192, 8193, 16385, 40000):
        t = H.random_times(rng, n, kind)
        if n:
            t = t - int(t[0]) + int(rng.integers(0, 1000))
        vf = H.random_values(rng, n, np.float64, float(rng.choice([0.0, 0.3])), specials=(n % 3 == 0 and n < 400))
        vi = H.random_values(rng, n, np.int64, float(rng.choice([0.0, 0.6])))
        cols = [(t, None), vf, vi]
        specs = [(op, c, [0.5] if (op, c) in (("Sum", 1), ("Count", 2), ("WindowStart", 0)) else None)
                 for c in (0, 1, 2) for op in WHOLE_ALL]
        fr = N.Frame.from_numpy(ctx, cols)
        got = fr.aggregate_whole(0, specs)
        want = R.aggregate_whole(R.Frame(cols), 0, specs)
        fr.close()
        span = float(t[-1] - t[0]) if n else 0.0
        for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
            assert gv.dtype == wv.dtype and np.array_equal(gm, wm), (kind, n, sp)
            if not gm.any():
                continue
            a, b = gv[0], wv[0]
            if sp[0] in ("Sum", "ArithmeticMean", "IntegralStep", "IntegralTrapezoid", "WeightedAverageStep",
                         "WeightedAverageLinear") and np.isfinite(b):
                scale = 1e4 * max(n, 1) * (max(span, 1.0) if "Integral" in sp[0] else 1.0)
                assert abs(a - b) <= 1e-12 * max(abs(b), scale), (kind, n, sp, a, b)
            else:
                assert H.same_value(a.item(), b.item()), (kind, n, sp, a, b)


@pytest.mark.parametrize("shift", [0, 1, 2])
def test_device_resident_inputs_zero_copy(ctx, shift):
    """BOWGPU_MEM_DEVICE frames: torch tensors already in HBM are wrapped (16-byte aligned values: zero copy; an odd
    element offset makes them 8-byte aligned only: device-to-device copy); the Arrow bit offset of the validity bitmap
    is honoured either way"""
    import torch
    from bow_b200 import native as N
    rng = np.random.default_rng(40 + shift)
    n, interval = 50000, 23
    t = (np.cumsum(rng.integers(0, 4, size=n)) + 7).astype(np.int64)
    v, m = H.random_values(rng, n, np.float64, 0.25, specials=True)
    pad = np.zeros(shift, dtype=np.int64)
    dt = torch.from_numpy(np.concatenate([pad, t])).cuda()
    dv = torch.from_numpy(np.concatenate([pad.astype(np.float64), v])).cuda()
    db = torch.from_numpy(N.pack_bits(m, shift)).cuda()
    arr = (N.Col * 2)()
    arr[0].values, arr[0].validity, arr[0].offset, arr[0].length, arr[0].null_count, arr[0].dtype = \
        dt.data_ptr(), None, shift, n, 0, N.INT64
    arr[1].values, arr[1].validity, arr[1].offset, arr[1].length, arr[1].null_count, arr[1].dtype = \
        dv.data_ptr(), db.data_ptr(), shift, n, -1, N.FLOAT64       # null_count unknown: counted on the device
    torch.cuda.synchronize()
    fr = N.Frame.from_col_descs(ctx, arr, 2, N.MEM_DEVICE, keep=[dt, dv, db])
    r = N.Rolling(fr, 0, interval, offset=5)
    specs = [("WindowStart", 0), ("Count", 1), ("Min", 1), ("Max", 1), ("First", 1), ("Last", 1)]
    got = r.aggregate(specs)
    want = R.RefRolling(R.Frame([(t, None), (v, m)]), 0, interval, offset=5).aggregate(specs)
    for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
        assert np.array_equal(gm, wm), sp
        a, b = gv[gm], wv[wm]
        same = a.view(np.int64) == b.view(np.int64)
        if a.dtype == np.float64:
            same |= np.isnan(a) & np.isnan(b)
        assert same.all(), sp
    r.close()
    fr.close()


@pytest.mark.parametrize("interval", [7, 3000, 10**7])
def test_exact_ops_with_special_values_across_tiles(ctx, interval):
    """Min / Max / First / Last / Count with NaN, +-0, +-Inf, 1e300 and subnormals in windows that span threads, tiles
    (8192 rows) and the whole column: the left-wins tie rule and the sticky leading NaN survive every stitch level"""
    from bow_b200 import native as N
    rng = np.random.default_rng(H.seed_of("specials", interval))
    n = 60000
    t = (np.cumsum(rng.integers(0, 3, size=n)) + 1).astype(np.int64)
    v, m = H.random_values(rng, n, np.float64, 0.3, specials=True)
    specs = [("WindowStart", 0), ("Count", 1), ("Min", 1), ("Max", 1), ("First", 1), ("Last", 1)]
    fr = N.Frame.from_numpy(ctx, [(t, None), (v, m)])
    got = N.Rolling(fr, 0, interval).aggregate(specs)
    want = R.RefRolling(R.Frame([(t, None), (v, m)]), 0, interval).aggregate(specs)
    gw = fr.aggregate_whole(0, specs[1:])
    ww = R.aggregate_whole(R.Frame([(t, None), (v, m)]), 0, specs[1:])
    for sp, (gv, gm), (wv, wm) in list(zip(specs, got, want)) + list(zip(specs[1:], gw, ww)):
        assert np.array_equal(gm, wm), sp
        a, b = gv[gm], wv[wm]
        same = a.view(np.int64) == b.view(np.int64)      # bit-exact: -0.0 and +0.0 are different answers
        if a.dtype == np.float64:
            same |= np.isnan(a) & np.isnan(b)
        assert same.all(), (sp, a[~same][:3], b[~same][:3])
    fr.close()


@pytest.mark.parametrize("n,kind", [(1000, "regular"), (5_000_000, "bursty"), (30_000_000, "regular"), (26_000_000, "sparse")])
def test_pipelined_host_aggregate(ctx, n, kind):
    """bowgpu_aggregate_host: host columns in, host results out, chunks of the window range processed by concurrent
    worker contexts (upload / kernels / download overlap).  Must equal the plain path and the oracle, including windows
    and inclusive rows at chunk cuts."""
    from bow_b200 import native as N
    rng = np.random.default_rng(H.seed_of("hostagg", n, kind))
    t = H.random_times(rng, n, kind)
    t = t - int(t[0]) + 12345
    v = H.random_values(rng, n, np.float64, 0.15)
    w = H.random_values(rng, n, np.int64, 0.0)
    unused = (np.zeros(n), None)                     # never read by the aggregations: must not be uploaded
    cols = [(t, None), v, unused, w]
    interval = 37 if kind != "sparse" else 11
    specs = [("WindowStart", 0), ("Count", 1), ("Min", 1), ("Last", 3), ("Sum", 3), ("ArithmeticMean", 1),
             ("IntegralTrapezoid", 1), ("WeightedAverageStep", 3)]
    got = N.aggregate_host(ctx, cols, 0, interval, specs, offset=5)
    want = R.RefRolling(R.Frame(cols), 0, interval, offset=5).aggregate(specs)
    scales = {c: H.term_scales(cols, c, interval, offset=5) for c in (1, 3)}
    for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
        assert gv.dtype == wv.dtype and np.array_equal(gm, wm), (n, kind, sp)
        if sp[0] in H.TOL_OPS and not (sp[0] == "Sum" and sp[1] == 3):   # Sum of the int64 column: bit-exact
            H.assert_in_tolerance_class(gv, wv, scales[sp[1]][sp[0]], f"{n} {kind} {sp}")
        else:
            assert np.array_equal(gv[gm].view(np.int64), wv[wm].view(np.int64)), (n, kind, sp)


@pytest.mark.parametrize("kind", ["regular", "bursty", "sparse"])
def test_host_aggregate_over_a_device_list_and_as_a_shard(ctx, kind):
    """bowgpu_aggregate_host_ex: ONE process, the chunks of the window range dealt to worker contexts on every listed
    device (here every visible GPU, and the same GPU listed twice), and the shard form one process per GPU
    uses.  Same results as the oracle."""
    import torch
    from bow_b200 import native as N
    from bow_b200 import parallel as PP
    rng = np.random.default_rng(H.seed_of("hostagg-multi", kind))
    n = 600_000
    t = H.random_times(rng, n, kind)
    t = t - int(t[0]) + 999
    v = H.random_values(rng, n, np.float64, 0.2)
    w = H.random_values(rng, n, np.int64, 0.0)
    cols = [(t, None), v, w]
    interval = 29
    specs = [("WindowStart", 0), ("Count", 1), ("Max", 1), ("First", 2), ("Sum", 2), ("ArithmeticMean", 1),
             ("IntegralTrapezoid", 1), ("WeightedAverageStep", 2)]
    want = R.RefRolling(R.Frame(cols), 0, interval, offset=4).aggregate(specs)
    scales = {c: H.term_scales(cols, c, interval, offset=4) for c in (1, 2)}

    def check(got, want, what):
        for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
            assert gv.dtype == wv.dtype and np.array_equal(gm, wm), (what, sp)
            if sp[0] in H.TOL_OPS and not (sp[0] == "Sum" and sp[1] == 2):
                H.assert_in_tolerance_class(gv, wv, scales[sp[1]][sp[0]], f"{what} {sp}")
            else:
                assert np.array_equal(gv[gm].view(np.int64), wv[wm].view(np.int64)), (what, sp)

    ngpu = torch.cuda.device_count()
    for devices in ([0, 0], list(range(ngpu)), None):
        got = N.aggregate_host(ctx, cols, 0, interval, specs, offset=4, devices=devices, workers_per_device=2, chunk_rows=40_000)
        check(got, want, f"devices={devices}")
    # the shard form: the frame cut in three by partition.plan, every piece through the pipelined call
    shards, s0 = PP.plan_for_columns(t, interval, 4, 3)
    per = []
    for sh in shards:
        local = PP.slice_cols(cols, sh.row_lo, sh.halo_hi)
        per.append(N.aggregate_host(ctx, local, 0, interval, specs, chunk_rows=30_000,
                                    shard=(s0 + sh.k_lo * interval, sh.num_windows)))
    check(PP.concat_outputs(per), want, "shards")


@pytest.mark.parametrize("vtype", [np.float64, np.int64])
def test_whole_aggregate_over_thousands_of_tiles(ctx, vtype):
    """aggregation.Aggregate over a Bow of 3e7 rows (3 662 tiles in ONE window): the tile records are joined through the
    skip records (seg_skip_build_kernel) instead of one by one; every aggregation against the oracle."""
    from bow_b200 import native as N
    rng = np.random.default_rng(H.seed_of("whole-big", str(vtype)))
    n = 30_000_123
    t = np.cumsum(rng.integers(0, 3, size=n)).astype(np.int64) + 5
    v = H.random_values(rng, n, vtype, 0.05)
    cols = [(t, None), v]
    specs = [(a, 0 if a == "WindowStart" else 1) for a in WHOLE_ALL]
    fr = N.Frame.from_numpy(ctx, cols)
    got = fr.aggregate_whole(0, specs)
    want = R.aggregate_whole(R.Frame(cols), 0, specs)
    av = np.abs(v[0].astype(np.float64))
    span = float(t[-1] - t[0])
    scale = {"Sum": float(av[v[1]].sum()) if v[1] is not None else float(av.sum())}
    scale["ArithmeticMean"] = scale["Sum"] / n
    scale["IntegralStep"] = scale["IntegralTrapezoid"] = float(av.max()) * span
    scale["WeightedAverageStep"] = scale["WeightedAverageLinear"] = float(av.max())
    for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
        assert gv.dtype == wv.dtype and np.array_equal(gm, wm), sp
        if sp[0] in H.TOL_OPS and not (sp[0] == "Sum" and vtype == np.int64):
            H.assert_in_tolerance_class(gv, wv, scale[sp[0]], str(sp))
        else:
            assert np.array_equal(bits(gv), bits(wv)), sp
    # the same rows as ONE rolling window
    r = N.Rolling(fr, 0, int(t[-1] - t[0]) + 10, offset=int(t[0]) % (int(t[-1] - t[0]) + 10))
    assert r.num_windows == 1
    g2 = r.aggregate([("WindowStart", 0), ("Count", 1), ("Min", 1), ("Max", 1), ("First", 1), ("Last", 1)])
    w2 = R.RefRolling(R.Frame(cols), 0, int(t[-1] - t[0]) + 10, offset=int(t[0]) % (int(t[-1] - t[0]) + 10)).aggregate(
        [("WindowStart", 0), ("Count", 1), ("Min", 1), ("Max", 1), ("First", 1), ("Last", 1)])
    for (gv, gm), (wv, wm) in zip(g2, w2):
        assert np.array_equal(gm, wm) and np.array_equal(bits(gv), bits(wv))
    r.close()
    fr.close()
