"""The reference-facing API (bow_b200.rolling, mirroring Metronlab/bow's Go packages) end to end on the
GPU: these tests read like the reference's own (rolling_test.go, aggregation_test.go,
interpolation_test.go, aggregation/*_test.go, interpolation/*_test.go) and replay their golden
vectors; plus the BASELINE.json configs[0] fixture against the oracle."""
import os

import numpy as np
import pytest

from bow_b200 import bow as B
from bow_b200 import rolling
from bow_b200.rolling import aggregation, interpolation, transformation
from oracle import refc as R
from tests.golden import reference_vectors as G

pytestmark = pytest.mark.gpu
TYPES = {"int64": B.Int64, "float64": B.Float64}


def two_col_bow(cols):
    return B.NewBowFromColBasedInterfaces([G.TIME, G.VALUE], [B.Int64, B.Float64], cols)


def rows_bow(rows):
    return two_col_bow([[r[0] for r in rows], [r[1] for r in rows]])


@pytest.mark.parametrize("name,opts,expected", G.ITERATE, ids=[c[0] for c in G.ITERATE])
def test_iterate(name, opts, expected):  # rolling_test.go:111-297
    o = rolling.Options(Offset=opts.get("offset", 0), Inclusive=opts.get("inclusive", False))
    r = rolling.IntervalRolling(two_col_bow(G.ITERATE_COLS), G.TIME, G.ITERATE_INTERVAL, o)
    got = []
    while r.HasNext():
        idx, w = r.Next()
        cols = w.Bow.ToColBased()
        got.append((idx, w.FirstValue, w.LastValue, w.FirstIndex, cols[0], cols[1]))
    assert got == expected
    idx, w = r.Next()
    assert w is None and idx == len(expected)


def test_unset_inclusive():  # aggregation_test.go:125-171
    r = rolling.IntervalRolling(two_col_bow(G.ITERATE_COLS), G.TIME, 5, rolling.Options(Inclusive=True))
    r.Next()
    r.Next()
    _, w = r.Next()      # window [20, 25] holds the inclusive row 25
    assert w.IsInclusive and w.Bow.ToColBased()[0] == [25]
    e = w.UnsetInclusive()
    assert not e.IsInclusive and e.Bow.NumRows() == 0 and w.IsInclusive


@pytest.mark.parametrize("agg,fixture,factor,vtype,expected,cite", G.AGGREGATIONS,
                         ids=[f"{c[0]}-{c[1]}-{c[2]}" for c in G.AGGREGATIONS])
def test_aggregations(agg, fixture, factor, vtype, expected, cite):  # aggregation/core_test.go:90-108 runTestCases
    b = rows_bow(G.FIXTURES[fixture])
    a = getattr(aggregation, agg)(G.VALUE)
    if factor is not None:
        a = a.SetTransformations(transformation.Factor(factor))
    out = rolling.IntervalRolling(b, G.TIME, 10).Aggregate(aggregation.WindowStart(G.TIME), a).Bow()
    want = B.NewBowFromColBasedInterfaces([G.TIME, G.VALUE], [B.Int64, TYPES[vtype]],
                                          [[r[0] for r in expected], [r[1] for r in expected]])
    assert out.Equal(want), f"{cite}\n{out}\n{want}"


@pytest.mark.parametrize("name,aggrs,names,types,cols", G.AGG_DRIVER, ids=[c[0] for c in G.AGG_DRIVER])
def test_aggregate_driver(name, aggrs, names, types, cols):  # aggregation_test.go:12-107
    """the reference uses ad-hoc closures (w.FirstValue, float64(NumRows), 2*float64(NumRows)); the same
    column plumbing is replayed with WindowStart, Count and Count x Factor(2) (int64 instead of float64)"""
    r = rolling.IntervalRolling(two_col_bow(G.AGG_DRIVER_COLS), G.TIME, 10)
    mk = {"time": lambda c: aggregation.WindowStart(c), "nrows": lambda c: aggregation.Count(c),
          "double": lambda c: aggregation.Count(c).SetTransformations(transformation.Factor(2))}
    lst = []
    for kind, col, rename in aggrs:
        a = mk[kind](col)
        lst.append(a.RenameOutput(rename) if rename else a)
    out = r.Aggregate(*lst).Bow()
    want = B.NewBowFromColBasedInterfaces(names, [B.Int64] * len(names), [[int(v) for v in c] for c in cols])
    assert out.Equal(want), f"{out}\n{want}"


def test_aggregate_host_transformation_and_cursor():
    """arbitrary transformation closures run on the host over the W-length result; an
    iterator already advanced with Next() aggregates only the remaining windows (aggregation.go:193-201)"""
    r = rolling.IntervalRolling(two_col_bow(G.AGG_DRIVER_COLS), G.TIME, 10)
    sq = aggregation.Sum(G.VALUE).SetTransformations(transformation.Factor(2), lambda x: None if x is None else x + 1)
    out = r.Aggregate(aggregation.WindowStart(G.TIME), sq).Bow()
    assert out.ToColBased() == [[10, 20], [(1.0 + 1.5 + 1.6) * 2 + 1, (2.5 + 2.9) * 2 + 1]]
    r.Next()
    # window 0 is skipped, so the new interval column starts with a null and the reference's own
    # newIntervalRolling rejects the result (rolling.go:89-94)
    with pytest.raises(B.BowError) as e:
        r.Aggregate(aggregation.WindowStart(G.TIME), aggregation.Count(G.VALUE)).Bow()
    assert str(e.value) == ("newIntervalRolling: the first value of the column should be convertible to int64, "
                            "got <nil>")


@pytest.mark.parametrize("name,kind,rows,offset,expected,cite", G.INTERPOLATIONS + G.INTERPOLATIONS_STEPNEXT,
                         ids=[c[0] for c in G.INTERPOLATIONS + G.INTERPOLATIONS_STEPNEXT])
def test_interpolations(name, kind, rows, offset, expected, cite):  # interpolation/*_test.go
    fn = interpolation.None_ if kind == "None" else getattr(interpolation, kind)
    r = rolling.IntervalRolling(rows_bow(rows), G.TIME, 2, rolling.Options(Offset=offset))
    out = r.Interpolate(interpolation.WindowStart(G.TIME), fn(G.VALUE)).Bow()
    assert out.Equal(rows_bow(expected)), f"{cite}\n{out}"


@pytest.mark.parametrize("name,times,offset,expected", G.INTERP_WINDOWSTART, ids=[c[0] for c in G.INTERP_WINDOWSTART])
def test_interp_windowstart(name, times, offset, expected):  # windowstart_test.go:13-64
    b = B.NewBowFromColBasedInterfaces([G.TIME], [B.Int64], [times])
    out = rolling.IntervalRolling(b, G.TIME, 2, rolling.Options(Offset=offset)).Interpolate(
        interpolation.WindowStart(G.TIME)).Bow()
    assert out.ToColBased() == [expected]


def test_interpolate_empty_bow():  # interpolation_test.go:49-63
    r = rolling.IntervalRolling(two_col_bow([[], []]), G.TIME, 2)
    out = r.Interpolate(interpolation.WindowStart(G.TIME), interpolation.Linear(G.VALUE))
    assert out.Bow().NumRows() == 0 and out.NumWindows() == 0


def test_chain_interpolate_aggregate_stays_on_device():
    """IntervalRolling -> Interpolate -> Aggregate: the interpolated frame is never downloaded"""
    b = rows_bow([(10, 10.0), (15, 15.0), (17, 17.0), (23, 11.0), (31, 0.5)])
    r = rolling.IntervalRolling(b, G.TIME, 5)
    ri = r.Interpolate(interpolation.WindowStart(G.TIME), interpolation.Linear(G.VALUE))
    assert ri.bow is None and ri.NumWindows() == 5
    out = ri.Aggregate(aggregation.WindowStart(G.TIME), aggregation.WeightedAverageLinear(G.VALUE),
                       aggregation.Count(G.VALUE).RenameOutput("n")).Bow()
    assert ri.bow is None and ri.frame is None     # fused: the interpolated frame was not even materialised
    assert out.ToColBased()[0] == [10, 15, 20, 25, 30] and out.ToColBased()[2] == [1, 2, 2, 1, 2]
    # asking for the interpolated Bow afterwards materialises it (and a second Aggregate runs on that frame)
    assert ri.Bow().NumRows() == 8 and ri.frame is not None
    out2 = ri.Aggregate(aggregation.WindowStart(G.TIME), aggregation.WeightedAverageLinear(G.VALUE),
                        aggregation.Count(G.VALUE).RenameOutput("n")).Bow()
    assert out2.Equal(out)
    # cross-check with the oracle on the same chain
    cols = [(np.array([10, 15, 17, 23, 31], dtype=np.int64), None), (np.array([10.0, 15.0, 17.0, 11.0, 0.5]), None)]
    ic = R.RefRolling(R.Frame(cols), 0, 5).interpolate(["WindowStart", "Linear"])
    want = R.RefRolling(R.Frame([(v, None if m.all() else m) for v, m in ic]), 0, 5).aggregate(
        [("WindowStart", 0), ("WeightedAverageLinear", 1)])
    got = np.array([np.nan if v is None else v for v in out.ToColBased()[1]])
    assert np.allclose(got[want[1][1]], want[1][0][want[1][1]], rtol=1e-12)
    assert [v is not None for v in out.ToColBased()[1]] == want[1][1].tolist()


def load_config1():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "config1_bow1_100000.npz"))
    cols = {}
    for name in ("Int64_ref", "Int64_no_nils_bow1", "Int64_bow1", "Float64_bow1"):
        v = z[name]
        m = np.unpackbits(z[name + "__valid"], bitorder="little")[:len(v)].astype(bool)
        cols[name] = (v, None if m.all() else m)
    return cols


@pytest.mark.parametrize("interval", [10, 100, 1000])
def test_config1_fixture(interval):
    """BASELINE.json configs[0]: benchmarks/bow1-100000-rows.parquet, IntervalRolling(Int64_ref, 10) +
    ArithmeticMean / Count / Min / Max (also at intervals 100 and 1000 for multi-row windows, and on the int64
    column), against the oracle."""
    c = load_config1()
    names = list(c)
    b = B.NewBow(*[B.NewSeriesFromNumpy(n, *c[n]) for n in names])
    r = rolling.IntervalRolling(b, "Int64_ref", interval)
    aggs, specs = [aggregation.WindowStart("Int64_ref")], [("WindowStart", 0)]
    for col in ("Float64_bow1", "Int64_bow1"):
        for a in ("ArithmeticMean", "Count", "Min", "Max", "Sum", "First", "Last"):
            aggs.append(getattr(aggregation, a)(col).RenameOutput(f"{a}_{col}"))
            specs.append((a, names.index(col)))
    out = r.Aggregate(*aggs).Bow()
    want = R.RefRolling(R.Frame([c[n] for n in names]), 0, interval).aggregate(specs)
    assert out.NumRows() == len(want[0][0]) == r.NumWindows()
    for j, (sp, (wv, wm)) in enumerate(zip(specs, want)):
        arr = out.Column(j)
        gm = np.asarray(arr.is_valid())
        gv = np.asarray(arr.fill_null(0))
        assert np.array_equal(gm, wm), sp
        if sp[0] in ("ArithmeticMean", "Sum"):
            assert np.allclose(gv, wv, rtol=1e-12, atol=0), sp
        else:
            assert np.array_equal(gv.view(np.int64), wv.view(np.int64)), sp


@pytest.mark.parametrize("name,rows,aggs,expected,cite", G.WHOLE_CASES, ids=[c[0] for c in G.WHOLE_CASES])
def test_whole_aggregate_api(name, rows, aggs, expected, cite):  # rolling/aggregation/whole_test.go
    b = rows_bow(rows)
    la = []
    for ctor, col, out in aggs:
        a = getattr(aggregation, ctor)(col)
        la.append(a.RenameOutput(out) if out else a)
    if isinstance(expected, str):
        with pytest.raises(B.BowError) as ei:
            aggregation.Aggregate(b, G.TIME, *la)
        assert str(ei.value) == expected, cite
        return
    got = aggregation.Aggregate(b, G.TIME, *la)
    want = B.NewBowFromColBasedInterfaces(expected["names"], [TYPES[t] for t in expected["types"]], expected["cols"])
    assert got.Equal(want), f"{cite}\nexpected: {want}\nactual: {got}"


def test_interpolating_the_interval_column_with_none_fails_like_the_eager_path():
    """Interpolate(None(time), ...) leaves null timestamps: the Rolling built on the interpolated Bow fails in
    newIntervalRolling (rolling.go:91-94, deferred error).  The lazy (fused) Interpolate must not hide that behind the
    parent's window lattice."""
    t = np.arange(0, 40, 3, dtype=np.int64) + 1          # no row on a window start: every window gets a start row
    b = B.NewBow(B.NewSeriesFromNumpy("time", t, None), B.NewSeriesFromNumpy("v", np.arange(len(t)) * 0.5, None))
    r = rolling.IntervalRolling(b, "time", 10).Interpolate(interpolation.None_("time"), interpolation.Linear("v"))
    with pytest.raises(B.BowError):
        r.NumWindows()
    with pytest.raises(B.BowError):
        r.Aggregate(aggregation.WindowStart("time"), aggregation.Count("v")).Bow()


def test_prev_row_must_match_the_bow_columns():
    t = np.arange(0, 40, 3, dtype=np.int64) + 1
    b = B.NewBow(B.NewSeriesFromNumpy("time", t, None), B.NewSeriesFromNumpy("v", np.arange(len(t)) * 0.5, None))
    short = B.NewBow(B.NewSeriesFromNumpy("time", np.array([0], dtype=np.int64), None))
    swapped = B.NewBow(B.NewSeriesFromNumpy("v", np.array([0.5]), None), B.NewSeriesFromNumpy("time", np.array([0], dtype=np.int64), None))
    for bad in (short, swapped):
        r = rolling.IntervalRolling(b, "time", 10, rolling.Options(PrevRow=bad))
        with pytest.raises(B.BowError, match="prevRow must have the same columns"):
            r.Interpolate(interpolation.WindowStart("time"), interpolation.StepPrevious("v")).Bow()
