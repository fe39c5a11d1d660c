"""The literal oracle (oracle/literal.py) replayed against every in-scope golden
vector of the reference's own tests (tests/golden/reference_vectors.py)."""
import pytest

from oracle import literal as L
from tests.golden import reference_vectors as G


def frame2(cols, time_type=L.INT64, value_type=L.FLOAT64):
    return L.Frame([G.TIME, G.VALUE], [time_type, value_type], [list(cols[0]), list(cols[1])])


def frame_rows(rows, value_type=L.FLOAT64):
    return L.Frame([G.TIME, G.VALUE], [L.INT64, value_type],
                   [[r[0] for r in rows], [r[1] for r in rows]])


def rows_of(f):
    cols = f.materialize()
    return list(zip(*cols)) if cols and cols[0] else []


@pytest.mark.parametrize("name,cols,interval,offset,expected", G.NUM_WINDOWS, ids=[c[0] for c in G.NUM_WINDOWS])
def test_num_windows(name, cols, interval, offset, expected):
    r = L.IntervalRolling.create(frame2(cols), G.TIME, interval, L.Options(offset=offset))
    assert r.num_windows == expected


def test_ctor_errors():
    b = frame2([[0], [1.0]])
    with pytest.raises(ValueError, match="enforceIntervalAndOffset: strictly positive interval required"):
        L.IntervalRolling.create(b, G.TIME, 0)
    with pytest.raises(ValueError, match="strictly positive"):
        L.IntervalRolling.create(b, G.TIME, -1)
    with pytest.raises(KeyError, match="no column 'badcol'"):
        L.IntervalRolling.create(b, "badcol", 1)
    with pytest.raises(TypeError, match="impossible to create a new intervalRolling on column of type float64"):
        L.IntervalRolling.create(L.Frame([G.TIME], [L.FLOAT64], [[0.0]]), G.TIME, 1)
    # empty bow gives valid finished iterator (rolling_test.go:100-108)
    r = L.IntervalRolling.create(frame2([[], []]), G.TIME, 1)
    assert r.next()[1] is None


@pytest.mark.parametrize("name,opts,expected", G.ITERATE, ids=[c[0] for c in G.ITERATE])
def test_iterate(name, opts, expected):
    r = L.IntervalRolling.create(frame2(G.ITERATE_COLS), G.TIME, G.ITERATE_INTERVAL, L.Options(**opts))
    got = []
    while r.has_next():
        wi, w = r.next()
        cols = w.bow.materialize()
        got.append((wi, w.first_value, w.last_value, w.first_index, cols[0], cols[1]))
    assert got == expected
    assert r.next()[1] is None


@pytest.mark.parametrize("agg,fixture,factor,vtype,expected,cite", G.AGGREGATIONS,
                         ids=[f"{c[0]}-{c[1]}-{c[2]}" for c in G.AGGREGATIONS])
def test_aggregations(agg, fixture, factor, vtype, expected, cite):
    b = frame_rows(G.FIXTURES[fixture])
    a = getattr(L, agg)(G.VALUE)
    if factor is not None:
        a = a.set_transformations(L.Factor(factor))
    out = L.IntervalRolling.create(b, G.TIME, 10).aggregate(L.WindowStart(G.TIME), a).bow
    assert out.names == [G.TIME, G.VALUE]
    assert out.types == [L.INT64, vtype]
    assert rows_of(out) == expected


def _driver_aggr(kind, col, rename):
    fns = {
        "time": (L.T_INT64, lambda c, w: w.first_value),
        "nrows": (L.T_FLOAT64, lambda c, w: float(w.bow.num_rows())),
        "double": (L.T_FLOAT64, lambda c, w: float(w.bow.num_rows()) * 2),
        "nil": (L.T_INT64, lambda c, w: None),
    }
    typ, fn = fns[kind]
    a = L.ColAggregation(col, False, typ, fn)
    return a.rename_output(rename) if rename else a


@pytest.mark.parametrize("name,aggrs,names,types,cols", G.AGG_DRIVER, ids=[c[0] for c in G.AGG_DRIVER])
def test_aggregate_driver(name, aggrs, names, types, cols):
    r = L.IntervalRolling.create(frame2(G.AGG_DRIVER_COLS), G.TIME, 10)
    out = r.aggregate(*[_driver_aggr(*a) for a in aggrs]).bow
    assert (out.names, out.types, out.materialize()) == (names, types, cols)


@pytest.mark.parametrize("name,aggrs,msg", G.AGG_DRIVER_ERRORS, ids=[c[0] for c in G.AGG_DRIVER_ERRORS])
def test_aggregate_driver_errors(name, aggrs, msg):
    r = L.IntervalRolling.create(frame2(G.AGG_DRIVER_COLS), G.TIME, 10)
    with pytest.raises((ValueError, KeyError)) as e:
        r.aggregate(*[_driver_aggr(*a) for a in aggrs])
    assert e.value.args[0] == msg


def test_unset_inclusive():
    g = G.UNSET_INCLUSIVE
    b = L.Frame([G.TIME, G.VALUE], [L.INT64, L.INT64], [list(c) for c in g["cols"]])
    w = L.Window(b, 0, 0, g["first_value"], g["last_value"], True)
    e = w.unset_inclusive()
    assert e.bow.materialize() == g["expected_cols"] and not e.is_inclusive
    assert w.is_inclusive and w.bow.materialize() == g["cols"]


@pytest.mark.parametrize("name,cols,offset,expected", G.INTERP_DRIVER, ids=[c[0] for c in G.INTERP_DRIVER])
def test_interpolate_driver(name, cols, offset, expected):
    ti = L.ColInterpolation(G.TIME, [L.INT64], lambda c, w, full, prev: w.first_value)
    vi = L.ColInterpolation(G.VALUE, [L.INT64, L.FLOAT64], lambda c, w, full, prev: 9.9)
    r = L.IntervalRolling.create(frame2(cols), G.TIME, 2, L.Options(offset=offset))
    assert r.interpolate(ti, vi).bow.materialize() == expected


def test_interpolate_driver_errors():
    ti = L.ColInterpolation(G.TIME, [L.INT64], lambda c, w, full, prev: w.first_value)
    vi = L.ColInterpolation(G.VALUE, [L.INT64, L.FLOAT64], lambda c, w, full, prev: 9.9)
    bad = L.ColInterpolation(G.VALUE, [L.INT64, "bool"], lambda c, w, full, prev: True)
    r = L.IntervalRolling.create(frame2([[10, 13], [1.0, 1.3]]), G.TIME, 2)
    with pytest.raises(TypeError) as e:
        r.interpolate(ti, bad)
    assert e.value.args[0] == G.INTERP_DRIVER_ERRORS[0][1]
    with pytest.raises(ValueError) as e:
        r.interpolate(vi)
    assert e.value.args[0] == G.INTERP_DRIVER_ERRORS[1][1]


@pytest.mark.parametrize("name,kind,rows,offset,expected,cite", G.INTERPOLATIONS + G.INTERPOLATIONS_STEPNEXT,
                         ids=[c[0] for c in G.INTERPOLATIONS + G.INTERPOLATIONS_STEPNEXT])
def test_interpolations(name, kind, rows, offset, expected, cite):
    r = L.IntervalRolling.create(frame_rows(rows), G.TIME, 2, L.Options(offset=offset))
    out = r.interpolate(L.InterpWindowStart(G.TIME), getattr(L, "Interp" + kind)(G.VALUE)).bow
    assert rows_of(out) == expected


@pytest.mark.parametrize("name,times,offset,expected", G.INTERP_WINDOWSTART,
                         ids=[c[0] for c in G.INTERP_WINDOWSTART])
def test_interp_windowstart(name, times, offset, expected):
    r = L.IntervalRolling.create(L.Frame([G.TIME], [L.INT64], [list(times)]), G.TIME, 2, L.Options(offset=offset))
    assert r.interpolate(L.InterpWindowStart(G.TIME)).bow.materialize() == [expected]


@pytest.mark.parametrize("vtype,msg", G.INTERP_TYPE_ERRORS)
def test_interp_type_errors(vtype, msg):
    b = L.Frame([G.TIME, G.VALUE], [L.INT64, vtype], [[10, 15], [None, None]])
    r = L.IntervalRolling.create(b, G.TIME, 2)
    with pytest.raises(TypeError) as e:
        r.interpolate(L.InterpWindowStart(G.TIME), L.InterpLinear(G.VALUE))
    assert e.value.args[0] == msg


@pytest.mark.parametrize("name,x,expected", G.FACTOR, ids=[c[0] for c in G.FACTOR])
def test_factor(name, x, expected):
    assert L.Factor(0.1)(x) == expected
    with pytest.raises(TypeError, match="factor: invalid type str"):
        L.Factor(0.1)("11")


@pytest.mark.parametrize("vtype", [L.INT64, L.FLOAT64])
def test_stepnext_is_pinned_on_the_reference_fillnext_goldens(vtype):
    """interpolation.StepNext does not exist upstream; its next-valid semantics are pinned on the reference's FillNext
    golden table instead (see G.STEPNEXT_VS_FILLNEXT): both oracles."""
    import numpy as np
    from oracle import refc as R
    from tests import helpers as H
    A = G.STEPNEXT_VS_FILLNEXT
    case = [c for c in G.FILL_CASES if c[0] == A["expected_case"]][0]
    expected = case[3]            # (the Float64 table of this case is the same, bowfill_test.go:269-288)
    conv = (lambda x: x) if vtype == L.INT64 else (lambda x: None if x is None else float(x))
    cols = [[conv(r[c]) for r in G.FILL_ROWS] for c in range(5)]
    names = ["t", "a", "b", "c", "d", "e"]
    fr = L.Frame(names, [L.INT64] + [vtype] * 5, [list(A["times"])] + cols)
    r = L.IntervalRolling.create(fr, "t", A["interval"], L.Options(offset=A["offset"]))
    out = r.interpolate(L.InterpWindowStart("t"), *[L.InterpStepNext(n) for n in names[1:]]).bow.materialize()
    syn = [i for i, t in enumerate(out[0]) if (t - A["offset"]) % A["interval"] == 0]    # the synthetic rows
    assert len(syn) == len(A["times"]) and len(out[0]) == 2 * len(A["times"])
    for k, i in enumerate(syn):
        assert [out[c + 1][i] for c in range(5)] == [conv(x) for x in expected[k]], (k, A["cite"])
    # the C oracle
    npcols = H.np_cols_from_lists([list(A["times"])] + cols, [L.INT64] + [vtype] * 5)
    got = R.RefRolling(R.Frame(npcols), 0, A["interval"], offset=A["offset"]).interpolate(["WindowStart"] + ["StepNext"] * 5)
    got = H.lists_from_np([(v, m) for v, m in got])
    for k, i in enumerate(syn):
        assert [got[c + 1][i] for c in range(5)] == [conv(x) for x in expected[k]], (k, "C oracle")
