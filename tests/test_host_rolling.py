"""Host logic of the reference-facing mirror (bow_b200.rolling): constructor validation, NumWindows,
deferred errors of Aggregate / Interpolate and their verbatim error strings.  None of this touches the
GPU (validation stays on the host, like in the Go shim), so it runs in the CPU container."""
import pytest

from bow_b200 import bow as B
from bow_b200 import rolling
from bow_b200.rolling import aggregation, interpolation, transformation
from tests.golden import reference_vectors as G


def two_col_bow(cols, time_type=B.Int64):
    return B.NewBowFromColBasedInterfaces([G.TIME, G.VALUE], [time_type, B.Float64], cols)


@pytest.mark.parametrize("name,cols,interval,offset,expected", G.NUM_WINDOWS, ids=[c[0] for c in G.NUM_WINDOWS])
def test_num_windows(name, cols, interval, offset, expected):  # rolling_test.go:20-68
    r = rolling.IntervalRolling(two_col_bow(cols), G.TIME, interval, rolling.Options(Offset=offset))
    assert r.NumWindows() == expected


def test_ctor_errors():  # rolling_test.go:70-109
    b = two_col_bow([[0], [1.0]])
    for itv in (0, -1):
        with pytest.raises(B.BowError) as e:
            rolling.IntervalRolling(b, G.TIME, itv)
        assert str(e.value) == "enforceIntervalAndOffset: strictly positive interval required"
    with pytest.raises(B.BowError) as e:
        rolling.IntervalRolling(b, "badcol", 1)
    assert str(e.value) == "no column 'badcol'"
    with pytest.raises(B.BowError) as e:
        rolling.IntervalRolling(B.NewBowFromColBasedInterfaces([G.TIME], [B.Float64], [[0.0]]), G.TIME, 1)
    assert str(e.value) == "impossible to create a new intervalRolling on column of type float64"
    with pytest.raises(B.BowError) as e:
        rolling.IntervalRolling(b, G.TIME, 1, rolling.Options(PrevRow=two_col_bow([[0, 1], [1.0, 2.0]])))
    assert str(e.value) == "enforcePrevRow: prevRow must have only one row, have 2"
    with pytest.raises(B.BowError) as e:
        rolling.IntervalRolling(two_col_bow([[None, 1], [1.0, 2.0]]), G.TIME, 1)
    assert str(e.value) == "the first value of the column should be convertible to int64, got <nil>"


def test_empty_bow_gives_finished_iterator():  # rolling_test.go:99-108
    r = rolling.IntervalRolling(two_col_bow([[], []]), G.TIME, 1)
    assert not r.HasNext()
    idx, w = r.Next()
    assert idx == 0 and w is None


def test_offset_normalisation():  # rolling.go:114-128
    b = two_col_bow([[12, 29], [1.0, 2.0]])
    for off, want in ((0, 0), (3, 3), (8, 3), (5, 0), (-2, 3), (-5, 0), (-7, 3)):
        assert rolling.IntervalRolling(b, G.TIME, 5, rolling.Options(Offset=off)).options.Offset == want


def test_aggregate_deferred_errors():  # aggregation_test.go:108-122 + aggregation.go:147-188
    r = rolling.IntervalRolling(two_col_bow(G.AGG_DRIVER_COLS), G.TIME, 10)
    cases = [
        ((aggregation.Sum(G.VALUE),), "intervalRolling.indexedAggregations: must keep interval column 'time'"),
        ((aggregation.WindowStart(G.TIME), aggregation.Sum("-")), "intervalRolling.indexedAggregations: no column '-'"),
        ((), "intervalRolling.indexedAggregations: at least one column aggregation is required"),
        ((aggregation.WindowStart(G.TIME), aggregation.Sum("")),
         "intervalRolling.indexedAggregations: aggregation 1 has no column name"),
        ((aggregation.WindowStart(G.TIME), rolling.NewColAggregation(G.VALUE, False, B.Float64, lambda c, w: 1.0)),
         "intervalRolling.aggregateWindows: aggregation 1: custom closures are not supported by the GPU backend"),
    ]
    for aggrs, msg in cases:
        out = r.Aggregate(*aggrs)          # never raises: the error is deferred (rolling.go:245-248)
        with pytest.raises(B.BowError) as e:
            out.Bow()
        assert str(e.value) == msg
        with pytest.raises(B.BowError):
            out.NumWindows()
        # an errored Rolling stays errored through further steps (aggregation.go:124-126)
        with pytest.raises(B.BowError) as e2:
            out.Aggregate(aggregation.WindowStart(G.TIME)).Bow()
        assert str(e2.value) == msg
    # the receiver itself is untouched (value semantics: rCopy := *r)
    assert r.err is None


def test_aggregation_descriptor():  # aggregation.go:40-121
    a = aggregation.ArithmeticMean(G.VALUE)
    assert a.InputName() == G.VALUE and a.OutputName() == "" and a.InputIndex() == -1
    b = a.RenameOutput("x").SetTransformations(transformation.Factor(2))
    assert b.OutputName() == "x" and len(b.Transformations()) == 1 and a.OutputName() == "" and not a.Transformations()
    assert aggregation.First(G.VALUE).GetReturnType(B.Int64, B.Int64) == B.Int64
    assert aggregation.First(G.VALUE).GetReturnType(B.Float64, B.Int64) == B.Float64
    assert aggregation.WindowStart(G.TIME).GetReturnType(B.Float64, B.Int64) == B.Int64
    assert aggregation.Sum(G.VALUE).GetReturnType(B.Int64, B.Int64) == B.Float64
    assert aggregation.Count(G.VALUE).GetReturnType(B.Float64, B.Int64) == B.Int64
    assert aggregation.IntegralTrapezoid(G.VALUE).NeedInclusiveWindow()
    assert aggregation.WeightedAverageLinear(G.VALUE).NeedInclusiveWindow()
    assert not aggregation.IntegralStep(G.VALUE).NeedInclusiveWindow()
    assert not aggregation.WeightedAverageStep(G.VALUE).NeedInclusiveWindow()


def test_interpolate_deferred_errors():  # interpolation_test.go:23-47, linear_test.go:128-163
    r = rolling.IntervalRolling(two_col_bow([[10, 13], [1.0, 1.3]]), G.TIME, 2)
    bad = rolling.NewColInterpolation(G.VALUE, [B.Int64, B.Boolean], lambda *a: True)
    with pytest.raises(B.BowError) as e:
        r.Interpolate(interpolation.WindowStart(G.TIME), bad).Bow()
    assert str(e.value) == G.INTERP_DRIVER_ERRORS[0][1]
    with pytest.raises(B.BowError) as e:
        r.Interpolate(interpolation.Linear(G.VALUE)).Bow()
    assert str(e.value) == G.INTERP_DRIVER_ERRORS[1][1]
    with pytest.raises(B.BowError) as e:
        r.Interpolate().Bow()
    assert str(e.value) == "at least one column interpolation is required"
    for typ, msg in ((B.String, G.INTERP_TYPE_ERRORS[0][1]), (B.Boolean, G.INTERP_TYPE_ERRORS[1][1])):
        b = B.NewBowFromColBasedInterfaces([G.TIME, G.VALUE], [B.Int64, typ], [[10, 15], [None, None]])
        rr = rolling.IntervalRolling(b, G.TIME, 2)
        with pytest.raises(B.BowError) as e:
            rr.Interpolate(interpolation.WindowStart(G.TIME), interpolation.Linear(G.VALUE)).Bow()
        assert str(e.value) == msg
    with pytest.raises(B.BowError) as e:   # GPU backend restriction (documented divergence)
        r.Interpolate(interpolation.Linear(G.VALUE), interpolation.WindowStart(G.TIME)).Bow()
    assert "one interpolation per column, in schema order" in str(e.value)


@pytest.mark.parametrize("name,x,expected", G.FACTOR, ids=[c[0] for c in G.FACTOR])
def test_factor(name, x, expected):  # factor_test.go:9-35
    assert transformation.Factor(0.1)(x) == expected


def test_factor_invalid_type():
    with pytest.raises(B.BowError, match="factor: invalid type str"):
        transformation.Factor(0.1)("11")


def test_bow_equal_skips_all_null_rows():  # bow.go:257-262
    a = B.NewBowFromColBasedInterfaces(["a", "b"], [B.Int64, B.Float64], [[1, None, 3], [1.0, None, 3.0]])
    b = B.NewBowFromColBasedInterfaces(["a", "b"], [B.Int64, B.Float64], [[1, 3], [1.0, 3.0]])
    c = B.NewBowFromColBasedInterfaces(["a", "b"], [B.Int64, B.Float64], [[1, 3], [1.0, 3.5]])
    assert a.Equal(b) and not a.Equal(c)
