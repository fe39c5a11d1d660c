"""aggregation.Aggregate over the whole Bow (rolling/aggregation/whole.go): the literal oracle and the C restatement
against the reference's golden vectors (whole_test.go) and against each other on randomised inputs."""
import numpy as np
import pytest

from oracle import literal as L
from oracle import refc as R
from tests import helpers as H
from tests.golden import reference_vectors as G

CTORS = dict(WindowStart=L.WindowStart, Count=L.Count, Sum=L.Sum, ArithmeticMean=L.ArithmeticMean, Min=L.Min, Max=L.Max,
             First=L.First, Last=L.Last, IntegralStep=L.IntegralStep, IntegralTrapezoid=L.IntegralTrapezoid,
             WeightedAverageStep=L.WeightedAverageStep, WeightedAverageLinear=L.WeightedAverageLinear)


def whole_frame(rows):
    return L.Frame(["time", "value"], [L.INT64, L.FLOAT64], [[r[0] for r in rows], [r[1] for r in rows]])


@pytest.mark.parametrize("name,rows,aggs,expected,cite", G.WHOLE_CASES, ids=[c[0] for c in G.WHOLE_CASES])
def test_whole_golden_literal(name, rows, aggs, expected, cite):
    b = whole_frame(rows)
    la = []
    for ctor, col, out in aggs:
        a = CTORS[ctor](col)
        la.append(a.rename_output(out) if out else a)
    if isinstance(expected, str):
        with pytest.raises((KeyError, ValueError)) as ei:
            L.whole_aggregate(b, "time", *la)
        assert ei.value.args[0] == expected, cite
        return
    got = L.whole_aggregate(b, "time", *la)
    assert got.names == expected["names"] and got.types == expected["types"], cite
    H.assert_cols_equal(got.cols, expected["cols"], cite)


@pytest.mark.parametrize("name,rows,aggs,expected,cite", [c for c in G.WHOLE_CASES if not isinstance(c[3], str)],
                         ids=[c[0] for c in G.WHOLE_CASES if not isinstance(c[3], str)])
def test_whole_golden_refc(name, rows, aggs, expected, cite):
    t = np.array([r[0] for r in rows], dtype=np.int64)
    v = np.array([0.0 if r[1] is None else r[1] for r in rows], dtype=np.float64)
    m = np.array([r[1] is not None for r in rows], dtype=bool)
    cols = [(t, None), (v, m)]
    specs = [(ctor, 0 if col == "time" else 1) for ctor, col, _ in aggs]
    got = R.aggregate_whole(R.Frame(cols), 0, specs)
    for (gv, gm), want in zip(got, expected["cols"]):
        assert [x if ok else None for x, ok in zip(gv.tolist(), gm.tolist())] == want, cite


ALL = ["WindowStart", "Count", "Sum", "ArithmeticMean", "Min", "Max", "First", "Last", "IntegralStep", "IntegralTrapezoid",
       "WeightedAverageStep", "WeightedAverageLinear"]


@pytest.mark.parametrize("kind", ["regular", "dense", "sparse"])
def test_whole_refc_matches_literal(kind):
    rng = np.random.default_rng(H.seed_of(kind))
    for trial in range(25):
        n = int(rng.integers(0, 60))
        t = H.random_times(rng, n, kind)
        vf = H.random_values(rng, n, np.float64, float(rng.choice([0.0, 0.3, 1.0])), specials=trial % 3 == 0)
        vi = H.random_values(rng, n, np.int64, float(rng.choice([0.0, 0.5])))
        cols = [(t, None), vf, vi]
        frame = H.literal_frame(cols, ["time", "f", "i"])
        specs, la = [], []
        for c, cname in ((0, "time"), (1, "f"), (2, "i")):
            for op in ALL:
                fac = [2.5] if (trial + c) % 4 == 0 and op in ("Sum", "Count", "First", "WindowStart") else None
                specs.append((op, c, fac))
                a = CTORS[op](cname).rename_output(f"{op}_{cname}")
                la.append(a.set_transformations(L.Factor(2.5)) if fac else a)
        want = L.whole_aggregate(frame, "time", *la)
        got = R.aggregate_whole(R.Frame(cols), 0, specs)
        for j, (gv, gm) in enumerate(got):
            g = [x if ok else None for x, ok in zip(gv.tolist(), gm.tolist())]
            H.assert_cols_equal([g], [want.cols[j]], f"{kind} trial {trial} spec {specs[j]}")
