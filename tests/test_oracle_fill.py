"""Whole-column fills (bowfill.go): literal oracle and C restatement against the reference's golden vectors
(bowfill_test.go) and against each other on randomised inputs."""
import numpy as np
import pytest

from oracle import literal as L
from oracle import refc as R
from tests import helpers as H
from tests.golden import reference_vectors as G

LIT = dict(FillMean=L.fill_mean, FillNext=L.fill_next, FillPrevious=L.fill_previous, FillLinear=L.fill_linear)
METHOD = dict(FillMean="Mean", FillNext="Next", FillPrevious="Previous", FillLinear="Linear")


def fresh(typ):
    conv = (lambda v: v) if typ == L.INT64 else (lambda v: None if v is None else float(v))
    cols = [[conv(r[c]) for r in G.FILL_ROWS] for c in range(5)]
    return L.Frame(list("abcde"), [typ] * 5, cols)


def expected_cols(rows, typ):
    conv = (lambda v: v) if typ == L.INT64 else (lambda v: None if v is None else float(v))
    return [[conv(r[c]) for r in rows] for c in range(5)]


@pytest.mark.parametrize("typ", [L.INT64, L.FLOAT64])
@pytest.mark.parametrize("name,method,args,exp_i,exp_f,cite", G.FILL_CASES, ids=[c[0] for c in G.FILL_CASES])
def test_fill_golden_literal(typ, name, method, args, exp_i, exp_f, cite):
    exp = exp_i if typ == L.INT64 or exp_f is None else exp_f
    if exp == "error":
        with pytest.raises((ValueError, TypeError)):
            LIT[method](fresh(typ), *args)
        return
    got = LIT[method](fresh(typ), *args)
    H.assert_cols_equal(got.materialize(), expected_cols(exp, typ), cite)


@pytest.mark.parametrize("typ", [L.INT64, L.FLOAT64])
@pytest.mark.parametrize("name,method,args,exp_i,exp_f,cite", [c for c in G.FILL_CASES if c[3] != "error"],
                         ids=[c[0] for c in G.FILL_CASES if c[3] != "error"])
def test_fill_golden_refc(typ, name, method, args, exp_i, exp_f, cite):
    exp = expected_cols(exp_i if typ == L.INT64 or exp_f is None else exp_f, typ)
    cols = H.np_cols_from_lists(fresh(typ).materialize(), [typ] * 5)
    fr = R.Frame(cols)
    targets = list(args) if method != "FillLinear" else [args[1]]
    for c in (targets or range(5)):
        v, m = R.fill(fr, METHOD[method], c, args[0] if method == "FillLinear" else -1)
        got = [x if ok else None for x, ok in zip(v.tolist(), m.tolist())]
        H.assert_cols_equal([got], [exp[c]], f"{cite} col {c}")


def test_go_round():
    for x, want in ((0.5, 1.0), (-0.5, -1.0), (1.5, 2.0), (2.5, 3.0), (-2.5, -3.0), (0.49999999999999994, 0.0),
                    (-0.49999999999999994, -0.0), (4503599627370497.0, 4503599627370497.0), (-7.5, -8.0), (1e300, 1e300)):
        assert H.same_value(L.go_round(x), want), x


@pytest.mark.parametrize("null_p", [0.0, 0.3, 0.9, 1.0])
def test_fill_refc_matches_literal(null_p):
    rng = np.random.default_rng(H.seed_of("fill", null_p))
    for trial in range(12):
        n = int(rng.integers(0, 80))
        ref = np.sort(rng.integers(-50, 50, size=n)).astype(np.int64)
        if trial % 2:
            ref = ref[::-1].copy()
        refm = rng.random(n) > 0.1 if trial % 3 == 0 else None
        vf = H.random_values(rng, n, np.float64, null_p, specials=trial % 4 == 0)
        vi = H.random_values(rng, n, np.int64, null_p)
        cols = [(ref, refm), vf, vi]
        frame = H.literal_frame(cols, ["ref", "f", "i"])
        fr = R.Frame(cols)
        for method, fn in (("Previous", L.fill_previous), ("Next", L.fill_next), ("Mean", L.fill_mean)):
            want = fn(frame, 1, 2).materialize()
            for c in (1, 2):
                v, m = R.fill(fr, method, c)
                H.assert_cols_equal([[x if ok else None for x, ok in zip(v.tolist(), m.tolist())]], [want[c]],
                                    f"{method} trial {trial} col {c}")
        if n and not L.is_col_empty(frame, 0):
            for c in (1, 2):
                want = L.fill_linear(frame, 0, c).materialize()
                v, m = R.fill(fr, "Linear", c, 0)
                H.assert_cols_equal([[x if ok else None for x, ok in zip(v.tolist(), m.tolist())]], [want[c]],
                                    f"Linear trial {trial} col {c}")
