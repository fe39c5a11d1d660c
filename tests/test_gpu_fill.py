"""Whole-column fills on the GPU (bowgpu_frame_fill / bowgpu_frame_fill_linear and the mirrored Bow.Fill* methods)
against the reference's golden vectors (bowfill_test.go) and the oracle.  Bit-exact: FillPrevious / FillNext copy
values, FillMean / FillLinear evaluate the reference's float64 expressions in the same order without FMA."""
import numpy as np
import pytest

from bow_b200 import bow as B
from oracle import literal as L
from oracle import refc as R
from tests import helpers as H
from tests.golden import reference_vectors as G

pytestmark = pytest.mark.gpu
METHOD = dict(FillMean="Mean", FillNext="Next", FillPrevious="Previous", FillLinear="Linear")


@pytest.fixture(scope="module")
def ctx():
    from bow_b200 import native as N
    c = N.Ctx(0)
    yield c
    c.close()


def fresh_bow(typ):
    conv = (lambda v: v) if typ == B.Int64 else (lambda v: None if v is None else float(v))
    return B.NewBowFromColBasedInterfaces(list("abcde"), [typ] * 5, [[conv(r[c]) for r in G.FILL_ROWS] for c in range(5)])


@pytest.mark.parametrize("typ", [B.Int64, B.Float64], ids=["int64", "float64"])
@pytest.mark.parametrize("name,method,args,exp_i,exp_f,cite", G.FILL_CASES, ids=[c[0] for c in G.FILL_CASES])
def test_fill_golden_api(typ, name, method, args, exp_i, exp_f, cite):
    b = fresh_bow(typ)
    exp = exp_i if typ == B.Int64 or exp_f is None else exp_f
    if exp == "error":
        with pytest.raises(B.BowError):
            getattr(b, method)(*args)
        return
    got = getattr(b, method)(*args)
    conv = (lambda v: v) if typ == B.Int64 else (lambda v: None if v is None else float(v))
    want = B.NewBowFromColBasedInterfaces(list("abcde"), [typ] * 5, [[conv(r[c]) for r in exp] for c in range(5)])
    assert got.ToColBased() == want.ToColBased(), f"{cite}\nexpected: {want}\nactual: {got}"
    assert [got.ColumnType(i) for i in range(5)] == [typ] * 5


def test_fill_keeps_metadata_and_errors():
    import pyarrow as pa
    rec = pa.RecordBatch.from_arrays([pa.array([1, None, 3], type=pa.int64()), pa.array([1.0, None, 3.0])], names=["int", "float"])
    b = B.Bow(rec.replace_schema_metadata({"k": "v"}))
    out = b.FillPrevious()       # bowfill_test.go:483-500
    assert out.ToColBased() == [[1, 1, 3], [1.0, 1.0, 3.0]] and out.Metadata() == b.Metadata()
    assert b.FillNext().ToColBased() == [[1, 3, 3], [1.0, 3.0, 3.0]]
    with pytest.raises(B.BowError, match="selectCols: colIndex '7' out of range"):
        b.FillMean(7)
    with pytest.raises(B.BowError, match="refColIndex and toFillColIndex are equal"):
        b.FillLinear(1, 1)


def assert_col(got, want, what):
    (gv, gm), (wv, wm) = got, want
    assert gv.dtype == wv.dtype, what
    assert np.array_equal(gm, wm), f"{what}: validity differs at {np.flatnonzero(gm != wm)[:8]}"
    a, b = gv[gm], wv[wm]
    same = a.view(np.int64) == b.view(np.int64)
    if gv.dtype == np.float64:
        same |= np.isnan(a) & np.isnan(b)
    assert same.all(), f"{what}: rows {np.flatnonzero(gm)[~same][:8]} got {a[~same][:4]} want {b[~same][:4]}"


@pytest.mark.parametrize("null_p", [0.0, 0.2, 0.97, 1.0])
@pytest.mark.parametrize("n", [0, 1, 31, 33, 8191, 8192, 8193, 70001])
def test_fill_random_vs_oracle(ctx, n, null_p):
    """all four methods on a float64 and an int64 column; long null runs cross 32-row words, 8192-row blocks and the
    column ends; FillLinear with ascending / descending reference columns, with and without nulls of their own"""
    from bow_b200 import native as N
    rng = np.random.default_rng(H.seed_of("gpufill", n, null_p))
    ref = np.sort(rng.integers(-10 * max(n, 1), 10 * max(n, 1), size=n)).astype(np.int64)
    if n % 2:
        ref = ref[::-1].copy()
    refm = (rng.random(n) > 0.05) if n % 3 == 0 else None
    vf = H.random_values(rng, n, np.float64, null_p, specials=n < 100)
    vi = H.random_values(rng, n, np.int64, null_p)
    if n > 20000 and null_p > 0.5:      # one null run longer than two blocks
        vf[1][1000:19000] = False
        vi[1][5:17000] = False
    cols = [(ref, refm), vf, vi]
    fr = N.Frame.from_numpy(ctx, cols)
    rfr = R.Frame(cols)
    for method in ("Previous", "Next", "Mean"):
        out = fr.fill(method, 1, 2)
        got = out.download()
        out.close()
        for c in (1, 2):
            assert_col(got[c], R.fill(rfr, method, c), f"{method} n={n} p={null_p} col {c}")
        assert_col(got[0], (ref, np.ones(n, bool) if refm is None else refm), "untouched column")
    if n and (refm is None or refm.any()):
        for c in (1, 2):
            out = fr.fill_linear(0, c)
            got = out.download()
            out.close()
            assert_col(got[c], R.fill(rfr, "Linear", c, 0), f"Linear n={n} p={null_p} col {c}")
    fr.close()


def test_fill_linear_unsorted_reference(ctx):
    from bow_b200 import native as N
    ref = np.array([1, 5, 3, 7], dtype=np.int64)
    v = (np.array([1.0, 0.0, 3.0, 4.0]), np.array([True, False, True, True]))
    fr = N.Frame.from_numpy(ctx, [(ref, None), v])
    with pytest.raises(N.BowGpuError, match="EUNSORTED"):
        fr.fill_linear(0, 1)
    # nulls in the reference column are skipped by IsColSorted (bowassertion.go:15-81)
    fr2 = N.Frame.from_numpy(ctx, [(ref, np.array([True, True, False, True])), v])
    out = fr2.fill_linear(0, 1)
    assert out.num_rows == 4


# ---- Bow.DropNils / Bow.IsColSorted -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name,cols,sel,expected,cite", G.DROP_CASES, ids=[c[0] for c in G.DROP_CASES])
def test_drop_nils_golden_api(name, cols, sel, expected, cite):
    names = list("abc")[:len(cols)]
    b = B.NewBowFromColBasedInterfaces(names, [B.Int64] * len(cols), cols)
    got = b.DropNils(*sel)
    assert got.ToColBased() == expected, cite


def test_is_col_sorted_golden_api():
    for typ, conv in ((B.Int64, lambda v: v), (B.Float64, lambda v: None if v is None else float(v))):
        b = B.NewBowFromColBasedInterfaces(list("abcde"), [typ] * 5, [[conv(r[c]) for r in G.SORTED_ROWS] for c in range(5)])
        assert [b.IsColSorted(c) for c in range(5)] == G.SORTED_EXPECTED      # bowassertion_test.go:24-33,45-54


@pytest.mark.parametrize("n", [1, 33, 8191, 8193, 100000])
def test_drop_nils_random_vs_oracle(ctx, n):
    from bow_b200 import native as N
    rng = np.random.default_rng(H.seed_of("dropnils", n))
    for null_p, sel in ((0.0, ()), (0.1, ()), (0.6, (1,)), (0.97, (0, 2)), (1.0, (2,))):
        t = (np.cumsum(rng.integers(0, 4, size=n)), rng.random(n) >= null_p / 2)
        cols = [(t[0].astype(np.int64), t[1]), H.random_values(rng, n, np.float64, null_p),
                H.random_values(rng, n, np.int64, null_p / 3)]
        fr = N.Frame.from_numpy(ctx, cols)
        out = fr.drop_nils(*sel)
        got = out.download()
        want = R.drop_nils(cols, sel)
        assert out.num_rows == len(want[0][0])
        for c in range(3):
            assert_col(got[c], (want[c][0], want[c][1]), f"DropNils n={n} p={null_p} sel={sel} col {c}")
        # the surviving time column can go straight into the rolling path: sortedness as the reference sees it
        assert out.is_col_sorted(0) == R.is_col_sorted(want[0][0], want[0][1]) if out.num_rows else True
        out.close()
        fr.close()
