"""Bow.SortByCol on the GPU (bowgpu_frame_sort_by_col and the mirrored Bow.SortByCol) against the reference's golden
vectors (bowsort_test.go:11-208) and the oracle.  Bit-exact: a permutation of rows.  Equal keys keep their input order
(see oracle/literal.py sort_by_col for why that is the pinned part of the reference's unstable sort)."""
import numpy as np
import pytest

from bow_b200 import bow as B
from bow_b200 import native as N
from oracle import refc as R
from tests import helpers as H
from tests.golden import reference_vectors as G

pytestmark = pytest.mark.gpu
TYPES = {"i": B.Int64, "f": B.Float64}


@pytest.fixture(scope="module")
def ctx():
    c = N.Ctx(0)
    yield c
    c.close()


def bow_of(rows, types):
    cols = [[(None if r[c] is None else (int(r[c]) if t == "i" else float(r[c]))) for r in rows] for c, t in enumerate(types)]
    return B.NewBowFromColBasedInterfaces([f"c{i}" for i in range(len(types))], [TYPES[t] for t in types], cols)


@pytest.mark.parametrize("name,types,rows,col,expected,cite", G.SORT_CASES, ids=[c[0] for c in G.SORT_CASES])
def test_sort_by_col_golden_api(name, types, rows, col, expected, cite):
    b = bow_of(rows, types)
    if expected == "error":
        with pytest.raises(B.BowError, match="column to sort by has 1 nil values"):
            b.SortByCol(col)
        return
    got = b.SortByCol(col)
    if expected == "same":
        assert got is b, cite
        return
    want = bow_of(expected, types)
    assert got.ToColBased() == want.ToColBased(), f"{cite}\nexpected: {want}\nactual: {got}"
    assert [got.ColumnType(i) for i in range(len(types))] == [TYPES[t] for t in types]


def test_sort_keeps_metadata():
    import pyarrow as pa
    rec = pa.RecordBatch.from_arrays([pa.array([1, 3, 2], type=pa.int64()), pa.array([.1, .3, .2])], names=["time", "value"])
    b = B.Bow(rec.replace_schema_metadata({"k": "v"}))
    out = b.SortByCol(0)        # bowsort_test.go:171-188
    assert out.ToColBased() == [[1, 2, 3], [.1, .2, .3]] and out.Metadata() == b.Metadata()
    with pytest.raises(B.BowError):
        b.SortByCol(5)


def check_against_oracle(ctx, cols, col, what):
    fr = N.Frame.from_numpy(ctx, cols)
    try:
        out = fr.sort_by_col(col)
        want = R.sort_by_col(cols, col)
        if want is None:
            assert out is None, what
            return
        assert out is not None, what
        got = out.download()
        out.close()
    finally:
        fr.close()
    for j, ((gv, gm), (wv, wm)) in enumerate(zip(got, want)):
        assert gv.dtype == wv.dtype and np.array_equal(gm, wm), f"{what} col {j}: validity"
        assert np.array_equal(gv[gm].view(np.int64), wv[wm].view(np.int64)), f"{what} col {j}: values"


KEYS = ["unique", "dups", "few", "float", "float_dups", "negzero", "narrow", "reversed", "nearly"]


@pytest.mark.parametrize("kind", KEYS)
@pytest.mark.parametrize("n", [2, 31, 33, 4095, 4096, 4097, 8193, 100_003])
def test_sort_by_col_matches_oracle(ctx, kind, n):
    rng = np.random.default_rng(H.seed_of("sort", kind, n))
    if kind == "unique":
        key = rng.permutation(n).astype(np.int64) * 1_000_003 - 5_000_000
    elif kind == "dups":
        key = rng.integers(-2 ** 62, 2 ** 62, size=max(1, n // 3))[rng.integers(0, max(1, n // 3), size=n)]
    elif kind == "few":
        key = rng.integers(-3, 4, size=n)
    elif kind == "float":
        key = rng.standard_normal(n) * 10.0 ** rng.integers(-300, 300, size=n)
    elif kind == "float_dups":
        key = rng.integers(-20, 20, size=n).astype(np.float64) / 4
    elif kind == "negzero":
        key = rng.choice(np.array([0.0, -0.0, 1.5, -1.5, np.inf, -np.inf]), size=n)
    elif kind == "narrow":        # only the low digit varies: seven passes are skipped
        key = (1 << 40) + rng.integers(0, 200, size=n)
    elif kind == "reversed":
        key = np.arange(n, 0, -1, dtype=np.int64) * 7
    else:                          # sorted but for one swap
        key = np.arange(n, dtype=np.int64)
        key[[0, n - 1]] = key[[n - 1, 0]]
    cols = [H.random_values(rng, n, np.float64, 0.3), (np.ascontiguousarray(key), None), H.random_values(rng, n, np.int64, 0.0),
            (np.arange(n, dtype=np.int64), None)]
    check_against_oracle(ctx, cols, 1, f"{kind} n={n}")


def test_sort_by_col_sorted_and_errors(ctx):
    n = 10_000
    t = np.arange(n, dtype=np.int64) // 3          # ascending with ties: sort.IsSorted
    fr = N.Frame.from_numpy(ctx, [(t, None), (np.ones(n), None)])
    assert fr.sort_by_col(0) is None
    with pytest.raises(N.BowGpuError):
        fr.sort_by_col(2)
    fr.close()
    v = np.array([1.0, np.nan, 2.0, 3.0])           # `<` never fires on NaN: sorted for sort.IsSorted
    fr = N.Frame.from_numpy(ctx, [(v, None)])
    assert fr.sort_by_col(0) is None
    fr.close()
    fr = N.Frame.from_numpy(ctx, [(np.array([3.0, np.nan, 2.0, 1.0]), None)])
    with pytest.raises(N.BowGpuError, match="NaN"):
        fr.sort_by_col(0)
    fr.close()
    m = np.ones(4, dtype=bool)
    m[2] = False
    fr = N.Frame.from_numpy(ctx, [(np.array([4, 3, 2, 1], dtype=np.int64), m)])
    with pytest.raises(N.BowGpuError, match="column to sort by has 1 nil values"):
        fr.sort_by_col(0)
    fr.close()


def test_sort_by_col_large_properties(ctx):
    """2e7 rows (size-independent checks): keys ascend, equal keys keep their row order, the payload follows its key"""
    n = int(2e7 * float(__import__("os").environ.get("BOW_TEST_SCALE", "1")))
    rng = np.random.default_rng(7)
    key = rng.integers(0, n // 4, size=n).astype(np.int64) * 1000 - 12345
    payload = key * 3 + 1
    m = rng.random(n) > 0.1
    fr = N.Frame.from_numpy(ctx, [(key, None), (payload.astype(np.float64), m), (np.arange(n, dtype=np.int64), None)])
    out = fr.sort_by_col(0)
    got = out.download()
    out.close()
    fr.close()
    k, (p, pm), idx = got[0][0], got[1], got[2][0]
    assert np.all(k[1:] >= k[:-1])
    assert np.all((k[1:] != k[:-1]) | (idx[1:] > idx[:-1]))
    assert np.array_equal(k, key[idx]) and np.array_equal(pm, m[idx])
    assert np.array_equal(p[pm], (k * 3 + 1).astype(np.float64)[pm])
