"""Bow.SortByCol: the oracle restatements against the reference's golden vectors (bowsort_test.go:11-208) and against
each other on random inputs with duplicate keys (equal keys keep their input order in both)."""
import numpy as np
import pytest

from oracle import literal as L
from oracle import refc as R
from tests import helpers as H
from tests.golden import reference_vectors as G

TYPES = {"i": L.INT64, "f": L.FLOAT64}


def columns_of(rows, types):
    return [[(None if r[c] is None else (int(r[c]) if t == "i" else float(r[c]))) for r in rows] for c, t in enumerate(types)]


@pytest.mark.parametrize("name,types,rows,col,expected,cite", G.SORT_CASES, ids=[c[0] for c in G.SORT_CASES])
def test_sort_by_col_golden(name, types, rows, col, expected, cite):
    typs = [TYPES[t] for t in types]
    cols = columns_of(rows, types)
    fr = L.Frame([f"c{i}" for i in range(len(types))], typs, cols)
    npc = H.np_cols_from_lists(cols, typs)
    if expected == "error":
        with pytest.raises(ValueError, match="column to sort by has 1 nil values"):
            L.sort_by_col(fr, col)
        with pytest.raises(ValueError, match="column to sort by has 1 nil values"):
            R.sort_by_col(npc, col)
        return
    got = L.sort_by_col(fr, col)
    if expected == "same":
        assert got is fr, cite
        assert R.sort_by_col(npc, col) is None, cite
        return
    want = columns_of(expected, types)
    assert got.materialize() == want, cite
    assert H.lists_from_np(R.sort_by_col(npc, col)) == want, cite


@pytest.mark.parametrize("seed", range(12))
def test_sort_by_col_restatements_agree(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 300))
    key = rng.integers(-5, 6, size=n) if seed % 2 else rng.integers(-2 ** 62, 2 ** 62, size=n)
    if seed % 3 == 0:
        key = key.astype(np.float64) / 2
        key[rng.random(n) < 0.1] = -0.0
    cols = [(key, None), H.random_values(rng, n, np.float64, 0.3), (np.arange(n, dtype=np.int64), None)]
    want = L.sort_by_col(H.literal_frame(cols), 0)
    got = R.sort_by_col(cols, 0)
    if got is None:
        assert want.materialize() == H.lists_from_np(cols)
        return
    for (gv, gm), wl in zip(got, want.materialize()):
        for g, ok, w in zip(gv.tolist(), gm.tolist(), wl):
            assert (w is None and not ok) or (ok and H.same_value(g, w))
    # stability: among equal keys the original row numbers ascend
    k, idx = got[0][0], got[2][0]
    assert np.all((k[1:] != k[:-1]) | (idx[1:] > idx[:-1]))
