"""Bow.DropNils / Bow.IsColSorted: the oracle restatements against the reference's golden vectors
(bow_test.go:165-282, bowassertion_test.go:11-55)."""
import numpy as np
import pytest

from oracle import literal as L
from oracle import refc as R
from tests import helpers as H
from tests.golden import reference_vectors as G


@pytest.mark.parametrize("name,cols,sel,expected,cite", G.DROP_CASES, ids=[c[0] for c in G.DROP_CASES])
def test_drop_nils_golden(name, cols, sel, expected, cite):
    names = list("abc")[:len(cols)]
    fr = L.Frame(names, [L.INT64] * len(cols), cols)
    assert L.drop_nils(fr, *sel).materialize() == expected, cite
    npc = H.np_cols_from_lists(cols, [L.INT64] * len(cols))
    got = R.drop_nils(npc, sel)
    assert H.lists_from_np([(v, m) for v, m in got]) == expected, cite


@pytest.mark.parametrize("typ", [L.INT64, L.FLOAT64])
def test_is_col_sorted_golden(typ):
    conv = (lambda v: v) if typ == L.INT64 else (lambda v: None if v is None else float(v))
    cols = [[conv(r[c]) for r in G.SORTED_ROWS] for c in range(5)]
    fr = L.Frame(list("abcde"), [typ] * 5, cols)
    npc = H.np_cols_from_lists(cols, [typ] * 5)
    for c, want in enumerate(G.SORTED_EXPECTED):
        assert L.is_col_sorted(fr, c) == want
        assert R.is_col_sorted(*npc[c]) == want
    assert not L.is_col_sorted(L.Frame(["a"], [typ], [[None, None]]), 0)      # empty column, bowassertion.go:16-18
    assert not R.is_col_sorted(np.zeros(2), np.zeros(2, dtype=bool))
