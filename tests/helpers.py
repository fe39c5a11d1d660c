"""Shared helpers for the parity tests: random frames, literal<->numpy conversions, comparisons."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

from oracle import literal as L

NpCol = Tuple[np.ndarray, Optional[np.ndarray]]


def np_cols_from_lists(cols: Sequence[Sequence], types: Sequence[str]) -> List[NpCol]:
    out = []
    for c, t in zip(cols, types):
        dt = np.int64 if t == L.INT64 else np.float64
        vals = np.array([0 if v is None else v for v in c], dtype=dt)
        mask = np.array([v is not None for v in c], dtype=bool)
        out.append((vals, mask if not mask.all() or len(c) == 0 else None))
    return out


def lists_from_np(cols: Sequence[NpCol]) -> List[list]:
    res = []
    for v, m in cols:
        lst = v.tolist()
        if m is not None:
            lst = [x if ok else None for x, ok in zip(lst, m.tolist())]
        res.append(lst)
    return res


def literal_frame(cols: Sequence[NpCol], names=None) -> L.Frame:
    names = names or [f"c{i}" for i in range(len(cols))]
    types = [L.INT64 if v.dtype == np.int64 else L.FLOAT64 for v, _ in cols]
    return L.Frame(list(names), types, lists_from_np(cols))


def same_value(a, b) -> bool:
    """bit-level equality for floats (distinguishes -0.0/+0.0, treats NaN == NaN)"""
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, float) or isinstance(b, float):
        a, b = float(a), float(b)
        if math.isnan(a) or math.isnan(b):
            return math.isnan(a) and math.isnan(b)
        return a == b and math.copysign(1.0, a) == math.copysign(1.0, b)
    return a == b


def assert_cols_equal(got: Sequence[list], want: Sequence[list], what=""):
    assert len(got) == len(want), (what, len(got), len(want))
    for j, (g, w) in enumerate(zip(got, want)):
        assert len(g) == len(w), (what, j, len(g), len(w))
        for i, (a, b) in enumerate(zip(g, w)):
            assert same_value(a, b), f"{what} col {j} row {i}: got {a!r} want {b!r}"


def random_times(rng: np.random.Generator, n: int, kind: str) -> np.ndarray:
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    if kind == "dense":          # many duplicates, tiny steps
        steps = rng.integers(0, 3, size=n)
    elif kind == "sparse":       # many empty windows
        steps = rng.integers(0, 40, size=n)
    elif kind == "bursty":
        steps = np.where(rng.random(n) < 0.1, rng.integers(20, 200, size=n), rng.integers(0, 2, size=n))
    else:
        steps = rng.integers(1, 10, size=n)
    start = int(rng.integers(-50, 50))
    return (start + np.cumsum(steps)).astype(np.int64)


SPECIALS = np.array([0.0, -0.0, math.nan, math.inf, -math.inf, 1.5, -2.25, 1e300, -1e300, 5e-324])


def random_values(rng: np.random.Generator, n: int, dtype, null_p: float, specials: bool = False) -> NpCol:
    if dtype == np.int64:
        v = rng.integers(-1000, 1000, size=n).astype(np.int64)
    else:
        v = rng.normal(size=n) * 10.0
        if specials and n:
            idx = rng.random(n) < 0.3
            v = np.where(idx, SPECIALS[rng.integers(0, len(SPECIALS), size=n)], v)
    mask = None
    if null_p > 0:
        mask = rng.random(n) >= null_p
    return v, mask


def seed_of(*key) -> int:
    """deterministic seed from a test's parameters (Python's hash() of strings changes from run to run)"""
    import zlib
    return zlib.crc32(repr(key).encode())


# ---- the stated tolerance class of the float reductions -------------------------------------------------------------
TOL = 1e-12
TOL_OPS = ("Sum", "ArithmeticMean", "IntegralStep", "IntegralTrapezoid", "WeightedAverageStep", "WeightedAverageLinear")


def term_scales(cols, vcol, interval, offset=0, inclusive=False, time_col=0):
    """sum of |term_i| per window for every float reduction of column `vcol` (the second argument of the parity bar
    |gpu - ref| <= 1e-12 * max(|ref|, sum|term_i|)): the oracle run on |v|.  Sum: sum|v|; ArithmeticMean: sum|v| / count;
    integrals: the integral of |v|; weighted averages: that over the window width."""
    from oracle import refc as R
    v, m = cols[vcol]
    av = np.abs(v.astype(np.float64))
    av = np.where(np.isfinite(av), av, 0.0)
    ref = R.RefRolling(R.Frame([cols[time_col], (av, m)]), 0, interval, offset=offset, inclusive=inclusive)
    out = ref.aggregate([("WindowStart", 0), ("Sum", 1), ("Count", 1), ("IntegralStep", 1), ("IntegralTrapezoid", 1)])
    s = out[1][0]
    cnt = np.maximum(out[2][0], 1).astype(np.float64)
    integ = np.maximum(np.where(out[3][1], np.abs(out[3][0]), 0.0), np.where(out[4][1], np.abs(out[4][0]), 0.0))
    return {"Sum": s, "ArithmeticMean": s / cnt, "IntegralStep": integ, "IntegralTrapezoid": integ,
            "WeightedAverageStep": integ / float(interval), "WeightedAverageLinear": integ / float(interval)}


def assert_in_tolerance_class(gv, wv, scale, what="", factor=1.0):
    """|gpu - ref| <= 1e-12 * max(|ref|, scale) elementwise (bit-identical values, incl. NaN / Inf, always pass);
    `factor`: product of the |transformation.Factor|s applied to the output"""
    gv, wv = np.asarray(gv, dtype=np.float64), np.asarray(wv, dtype=np.float64)
    same = (gv.view(np.int64) == wv.view(np.int64)) | (np.isnan(gv) & np.isnan(wv))
    with np.errstate(invalid="ignore"):
        ok = np.abs(gv - wv) <= TOL * np.maximum(np.abs(wv), np.asarray(scale, dtype=np.float64) * factor)
    bad = np.flatnonzero(~same & ~ok)
    assert bad.size == 0, f"{what}: {bad[:10]} got {gv[bad[:10]]} want {wv[bad[:10]]}"
