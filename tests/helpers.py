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
