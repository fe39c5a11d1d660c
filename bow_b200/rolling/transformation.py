"""rolling/transformation (reference rolling/transformation/factor.go:5-20)."""
from __future__ import annotations

from typing import Any, Callable

from ..bow import BowError

Func = Callable[[Any], Any]


class _Factor:
    """transformation.Factor(n): float64 -> x*n, int64 -> int64(float64(x)*n), nil -> nil.
    Carries `n` so that Rolling.Aggregate can fuse it into the device epilogue."""

    def __init__(self, n: float):
        self.n = float(n)

    def __call__(self, x):
        if x is None:
            return None
        if isinstance(x, float):
            return x * self.n
        if isinstance(x, int) and not isinstance(x, bool):
            return int(float(x) * self.n)  # truncation toward zero, like Go's int64(float64)
        raise BowError(f"factor: invalid type {type(x).__name__}")


def Factor(n: float) -> Func:
    return _Factor(n)
