"""rolling/aggregation: the built-in ColAggregation constructors (reference rolling/aggregation/*.go).

Each constructor returns the same descriptor as the reference (input column, NeedInclusiveWindow,
return type) plus the kernel opcode the C ABI understands (include/bowgpu.h BOWGPU_AGG_*); the Go
closure is replaced by the CUDA segmented-reduction family.  Mode (mode.go) and the whole-frame
Aggregate (whole.go) are out of scope."""
from __future__ import annotations

from .. import bow as B
from .. import native as N
from . import ColAggregation


def _mk(col: str, inclusive: bool, typ: B.Type, op: str) -> ColAggregation:
    return ColAggregation(col, inclusive, typ, None, kernel_op=N.AGG[op])


def WindowStart(col: str) -> ColAggregation:            # windowstart.go:8-13
    return _mk(col, False, B.IteratorDependent, "WindowStart")


def Count(col: str) -> ColAggregation:                  # count.go:8-20
    return _mk(col, False, B.Int64, "Count")


def Sum(col: str) -> ColAggregation:                    # sum.go:8-25
    return _mk(col, False, B.Float64, "Sum")


def ArithmeticMean(col: str) -> ColAggregation:         # arithmeticmean.go:8-30
    return _mk(col, False, B.Float64, "ArithmeticMean")


def Min(col: str) -> ColAggregation:                    # minmax.go:8-31
    return _mk(col, False, B.Float64, "Min")


def Max(col: str) -> ColAggregation:                    # minmax.go:33-56
    return _mk(col, False, B.Float64, "Max")


def First(col: str) -> ColAggregation:                  # firstlast.go:8-21
    return _mk(col, False, B.InputDependent, "First")


def Last(col: str) -> ColAggregation:                   # firstlast.go:23-36
    return _mk(col, False, B.InputDependent, "Last")


def IntegralStep(col: str) -> ColAggregation:           # integral.go:40-69
    return _mk(col, False, B.Float64, "IntegralStep")


def IntegralTrapezoid(col: str) -> ColAggregation:      # integral.go:8-38
    return _mk(col, True, B.Float64, "IntegralTrapezoid")


def WeightedAverageStep(col: str) -> ColAggregation:    # weightedmean.go:8-20
    return _mk(col, False, B.Float64, "WeightedAverageStep")


def WeightedAverageLinear(col: str) -> ColAggregation:  # weightedmean.go:22-34
    return _mk(col, True, B.Float64, "WeightedAverageLinear")
