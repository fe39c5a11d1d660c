"""rolling/aggregation: the built-in ColAggregation constructors (reference rolling/aggregation/*.go).

Each constructor returns the same descriptor as the reference (input column, NeedInclusiveWindow,
return type) plus the kernel opcode the C ABI understands (include/bowgpu.h BOWGPU_AGG_*); the Go
closure is replaced by the CUDA segmented-reduction family.  `Aggregate` is the whole-frame variant
(whole.go).  Mode (mode.go) is out of scope."""
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


def Aggregate(b: "B.Bow", intervalColName: str, *aggrs: ColAggregation) -> "B.Bow":
    """aggregation.Aggregate (rolling/aggregation/whole.go:12-93): the whole dataframe as ONE window.
    Go's `(Bow, error)` becomes a returned Bow or a raised BowError carrying the reference's message."""
    import numpy as np

    from . import _cols_from_bow, _gpu_error
    from . import transformation as _tr
    from ..runtime import default_ctx
    if b is None:
        raise B.BowError("nil bow")
    if len(aggrs) == 0:
        raise B.BowError("at least one column aggregation is required")
    names = [b.ColumnName(i) for i in range(b.NumCols())]
    intervalColIndex = b.ColumnIndex(intervalColName)
    specs = []
    for i, a in enumerate(aggrs):
        if a.InputName() == "":
            raise B.BowError(f"column aggregation {i}: no input name")
        if a.InputName() not in names:
            raise B.BowError(f"column aggregation {i}: no column '{a.InputName()}'")
        a.SetInputIndex(names.index(a.InputName()))
        if a.kernelOp() is None:
            raise B.BowError(f"column aggregation {i}: custom closures are not supported by the GPU backend")
        trans = a.Transformations()
        if not all(isinstance(t, _tr._Factor) for t in trans) or len(trans) > 4:
            raise B.BowError(f"column aggregation {i}: only transformation.Factor runs on the GPU backend")
        specs.append((a.kernelOp(), a.InputIndex(), [t.n for t in trans]))
    try:
        arr, keep = _cols_from_bow(b)
        frame = N.Frame.from_col_descs(default_ctx(), arr, b.NumCols(), N.MEM_HOST, keep)
        try:
            res = frame.aggregate_whole(intervalColIndex, specs)
        finally:
            frame.close()
    except N.BowGpuError as e:
        raise _gpu_error(e)
    series = []
    for a, (vals, mask) in zip(aggrs, res):
        name = a.OutputName() or names[a.InputIndex()]
        series.append(B.NewSeriesFromNumpy(name, np.asarray(vals), np.asarray(mask)))
    return B.NewBow(*series)
