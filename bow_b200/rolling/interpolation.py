"""rolling/interpolation: the built-in ColInterpolation constructors (reference rolling/interpolation/*.go).
`None` is spelled `None_` (Python keyword).  The north-star's `StepNext` does not exist upstream
(SURVEY section 0); it is provided as the mirror image of StepPrevious over the reference's GetNextValues getter.
Its next-valid semantics are pinned on the reference's own FillNext golden table
(tests/golden/reference_vectors.py STEPNEXT_VS_FILLNEXT); there is no upstream StepNext to compare the rest with."""
from __future__ import annotations

from .. import bow as B
from .. import native as N
from . import ColInterpolation


def WindowStart(colName: str) -> ColInterpolation:     # windowstart.go:8-14
    return ColInterpolation(colName, [B.Int64], None, kernel_op=N.INTERP["WindowStart"])


def StepNext(colName: str) -> ColInterpolation:
    """Named by the north-star, not in the reference: StepPrevious mirrored over Bow.GetNextValues
    (bowgetters.go:111-123) - the next valid value at or after the window's first row, else nil: what Bow.FillNext
    (bowfill.go:154-158) leaves at that row, which is how its goldens pin it."""
    return ColInterpolation(colName, [B.Int64, B.Float64, B.Boolean, B.String], None, kernel_op=N.INTERP["StepNext"])


def Linear(colName: str) -> ColInterpolation:          # linear.go:8-38
    return ColInterpolation(colName, [B.Int64, B.Float64], None, kernel_op=N.INTERP["Linear"])


def StepPrevious(colName: str) -> ColInterpolation:    # stepprevious.go:8-26
    return ColInterpolation(colName, [B.Int64, B.Float64, B.Boolean, B.String], None,
                            kernel_op=N.INTERP["StepPrevious"])


def None_(colName: str) -> ColInterpolation:           # none.go:8-14
    return ColInterpolation(colName, [B.Int64, B.Float64, B.Boolean, B.String], None, kernel_op=N.INTERP["None_"])
