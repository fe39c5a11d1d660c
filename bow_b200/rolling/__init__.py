"""Host-side mirror of the reference's `rolling` package on top of the C ABI (include/bowgpu.h).

Same exported names, argument meaning and error strings as Metronlab/bow's Go API
(rolling/rolling.go, rolling/aggregation.go, rolling/interpolation.go, rolling/window.go), so the
parity tests read like the reference's own tests.  Go's `(value, error)` returns become Python
exceptions (`bow.BowError` carrying the Go error string); the fluent steps keep the reference's
DEFERRED error: Aggregate / Interpolate never raise, they return a Rolling carrying the error, which
`.Bow()` / `.NumWindows()` surface (rolling.go:241-248).

This layer holds validation and schema plumbing only.  Every row of data is processed by the CUDA
kernels behind the C ABI; there is no CPU fallback — when the library or a B200 is missing the
first device call raises.
"""
from __future__ import annotations

import copy
import ctypes as C
from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Sequence

import numpy as np
import pyarrow as pa

from .. import bow as B
from .. import native as N
from ..bow import BowError
from . import transformation as _tr

__all__ = ["IntervalRolling", "Options", "Window", "Rolling", "ColAggregation", "ColInterpolation",
           "NewColAggregation", "NewColInterpolation"]


# ---------------------------------------------------------------------------------------------------
# Options / Window (rolling.go:49-53, window.go:12-31)
# ---------------------------------------------------------------------------------------------------
@dataclass
class Options:
    Offset: int = 0
    Inclusive: bool = False
    PrevRow: Optional[B.Bow] = None


@dataclass
class Window:
    Bow: B.Bow
    FirstIndex: int
    IntervalColIndex: int
    FirstValue: int
    LastValue: int
    IsInclusive: bool

    def UnsetInclusive(self) -> "Window":  # window.go:23-31
        if not self.IsInclusive:
            return self
        w = copy.copy(self)
        w.IsInclusive = False
        w.Bow = self.Bow.NewSlice(0, self.Bow.NumRows() - 1)
        return w


# ---------------------------------------------------------------------------------------------------
# plugin descriptors (aggregation.go:11-121, interpolation.go:10-28)
# ---------------------------------------------------------------------------------------------------
class ColAggregation:
    """rolling.ColAggregation.  Built-in constructors (rolling/aggregation.py) additionally carry the
    kernel opcode; a custom closure has none and cannot run on the device."""

    def __init__(self, inputName: str, needInclusiveWindow: bool, typ: B.Type, fn: Optional[Callable],
                 kernel_op: Optional[int] = None):
        self.inputName = inputName
        self.inputIndex = -1
        self.needInclusiveWindow = needInclusiveWindow
        self.aggregationFn = fn
        self.transformationFns: List[Callable] = []
        self.outputName = ""
        self.typ = typ
        self._kernel_op = kernel_op

    def InputName(self) -> str:
        return self.inputName

    def InputIndex(self) -> int:
        return self.inputIndex

    def SetInputIndex(self, i: int) -> None:
        self.inputIndex = i

    def OutputName(self) -> str:
        return self.outputName

    def RenameOutput(self, name: str) -> "ColAggregation":
        c = copy.copy(self)
        c.outputName = name
        return c

    def NeedInclusiveWindow(self) -> bool:
        return self.needInclusiveWindow

    def Type(self) -> B.Type:
        return self.typ

    def GetReturnType(self, inputType: B.Type, iteratorType: B.Type) -> B.Type:  # aggregation.go:110-121
        if self.typ in (B.Int64, B.Float64, B.Boolean, B.String):
            return self.typ
        if self.typ == B.InputDependent:
            return inputType
        if self.typ == B.IteratorDependent:
            return iteratorType
        raise BowError(f"invalid return type {self.typ}")

    def Func(self) -> Optional[Callable]:
        return self.aggregationFn

    def Transformations(self) -> List[Callable]:
        return self.transformationFns

    def SetTransformations(self, *transformations: Callable) -> "ColAggregation":
        c = copy.copy(self)
        c.transformationFns = list(transformations)
        return c

    def kernelOp(self) -> Optional[int]:
        return self._kernel_op


def NewColAggregation(inputName: str, needInclusiveWindow: bool, typ: B.Type, fn: Callable) -> ColAggregation:
    """aggregation.go:53-61.  A custom closure is accepted for API compatibility but Aggregate reports an
    error for it: a host closure cannot execute on the GPU and there is no CPU fallback."""
    return ColAggregation(inputName, needInclusiveWindow, typ, fn)


class ColInterpolation:
    def __init__(self, colName: str, inputTypes: Sequence[B.Type], fn: Optional[Callable],
                 kernel_op: Optional[int] = None):
        self.colName = colName
        self.inputTypes = list(inputTypes)
        self.fn = fn
        self.colIndex = -1
        self._kernel_op = kernel_op


def NewColInterpolation(colName: str, inputTypes: Sequence[B.Type], fn: Callable) -> ColInterpolation:
    """interpolation.go:22-28 (custom closures: see NewColAggregation)."""
    return ColInterpolation(colName, inputTypes, fn)


# ---------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------
def _go_div(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _enforce_interval_and_offset(interval: int, offset: int) -> int:  # rolling.go:114-128
    if interval <= 0:
        raise BowError("strictly positive interval required")
    if offset >= interval or offset <= -interval:
        offset = offset - _go_div(offset, interval) * interval
    if offset < 0:
        offset += interval
    return offset


def _enforce_prev_row(prev: Optional[B.Bow]) -> Optional[B.Bow]:  # rolling.go:130-141
    if prev is None or prev.NumRows() == 0:
        return None
    if prev.NumRows() != 1:
        raise BowError(f"prevRow must have only one row, have {prev.NumRows()}")
    return prev


def _arrow_col_desc(arr: pa.Array, keep: list) -> N.Col:
    """bowgpu_col of one Arrow array (zero copy: buffer addresses + offset)."""
    c = N.Col()
    typ = B._from_arrow(arr.type)
    if typ not in (B.Int64, B.Float64):
        raise BowError(f"column of type {typ}: only int64 / float64 columns can be processed by the GPU backend")
    bufs = arr.buffers()
    keep.append(arr)
    c.values = bufs[1].address if bufs[1] is not None else None
    c.validity = bufs[0].address if (bufs[0] is not None and arr.null_count) else None
    c.offset = arr.offset
    c.length = len(arr)
    c.null_count = arr.null_count
    c.dtype = N.INT64 if typ == B.Int64 else N.FLOAT64
    if len(arr) and not c.values:
        raise BowError("column without a values buffer")
    return c


def _cols_from_bow(b: B.Bow):
    keep: list = []
    arr = (N.Col * max(1, b.NumCols()))()
    for j in range(b.NumCols()):
        arr[j] = _arrow_col_desc(b.Column(j), keep)
    return arr, keep


def _gpu_error(e: N.BowGpuError) -> BowError:
    return BowError(str(e))


class Rolling:
    """rolling.Rolling / intervalRolling (rolling.go:14-43)."""

    def __init__(self):
        self.bow: Optional[B.Bow] = None          # host copy (None while the data only lives on the device)
        self.frame: Optional[N.Frame] = None      # device copy (None until first needed)
        self.names: List[str] = []
        self.types: List[B.Type] = []
        self.metadata = None
        self.intervalColIndex = -1
        self.interval = 0
        self.options = Options()
        self.numWindows = 0
        self.currWindowFirstValue = 0
        self.currWindowIndex = 0
        self.err: Optional[BowError] = None
        self._nrows = 0
        self._bounds = None
        self._handle: Optional[N.Rolling] = None
        self._lazy = None                          # (parent Rolling, interpolation opcodes) of a pending Interpolate

    # -- device plumbing ----------------------------------------------------------------------------------
    def _ensure_frame(self) -> N.Frame:
        self._materialize()
        if self.frame is None:
            from ..runtime import default_ctx
            arr, keep = _cols_from_bow(self.bow)
            self.frame = N.Frame.from_col_descs(default_ctx(), arr, self.bow.NumCols(), N.MEM_HOST, keep)
        return self.frame

    def _ensure_handle(self) -> N.Rolling:
        self._materialize()
        if self._handle is None:
            fr = self._ensure_frame()
            prev = None
            if self.options.PrevRow is not None:
                prev = []
                pr = self.options.PrevRow
                # the C ABI reads one PrevRow cell per frame column, by position: same columns, same types
                if pr.NumCols() != len(self.types) or any(pr.ColumnType(j) != self.types[j] for j in range(pr.NumCols())):
                    raise BowError("newIntervalRolling: prevRow must have the same columns (number, order and types) "
                                   "as the bow")
                for j in range(pr.NumCols()):
                    a = pr.Column(j)
                    dt = np.int64 if pr.ColumnType(j) == B.Int64 else np.float64
                    v = a.to_pylist()[0]
                    prev.append((np.array([0 if v is None else v], dtype=dt), np.array([v is not None])))
            self._handle = N.Rolling(fr, self.intervalColIndex, self.interval, self.options.Offset,
                                     self.options.Inclusive, prev)
        return self._handle

    def _ensure_bow(self) -> B.Bow:
        self._materialize()
        if self.bow is None:
            cols = self.frame.download()
            series = [B.NewSeriesFromNumpy(n, v, m) for n, (v, m) in zip(self.names, cols)]
            rec = pa.RecordBatch.from_arrays([s.Array for s in series], names=self.names)
            if self.metadata:
                rec = rec.replace_schema_metadata(self.metadata)
            self.bow = B.Bow(rec)
        return self.bow

    def _copy(self) -> "Rolling":
        c = copy.copy(self)
        c.options = copy.copy(self.options)
        return c

    def _set_error(self, err: BowError) -> "Rolling":  # rolling.go:245-248
        self.err = err
        return self

    # -- public API ---------------------------------------------------------------------------------------
    def NumWindows(self) -> int:  # rolling.go:156
        if self.err is not None:
            raise self.err
        return self.numWindows

    def Bow(self) -> B.Bow:  # rolling.go:241-243
        if self.err is not None:
            raise self.err
        return self._ensure_bow()

    def _get_bounds(self):
        if self._bounds is None:
            try:
                self._bounds = self._ensure_handle().bounds()
            except N.BowGpuError as e:
                raise _gpu_error(e)
        return self._bounds

    def HasNext(self) -> bool:  # rolling.go:162-173
        self._materialize()
        return self._nrows > 0 and self.currWindowIndex < self.numWindows

    def Next(self):  # rolling.go:177-239 -> (windowIndex, Window | None)
        if not self.HasNext():
            return self.currWindowIndex, None
        first, inc = self._get_bounds()
        k = self.currWindowIndex
        lo, hi = int(first[k]), int(first[k + 1]) + int(inc[k])
        early, kept = self._ensure_handle().early_rows()
        if k == 0 and early and not kept:
            hi = lo
        fv = self.currWindowFirstValue
        b = self._ensure_bow()
        w = Window(Bow=b.NewSlice(lo, hi) if hi > lo else b.NewEmptySlice(), FirstIndex=lo,
                   IntervalColIndex=self.intervalColIndex, FirstValue=fv, LastValue=fv + self.interval,
                   IsInclusive=bool(inc[k]))
        self.currWindowFirstValue = fv + self.interval
        self.currWindowIndex += 1
        return k, w

    # -- Aggregate (aggregation.go:123-238) ------------------------------------------------------------------
    def Aggregate(self, *aggrs: ColAggregation) -> "Rolling":
        if self.err is not None:
            return self
        rc = self._copy()
        try:
            newIntervalCol = rc._indexed_aggregations(aggrs)
        except BowError as e:
            return rc._set_error(BowError(f"intervalRolling.indexedAggregations: {e}"))
        try:
            out = rc._aggregate_windows(aggrs)
        except BowError as e:
            return rc._set_error(BowError(f"intervalRolling.aggregateWindows: {e}"))
        try:
            return _new_interval_rolling(out, newIntervalCol, rc.interval, rc.options)
        except BowError as e:
            return rc._set_error(BowError(f"newIntervalRolling: {e}"))

    def _indexed_aggregations(self, aggrs) -> int:  # aggregation.go:147-188
        if len(aggrs) == 0:
            raise BowError("at least one column aggregation is required")
        newIntervalCol = -1
        for i, a in enumerate(aggrs):
            if a.InputName() == "":
                raise BowError(f"aggregation {i} has no column name")
            if a.InputName() not in self.names:
                raise BowError(f"no column '{a.InputName()}'")
            if self.names.count(a.InputName()) > 1:
                raise BowError(f"several columns '{a.InputName()}'")
            readIndex = self.names.index(a.InputName())
            a.SetInputIndex(readIndex)  # mutates the caller's object, like the reference (aggregation.go:181)
            if a.NeedInclusiveWindow():
                self.options.Inclusive = True
            if readIndex == self.intervalColIndex:
                newIntervalCol = i
        if newIntervalCol == -1:
            raise BowError(f"must keep interval column '{self.names[self.intervalColIndex]}'")
        return newIntervalCol

    def _aggregate_windows(self, aggrs) -> B.Bow:  # aggregation.go:190-238
        specs, host_trans = [], []
        for i, a in enumerate(aggrs):
            op = a.kernelOp()
            if op is None:
                raise BowError(f"aggregation {i}: custom closures are not supported by the GPU backend")
            trans = a.Transformations()
            if all(isinstance(t, _tr._Factor) for t in trans) and len(trans) <= 4:
                specs.append((op, a.InputIndex(), [t.n for t in trans]))
                host_trans.append(None)
            else:  # arbitrary host closures run over the W-length result
                specs.append((op, a.InputIndex(), []))
                host_trans.append(list(trans))
        # the Inclusive flag forced by validateAggregation takes part in the iteration (aggregation.go:183-185)
        self._handle = None
        self._bounds = None
        try:
            if self._lazy is not None and self.currWindowIndex == 0:
                parent, ops = self._lazy   # Aggregate right after Interpolate: fused, nothing is materialised
                res = parent._ensure_handle().interpolate_aggregate(ops, specs)
            elif self.frame is None and self.bow is not None and self.currWindowIndex == 0 and self.numWindows > 0:
                res = self._aggregate_host(specs)   # host Bow, nothing on the device yet: one pipelined call
            else:
                res = self._ensure_handle().aggregate(specs)
        except N.BowGpuError as e:
            raise _gpu_error(e)
        series = []
        for i, (a, (vals, mask)) in enumerate(zip(aggrs, res)):
            if self.currWindowIndex > 0:  # an iterator advanced with Next() only aggregates what is left
                vals, mask = vals.copy(), mask.copy()
                vals[:self.currWindowIndex] = 0
                mask[:self.currWindowIndex] = False
            if host_trans[i]:
                lst = [v if ok else None for v, ok in zip(vals.tolist(), mask.tolist())]
                for t in host_trans[i]:
                    lst = [t(v) for v in lst]
                typ = a.GetReturnType(self.types[a.InputIndex()], self.types[self.intervalColIndex])
                conv = float if typ == B.Float64 else int
                mask = np.array([v is not None for v in lst], dtype=bool)
                vals = np.array([0 if v is None else conv(v) for v in lst],
                                dtype=np.float64 if typ == B.Float64 else np.int64)
            name = a.OutputName() or self.names[a.InputIndex()]
            series.append(B.NewSeriesFromNumpy(name, vals, mask))
        return B.NewBow(*series)

    # -- Interpolate (interpolation.go:30-161) -----------------------------------------------------------------
    def Interpolate(self, *interps: ColInterpolation) -> "Rolling":
        if self.err is not None:
            return self
        rc = self._copy()
        if len(interps) == 0:
            return rc._set_error(BowError("at least one column interpolation is required"))
        interps = [copy.copy(i) for i in interps]
        newIntervalCol = -1
        for i, it in enumerate(interps):
            try:
                isInterval = self._validate_interpolation(it, i)
            except BowError as e:
                return rc._set_error(BowError(f"intervalRolling.validateInterpolation: {e}"))
            if isInterval:
                newIntervalCol = i
        if newIntervalCol == -1:
            return rc._set_error(BowError(f"must keep interval column '{self.names[self.intervalColIndex]}'"))
        try:
            out = rc._interpolate_windows(interps)
        except BowError as e:
            return rc._set_error(BowError(f"intervalRolling.interpolateWindows: {e}"))
        try:
            return out
        except BowError as e:  # pragma: no cover
            return rc._set_error(BowError(f"newIntervalRolling: {e}"))

    def _validate_interpolation(self, it: ColInterpolation, newIndex: int) -> bool:  # interpolation.go:71-96
        if it.colName == "":
            raise BowError(f"interpolation {newIndex} has no column name")
        if it.colName not in self.names:
            raise BowError(f"no column '{it.colName}'")
        if self.names.count(it.colName) > 1:
            raise BowError(f"several columns '{it.colName}'")
        it.colIndex = self.names.index(it.colName)
        typ = self.types[it.colIndex]
        if typ not in it.inputTypes:
            raise BowError(f"accepts types [{' '.join(str(t) for t in it.inputTypes)}], got type {typ}")
        return it.colIndex == self.intervalColIndex

    def _interpolate_windows(self, interps) -> "Rolling":
        # The reference appends [start row] ++ window by column POSITION (bowappend.go:28-47), which is only
        # well defined when the interpolations name every column in schema order; the GPU backend requires it.
        if [it.colIndex for it in interps] != list(range(len(self.names))):
            raise BowError("the GPU backend requires one interpolation per column, in schema order")
        ops = []
        for i, it in enumerate(interps):
            if it._kernel_op is None:
                raise BowError(f"interpolation {i}: custom closures are not supported by the GPU backend")
            ops.append(it._kernel_op)
        r = Rolling()
        r.names, r.types, r.metadata = list(self.names), list(self.types), self.metadata
        r.intervalColIndex = self.intervalColIndex
        r.interval = self.interval
        r.options = copy.copy(self.options)
        # Lazy: when Aggregate follows directly, the interpolated frame is never materialised
        # (bowgpu_rolling_interpolate_aggregate).  Only taken when the host can tell that the interpolated frame keeps
        # the window lattice: no user Inclusive, first row not before the first window start, at least one window.
        first_time = self._first_time()
        # ... and the interval column is interpolated by WindowStart (the condition of the fused C entry point): any other
        # interpolation of it (None leaves null timestamps, newIntervalRolling then fails, rolling.go:91-94) goes
        # through the eager path so that NumWindows / Bow / the deferred error behave like the reference's.
        if (not self.options.Inclusive and self.numWindows > 0 and first_time is not None
                and self.currWindowFirstValue <= first_time and self.currWindowIndex == 0
                and ops[self.intervalColIndex] == N.INTERP["WindowStart"]):
            r._lazy = (self, ops)
            r.numWindows = self.numWindows
            r.currWindowFirstValue = self.currWindowFirstValue
            r._nrows = -1
            return r
        try:
            frame = self._ensure_handle().interpolate(ops)
        except N.BowGpuError as e:
            raise _gpu_error(e)
        r.frame = frame
        r._nrows = frame.num_rows
        if r._nrows == 0:  # interpolation.go:59-61
            r.numWindows = 0
            r.currWindowFirstValue = 0
            return r
        try:
            h = r._ensure_handle()
        except N.BowGpuError as e:
            raise BowError(f"newIntervalRolling: {_gpu_error(e)}")
        r.numWindows = h.num_windows
        r.currWindowFirstValue = h.first_window_start
        return r

    def _aggregate_host(self, specs):
        """bowgpu_aggregate_host: upload, kernels and download of window-range chunks overlap; only the columns the
        aggregations read cross the bus"""
        import ctypes as C
        from ..runtime import default_ctx
        ctx = default_ctx()
        arr, keep = _cols_from_bow(self.bow)
        W = self.numWindows
        sarr = N.make_specs(specs)
        outs = (N.OutCol * len(specs))()
        bufs = []
        for j in range(len(specs)):
            v = np.zeros(max(W, 1), dtype=np.int64)
            b = np.zeros((W + 7) // 8 + 1, dtype=np.uint8)
            bufs.append((v, b))
            outs[j].values, outs[j].validity = v.ctypes.data, b.ctypes.data
        got = C.c_int64()
        ctx.check(N.lib().bowgpu_aggregate_host(ctx.h, arr, self.bow.NumCols(), self.intervalColIndex, self.interval,
                                                self.options.Offset, int(self.options.Inclusive), sarr, len(specs), outs, W,
                                                C.byref(got)))
        if got.value != W:
            raise BowError(f"window count mismatch: {got.value} != {W}")
        res = []
        for j, (v, b) in enumerate(bufs):
            vals = v[:W] if outs[j].dtype == N.INT64 else v[:W].view(np.float64)
            res.append((vals, N.unpack_bits(b, W)))
        return res

    def _first_time(self):
        if self.bow is None or self.bow.NumRows() == 0:
            return None
        return self.bow.Column(self.intervalColIndex)[0].as_py()

    def _materialize(self):
        """runs the pending Interpolate of a lazy Rolling (the interpolated frame was asked for after all)"""
        if self._lazy is None:
            return
        parent, ops = self._lazy
        self._lazy = None
        try:
            self.frame = parent._ensure_handle().interpolate(ops)
        except N.BowGpuError as e:
            raise _gpu_error(e)
        self._nrows = self.frame.num_rows


def _new_interval_rolling(b: B.Bow, intervalColIndex: int, interval: int, options: Options) -> Rolling:
    """newIntervalRolling, rolling.go:69-112 (host arithmetic only: nothing touches the GPU here)."""
    if b.ColumnType(intervalColIndex) != B.Int64:
        raise BowError(f"impossible to create a new intervalRolling on column of type {b.ColumnType(intervalColIndex)}")
    options = copy.copy(options)
    try:
        options.Offset = _enforce_interval_and_offset(interval, options.Offset)
    except BowError as e:
        raise BowError(f"enforceIntervalAndOffset: {e}")
    try:
        options.PrevRow = _enforce_prev_row(options.PrevRow)
    except BowError as e:
        raise BowError(f"enforcePrevRow: {e}")
    r = Rolling()
    r.bow = b
    r.names = [b.ColumnName(i) for i in range(b.NumCols())]
    r.types = [b.ColumnType(i) for i in range(b.NumCols())]
    r.metadata = b.Record.schema.metadata
    r.intervalColIndex = intervalColIndex
    r.interval = interval
    r.options = options
    r._nrows = b.NumRows()
    first = 0
    tcol = b.Column(intervalColIndex)
    if b.NumRows() > 0:
        v = tcol[0].as_py()
        if v is None:
            raise BowError("the first value of the column should be convertible to int64, got <nil>")
        first = _go_div(v, interval) * interval + options.Offset  # rolling.go:96-99
        if first > v:
            first -= interval
    r.currWindowFirstValue = first
    # countWindows, rolling.go:143-154
    nw = 0
    if b.NumRows() > 0:
        i = b.NumRows() - 1
        if tcol.null_count:
            while i >= 0 and not tcol[i].is_valid:
                i -= 1
        if i >= 0:
            last = tcol[i].as_py()
            if first <= last:
                nw = (last - first) // interval + 1
    r.numWindows = nw
    return r


def IntervalRolling(b: B.Bow, colName: str, interval: int, options: Optional[Options] = None) -> Rolling:
    """rolling.IntervalRolling (rolling.go:60-67)."""
    colIndex = b.ColumnIndex(colName)
    return _new_interval_rolling(b, colIndex, interval, options or Options())
