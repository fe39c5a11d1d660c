"""TEST INFRASTRUCTURE ONLY — literal CPU restatement of Bow's interval-rolling path.

This module is the *slow, line-by-line* oracle: every function follows one
function of the reference (Metronlab/bow, pure Go) and cites the file:line it
restates.  It exists to pin `oracle/ref.c` (the fast C restatement) and, through
it, the CUDA path.  It is pure-Python loops over Python lists and is only meant
for small cases (<= ~1e5 rows).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may
import anything under `oracle/`.  The product (`bow_b200/`) never does.

Parity status: PINNED — replays every in-scope golden vector of the reference's
own tests (see tests/golden/reference_vectors.py and tests/test_oracle_golden.py).
`interpolation.StepNext` named by the north-star does not exist upstream; InterpStepNext
restates it as StepPrevious mirrored over Bow.GetNextValues — PARITY UNPINNED for that one
function (no upstream code, no golden vectors).

Arithmetic conventions restated from Go/amd64:
  * int64 `/` and `%` truncate toward zero                (go spec; rolling.go:96,119)
  * float64 is IEEE double, no fused multiply-add          (Go/amd64 never fuses)
  * int64(float64) truncates toward zero; NaN/out-of-range give INT64_MIN
    (CVTTSD2SI "integer indefinite")                       (bowconvert.go:28-29)
  * float64(int64) rounds to nearest even                  (bowgetters.go:225)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Any, Callable, List, Optional, Sequence

INT64 = "int64"
FLOAT64 = "float64"
INT64_MIN = -(1 << 63)
INT64_MAX = (1 << 63) - 1

# return-type tags of ColAggregation (bowtypes.go:17-32)
T_INT64 = "int64"
T_FLOAT64 = "float64"
T_INPUT = "input-dependent"
T_ITER = "iterator-dependent"


def go_div(a: int, b: int) -> int:
    """Go integer division (truncates toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def go_mod(a: int, b: int) -> int:
    """Go remainder: sign follows the dividend."""
    return a - go_div(a, b) * b


def wrap64(x: int) -> int:
    """two's complement wrap-around of an int64 expression"""
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def f64_to_i64(x: float) -> int:
    """Go int64(float64) on amd64 (bowconvert.go:28-29)."""
    if x != x or x >= 9223372036854775808.0 or x < -9223372036854775808.0:
        return INT64_MIN
    return int(x)


# ----------------------------------------------------------------------------
# minimal Bow (bow.go:16-98): column-major, None == null
# ----------------------------------------------------------------------------
@dataclass
class Frame:
    names: List[str]
    types: List[str]
    cols: List[List[Any]]
    lo: int = 0          # zero-copy slice [lo, hi) over `cols` (bow.go:279-283)
    hi: Optional[int] = None

    def __post_init__(self):
        if self.hi is None:
            self.hi = len(self.cols[0]) if self.cols else 0

    def num_rows(self) -> int:
        return self.hi - self.lo

    def num_cols(self) -> int:
        return len(self.names)

    def column_index(self, name: str) -> int:          # bowgetters.go:320-332
        idx = [i for i, n in enumerate(self.names) if n == name]
        if not idx:
            raise KeyError(f"no column '{name}'")
        if len(idx) > 1:
            raise KeyError(f"several columns '{name}'")
        return idx[0]

    def new_slice(self, i: int, j: int) -> "Frame":    # bow.go:279-283
        return Frame(self.names, self.types, self.cols, self.lo + i, self.lo + j)

    def new_empty_slice(self) -> "Frame":
        return Frame(self.names, self.types, self.cols, self.lo, self.lo)

    # -- getters ----------------------------------------------------------
    def get_value(self, c: int, r: int):               # bowgetters.go:46-63
        return self.cols[c][self.lo + r]

    def get_int64(self, c: int, r: int):               # bowgetters.go:155-184
        if r < 0 or r >= self.num_rows():
            return 0, False
        v = self.cols[c][self.lo + r]
        if self.types[c] == INT64:
            return (v if v is not None else 0), v is not None
        return (f64_to_i64(v) if v is not None else 0), v is not None

    def get_float64(self, c: int, r: int):             # bowgetters.go:218-247
        if r < 0 or r >= self.num_rows():
            return 0.0, False
        v = self.cols[c][self.lo + r]
        if v is None:
            return 0.0, False
        return float(v), True

    def get_prev_int64(self, c: int, r: int):          # bowgetters.go:189-199
        while 0 <= r < self.num_rows():
            v, ok = self.get_int64(c, r)
            if ok:
                return v, r
            r -= 1
        return 0, -1

    def get_next_float64(self, c: int, r: int):        # bowgetters.go:266-276
        while 0 <= r < self.num_rows():
            v, ok = self.get_float64(c, r)
            if ok:
                return v, r
            r += 1
        return 0.0, -1

    def get_prev_float64(self, c: int, r: int):        # bowgetters.go:252-262
        while 0 <= r < self.num_rows():
            v, ok = self.get_float64(c, r)
            if ok:
                return v, r
            r -= 1
        return 0.0, -1

    def get_next_float64s(self, c1: int, c2: int, r: int):   # bowgetters.go:299-311
        while 0 <= r < self.num_rows():
            v1, r = self.get_next_float64(c1, r)
            v2, r2 = self.get_next_float64(c2, r)
            if r == r2:
                return v1, v2, r
            r += 1
        return 0.0, 0.0, -1

    def get_prev_float64s(self, c1: int, c2: int, r: int):   # bowgetters.go:282-294
        while 0 <= r < self.num_rows():
            v1, r = self.get_prev_float64(c1, r)
            v2, r2 = self.get_prev_float64(c2, r)
            if r == r2:
                return v1, v2, r
            r -= 1
        return 0.0, 0.0, -1

    def get_next_value(self, c: int, r: int):          # bowgetters.go:80-91
        while 0 <= r < self.num_rows():
            v = self.get_value(c, r)
            if v is not None:
                return v, r
            r += 1
        return None, -1

    def get_prev_value(self, c: int, r: int):          # bowgetters.go:67-78
        while 0 <= r < self.num_rows():
            v = self.get_value(c, r)
            if v is not None:
                return v, r
            r -= 1
        return None, -1

    def get_next_values(self, c1: int, c2: int, r: int):     # bowgetters.go:111-123
        while 0 <= r < self.num_rows():
            v1, r = self.get_next_value(c1, r)
            v2, r2 = self.get_next_value(c2, r)
            if r == r2:
                return v1, v2, r
            r += 1
        return None, None, -1

    def get_prev_values(self, c1: int, c2: int, r: int):     # bowgetters.go:93-107
        while 0 <= r < self.num_rows():
            v1, r = self.get_prev_value(c1, r)
            v2, r2 = self.get_prev_value(c2, r)
            if r == r2:
                return v1, v2, r
            r -= 1
        return None, None, -1

    def materialize(self) -> List[List[Any]]:
        return [c[self.lo:self.hi] for c in self.cols]


def append_frames(frames: Sequence[Optional[Frame]]) -> Optional[Frame]:
    """bow.AppendBows (bowappend.go:14-103): positional column match, schema of the first."""
    frames = list(frames)
    if len(frames) == 0:
        return None
    if len(frames) == 1:
        return frames[0]
    ref = frames[0]
    cols = [[] for _ in range(ref.num_cols())]
    for f in frames:
        for c in range(ref.num_cols()):
            if f.types[c] != ref.types[c]:
                raise TypeError(f"incompatible types '{ref.types[c]}' and '{f.types[c]}'")
            cols[c].extend(f.cols[c][f.lo:f.hi])
    return Frame(list(ref.names), list(ref.types), cols)


# ----------------------------------------------------------------------------
# rolling/window.go
# ----------------------------------------------------------------------------
@dataclass
class Window:                                            # window.go:12-19
    bow: Frame
    first_index: int
    interval_col_index: int
    first_value: int
    last_value: int
    is_inclusive: bool

    def unset_inclusive(self) -> "Window":               # window.go:23-31
        if not self.is_inclusive:
            return self
        return replace(self, is_inclusive=False,
                       bow=self.bow.new_slice(0, self.bow.num_rows() - 1))


@dataclass
class Options:                                           # rolling.go:49-53
    offset: int = 0
    inclusive: bool = False
    prev_row: Optional[Frame] = None


# ----------------------------------------------------------------------------
# rolling/rolling.go
# ----------------------------------------------------------------------------
def enforce_interval_and_offset(interval: int, offset: int) -> int:   # rolling.go:114-128
    if interval <= 0:
        raise ValueError("enforceIntervalAndOffset: strictly positive interval required")
    if offset >= interval or offset <= -interval:
        offset = go_mod(offset, interval)
    if offset < 0:
        offset += interval
    return offset


def enforce_prev_row(prev_row: Optional[Frame]) -> Optional[Frame]:   # rolling.go:130-141
    if prev_row is None or prev_row.num_rows() == 0:
        return None
    if prev_row.num_rows() != 1:
        raise ValueError(f"enforcePrevRow: prevRow must have only one row, have {prev_row.num_rows()}")
    return prev_row


def count_windows(b: Frame, col: int, first_window_start: int, interval: int) -> int:  # rolling.go:143-154
    if b.num_rows() == 0:
        return 0
    last, idx = b.get_prev_int64(col, b.num_rows() - 1)
    if idx == -1 or first_window_start > last:
        return 0
    return go_div(last - first_window_start, interval) + 1


class NewRollingError(ValueError):
    """Aggregate/Interpolate computed `frame` but re-wrapping it in a Rolling failed
    (e.g. trailing null timestamps leave every window unset -> first output time is nil)."""

    def __init__(self, msg, frame):
        super().__init__(msg)
        self.frame = frame


class IntervalRolling:
    """rolling.intervalRolling (rolling.go:31-43) with its iterator and drivers."""

    def __init__(self, b: Frame, interval_col: int, interval: int, options: Options):
        # newIntervalRolling, rolling.go:69-112
        if b.types[interval_col] != INT64:
            raise TypeError(
                f"impossible to create a new intervalRolling on column of type {b.types[interval_col]}")
        options = Options(enforce_interval_and_offset(interval, options.offset),
                          options.inclusive, enforce_prev_row(options.prev_row))
        first = 0
        if b.num_rows() > 0:
            v, ok = b.get_int64(interval_col, 0)
            if not ok:
                raise ValueError("the first value of the column should be convertible to int64, got <nil>")
            first = wrap64(go_div(v, interval) * interval + options.offset)
            if first > v:
                first = wrap64(first - interval)
        self.bow = b
        self.interval_col = interval_col
        self.interval = interval
        self.options = options
        self.num_windows = count_windows(b, interval_col, first, interval)
        self.curr_window_first_value = first
        self.curr_row_index = 0
        self.curr_window_index = 0

    @classmethod
    def create(cls, b: Frame, col_name: str, interval: int, options: Options = None):   # rolling.go:60-67
        return cls(b, b.column_index(col_name), interval, options or Options())

    def copy(self) -> "IntervalRolling":
        c = object.__new__(IntervalRolling)
        c.__dict__.update(self.__dict__)
        c.options = replace(self.options)
        return c

    def has_next(self) -> bool:                          # rolling.go:162-173
        if self.curr_row_index >= self.bow.num_rows():
            return False
        last, ok = self.bow.get_int64(self.interval_col, self.bow.num_rows() - 1)
        if not ok:
            return False
        return self.curr_window_first_value <= last

    def next(self):                                      # rolling.go:177-239
        if not self.has_next():
            return self.curr_window_index, None
        first_value = self.curr_window_first_value
        last_value = wrap64(first_value + self.interval)
        is_inclusive = False
        first_row = self.curr_row_index
        last_row = -1
        row = first_row
        n = self.bow.num_rows()
        while row < n:
            val, ok = self.bow.get_int64(self.interval_col, row)
            if not ok:
                row += 1
                continue
            if val < first_value:
                row += 1
                continue
            if val > last_value:
                break
            if val == last_value:
                if is_inclusive:
                    break
                if not self.options.inclusive:
                    break
                is_inclusive = True
            last_row = row
            row += 1
        self.curr_row_index = row - 1 if is_inclusive else row
        self.curr_window_first_value = last_value
        wi = self.curr_window_index
        self.curr_window_index += 1
        b = self.bow.new_empty_slice() if last_row == -1 else self.bow.new_slice(first_row, last_row + 1)
        return wi, Window(b, first_row, self.interval_col, first_value, last_value, is_inclusive)

    # ---- Aggregate driver: rolling/aggregation.go:123-238 -------------------
    def aggregate(self, *aggrs: "ColAggregation") -> "IntervalRolling":
        r = self.copy()
        if len(aggrs) == 0:
            raise ValueError("intervalRolling.indexedAggregations: at least one column aggregation is required")
        new_interval_col = -1
        for i, a in enumerate(aggrs):                    # validateAggregation :171-188
            if a.input_name == "":
                raise ValueError(f"intervalRolling.indexedAggregations: aggregation {i} has no column name")
            try:
                idx = r.bow.column_index(a.input_name)
            except KeyError as e:
                raise KeyError(f"intervalRolling.indexedAggregations: {e.args[0]}")
            a.input_index = idx
            if a.need_inclusive:
                r.options.inclusive = True
            if idx == r.interval_col:
                new_interval_col = i
        if new_interval_col == -1:
            raise ValueError("intervalRolling.indexedAggregations: must keep interval column "
                             f"'{r.bow.names[r.interval_col]}'")
        names, types, cols = [], [], []
        for a in aggrs:                                  # aggregateWindows :190-238
            rc = r.copy()
            typ = a.return_type(rc.bow.types[a.input_index], rc.bow.types[rc.interval_col])
            buf = [None] * rc.num_windows                # bow.NewBuffer: all null
            while rc.has_next():
                wi, w = rc.next()
                if not a.need_inclusive and w.is_inclusive:
                    w = w.unset_inclusive()
                val = a.fn(a.input_index, w)
                for tr in a.transformations:
                    val = tr(val)
                if val is None:
                    continue
                buf[wi] = convert(val, typ)              # Buffer.SetOrDrop, bowbuffer.go:60-80
            names.append(a.output_name or rc.bow.names[a.input_index])
            types.append(typ)
            cols.append(buf)
        out = Frame(names, types, cols)
        try:                                             # aggregation.go:139-142
            return IntervalRolling(out, new_interval_col, r.interval, r.options)
        except (ValueError, TypeError) as e:
            raise NewRollingError(f"newIntervalRolling: {e.args[0]}", out)

    # ---- Interpolate driver: rolling/interpolation.go:30-161 ----------------
    def interpolate(self, *interps: "ColInterpolation") -> "IntervalRolling":
        r = self.copy()
        if len(interps) == 0:
            raise ValueError("at least one column interpolation is required")
        new_interval_col = -1
        for i, it in enumerate(interps):                 # validateInterpolation :71-96
            if it.col_name == "":
                raise ValueError(f"intervalRolling.validateInterpolation: interpolation {i} has no column name")
            try:
                it.col_index = r.bow.column_index(it.col_name)
            except KeyError as e:
                raise KeyError(f"intervalRolling.validateInterpolation: {e.args[0]}")
            ct = r.bow.types[it.col_index]
            if ct not in it.input_types:
                raise TypeError("intervalRolling.validateInterpolation: accepts types "
                                f"[{' '.join(it.input_types)}], got type {ct}")
            if it.col_index == r.interval_col:
                new_interval_col = i
        if new_interval_col == -1:
            raise ValueError(f"must keep interval column '{r.bow.names[r.interval_col]}'")
        rc = r.copy()                                    # interpolateWindows :98-116
        bows: List[Optional[Frame]] = [None] * rc.num_windows
        while rc.has_next():
            wi, w = rc.next()
            bows[wi] = rc._interpolate_window(interps, w)
        b = append_frames(bows)
        if b is None:
            b = r.bow.new_empty_slice()
        try:                                             # interpolation.go:63-66
            return IntervalRolling(b, new_interval_col, r.interval, r.options)
        except (ValueError, TypeError) as e:
            raise NewRollingError(f"newIntervalRolling: {e.args[0]}", b)

    def _interpolate_window(self, interps, window: Window) -> Frame:   # interpolation.go:118-161
        first_col_value = -1
        if window.bow.num_rows() > 0:
            v, i = window.bow.get_next_float64(self.interval_col, 0)
            if i > -1:
                first_col_value = f64_to_i64(v)
        if first_col_value == window.first_value:
            for it in interps:
                it.fn(it.col_index, window, self.bow, self.options.prev_row)
            return window.bow
        names, types, cols = [], [], []
        for it in interps:
            ct = window.bow.types[it.col_index]
            v = it.fn(it.col_index, window, self.bow, self.options.prev_row)
            names.append(window.bow.names[it.col_index])
            types.append(ct)
            cols.append([convert(v, ct)])
        return append_frames([Frame(names, types, cols), window.bow])


def convert(val, typ: str):
    """Type.Convert via ToInt64 / ToFloat64 (bowconvert.go:11-73) for the in-scope types."""
    if val is None:
        return None
    if typ == INT64:
        if isinstance(val, bool):
            return int(val)
        if isinstance(val, int):
            return val
        return f64_to_i64(val)
    if typ == FLOAT64:
        return float(val)
    raise TypeError(typ)


# ----------------------------------------------------------------------------
# rolling/aggregation.go:11-121 — plugin descriptor
# ----------------------------------------------------------------------------
@dataclass
class ColAggregation:
    input_name: str
    need_inclusive: bool
    typ: str
    fn: Callable[[int, Window], Any]
    input_index: int = -1
    output_name: str = ""
    transformations: list = field(default_factory=list)

    def rename_output(self, name: str) -> "ColAggregation":        # aggregation.go:82-86
        return replace(self, output_name=name)

    def set_transformations(self, *tr) -> "ColAggregation":        # aggregation.go:104-108
        return replace(self, transformations=list(tr))

    def return_type(self, input_type: str, iterator_type: str) -> str:   # aggregation.go:110-121
        if self.typ == T_INPUT:
            return input_type
        if self.typ == T_ITER:
            return iterator_type
        return self.typ


@dataclass
class ColInterpolation:                                   # interpolation.go:10-28
    col_name: str
    input_types: List[str]
    fn: Callable
    col_index: int = -1


# ----------------------------------------------------------------------------
# rolling/aggregation/*.go
# ----------------------------------------------------------------------------
def WindowStart(col: str) -> ColAggregation:              # windowstart.go:8-13
    return ColAggregation(col, False, T_ITER, lambda c, w: w.first_value)


def Count(col: str) -> ColAggregation:                    # count.go:8-20
    def fn(c, w):
        count = 0
        for i in range(w.bow.num_rows()):
            if w.bow.get_value(c, i) is not None:
                count += 1
        return count
    return ColAggregation(col, False, T_INT64, fn)


def Sum(col: str) -> ColAggregation:                      # sum.go:8-25
    def fn(c, w):
        if w.bow.num_rows() == 0:
            return 0.0
        s = 0.0
        for i in range(w.bow.num_rows()):
            v, ok = w.bow.get_float64(c, i)
            if not ok:
                continue
            s += v
        return s
    return ColAggregation(col, False, T_FLOAT64, fn)


def ArithmeticMean(col: str) -> ColAggregation:           # arithmeticmean.go:8-30
    def fn(c, w):
        if w.bow.num_rows() == 0:
            return None
        s, count = 0.0, 0
        for i in range(w.bow.num_rows()):
            v, ok = w.bow.get_float64(c, i)
            if not ok:
                continue
            s += v
            count += 1
        if count == 0:
            return None
        return s / float(count)
    return ColAggregation(col, False, T_FLOAT64, fn)


def Min(col: str) -> ColAggregation:                      # minmax.go:8-31
    def fn(c, w):
        if w.bow.num_rows() == 0:
            return None
        m = None
        for i in range(w.bow.num_rows()):
            v, ok = w.bow.get_float64(c, i)
            if not ok:
                continue
            if m is not None:
                if v < m:
                    m = v
                continue
            m = v
        return m
    return ColAggregation(col, False, T_FLOAT64, fn)


def Max(col: str) -> ColAggregation:                      # minmax.go:33-56
    def fn(c, w):
        if w.bow.num_rows() == 0:
            return None
        m = None
        for i in range(w.bow.num_rows()):
            v, ok = w.bow.get_float64(c, i)
            if not ok:
                continue
            if m is not None:
                if v > m:
                    m = v
                continue
            m = v
        return m
    return ColAggregation(col, False, T_FLOAT64, fn)


def First(col: str) -> ColAggregation:                    # firstlast.go:8-21
    def fn(c, w):
        if w.bow.num_rows() == 0:
            return None
        v, i = w.bow.get_next_value(c, 0)
        return None if i == -1 else v
    return ColAggregation(col, False, T_INPUT, fn)


def Last(col: str) -> ColAggregation:                     # firstlast.go:23-36
    def fn(c, w):
        if w.bow.num_rows() == 0:
            return None
        v, i = w.bow.get_prev_value(c, w.bow.num_rows() - 1)
        return None if i == -1 else v
    return ColAggregation(col, False, T_INPUT, fn)


def _integral_trapezoid(c, w):                            # integral.go:8-38
    if w.bow.num_rows() == 0:
        return None
    s, ok = 0.0, False
    t0, v0, row = w.bow.get_next_float64s(w.interval_col_index, c, 0)
    if row < 0:
        return None
    while row >= 0:
        t1, v1, nxt = w.bow.get_next_float64s(w.interval_col_index, c, row + 1)
        if nxt < 0:
            break
        s += (v0 + v1) / 2 * (t1 - t0)
        ok = True
        t0, v0, row = t1, v1, nxt
    return s if ok else None


def _integral_step(c, w):                                 # integral.go:40-69
    if w.bow.num_rows() == 0:
        return None
    s, ok = 0.0, False
    t0, v0, row = w.bow.get_next_float64s(w.interval_col_index, c, 0)
    while row >= 0:
        t1, v1, nxt = w.bow.get_next_float64s(w.interval_col_index, c, row + 1)
        if nxt < 0:
            t1 = float(w.last_value)
        s += v0 * (t1 - t0)
        ok = True
        if nxt < 0:
            break
        t0, v0, row = t1, v1, nxt
    return s if ok else None


def IntegralTrapezoid(col: str) -> ColAggregation:
    return ColAggregation(col, True, T_FLOAT64, _integral_trapezoid)


def IntegralStep(col: str) -> ColAggregation:
    return ColAggregation(col, False, T_FLOAT64, _integral_step)


def go_fdiv(a: float, b: float) -> float:
    """Go float64 division: x/0 is +-Inf or NaN, never a panic."""
    if b == 0.0:
        if a == 0.0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def WeightedAverageStep(col: str) -> ColAggregation:      # weightedmean.go:8-20
    def fn(c, w):
        v = _integral_step(c, w)
        if v is None:
            return None
        return go_fdiv(v, float(w.last_value - w.first_value))
    return ColAggregation(col, False, T_FLOAT64, fn)


def WeightedAverageLinear(col: str) -> ColAggregation:    # weightedmean.go:22-34
    def fn(c, w):
        v = _integral_trapezoid(c, w)
        if v is None:
            return None
        return go_fdiv(v, float(w.last_value - w.first_value))
    return ColAggregation(col, True, T_FLOAT64, fn)


# ----------------------------------------------------------------------------
# rolling/transformation/factor.go:7-20
# ----------------------------------------------------------------------------
def Factor(n: float):
    def tr(x):
        if x is None:
            return None
        if isinstance(x, bool) or not isinstance(x, (int, float)):
            raise TypeError(f"factor: invalid type {type(x).__name__}")
        if isinstance(x, float):
            return x * n
        return f64_to_i64(float(x) * n)
    return tr


# ----------------------------------------------------------------------------
# rolling/interpolation/*.go
# ----------------------------------------------------------------------------
def InterpWindowStart(col: str) -> ColInterpolation:      # interpolation/windowstart.go:8-14
    return ColInterpolation(col, [INT64], lambda c, w, full, prev: w.first_value)


def InterpNone(col: str) -> ColInterpolation:             # interpolation/none.go:8-14
    return ColInterpolation(col, [INT64, FLOAT64, "bool"], lambda c, w, full, prev: None)


def InterpStepPrevious(col: str) -> ColInterpolation:     # interpolation/stepprevious.go:8-26
    state = {"prev": None}

    def fn(c, w, full, prev_row):
        if w.first_index == 0 and prev_row is not None:
            state["prev"] = prev_row.get_value(c, prev_row.num_rows() - 1)
        _, v, _ = full.get_prev_values(w.interval_col_index, c, w.first_index - 1)
        if v is not None:
            state["prev"] = v
        return state["prev"]
    return ColInterpolation(col, [INT64, FLOAT64, "bool", "utf8"], fn)


def InterpStepNext(col: str) -> ColInterpolation:
    """NOT in the reference (rolling/interpolation/ holds Linear, None, StepPrevious, WindowStart): named by the
    north-star.  Defined as the mirror image of StepPrevious (stepprevious.go:8-26) over the reference's own getter
    Bow.GetNextValues (bowgetters.go:111-123): the value of the first row at or after the window's first row where
    interval column and value are both valid, else nil; PrevRow plays no part.  PARITY UNPINNED (no golden vectors)."""
    def fn(c, w, full, prev_row):
        _, v, _ = full.get_next_values(w.interval_col_index, c, w.first_index)
        return v
    return ColInterpolation(col, [INT64, FLOAT64, "bool", "utf8"], fn)


def InterpLinear(col: str) -> ColInterpolation:           # interpolation/linear.go:8-38
    state = {"t0": 0.0, "v0": 0.0, "valid": False}

    def fn(c, w, full, prev_row):
        if w.first_index == 0 and prev_row is not None:
            state["t0"], vt = prev_row.get_float64(w.interval_col_index, prev_row.num_rows() - 1)
            state["v0"], vv = prev_row.get_float64(c, prev_row.num_rows() - 1)
            state["valid"] = vt and vv
        t0, v0, prev_index = full.get_prev_float64s(w.interval_col_index, c, w.first_index - 1)
        if prev_index == -1:
            if not state["valid"]:
                return None
            t0, v0 = state["t0"], state["v0"]
        t2, v2, next_index = full.get_next_float64s(w.interval_col_index, c, w.first_index)
        if next_index == -1:
            return None
        d = t2 - t0
        num = float(w.first_value) - t0
        if d == 0.0:                                       # Go float division never panics
            coef = math.nan if (num == 0.0 or num != num) else math.copysign(math.inf, num) * math.copysign(1.0, d)
        else:
            coef = num / d
        return ((v2 - v0) * coef) + v0
    return ColInterpolation(col, [INT64, FLOAT64], fn)


# ----------------------------------------------------------------------------
# rolling/aggregation/whole.go:12-93 — aggregation.Aggregate: ONE window over the whole Bow
# ----------------------------------------------------------------------------
def whole_aggregate(b: Frame, interval_col_name: str, *aggrs: ColAggregation) -> Frame:
    if b is None:
        raise ValueError("nil bow")
    if len(aggrs) == 0:
        raise ValueError("at least one column aggregation is required")
    interval_col = b.column_index(interval_col_name)
    names, types, cols = [], [], []
    for i, a in enumerate(aggrs):
        if a.input_name == "":
            raise ValueError(f"column aggregation {i}: no input name")
        try:
            a.input_index = b.column_index(a.input_name)
        except KeyError as e:
            raise KeyError(f"column aggregation {i}: {e.args[0]}")
        name = a.output_name or b.names[a.input_index]
        typ = a.return_type(b.types[a.input_index], b.types[a.input_index])     # whole.go:44-46
        if b.num_rows() == 0:
            buf = []
        else:
            buf = [None]
            first_value, first_index = b.get_next_float64(interval_col, 0)       # whole.go:54-57
            if first_index == -1:
                first_value = -1.0
            last_value, last_index = b.get_prev_float64(interval_col, b.num_rows() - 1)
            if last_index == -1:
                last_value = -1.0
            w = Window(b, 0, interval_col, f64_to_i64(first_value), f64_to_i64(last_value), True)
            val = a.fn(a.input_index, w)
            for tr in a.transformations:
                val = tr(val)
            # Buffer.SetOrDropStrict (bowbuffer.go:82-104): plain type assertion, no conversion
            if typ == INT64 and isinstance(val, int) and not isinstance(val, bool):
                buf[0] = val
            elif typ == FLOAT64 and isinstance(val, float):
                buf[0] = val
        names.append(name)
        types.append(typ)
        cols.append(buf)
    return Frame(names, types, cols)


# ----------------------------------------------------------------------------
# bowfill.go:14-288, bowassertion.go:15-87 — whole-column fills (Int64 / Float64 columns)
# ----------------------------------------------------------------------------
def is_col_empty(b: Frame, c: int) -> bool:               # bowassertion.go:84-87
    return all(b.get_value(c, r) is None for r in range(b.num_rows()))


def is_col_sorted(b: Frame, c: int) -> bool:              # bowassertion.go:15-81 (nil values are skipped)
    if is_col_empty(b, c):
        return False
    vals = [b.get_value(c, r) for r in range(b.num_rows()) if b.get_value(c, r) is not None]
    order, curr = 0, vals[0]
    for nxt in vals[1:]:
        if order == 0:
            if curr < nxt:
                order = 1
            elif curr > nxt:
                order = -1
        if (order == 1 and nxt < curr) or (order == -1 and nxt > curr):
            return False
        curr = nxt
    return True


def go_round(x: float) -> float:
    """math.Round: half away from zero; NaN and +-Inf pass through"""
    if x != x or x in (math.inf, -math.inf):
        return x
    a = abs(x)
    if a >= 2.0 ** 52:
        return x
    t = float(math.floor(a))
    if a - t >= 0.5:          # exact: both are doubles below 2^52
        t += 1.0
    return math.copysign(t, x)


def _select_cols(b: Frame, col_indices) -> List[bool]:    # bowfill.go:266-288
    sel = [False] * b.num_cols()
    if len(col_indices) == 0:
        return [True] * b.num_cols()
    for c in col_indices:
        if c < 0 or c > b.num_cols() - 1:
            raise ValueError(f"selectCols: colIndex '{c}' out of range")
        sel[c] = True
    return sel


def _fill(method: str, b: Frame, *col_indices: int) -> Frame:   # bowfill.go:166-253
    sel = _select_cols(b, col_indices)
    cols = []
    for c in range(b.num_cols()):
        col = [b.get_value(c, r) for r in range(b.num_rows())]
        if sel[c] and any(v is None for v in col):
            for r in range(b.num_rows()):
                if col[r] is not None:
                    continue
                rr = r - 1 if method == "Previous" else r + 1       # getFillRowIndex, bowfill.go:255-264
                while 0 <= rr < b.num_rows() and b.get_value(c, rr) is None:
                    rr += -1 if method == "Previous" else 1
                if 0 <= rr < b.num_rows():
                    col[r] = b.get_value(c, rr)
        cols.append(col)
    return Frame(list(b.names), list(b.types), cols)


def fill_previous(b: Frame, *col_indices: int) -> Frame:  # bowfill.go:160-164
    return _fill("Previous", b, *col_indices)


def fill_next(b: Frame, *col_indices: int) -> Frame:      # bowfill.go:154-158
    return _fill("Next", b, *col_indices)


def fill_mean(b: Frame, *col_indices: int) -> Frame:      # bowfill.go:104-152
    sel = _select_cols(b, col_indices)
    for c in range(b.num_cols()):
        if sel[c] and b.types[c] not in (INT64, FLOAT64):
            raise TypeError(f"column '{b.names[c]}' is of unsupported type '{b.types[c]}'")
    cols = []
    for c in range(b.num_cols()):
        col = [b.get_value(c, r) for r in range(b.num_rows())]
        if sel[c] and any(v is None for v in col):
            for r in range(b.num_rows()):
                if col[r] is not None:
                    continue
                prev_val, prev_row = b.get_prev_float64(c, r - 1)
                next_val, next_row = b.get_next_float64(c, r + 1)
                if prev_row > -1 and next_row > -1:
                    m = (prev_val + next_val) / 2
                    col[r] = f64_to_i64(go_round(m)) if b.types[c] == INT64 else m
        cols.append(col)
    return Frame(list(b.names), list(b.types), cols)


def fill_linear(b: Frame, ref: int, to_fill: int) -> Frame:     # bowfill.go:14-102
    if ref < 0 or ref > b.num_cols() - 1:
        raise ValueError("refColIndex is out of range")
    if to_fill < 0 or to_fill > b.num_cols() - 1:
        raise ValueError("toFillColIndex is out of range")
    if ref == to_fill:
        raise ValueError("refColIndex and toFillColIndex are equal")
    if b.types[ref] not in (INT64, FLOAT64):
        raise TypeError(f"refColIndex '{ref}' is of type '{b.types[ref]}'")
    if is_col_empty(b, ref):
        return b
    if not is_col_sorted(b, ref):
        raise ValueError(f"refColIndex '{ref}' is empty or not sorted")
    if b.types[to_fill] not in (INT64, FLOAT64):
        raise TypeError(f"toFillColIndex '{to_fill}' is of unsupported type '{b.types[to_fill]}'")
    col = [b.get_value(to_fill, r) for r in range(b.num_rows())]
    if all(v is not None for v in col):
        return b
    is_int = b.types[to_fill] == INT64
    for r in range(b.num_rows()):
        if col[r] is not None:
            continue
        prev_to_fill, row_prev = b.get_prev_float64(to_fill, r - 1)
        next_to_fill, row_next = b.get_next_float64(to_fill, r + 1)
        row_ref, valid1 = b.get_float64(ref, r)
        prev_ref, valid2 = b.get_float64(ref, row_prev)
        next_ref, valid3 = b.get_float64(ref, row_next)
        if not (valid1 and valid2 and valid3):
            continue
        # (bowfill.go:76-83 stores prevToFill when nextRef == prevRef, then falls through and overwrites it)
        tmp = row_ref - prev_ref
        tmp = go_fdiv(tmp, next_ref - prev_ref)
        tmp *= next_to_fill - prev_to_fill
        tmp += prev_to_fill
        col[r] = f64_to_i64(go_round(tmp)) if is_int else tmp
    cols = [[b.get_value(c, r) for r in range(b.num_rows())] if c != to_fill else col for c in range(b.num_cols())]
    return Frame(list(b.names), list(b.types), cols)


def drop_nils(b: Frame, *col_indices: int) -> Frame:       # bow.go:188-224
    sel = _select_cols(b, col_indices)
    keep = [r for r in range(b.num_rows())
            if not any(sel[c] and b.get_value(c, r) is None for c in range(b.num_cols()))]
    return Frame(list(b.names), list(b.types), [[b.get_value(c, r) for r in keep] for c in range(b.num_cols())])


def sort_by_col(b: Frame, col_index: int) -> Frame:        # bowsort.go:10-47
    """SortByCol.  The reference calls sort.Sort (bowsort.go:24) over Buffer.Less / Swap (bowbuffer.go:126-139,202-206),
    which is NOT stable: the order among equal keys depends on the algorithm of the Go toolchain that built the
    program (quicksort + insertion sort before go1.19, pattern-defeating quicksort since), not on anything in the
    reference's sources.  Both toolchains run a plain insertion sort on slices of at most 12 elements, which is what
    is transliterated here for every length: equal keys keep their input order.  That is the order every golden
    vector of the reference shows (bowsort_test.go:11-208, at most 4 rows); on longer inputs with duplicate keys it
    is one valid outcome of the reference's contract, not necessarily the Go runtime's — PARITY UNPINNED there."""
    nulls = sum(1 for r in range(b.num_rows()) if b.get_value(col_index, r) is None)
    if nulls != 0:
        raise ValueError(f"column to sort by has {nulls} nil values")        # bowsort.go:11-15
    keys = [b.get_value(col_index, r) for r in range(b.num_rows())]
    indices = list(range(b.num_rows()))
    if all(not (keys[i] < keys[i - 1]) for i in range(1, len(keys))):      # sort.IsSorted, bowsort.go:18-21
        return b
    for i in range(1, len(keys)):                                          # insertionSort(data, 0, n)
        j = i
        while j > 0 and keys[j] < keys[j - 1]:                             # data.Less(j, j-1)
            keys[j], keys[j - 1] = keys[j - 1], keys[j]                    # data.Swap(j, j-1): value and index
            indices[j], indices[j - 1] = indices[j - 1], indices[j]
            j -= 1
    cols = [keys if c == col_index else [b.get_value(c, r) for r in indices] for c in range(b.num_cols())]
    return Frame(list(b.names), list(b.types), cols)
