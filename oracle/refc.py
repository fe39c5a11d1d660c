"""TEST INFRASTRUCTURE ONLY — ctypes wrapper around oracle/libbowref.so (oracle/ref.c).

Columns are (numpy values[int64|float64], numpy bool valid-mask or None).  The
wrapper packs Arrow-style LSB-first validity bitmaps (optionally with a non-zero
element offset, to exercise sliced inputs) and calls the C restatement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbowref.so")

FLOAT64, INT64 = 1, 2
AGG = dict(WindowStart=0, Count=1, Sum=2, ArithmeticMean=3, Min=4, Max=5, First=6, Last=7,
           IntegralStep=8, IntegralTrapezoid=9, WeightedAverageStep=10, WeightedAverageLinear=11)
INTERP = dict(WindowStart=0, Linear=1, StepPrevious=2, None_=3, StepNext=4)
ERRORS = {1: "EINVAL", 2: "ETYPE", 3: "EFIRSTNULL", 4: "EPREVROW", 5: "ENOINTERVALCOL", 6: "ECAPACITY"}


class RefError(RuntimeError):
    def __init__(self, code):
        super().__init__(ERRORS.get(code, str(code)))
        self.code = code


class Col(C.Structure):
    _fields_ = [("values", C.c_void_p), ("validity", C.c_void_p), ("offset", C.c_int64),
                ("length", C.c_int64), ("dtype", C.c_int32), ("_pad", C.c_int32)]


class Rolling(C.Structure):
    _fields_ = [("cols", C.POINTER(Col)), ("ncols", C.c_int32), ("time_col", C.c_int32), ("nrows", C.c_int64),
                ("interval", C.c_int64), ("offset", C.c_int64), ("inclusive", C.c_int32), ("_pad", C.c_int32),
                ("prev_row", C.POINTER(Col)), ("num_windows", C.c_int64), ("curr_window_first_value", C.c_int64),
                ("curr_row_index", C.c_int64), ("curr_window_index", C.c_int64)]


class AggSpec(C.Structure):
    _fields_ = [("op", C.c_int32), ("col", C.c_int32), ("nfactors", C.c_int32), ("_pad", C.c_int32),
                ("factors", C.c_double * 4)]


class OutCol(C.Structure):
    _fields_ = [("values", C.c_void_p), ("validity", C.c_void_p), ("dtype", C.c_int32), ("_pad", C.c_int32)]


_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "ref.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libbowref.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.bowref_interpolate.restype = C.c_int64
        _lib.bowref_windows.restype = C.c_int64
        assert _lib.bowref_sizeof_rolling() == C.sizeof(Rolling)
    return _lib


def pack_bits(mask: np.ndarray, offset: int = 0) -> np.ndarray:
    """LSB-first validity bitmap with `offset` leading garbage bits (set to 1 to catch misuse)."""
    full = np.concatenate([np.ones(offset, dtype=bool), mask.astype(bool)])
    return np.packbits(full, bitorder="little")


def unpack_bits(bitmap: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(bitmap, bitorder="little")[:n].astype(bool)


class Frame:
    """Column set handed to the C oracle.  cols: list of (values ndarray, mask ndarray|None)."""

    def __init__(self, cols: Sequence[Tuple[np.ndarray, Optional[np.ndarray]]], offset: int = 0):
        self.keep = []
        self.n = len(cols[0][0]) if cols else 0
        self.ncols = len(cols)
        self.dtypes = []
        self.arr = (Col * max(1, len(cols)))()
        for j, (v, m) in enumerate(cols):
            v = np.ascontiguousarray(v)
            assert v.dtype in (np.int64, np.float64), v.dtype
            if offset:
                v = np.concatenate([np.full(offset, 123456789, dtype=v.dtype), v])
            bm = None
            if m is not None:
                bm = pack_bits(np.asarray(m, dtype=bool), offset)
            self.keep += [v, bm]
            self.arr[j].values = v.ctypes.data
            self.arr[j].validity = bm.ctypes.data if bm is not None else None
            self.arr[j].offset = offset
            self.arr[j].length = self.n
            self.arr[j].dtype = INT64 if v.dtype == np.int64 else FLOAT64
            self.dtypes.append(self.arr[j].dtype)


class RefRolling:
    def __init__(self, frame: Frame, time_col: int, interval: int, offset: int = 0, inclusive: bool = False,
                 prev_row: Optional[Frame] = None):
        self.frame, self.prev = frame, prev_row
        self.r = Rolling()
        rc = lib().bowref_rolling_init(C.byref(self.r), frame.arr, frame.ncols, time_col, C.c_int64(interval),
                                       C.c_int64(offset), int(inclusive), prev_row.arr if prev_row else None)
        if rc:
            raise RefError(rc)

    @property
    def num_windows(self) -> int:
        return self.r.num_windows

    @property
    def first_window_start(self) -> int:
        return self.r.curr_window_first_value

    def windows(self):
        """-> dict of arrays: first_index, lo, hi, first_value, is_inclusive (one entry per window)"""
        W = self.num_windows
        fi, lo, hi, fv = (np.zeros(W, dtype=np.int64) for _ in range(4))
        inc = np.zeros(W, dtype=np.uint8)
        n = lib().bowref_windows(C.byref(self.r), fi.ctypes.data_as(C.c_void_p), lo.ctypes.data_as(C.c_void_p),
                                 hi.ctypes.data_as(C.c_void_p), fv.ctypes.data_as(C.c_void_p),
                                 inc.ctypes.data_as(C.c_void_p))
        assert n <= W, (n, W)   # n < W only with trailing null timestamps (HasNext, rolling.go:167-170)
        return dict(first_index=fi[:n], lo=lo[:n], hi=hi[:n], first_value=fv[:n], is_inclusive=inc[:n].astype(bool))

    def aggregate(self, specs: Sequence[tuple]):
        """specs: (op name|code, col index[, [factors]]) -> list of (values ndarray, valid mask ndarray)"""
        W = self.num_windows
        arr = (AggSpec * len(specs))()
        outs = (OutCol * len(specs))()
        bufs = []
        for j, s in enumerate(specs):
            op = AGG[s[0]] if isinstance(s[0], str) else s[0]
            arr[j].op, arr[j].col = op, s[1]
            fs = list(s[2]) if len(s) > 2 and s[2] else []
            arr[j].nfactors = len(fs)
            for k, f in enumerate(fs):
                arr[j].factors[k] = f
            v = np.zeros(max(W, 1), dtype=np.int64)
            b = np.zeros((W + 7) // 8 + 1, dtype=np.uint8)
            bufs.append((v, b))
            outs[j].values, outs[j].validity = v.ctypes.data, b.ctypes.data
        rc = lib().bowref_aggregate(C.byref(self.r), arr, len(specs), outs)
        if rc:
            raise RefError(rc)
        res = []
        for j, (v, b) in enumerate(bufs):
            vals = v[:W] if outs[j].dtype == INT64 else v[:W].view(np.float64)
            res.append((vals, unpack_bits(b, W)))
        return res

    def interpolate(self, ops: Sequence):
        codes = (C.c_int32 * len(ops))(*[INTERP[o] if isinstance(o, str) else o for o in ops])
        n_out = lib().bowref_interpolate(C.byref(self.r), codes, len(ops), None, None, C.c_int64(0))
        if n_out < 0:
            raise RefError(-n_out)
        vals = [np.zeros(max(n_out, 1), dtype=np.int64) for _ in ops]
        bms = [np.zeros((n_out + 7) // 8 + 1, dtype=np.uint8) for _ in ops]
        pv = (C.c_void_p * len(ops))(*[v.ctypes.data for v in vals])
        pb = (C.c_void_p * len(ops))(*[b.ctypes.data for b in bms])
        n2 = lib().bowref_interpolate(C.byref(self.r), codes, len(ops), pv, pb, C.c_int64(n_out))
        assert n2 == n_out, (n2, n_out)
        res = []
        for j in range(len(ops)):
            v = vals[j][:n_out] if self.frame.dtypes[j] == INT64 else vals[j][:n_out].view(np.float64)
            res.append((v, unpack_bits(bms[j], n_out)))
        return res


def aggregate_whole(frame: Frame, time_col: int, specs: Sequence[tuple]):
    """aggregation.Aggregate over the whole frame (rolling/aggregation/whole.go) -> list of (values, valid mask)
    with one entry each (zero entries for an empty frame)."""
    n_out = 1 if frame.n else 0
    arr = (AggSpec * len(specs))()
    outs = (OutCol * len(specs))()
    bufs = []
    for j, s in enumerate(specs):
        op = AGG[s[0]] if isinstance(s[0], str) else s[0]
        arr[j].op, arr[j].col = op, s[1]
        fs = list(s[2]) if len(s) > 2 and s[2] else []
        arr[j].nfactors = len(fs)
        for k, f in enumerate(fs):
            arr[j].factors[k] = f
        v = np.zeros(1, dtype=np.int64)
        b = np.zeros(1, dtype=np.uint8)
        bufs.append((v, b))
        outs[j].values, outs[j].validity = v.ctypes.data, b.ctypes.data
    rc = lib().bowref_aggregate_whole(frame.arr, frame.ncols, time_col, arr, len(specs), outs)
    if rc:
        raise RefError(rc)
    res = []
    for j, (v, b) in enumerate(bufs):
        vals = v[:n_out] if outs[j].dtype == INT64 else v[:n_out].view(np.float64)
        res.append((vals, unpack_bits(b, n_out)))
    return res


FILL = dict(Previous=0, Next=1, Mean=2, Linear=3)


def fill(frame: Frame, method, col: int, ref_col: int = -1):
    """FillPrevious / FillNext / FillMean / FillLinear of ONE column (bowfill.go) -> (values, valid mask)"""
    n = frame.n
    v = np.zeros(max(n, 1), dtype=np.int64)
    b = np.zeros((n + 7) // 8 + 1, dtype=np.uint8)
    m = FILL[method] if isinstance(method, str) else method
    rc = lib().bowref_fill(frame.arr, frame.ncols, m, col, ref_col, v.ctypes.data_as(C.c_void_p),
                           b.ctypes.data_as(C.c_void_p))
    if rc:
        raise RefError(rc)
    vals = v[:n] if frame.dtypes[col] == INT64 else v[:n].view(np.float64)
    return vals, unpack_bits(b, n)


# ---- numpy restatements (byte / index work, no C needed) ------------------------------------------------------------
def drop_nils(cols, selected=None):
    """Bow.DropNils (bow.go:188-224): rows with a nil in a selected column (default: any column) are dropped.
    cols: list of (values, mask | None) -> same layout"""
    n = len(cols[0][0]) if cols else 0
    keep = np.ones(n, dtype=bool)
    for j, (_, m) in enumerate(cols):
        if (selected is None or len(selected) == 0 or j in selected) and m is not None:
            keep &= np.asarray(m, dtype=bool)
    return [(np.asarray(v)[keep], np.ones(int(keep.sum()), dtype=bool) if m is None else np.asarray(m, dtype=bool)[keep])
            for v, m in cols]


def is_col_sorted(values, mask=None) -> bool:
    """Bow.IsColSorted (bowassertion.go:15-81): nil values skipped, ascending or descending, empty column -> False"""
    v = np.asarray(values)
    if mask is not None:
        v = v[np.asarray(mask, dtype=bool)]
    if len(v) == 0:
        return False
    lt, gt = bool((v[1:] < v[:-1]).any()), bool((v[1:] > v[:-1]).any())
    return not (lt and gt)


def sort_by_col(cols, col):
    """Bow.SortByCol (bowsort.go:10-47) with equal keys kept in input order (see oracle/literal.py sort_by_col: the
    reference's sort.Sort leaves that order to the Go toolchain — parity unpinned for duplicate keys beyond the golden
    vectors).  cols: list of (values, mask | None) -> None when already sorted (the reference returns b itself), else
    the same layout.  Raises ValueError for a sort column with nils (bowsort.go:11-15)."""
    v, m = cols[col]
    v = np.asarray(v)
    if m is not None and not np.asarray(m, dtype=bool).all():
        raise ValueError(f"column to sort by has {int((~np.asarray(m, dtype=bool)).sum())} nil values")
    if len(v) < 2 or not (v[1:] < v[:-1]).any():        # sort.IsSorted over Buffer.Less
        return None
    order = np.argsort(v, kind="stable")                # (-0.0 == 0.0 for numpy as for Go's `<`)
    out = []
    for vv, mm in cols:
        vv = np.asarray(vv)
        out.append((vv[order], np.ones(len(vv), dtype=bool) if mm is None else np.asarray(mm, dtype=bool)[order]))
    return out
