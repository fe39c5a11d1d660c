"""ctypes binding of libbowgpu.so (include/bowgpu.h) — the same C ABI the Go cgo shim binds.

This module holds no algorithm: it marshals Arrow-layout buffers (numpy / pyarrow / raw device
pointers) into `bowgpu_col` descriptors and calls the library.  If the library is missing or no
B200 is visible it raises — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BOWGPU_LIB") or os.path.join(HERE, "libbowgpu.so")  # BOWGPU_LIB: tuning builds only

FLOAT64, INT64 = 1, 2
MEM_HOST, MEM_DEVICE = 0, 1

AGG = dict(WindowStart=0, Count=1, Sum=2, ArithmeticMean=3, Min=4, Max=5, First=6, Last=7,
           IntegralStep=8, IntegralTrapezoid=9, WeightedAverageStep=10, WeightedAverageLinear=11)
INTERP = dict(WindowStart=0, Linear=1, StepPrevious=2, None_=3, StepNext=4)
FILL = dict(Previous=0, Next=1, Mean=2, Linear=3)
STATUS = {0: "OK", 1: "EINVAL", 2: "ETYPE", 3: "EFIRSTNULL", 4: "EPREVROW", 5: "ENOINTERVALCOL", 6: "ECAPACITY",
          7: "EUNSORTED", 8: "ENULLTIME", 9: "ECUDA", 10: "ENOMEM", 11: "EUNSUPPORTED", 12: "EIO"}

# every symbol include/bowgpu.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "bowgpu_abi_version", "bowgpu_ctx_create", "bowgpu_ctx_destroy", "bowgpu_last_error", "bowgpu_status_string",
    "bowgpu_ctx_synchronize", "bowgpu_ctx_enable_timing", "bowgpu_ctx_last_timing", "bowgpu_ctx_sm_count", "bowgpu_ctx_trim",
    "bowgpu_frame_create", "bowgpu_frame_destroy", "bowgpu_frame_num_rows", "bowgpu_frame_num_cols",
    "bowgpu_frame_col_dtype", "bowgpu_frame_col_has_validity", "bowgpu_frame_col_device_ptrs",
    "bowgpu_frame_download", "bowgpu_frame_download_range", "bowgpu_frame_generate",
    "bowgpu_rolling_create", "bowgpu_rolling_create_shard", "bowgpu_rolling_destroy", "bowgpu_rolling_num_windows",
    "bowgpu_rolling_first_window_start", "bowgpu_rolling_inclusive", "bowgpu_rolling_early_rows",
    "bowgpu_rolling_bounds", "bowgpu_rolling_aggregate", "bowgpu_agg_return_type", "bowgpu_agg_needs_inclusive",
    "bowgpu_rolling_interpolate", "bowgpu_frame_aggregate_whole", "bowgpu_frame_fill", "bowgpu_frame_fill_linear",
    "bowgpu_rolling_interpolate_aggregate", "bowgpu_frame_drop_nils", "bowgpu_frame_is_col_sorted",
    "bowgpu_aggregate_host", "bowgpu_frame_sort_by_col", "bowgpu_aggregate_host_ex", "bowgpu_interpolate_aggregate_host",
    "bowgpu_parquet_open", "bowgpu_parquet_close", "bowgpu_parquet_num_rows", "bowgpu_parquet_num_cols",
    "bowgpu_parquet_col_name", "bowgpu_parquet_col_dtype", "bowgpu_parquet_col_physical_type", "bowgpu_parquet_read", "bowgpu_parquet_plan",
    "bowgpu_frame_col_null_count",
]


class BowGpuError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        self.code = code
        self.status = STATUS.get(code, str(code))
        super().__init__(f"{self.status}: {msg}" if msg else self.status)


class Col(C.Structure):
    _fields_ = [("values", C.c_void_p), ("validity", C.c_void_p), ("offset", C.c_int64), ("length", C.c_int64),
                ("null_count", C.c_int64), ("dtype", C.c_int32), ("_pad", C.c_int32)]


class AggSpec(C.Structure):
    _fields_ = [("op", C.c_int32), ("col", C.c_int32), ("nfactors", C.c_int32), ("_pad", C.c_int32),
                ("factors", C.c_double * 4)]


class OutCol(C.Structure):
    _fields_ = [("values", C.c_void_p), ("validity", C.c_void_p), ("dtype", C.c_int32), ("_pad", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("main_ms", C.c_float), ("launches", C.c_int32),
                ("main_launches", C.c_int32)]


class HostOpts(C.Structure):
    _fields_ = [("devices", C.POINTER(C.c_int32)), ("ndevices", C.c_int32), ("workers_per_device", C.c_int32),
                ("chunk_rows", C.c_int64), ("shard", C.c_int32), ("_pad", C.c_int32), ("s0", C.c_int64),
                ("num_windows", C.c_int64)]


class GenSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ncols", C.c_int32), ("nrows", C.c_int64), ("row0", C.c_int64),
                ("t0", C.c_int64), ("step", C.c_int64), ("seed", C.c_uint64), ("null_mask", C.c_uint32),
                ("int_mask", C.c_uint32), ("null_mod", C.c_uint32), ("_pad", C.c_uint32)]


_lib = None


def lib():
    """Loads libbowgpu.so (built by __graft_entry__.build()).  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(bow_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.bowgpu_last_error.restype = C.c_char_p
        L.bowgpu_status_string.restype = C.c_char_p
        for name in ("bowgpu_frame_num_rows", "bowgpu_rolling_num_windows", "bowgpu_rolling_first_window_start",
                     "bowgpu_rolling_early_rows"):
            getattr(L, name).restype = C.c_int64
        L.bowgpu_ctx_destroy.restype = None
        L.bowgpu_frame_destroy.restype = None
        L.bowgpu_rolling_destroy.restype = None
        L.bowgpu_ctx_create.argtypes = [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
        L.bowgpu_frame_create.argtypes = [C.c_void_p, C.POINTER(Col), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.bowgpu_frame_generate.argtypes = [C.c_void_p, C.POINTER(GenSpec), C.POINTER(C.c_void_p)]
        L.bowgpu_frame_download.argtypes = [C.c_void_p, C.POINTER(OutCol), C.c_int32]
        L.bowgpu_frame_download_range.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(OutCol), C.c_int32]
        L.bowgpu_frame_col_device_ptrs.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p),
                                                   C.POINTER(C.c_void_p)]
        L.bowgpu_rolling_create.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.POINTER(Col),
                                            C.POINTER(C.c_void_p)]
        L.bowgpu_rolling_create_shard.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                                  C.POINTER(Col), C.POINTER(C.c_void_p)]
        L.bowgpu_rolling_bounds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bowgpu_rolling_aggregate.argtypes = [C.c_void_p, C.POINTER(AggSpec), C.c_int32, C.POINTER(OutCol),
                                               C.c_int32]
        L.bowgpu_rolling_interpolate_aggregate.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(AggSpec),
                                                           C.c_int32, C.POINTER(OutCol), C.c_int32]
        L.bowgpu_aggregate_host.argtypes = [C.c_void_p, C.POINTER(Col), C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32,
                                            C.POINTER(AggSpec), C.c_int32, C.POINTER(OutCol), C.c_int64,
                                            C.POINTER(C.c_int64)]
        L.bowgpu_aggregate_host_ex.argtypes = L.bowgpu_aggregate_host.argtypes + [C.POINTER(HostOpts)]
        L.bowgpu_interpolate_aggregate_host.argtypes = [C.c_void_p, C.POINTER(Col), C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                                        C.POINTER(Col), C.POINTER(C.c_int32), C.c_int32, C.POINTER(AggSpec),
                                                        C.c_int32, C.POINTER(OutCol), C.c_int64, C.POINTER(C.c_int64),
                                                        C.POINTER(HostOpts)]
        L.bowgpu_frame_drop_nils.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]
        L.bowgpu_frame_is_col_sorted.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        L.bowgpu_frame_sort_by_col.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.bowgpu_frame_fill.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]
        L.bowgpu_frame_fill_linear.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.bowgpu_frame_aggregate_whole.argtypes = [C.c_void_p, C.c_int32, C.POINTER(AggSpec), C.c_int32, C.POINTER(OutCol),
                                                   C.c_int32]
        L.bowgpu_rolling_interpolate.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32,
                                                 C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.bowgpu_rolling_early_rows.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        for name in ("bowgpu_ctx_destroy", "bowgpu_frame_destroy", "bowgpu_rolling_destroy", "bowgpu_last_error",
                     "bowgpu_ctx_synchronize", "bowgpu_ctx_sm_count", "bowgpu_ctx_trim", "bowgpu_frame_num_rows",
                     "bowgpu_frame_num_cols", "bowgpu_rolling_num_windows", "bowgpu_rolling_first_window_start",
                     "bowgpu_rolling_inclusive"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.bowgpu_ctx_enable_timing.argtypes = [C.c_void_p, C.c_int32]
        L.bowgpu_ctx_last_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
        L.bowgpu_frame_col_dtype.argtypes = [C.c_void_p, C.c_int32]
        L.bowgpu_frame_col_has_validity.argtypes = [C.c_void_p, C.c_int32]
        L.bowgpu_frame_col_null_count.argtypes = [C.c_void_p, C.c_int32]
        L.bowgpu_frame_col_null_count.restype = C.c_int64
        L.bowgpu_parquet_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_int32]
        L.bowgpu_parquet_close.argtypes = [C.c_void_p]
        L.bowgpu_parquet_close.restype = None
        L.bowgpu_parquet_num_rows.argtypes = [C.c_void_p]
        L.bowgpu_parquet_num_rows.restype = C.c_int64
        L.bowgpu_parquet_num_cols.argtypes = [C.c_void_p]
        L.bowgpu_parquet_col_name.argtypes = [C.c_void_p, C.c_int32]
        L.bowgpu_parquet_col_name.restype = C.c_char_p
        L.bowgpu_parquet_col_dtype.argtypes = [C.c_void_p, C.c_int32]
        L.bowgpu_parquet_col_physical_type.argtypes = [C.c_void_p, C.c_int32]
        L.bowgpu_parquet_plan.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int64), C.c_char_p, C.c_int32]
        L.bowgpu_parquet_read.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]
        assert L.bowgpu_abi_version() == 1
        _lib = L
    return _lib


def pack_bits(mask: np.ndarray, offset: int = 0) -> np.ndarray:
    """LSB-first Arrow validity bitmap with `offset` leading bits (set, so misuse shows)."""
    full = np.concatenate([np.ones(offset, dtype=bool), np.asarray(mask, dtype=bool)])
    return np.packbits(full, bitorder="little")


def unpack_bits(bitmap: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(np.asarray(bitmap, dtype=np.uint8), bitorder="little")[:n].astype(bool)


NpCol = Tuple[np.ndarray, Optional[np.ndarray]]


def cols_from_numpy(cols: Sequence[NpCol], offset: int = 0):
    """-> (ctypes Col array, keep-alive list).  cols: (values int64|float64, bool mask | None)."""
    keep = []
    arr = (Col * max(1, len(cols)))()
    for j, (v, m) in enumerate(cols):
        v = np.ascontiguousarray(v)
        assert v.dtype in (np.int64, np.float64), v.dtype
        n = len(v)
        if offset:
            v = np.concatenate([np.full(offset, 123456789, dtype=v.dtype), v])
        bm = pack_bits(m, offset) if m is not None else None
        keep += [v, bm]
        arr[j].values = v.ctypes.data
        arr[j].validity = bm.ctypes.data if bm is not None else None
        arr[j].offset = offset
        arr[j].length = n
        arr[j].null_count = int(n - np.count_nonzero(m)) if m is not None else 0
        arr[j].dtype = INT64 if v.dtype == np.int64 else FLOAT64
    return arr, keep


class Ctx:
    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.h = C.c_void_p()
        rc = lib().bowgpu_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self.h))
        if rc:
            raise BowGpuError(rc, "bowgpu_ctx_create failed (is a B200 visible? bow_b200 has no CPU fallback)")
        self.device = device
        self._children = weakref.WeakSet()   # frames / rollings created on this ctx: closed before the ctx itself

    def _adopt(self, obj):
        self._children.add(obj)
        return obj

    def check(self, rc: int):
        if rc:
            raise BowGpuError(rc, lib().bowgpu_last_error(self.h).decode())

    def synchronize(self):
        self.check(lib().bowgpu_ctx_synchronize(self.h))

    def enable_timing(self, on: bool = True):
        self.check(lib().bowgpu_ctx_enable_timing(self.h, int(on)))

    def last_timing(self) -> Timing:
        t = Timing()
        self.check(lib().bowgpu_ctx_last_timing(self.h, C.byref(t)))
        return t

    def trim(self):
        """returns cached column buffers of destroyed frames to the driver"""
        self.check(lib().bowgpu_ctx_trim(self.h))

    @property
    def sm_count(self) -> int:
        return lib().bowgpu_ctx_sm_count(self.h)

    def close(self):
        if self.h:
            for child in list(self._children):   # a frame outliving its ctx would free into a destroyed pool
                child.close()
            lib().bowgpu_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def make_host_opts(devices: Optional[Sequence[int]] = None, workers_per_device: int = 0, chunk_rows: int = 0,
                   shard: Optional[Tuple[int, int]] = None):
    """-> (HostOpts, keep-alive) for the one-shot host calls; shard = (s0, num_windows) of a range-partitioned shard"""
    o = HostOpts()
    keep = None
    if devices:
        keep = (C.c_int32 * len(devices))(*devices)
        o.devices, o.ndevices = keep, len(devices)
    o.workers_per_device, o.chunk_rows = workers_per_device, chunk_rows
    if shard is not None:
        o.shard, o.s0, o.num_windows = 1, int(shard[0]), int(shard[1])
    return o, keep


def _host_outs(specs, W):
    outs = (OutCol * len(specs))()
    bufs = []
    for j in range(len(specs)):
        v = np.full(max(W, 1), -7, dtype=np.int64)
        b = np.full((W + 7) // 8 + 1, 0xAA, dtype=np.uint8)
        bufs.append((v, b))
        outs[j].values, outs[j].validity = v.ctypes.data, b.ctypes.data
    return outs, bufs


def _host_results(outs, bufs, W):
    res = []
    for j, (v, b) in enumerate(bufs):
        vals = v[:W] if outs[j].dtype == INT64 else v[:W].view(np.float64)
        res.append((vals, unpack_bits(b, W)))
    return res


def _count_windows(t, interval, offset):   # countWindows on the host (rolling.go:96-99,143-154), like the Go side
    from . import partition as P
    return P.num_windows(int(t[0]), int(t[-1]), interval, P.normalise_offset(interval, offset)) if len(t) else 0


def aggregate_host(ctx: "Ctx", cols: Sequence[NpCol], time_col: int, interval: int, specs: Sequence[tuple],
                   offset: int = 0, inclusive: bool = False, num_windows: Optional[int] = None,
                   devices: Optional[Sequence[int]] = None, workers_per_device: int = 0, chunk_rows: int = 0,
                   shard: Optional[Tuple[int, int]] = None):
    """One-shot pipelined IntervalRolling -> Aggregate from host columns to host results (bowgpu_aggregate_host_ex),
    spread over `devices` (default: the ctx's GPU) -> list of (values ndarray, valid mask ndarray)"""
    arr, keep = cols_from_numpy(cols)
    if shard is not None:
        num_windows = shard[1]
    W = _count_windows(cols[time_col][0], interval, offset) if num_windows is None else num_windows
    sarr = make_specs(specs)
    outs, bufs = _host_outs(specs, W)
    opts, okeep = make_host_opts(devices, workers_per_device, chunk_rows, shard)
    got = C.c_int64()
    ctx.check(lib().bowgpu_aggregate_host_ex(ctx.h, arr, len(cols), time_col, interval, offset, int(inclusive), sarr,
                                             len(specs), outs, W, C.byref(got), C.byref(opts)))
    assert got.value == W, (got.value, W)
    return _host_results(outs, bufs, W)


def interpolate_aggregate_host(ctx: "Ctx", cols: Sequence[NpCol], time_col: int, interval: int, ops: Sequence,
                               specs: Sequence[tuple], offset: int = 0, prev_row: Optional[Sequence[NpCol]] = None,
                               devices: Optional[Sequence[int]] = None, workers_per_device: int = 0, chunk_rows: int = 0):
    """One-shot pipelined IntervalRolling -> Interpolate -> Aggregate from host columns to host results
    (bowgpu_interpolate_aggregate_host) -> list of (values ndarray, valid mask ndarray)"""
    arr, keep = cols_from_numpy(cols)
    W = _count_windows(cols[time_col][0], interval, offset)
    sarr = make_specs(specs)
    outs, bufs = _host_outs(specs, W)
    opts, okeep = make_host_opts(devices, workers_per_device, chunk_rows)
    codes = (C.c_int32 * len(ops))(*[INTERP[o] if isinstance(o, str) else int(o) for o in ops])
    parr, pkeep = (None, None) if prev_row is None else cols_from_numpy(prev_row)
    got = C.c_int64()
    ctx.check(lib().bowgpu_interpolate_aggregate_host(ctx.h, arr, len(cols), time_col, interval, offset, parr, codes, len(ops),
                                                      sarr, len(specs), outs, W, C.byref(got), C.byref(opts)))
    assert got.value == W, (got.value, W)
    return _host_results(outs, bufs, W)


class Frame:
    """Device-resident Bow."""

    def __init__(self, ctx: Ctx, handle: C.c_void_p, keep=None):
        self.ctx, self.h, self.keep = ctx, handle, keep
        ctx._adopt(self)

    @classmethod
    def from_numpy(cls, ctx: Ctx, cols: Sequence[NpCol], offset: int = 0) -> "Frame":
        arr, keep = cols_from_numpy(cols, offset)
        h = C.c_void_p()
        ctx.check(lib().bowgpu_frame_create(ctx.h, arr, len(cols), MEM_HOST, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_col_descs(cls, ctx: Ctx, arr, ncols: int, mem: int, keep=None) -> "Frame":
        h = C.c_void_p()
        ctx.check(lib().bowgpu_frame_create(ctx.h, arr, ncols, mem, C.byref(h)))
        return cls(ctx, h, keep if mem == MEM_DEVICE else None)

    @classmethod
    def generate(cls, ctx: Ctx, nrows: int, ncols: int = 1, row0: int = 0, t0: int = 1_700_000_000_000_000_000,
                 step: int = 1_000_000_000, seed: int = 42, null_mask: int = 0, int_mask: int = 0,
                 null_mod: int = 10, kind: int = 0) -> "Frame":
        g = GenSpec(kind, ncols, nrows, row0, t0, step, seed, null_mask, int_mask, null_mod, 0)
        h = C.c_void_p()
        ctx.check(lib().bowgpu_frame_generate(ctx.h, C.byref(g), C.byref(h)))
        return cls(ctx, h)

    @property
    def num_rows(self) -> int:
        return lib().bowgpu_frame_num_rows(self.h)

    @property
    def num_cols(self) -> int:
        return lib().bowgpu_frame_num_cols(self.h)

    def dtype(self, j: int) -> int:
        return lib().bowgpu_frame_col_dtype(self.h, j)

    def null_count(self, j: int) -> int:
        return lib().bowgpu_frame_col_null_count(self.h, j)

    def device_ptrs(self, j: int) -> Tuple[int, int]:
        v, b = C.c_void_p(), C.c_void_p()
        self.ctx.check(lib().bowgpu_frame_col_device_ptrs(self.h, j, C.byref(v), C.byref(b)))
        return v.value or 0, b.value or 0

    def download(self, row0: int = 0, nrows: Optional[int] = None) -> List[NpCol]:
        n = self.num_rows - row0 if nrows is None else nrows
        nc = self.num_cols
        outs = (OutCol * max(1, nc))()
        bufs = []
        for j in range(nc):
            v = np.zeros(max(n, 1), dtype=np.int64)
            b = np.zeros((n + 7) // 8 + 1, dtype=np.uint8)
            bufs.append((v, b))
            outs[j].values, outs[j].validity = v.ctypes.data, b.ctypes.data
        self.ctx.check(lib().bowgpu_frame_download_range(self.h, row0, n, outs, nc))
        res = []
        for j, (v, b) in enumerate(bufs):
            vals = v[:n] if outs[j].dtype == INT64 else v[:n].view(np.float64)
            res.append((vals, unpack_bits(b, n)))
        return res

    def fill(self, method, *cols: int) -> "Frame":
        """Bow.FillPrevious / FillNext / FillMean (bowfill.go); no column index = every column"""
        m = FILL[method] if isinstance(method, str) else int(method)
        arr = (C.c_int32 * max(1, len(cols)))(*cols)
        h = C.c_void_p()
        self.ctx.check(lib().bowgpu_frame_fill(self.h, m, arr, len(cols), C.byref(h)))
        return Frame(self.ctx, h)

    def drop_nils(self, *cols: int) -> "Frame":
        """Bow.DropNils (bow.go:188-224); no column index = any column"""
        arr = (C.c_int32 * max(1, len(cols)))(*cols)
        h = C.c_void_p()
        self.ctx.check(lib().bowgpu_frame_drop_nils(self.h, arr, len(cols), C.byref(h)))
        return Frame(self.ctx, h)

    def is_col_sorted(self, col: int) -> bool:
        """Bow.IsColSorted (bowassertion.go:15-81)"""
        out = C.c_int32()
        self.ctx.check(lib().bowgpu_frame_is_col_sorted(self.h, col, C.byref(out)))
        return bool(out.value)

    def sort_by_col(self, col: int) -> Optional["Frame"]:
        """Bow.SortByCol (bowsort.go:10-47); None = already sorted (the reference returns the same Bow)"""
        h = C.c_void_p()
        self.ctx.check(lib().bowgpu_frame_sort_by_col(self.h, col, C.byref(h)))
        return Frame(self.ctx, h) if h.value else None

    def fill_linear(self, ref_col: int, tofill_col: int) -> "Frame":
        """Bow.FillLinear (bowfill.go:14-102)"""
        h = C.c_void_p()
        self.ctx.check(lib().bowgpu_frame_fill_linear(self.h, ref_col, tofill_col, C.byref(h)))
        return Frame(self.ctx, h)

    def aggregate_whole(self, time_col: int, specs: Sequence[tuple]):
        """aggregation.Aggregate over the whole frame (rolling/aggregation/whole.go) -> list of (values, valid mask)
        with one entry each (none for an empty frame)"""
        n_out = 1 if self.num_rows else 0
        arr = make_specs(specs)
        outs = (OutCol * len(specs))()
        bufs = []
        for j in range(len(specs)):
            v = np.full(1, -7, dtype=np.int64)
            b = np.full(1, 0xAA, dtype=np.uint8)
            bufs.append((v, b))
            outs[j].values, outs[j].validity = v.ctypes.data, b.ctypes.data
        self.ctx.check(lib().bowgpu_frame_aggregate_whole(self.h, time_col, arr, len(specs), outs, MEM_HOST))
        res = []
        for j, (v, b) in enumerate(bufs):
            vals = v[:n_out] if outs[j].dtype == INT64 else v[:n_out].view(np.float64)
            res.append((vals, unpack_bits(b, n_out)))
        return res

    def aggregate_whole_device(self, time_col: int, specs_arr, nspecs: int, outs) -> None:
        """same, results left in device buffers (asynchronous)"""
        self.ctx.check(lib().bowgpu_frame_aggregate_whole(self.h, time_col, specs_arr, nspecs, outs, MEM_DEVICE))

    def close(self):
        if self.h:
            lib().bowgpu_frame_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class ParquetFile:
    """An open Parquet file: footer parsed on the host (bowparquet.go:44-66); read() decodes chosen columns on the GPU."""

    def __init__(self, path: str):
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib().bowgpu_parquet_open(os.fsencode(path), C.byref(h), err, 512)
        if rc:
            raise BowGpuError(rc, err.value.decode(errors="replace"))
        self.h = h
        L = lib()
        self.num_rows = L.bowgpu_parquet_num_rows(h)
        nc = L.bowgpu_parquet_num_cols(h)
        self.names = [L.bowgpu_parquet_col_name(h, j).decode() for j in range(nc)]
        self.dtypes = [L.bowgpu_parquet_col_dtype(h, j) for j in range(nc)]          # INT64 / FLOAT64 / 0 (no GPU type)
        self.physical = [L.bowgpu_parquet_col_physical_type(h, j) for j in range(nc)]  # parquet.Type

    def plan(self, cols: Optional[Sequence[int]] = None) -> dict:
        """The host-side page walk only: pages, uploaded bytes, scratch bytes, dictionary-index entries."""
        if cols is None:
            cols = [j for j, d in enumerate(self.dtypes) if d]
        arr = (C.c_int32 * max(1, len(cols)))(*cols)
        out = (C.c_int64 * 4)()
        err = C.create_string_buffer(512)
        rc = lib().bowgpu_parquet_plan(self.h, arr, len(cols), out, err, 512)
        if rc:
            raise BowGpuError(rc, err.value.decode(errors="replace"))
        return {"pages": out[0], "image_bytes": out[1], "scratch_bytes": out[2], "aux_entries": out[3]}

    def read(self, ctx: "Ctx", cols: Optional[Sequence[int]] = None) -> "Frame":
        """cols = leaf column indices (default: every column with a GPU type)."""
        if cols is None:
            cols = [j for j, d in enumerate(self.dtypes) if d]
        arr = (C.c_int32 * max(1, len(cols)))(*cols)
        h = C.c_void_p()
        ctx.check(lib().bowgpu_parquet_read(ctx.h, self.h, arr, len(cols), C.byref(h)))
        return Frame(ctx, h)

    def close(self):
        if self.h:
            lib().bowgpu_parquet_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def make_specs(specs: Sequence[tuple]):
    """specs: (op name|code, col index[, [factors]])"""
    arr = (AggSpec * len(specs))()
    for j, s in enumerate(specs):
        arr[j].op = AGG[s[0]] if isinstance(s[0], str) else int(s[0])
        arr[j].col = s[1]
        fs = list(s[2]) if len(s) > 2 and s[2] else []
        arr[j].nfactors = len(fs)
        for k, f in enumerate(fs):
            arr[j].factors[k] = f
    return arr


class Rolling:
    """intervalRolling on a device frame (bowgpu_rolling)."""

    def __init__(self, frame: Frame, time_col: int, interval: int, offset: int = 0, inclusive: bool = False,
                 prev_row: Optional[Sequence[NpCol]] = None, shard: Optional[Tuple[int, int]] = None):
        """shard = (s0, num_windows): range-partitioned variant (bowgpu_rolling_create_shard)."""
        self.frame, self.ctx = frame, frame.ctx
        self.h = C.c_void_p()
        self.ctx._adopt(self)
        parr, self._keep = (None, None)
        if prev_row is not None:
            parr, self._keep = cols_from_numpy(prev_row)
        if shard is not None:
            self.ctx.check(lib().bowgpu_rolling_create_shard(frame.h, time_col, interval, shard[0], shard[1],
                                                             int(inclusive), parr, C.byref(self.h)))
        else:
            self.ctx.check(lib().bowgpu_rolling_create(frame.h, time_col, interval, offset, int(inclusive), parr,
                                                       C.byref(self.h)))

    @property
    def num_windows(self) -> int:
        return lib().bowgpu_rolling_num_windows(self.h)

    @property
    def first_window_start(self) -> int:
        return lib().bowgpu_rolling_first_window_start(self.h)

    def early_rows(self) -> Tuple[int, bool]:
        kept = C.c_int32()
        n = lib().bowgpu_rolling_early_rows(self.h, C.byref(kept))
        return n, bool(kept.value)

    def bounds(self):
        """-> (first[W+1] int64, inclusive[W] bool)"""
        W = self.num_windows
        first = np.zeros(W + 1, dtype=np.int64)
        inc = np.zeros((W + 7) // 8 + 1, dtype=np.uint8)
        self.ctx.check(lib().bowgpu_rolling_bounds(self.h, first.ctypes.data, inc.ctypes.data))
        return first, unpack_bits(inc, W)

    def aggregate(self, specs: Sequence[tuple]):
        """-> list of (values ndarray, valid mask ndarray), host memory"""
        W = self.num_windows
        arr = make_specs(specs)
        outs = (OutCol * len(specs))()
        bufs = []
        for j in range(len(specs)):
            v = np.full(max(W, 1), -7, dtype=np.int64)          # poisoned: every slot must be written
            b = np.full((W + 7) // 8 + 1, 0xAA, dtype=np.uint8)
            bufs.append((v, b))
            outs[j].values, outs[j].validity = v.ctypes.data, b.ctypes.data
        self.ctx.check(lib().bowgpu_rolling_aggregate(self.h, arr, len(specs), outs, MEM_HOST))
        res = []
        for j, (v, b) in enumerate(bufs):
            vals = v[:W] if outs[j].dtype == INT64 else v[:W].view(np.float64)
            res.append((vals, unpack_bits(b, W)))
        return res

    def aggregate_device(self, specs_arr, nspecs: int, outs) -> None:
        """Asynchronous: results stay on the device (outs: OutCol array of device pointers)."""
        self.ctx.check(lib().bowgpu_rolling_aggregate(self.h, specs_arr, nspecs, outs, MEM_DEVICE))

    def interpolate_aggregate(self, ops: Sequence, specs: Sequence[tuple]):
        """Interpolate(ops).Aggregate(specs) without materialising the interpolated frame
        -> list of (values ndarray, valid mask ndarray), host memory"""
        W = self.num_windows
        codes = (C.c_int32 * len(ops))(*[INTERP[o] if isinstance(o, str) else int(o) for o in ops])
        arr = make_specs(specs)
        outs = (OutCol * len(specs))()
        bufs = []
        for j in range(len(specs)):
            v = np.full(max(W, 1), -7, dtype=np.int64)
            b = np.full((W + 7) // 8 + 1, 0xAA, dtype=np.uint8)
            bufs.append((v, b))
            outs[j].values, outs[j].validity = v.ctypes.data, b.ctypes.data
        self.ctx.check(lib().bowgpu_rolling_interpolate_aggregate(self.h, codes, len(ops), arr, len(specs), outs, MEM_HOST))
        res = []
        for j, (v, b) in enumerate(bufs):
            vals = v[:W] if outs[j].dtype == INT64 else v[:W].view(np.float64)
            res.append((vals, unpack_bits(b, W)))
        return res

    def interpolate_aggregate_device(self, ops: Sequence, specs_arr, nspecs: int, outs) -> None:
        codes = (C.c_int32 * len(ops))(*[INTERP[o] if isinstance(o, str) else int(o) for o in ops])
        self.ctx.check(lib().bowgpu_rolling_interpolate_aggregate(self.h, codes, len(ops), specs_arr, nspecs, outs, MEM_DEVICE))

    def interpolate(self, ops: Sequence) -> Frame:
        codes = (C.c_int32 * len(ops))(*[INTERP[o] if isinstance(o, str) else int(o) for o in ops])
        h, n_out = C.c_void_p(), C.c_int64()
        self.ctx.check(lib().bowgpu_rolling_interpolate(self.h, codes, len(ops), C.byref(h), C.byref(n_out)))
        return Frame(self.ctx, h)

    def close(self):
        if self.h:
            lib().bowgpu_rolling_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
