"""Minimal host-side `bow.Bow` for the interval-rolling path: a thin wrapper over an Arrow record
batch, exactly like the reference's `type bow struct{ arrow.Record }` (reference bow.go:96-98).

Only what the rolling path and its tests touch is mirrored (constructors from column / row based
data, schema accessors, zero-copy slicing, Equal).  Everything else of the reference's Bow surface
(joins, fills, parquet ...) is out of scope.  Only Int64 and Float64 columns can be
sent to the GPU; Boolean / String columns may exist in a Bow but cannot take part in a rolling call.
"""
from __future__ import annotations

import enum
from typing import Any, List, Optional, Sequence

import numpy as np
import pyarrow as pa


class BowError(Exception):
    """Carries the reference's error strings (Go `error` values)."""


class Type(enum.IntEnum):  # bowtypes.go:17-32
    Unknown = 0
    Float64 = 1
    Int64 = 2
    Boolean = 3
    String = 4
    InputDependent = 5
    IteratorDependent = 6

    def __str__(self) -> str:  # Type.String(), used by "%v" in the reference's error messages
        return {0: "undefined", 1: "float64", 2: "int64", 3: "bool", 4: "utf8"}.get(int(self), self.name)


Unknown, Float64, Int64, Boolean, String = Type.Unknown, Type.Float64, Type.Int64, Type.Boolean, Type.String
InputDependent, IteratorDependent = Type.InputDependent, Type.IteratorDependent

_TO_ARROW = {Type.Float64: pa.float64(), Type.Int64: pa.int64(), Type.Boolean: pa.bool_(), Type.String: pa.string()}


def _from_arrow(t: pa.DataType) -> Type:
    for k, v in _TO_ARROW.items():
        if v == t:
            return k
    return Type.Unknown


class Series:
    """bow.Series (bowseries.go:14-25): a named Arrow array."""

    def __init__(self, name: str, array: pa.Array):
        self.Name, self.Array = name, array


def NewSeries(name: str, typ: Type, data: Sequence, valid: Optional[Sequence[bool]] = None) -> Series:
    """bowseries.go:27-30"""
    mask = None if valid is None else ~np.asarray(valid, dtype=bool)
    arr = pa.array(list(data) if not isinstance(data, np.ndarray) else data, type=_TO_ARROW[typ],
                   mask=mask, from_pandas=False)
    return Series(name, arr)


def NewSeriesFromNumpy(name: str, values: np.ndarray, valid: Optional[np.ndarray] = None) -> Series:
    """Zero-copy wrap of a numpy int64 / float64 buffer (+ optional bool validity), the analogue of
    bow.NewSeriesFromBuffer (bowseries.go:32-34)."""
    values = np.ascontiguousarray(values)
    assert values.dtype in (np.int64, np.float64)
    bufs = [None, pa.py_buffer(values)]
    nulls = 0
    if valid is not None:
        valid = np.asarray(valid, dtype=bool)
        nulls = int(len(valid) - np.count_nonzero(valid))
        if nulls:
            bufs[0] = pa.py_buffer(np.packbits(valid, bitorder="little"))
    typ = pa.int64() if values.dtype == np.int64 else pa.float64()
    return Series(name, pa.Array.from_buffers(typ, len(values), bufs, null_count=nulls))


class Bow:
    def __init__(self, record: pa.RecordBatch):
        self.Record = record

    # -- schema -------------------------------------------------------------------------------------
    def NumRows(self) -> int:
        return self.Record.num_rows

    def NumCols(self) -> int:
        return self.Record.num_columns

    def ColumnName(self, i: int) -> str:
        return self.Record.schema.field(i).name

    def ColumnType(self, i: int) -> Type:
        return _from_arrow(self.Record.schema.field(i).type)

    def ColumnIndex(self, name: str) -> int:  # bowgetters.go:318-331
        idx = [i for i, f in enumerate(self.Record.schema) if f.name == name]
        if not idx:
            raise BowError(f"no column '{name}'")
        if len(idx) > 1:
            raise BowError(f"several columns '{name}'")
        return idx[0]

    def Column(self, i: int) -> pa.Array:
        return self.Record.column(i)

    def Metadata(self) -> dict:
        return dict(self.Record.schema.metadata or {})

    # -- cells / slices -------------------------------------------------------------------------------
    def GetValue(self, col: int, row: int) -> Any:  # bowgetters.go:46-63
        return self.Record.column(col)[row].as_py()

    def NewSlice(self, i: int, j: int) -> "Bow":  # bow.go:279-283 (zero copy)
        return Bow(self.Record.slice(i, j - i))

    def NewEmptySlice(self) -> "Bow":
        return Bow(self.Record.slice(0, 0))

    def ToColBased(self) -> List[list]:
        return [self.Record.column(i).to_pylist() for i in range(self.NumCols())]

    # -- comparator of the reference's tests (bow.go:227-275) --------------------------------------------
    def Equal(self, other: "Bow") -> bool:
        a, b = self.Record, other.Record
        if a.schema.names != b.schema.names or [f.type for f in a.schema] != [f.type for f in b.schema]:
            return False
        if (a.schema.metadata or {}) != (b.schema.metadata or {}):
            return False

        def rows(r):  # rows whose every cell is null are skipped (bow.go:257-262)
            cols = [r.column(i).to_pylist() for i in range(r.num_columns)]
            out = []
            for i in range(r.num_rows):
                row = {r.schema.field(j).name: cols[j][i] for j in range(r.num_columns) if cols[j][i] is not None}
                if row:
                    out.append(row)
            return out

        ra, rb = rows(a), rows(b)
        if len(ra) != len(rb):
            return False
        for x, y in zip(ra, rb):
            if x.keys() != y.keys():
                return False
            for k in x:
                u, v = x[k], y[k]
                if isinstance(u, float) and isinstance(v, float) and u != u and v != v:
                    continue
                if u != v or type(u) is not type(v):
                    return False
        return True

    def __str__(self) -> str:
        return self.Record.to_pandas().to_string() if self.NumRows() else f"<empty Bow {self.Record.schema.names}>"

    # -- whole-column fills on the GPU (bowfill.go:14-288); Go's (Bow, error) -> returned Bow / raised BowError ----
    def _gpu_fill(self, run) -> "Bow":
        from . import native as N
        from .rolling import _cols_from_bow
        from .runtime import default_ctx
        try:
            arr, keep = _cols_from_bow(self)
            frame = N.Frame.from_col_descs(default_ctx(), arr, self.NumCols(), N.MEM_HOST, keep)
            try:
                out = run(frame)
                if out is None:  # nothing to do: the reference returns the receiver itself
                    return self
                try:
                    cols = out.download()
                finally:
                    out.close()
            finally:
                frame.close()
        except N.BowGpuError as e:
            raise BowError(str(e).split(": ", 1)[-1])
        series = [NewSeriesFromNumpy(self.ColumnName(j), v, m) for j, (v, m) in enumerate(cols)]
        rec = pa.RecordBatch.from_arrays([s_.Array for s_ in series], names=[s_.Name for s_ in series])
        if self.Record.schema.metadata:   # NewBowWithMetadata(b.Metadata(), ...), bowfill.go:101,151,252
            rec = rec.replace_schema_metadata(self.Record.schema.metadata)
        return Bow(rec)

    def _check_fill_types(self, colIndices, numeric_only: bool):
        n = self.NumCols()
        for c in colIndices:  # selectCols, bowfill.go:266-288
            if c < 0 or c > n - 1:
                raise BowError(f"selectCols: colIndex '{c}' out of range")
        for c in (colIndices or range(n)):
            t = self.ColumnType(c)
            if t not in (Type.Int64, Type.Float64):
                if numeric_only:
                    raise BowError(f"column '{self.ColumnName(c)}' is of unsupported type '{t}'")
                raise BowError(f"column '{self.ColumnName(c)}': only Int64 / Float64 columns run on the GPU backend")
        for c in range(n):
            if self.ColumnType(c) not in (Type.Int64, Type.Float64):
                raise BowError(f"column '{self.ColumnName(c)}': only Int64 / Float64 columns run on the GPU backend")

    def DropNils(self, *colIndices: int) -> "Bow":  # bow.go:188-224
        self._check_fill_types(colIndices, False)
        return self._gpu_fill(lambda f: f.drop_nils(*colIndices))

    def SortByCol(self, colIndex: int) -> "Bow":  # bowsort.go:10-47 (equal keys keep their input order)
        if colIndex < 0 or colIndex > self.NumCols() - 1:
            raise BowError(f"no column {colIndex}")
        self._check_fill_types((), False)
        return self._gpu_fill(lambda f: f.sort_by_col(colIndex))

    def IsColSorted(self, colIndex: int) -> bool:  # bowassertion.go:15-81
        from . import native as N
        from .rolling import _cols_from_bow
        from .runtime import default_ctx
        if self.ColumnType(colIndex) not in (Type.Int64, Type.Float64):
            return False
        self._check_fill_types((), False)
        arr, keep = _cols_from_bow(self)
        frame = N.Frame.from_col_descs(default_ctx(), arr, self.NumCols(), N.MEM_HOST, keep)
        try:
            return frame.is_col_sorted(colIndex)
        finally:
            frame.close()

    def FillPrevious(self, *colIndices: int) -> "Bow":  # bowfill.go:160-164
        self._check_fill_types(colIndices, False)
        return self._gpu_fill(lambda f: f.fill("Previous", *colIndices))

    def FillNext(self, *colIndices: int) -> "Bow":  # bowfill.go:154-158
        self._check_fill_types(colIndices, False)
        return self._gpu_fill(lambda f: f.fill("Next", *colIndices))

    def FillMean(self, *colIndices: int) -> "Bow":  # bowfill.go:104-152
        self._check_fill_types(colIndices, True)
        return self._gpu_fill(lambda f: f.fill("Mean", *colIndices))

    def FillLinear(self, refColIndex: int, toFillColIndex: int) -> "Bow":  # bowfill.go:14-102
        n = self.NumCols()
        if refColIndex < 0 or refColIndex > n - 1:
            raise BowError("refColIndex is out of range")
        if toFillColIndex < 0 or toFillColIndex > n - 1:
            raise BowError("toFillColIndex is out of range")
        if refColIndex == toFillColIndex:
            raise BowError("refColIndex and toFillColIndex are equal")
        if self.ColumnType(refColIndex) not in (Type.Int64, Type.Float64):
            raise BowError(f"refColIndex '{refColIndex}' is of type '{self.ColumnType(refColIndex)}'")
        if self.ColumnType(toFillColIndex) not in (Type.Int64, Type.Float64):
            raise BowError(f"toFillColIndex '{toFillColIndex}' is of unsupported type '{self.ColumnType(toFillColIndex)}'")
        self._check_fill_types((), False)
        return self._gpu_fill(lambda f: f.fill_linear(refColIndex, toFillColIndex))


def NewBow(*series: Series) -> Bow:
    """bow.go:109-116: fresh schema, no metadata, every field nullable (bowrecord.go:36-40)"""
    names = [s.Name for s in series]
    arrays = [s.Array for s in series]
    n = {len(a) for a in arrays}
    if len(n) > 1:
        raise BowError("bow.NewBow: series have different lengths")
    return Bow(pa.RecordBatch.from_arrays(arrays, names=names))


def NewBowFromColBasedInterfaces(colNames: Sequence[str], colTypes: Sequence[Type], colBasedData: Sequence[Sequence]) -> Bow:
    """bow.go:125-148"""
    if len(colNames) != len(colTypes) or len(colNames) != len(colBasedData):
        raise BowError("bow.NewBowFromColBasedInterfaces: mismatch between colNames, colTypes and colBasedData lengths")
    series = []
    for name, typ, data in zip(colNames, colTypes, colBasedData):
        conv = []
        for v in data:
            if v is None:
                conv.append(None)
            elif typ == Type.Float64:
                conv.append(float(v))
            elif typ == Type.Int64:
                conv.append(int(v))
            else:
                conv.append(v)
        series.append(Series(name, pa.array(conv, type=_TO_ARROW[typ])))
    return NewBow(*series)


def NewBowFromRowBasedInterfaces(colNames: Sequence[str], colTypes: Sequence[Type], rows: Sequence[Sequence]) -> Bow:
    """bow.go:150-170"""
    cols = [[r[j] for r in rows] for j in range(len(colNames))]
    return NewBowFromColBasedInterfaces(colNames, colTypes, cols)


def NewBowFromParquet(path: str, verbose: bool = False, colNames: Optional[Sequence[str]] = None) -> Bow:
    """bowparquet.go:44-155: loads a parquet file into a new Bow.  The footer and the page headers are walked on the
    host; the column data is decompressed and decoded on the GPU (csrc/parquet.cu) and downloaded.  The reference maps
    BOOLEAN / INT64 / DOUBLE / BYTE_ARRAY leaves (bowparquet.go:20-25); the GPU backend has Int64 and Float64 only, so a
    file holding other columns needs `colNames` (the numeric columns to load) — there is no CPU decode to fall back to.
    Key-value metadata of the footer (bowparquet.go:118-134) is not carried over."""
    from . import native as N
    from .runtime import default_ctx
    try:
        with N.ParquetFile(path) as pf:
            if colNames is None:
                idx = list(range(len(pf.names)))
            else:
                idx = []
                for name in colNames:
                    if name not in pf.names:
                        raise BowError(f"bow.NewBowFromParquet: no column '{name}' in {path}")
                    idx.append(pf.names.index(name))
            for j in idx:
                if not pf.dtypes[j]:
                    raise BowError(f"bow.NewBowFromParquet: column '{pf.names[j]}' (parquet type {pf.physical[j]}): only "
                                   "INT64 / DOUBLE columns run on the GPU backend; name the columns to load in colNames")
            frame = pf.read(default_ctx(), idx)
            try:
                cols = frame.download()
            finally:
                frame.close()
            names = [pf.names[j] for j in idx]
            nrows = pf.num_rows
    except N.BowGpuError as e:
        raise BowError("bow.NewBowFromParquet: " + str(e).split(": ", 1)[-1])
    series = [NewSeriesFromNumpy(name, v, m) for name, (v, m) in zip(names, cols)]
    b = NewBow(*series)
    if verbose:
        print(f"bow.NewBowFromParquet: {path} successfully read: {nrows} rows\n{b.Record.schema}")
    return b
