"""TEST INFRASTRUCTURE ONLY — CPU restatement of the Parquet read path behind bow.NewBowFromParquet
(/root/reference/bowparquet.go:44-155).

The reference delegates the decoding to github.com/xitongsys/parquet-go v1.6.2 (go.mod:13), which is NOT vendored
under /root/reference; what is restated here is therefore the published Apache Parquet format that library implements —
Thrift compact protocol (FileMetaData, PageHeader), Snappy block format, RLE / bit-packed hybrid levels, PLAIN and
dictionary encodings — for the column kinds the GPU path takes (flat schema, INT64 / DOUBLE leaves), followed by the
reference's own step: values into a bow.Buffer, nil where the definition level is 0 (bowparquet.go:97-107,
`buf.SetOrDrop(i, v)`: value 0 in null slots).

Pinned (tests/test_oracle_parquet.py) on the files the reference's own writer produced (tests/golden/parquet, copied from
/root/reference/benchmarks) against an independent implementation (pyarrow) and against the round-1 fixture
tests/golden/config1_bow1_100000.npz.  Plain Python / numpy; small files only."""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np


# ---- Thrift compact protocol: a struct becomes {field id: value}, lists become Python lists -----------------------------
class _T:
    def __init__(self, b: bytes, pos: int = 0):
        self.b, self.p = b, pos

    def varint(self) -> int:
        v = s = 0
        while True:
            x = self.b[self.p]
            self.p += 1
            v |= (x & 0x7F) << s
            if not x & 0x80:
                return v
            s += 7

    def zigzag(self) -> int:
        v = self.varint()
        return (v >> 1) ^ -(v & 1)

    def value(self, t: int):
        if t in (1, 2):
            return t == 1
        if t == 3:
            self.p += 1
            return self.b[self.p - 1]
        if t in (4, 5, 6):
            return self.zigzag()
        if t == 7:
            self.p += 8
            return struct.unpack_from("<d", self.b, self.p - 8)[0]
        if t == 8:
            n = self.varint()
            self.p += n
            return bytes(self.b[self.p - n:self.p])
        if t in (9, 10):
            h = self.b[self.p]
            self.p += 1
            n, et = h >> 4, h & 0x0F
            if n == 15:
                n = self.varint()
            out = []
            for _ in range(n):
                if et in (1, 2):
                    out.append(self.b[self.p] == 1)
                    self.p += 1
                else:
                    out.append(self.value(et))
            return out
        if t == 12:
            return self.struct()
        raise ValueError(f"thrift type {t}")

    def struct(self) -> Dict[int, object]:
        out, fid = {}, 0
        while True:
            h = self.b[self.p]
            self.p += 1
            if h == 0:
                return out
            t, d = h & 0x0F, h >> 4
            fid = fid + d if d else self.zigzag()
            out[fid] = self.value(t)


# ---- Snappy block format -----------------------------------------------------------------------------------------------------
def snappy_decompress(src: bytes) -> bytes:
    t = _T(src)
    n = t.varint()
    p, out = t.p, bytearray()
    while p < len(src):
        tag = src[p]
        p += 1
        kind = tag & 3
        if kind == 0:
            ln = (tag >> 2) + 1
            if ln > 60:
                nb = ln - 60
                ln = int.from_bytes(src[p:p + nb], "little") + 1
                p += nb
            out += src[p:p + ln]
            p += ln
            continue
        if kind == 1:
            ln, off = 4 + ((tag >> 2) & 7), ((tag >> 5) << 8) | src[p]
            p += 1
        elif kind == 2:
            ln, off = (tag >> 2) + 1, int.from_bytes(src[p:p + 2], "little")
            p += 2
        else:
            ln, off = (tag >> 2) + 1, int.from_bytes(src[p:p + 4], "little")
            p += 4
        if off <= 0 or off > len(out):
            raise ValueError("snappy: bad offset")
        for _ in range(ln):            # byte by byte: overlapping copies repeat their period
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy: length mismatch")
    return bytes(out)


# ---- RLE / bit-packed hybrid ------------------------------------------------------------------------------------------------
def hybrid_decode(buf: bytes, bw: int, count: int) -> np.ndarray:
    out = np.zeros(count, dtype=np.int64)
    t, n = _T(buf), 0
    while n < count and t.p < len(buf):
        h = t.varint()
        if h & 1:
            groups = h >> 1
            nbytes = groups * bw
            bits = np.unpackbits(np.frombuffer(buf[t.p:t.p + nbytes], dtype=np.uint8), bitorder="little")
            t.p += nbytes
            nv = min(groups * 8, count - n, len(bits) // bw if bw else groups * 8)
            if bw:
                vals = bits[:nv * bw].reshape(nv, bw).astype(np.int64) @ (1 << np.arange(bw, dtype=np.int64))
            else:
                vals = np.zeros(nv, dtype=np.int64)
            out[n:n + nv] = vals
            n += nv
        else:
            nb = (bw + 7) // 8
            v = int.from_bytes(buf[t.p:t.p + nb], "little")
            t.p += nb
            nv = min(h >> 1, count - n)
            out[n:n + nv] = v
            n += nv
    return out[:n]


# ---- the file ------------------------------------------------------------------------------------------------------------------
def read_parquet(path: str, columns: Optional[List[str]] = None) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """-> {column name: (values with 0 in null slots, bool validity)} for the INT64 / DOUBLE leaves (all, or `columns`)"""
    b = open(path, "rb").read()
    if b[:4] != b"PAR1" or b[-4:] != b"PAR1":
        raise ValueError("not a parquet file")
    flen = int.from_bytes(b[-8:-4], "little")
    meta = _T(b, len(b) - 8 - flen).struct()
    schema = meta[2]
    leaves = schema[1:]
    assert schema[0].get(5, 0) == len(leaves), "flat schemas only"
    out = {}
    for ci, se in enumerate(leaves):
        name, ptype, optional = se[4].decode(), se.get(1), se.get(3, 0) == 1
        if ptype not in (2, 5) or (columns is not None and name not in columns):
            continue
        dt = np.dtype("<i8") if ptype == 2 else np.dtype("<f8")
        vals_all, valid_all = [], []
        for rg in meta[4]:
            cm = rg[1][ci][3]
            codec, nvals, total = cm[4], cm[5], cm[7]
            if nvals == 0:
                continue
            pos = cm[9]
            if cm.get(11) and 0 < cm[11] < pos:
                pos = cm[11]
            end, seen, dictionary = pos + total, 0, None
            while seen < nvals and pos < end:
                t = _T(b, pos)
                ph = t.struct()
                body = b[t.p:t.p + ph[3]]
                pos = t.p + ph[3]

                def inflate(x):
                    return snappy_decompress(x) if codec == 1 else x
                if ph[1] == 2:                       # dictionary page (PLAIN values)
                    dictionary = np.frombuffer(inflate(body), dtype=dt, count=ph[7][1])
                    continue
                if ph[1] == 0:                       # data page v1: [u32 length + hybrid levels] values
                    h = ph[5]
                    n, enc = h[1], h[2]
                    raw = inflate(body)
                    if optional:
                        ln = int.from_bytes(raw[:4], "little")
                        levels = hybrid_decode(raw[4:4 + ln], 1, n)
                        raw = raw[4 + ln:]
                    else:
                        levels = np.ones(n, dtype=np.int64)
                elif ph[1] == 3:                     # data page v2: levels in front, never compressed
                    h = ph[8]
                    n, enc, dl, rl = h[1], h[4], h.get(5, 0), h.get(6, 0)
                    levels = hybrid_decode(body[rl:rl + dl], 1, n) if optional else np.ones(n, dtype=np.int64)
                    raw = body[rl + dl:]
                    if h.get(7, True):
                        raw = inflate(raw)
                else:
                    continue
                valid = levels.astype(bool)
                nn = int(valid.sum())
                if enc in (2, 8):                    # dictionary indices: [bit width] hybrid
                    idx = hybrid_decode(raw[1:], raw[0], nn)
                    present = dictionary[idx]
                else:
                    present = np.frombuffer(raw, dtype=dt, count=nn)
                vals = np.zeros(n, dtype=dt)        # bow.NewBuffer: zero in null slots
                vals[valid] = present
                vals_all.append(vals)
                valid_all.append(valid)
                seen += n
        out[name] = (np.concatenate(vals_all) if vals_all else np.zeros(0, dtype=dt),
                     np.concatenate(valid_all) if valid_all else np.zeros(0, dtype=bool))
    return out
