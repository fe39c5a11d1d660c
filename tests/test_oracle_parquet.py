"""The Parquet oracle (oracle/parquet_ref.py) pinned: the files the reference's own writer produced decode to what an
independent implementation (pyarrow) yields and to the round-1 fixture; further encodings on files written here."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from oracle import parquet_ref as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "parquet")


def arrow_cols(path):
    t = pq.read_table(path)
    out = {}
    for name in t.column_names:
        col = t.column(name).combine_chunks()
        if col.type not in (pa.int64(), pa.float64()):
            continue
        out[name] = (np.asarray(col.fill_null(0)), np.asarray(col.is_valid()))
    return out


def same(got, want):
    assert list(got) == list(want)
    for name in want:
        (gv, gm), (wv, wm) = got[name], want[name]
        assert gv.dtype == wv.dtype and np.array_equal(gm, wm), name
        assert np.array_equal(gv.view(np.int64), wv.view(np.int64)), name


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLD) if f.endswith(".parquet")))
def test_reference_written_files(name):
    path = os.path.join(GOLD, name)
    same(P.read_parquet(path), arrow_cols(path))


def test_config0_file_equals_the_round1_fixture():
    z = np.load(os.path.join(os.path.dirname(GOLD), "config1_bow1_100000.npz"))
    got = P.read_parquet(os.path.join(GOLD, "refwriter_bow1_100000_rows.parquet"))
    for name in ("Int64_ref", "Int64_no_nils_bow1", "Int64_bow1", "Float64_bow1"):
        gv, gm = got[name]
        assert np.array_equal(gv.view(np.int64), z[name].view(np.int64))
        assert np.array_equal(np.packbits(gm, bitorder="little"), z[name + "__valid"])


@pytest.mark.parametrize("opts", [dict(compression="NONE", use_dictionary=False), dict(compression="SNAPPY", use_dictionary=True),
                                  dict(compression="SNAPPY", use_dictionary=False, data_page_version="2.0", data_page_size=2048),
                                  dict(compression="NONE", use_dictionary=True, data_page_version="2.0", row_group_size=1500)])
def test_written_files(tmp_path, opts):
    rng = np.random.default_rng(4)
    n = 5000
    t = pa.table({"t": pa.array(np.cumsum(rng.integers(0, 9, n)).astype(np.int64)),
                  "v": pa.array(rng.normal(size=n), mask=rng.random(n) < 0.3),
                  "k": pa.array(rng.integers(0, 5, n).astype(np.int64) * 7, mask=rng.random(n) < 0.5),
                  "allnull": pa.array(np.zeros(n), mask=np.ones(n, dtype=bool)),
                  "s": pa.array(["x"] * n)})
    path = str(tmp_path / "f.parquet")
    pq.write_table(t, path, **opts)
    same(P.read_parquet(path), arrow_cols(path))


def test_snappy_overlapping_copy():
    # literal "ab" then a copy of length 10 at offset 2: "abababababab"
    blob = bytes([12, (2 - 1) << 2, ord("a"), ord("b"), ((10 - 1) << 2) | 2, 2, 0])
    assert P.snappy_decompress(blob) == b"abababababab"
