"""Parquet ingest, host side (no GPU): the footer walk of bowgpu_parquet_open against pyarrow's metadata, on the files the
reference's own writer produced (tests/golden/parquet, copied from /root/reference/benchmarks) and on files written here."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from bow_b200 import native as N

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "parquet")
PHYS = {"BOOLEAN": 0, "INT32": 1, "INT64": 2, "INT96": 3, "FLOAT": 4, "DOUBLE": 5, "BYTE_ARRAY": 6, "FIXED_LEN_BYTE_ARRAY": 7}


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLD) if f.endswith(".parquet")))
def test_footer_of_reference_files(name):
    path = os.path.join(GOLD, name)
    meta = pq.ParquetFile(path).metadata
    with N.ParquetFile(path) as pf:
        assert pf.num_rows == meta.num_rows
        assert pf.names == [meta.schema.column(j).name for j in range(meta.num_columns)]
        assert pf.physical == [PHYS[meta.schema.column(j).physical_type] for j in range(meta.num_columns)]
        want = [N.INT64 if t == 2 else N.FLOAT64 if t == 5 else 0 for t in pf.physical]
        assert pf.dtypes == want  # mapParquetToBowTypes (bowparquet.go:20-25) restricted to the GPU types


def test_footer_of_written_file(tmp_path):
    t = pa.table({"t": pa.array(np.arange(1000, dtype=np.int64)), "v": pa.array(np.linspace(0, 1, 1000)),
                  "i32": pa.array(np.arange(1000, dtype=np.int32)), "s": pa.array(["x"] * 1000)})
    path = str(tmp_path / "a.parquet")
    pq.write_table(t, path, row_group_size=300)
    with N.ParquetFile(path) as pf:
        assert pf.num_rows == 1000 and pf.names == ["t", "v", "i32", "s"]
        assert pf.dtypes == [N.INT64, N.FLOAT64, 0, 0]


def test_open_errors(tmp_path):
    with pytest.raises(N.BowGpuError) as e:
        N.ParquetFile(str(tmp_path / "missing.parquet"))
    assert e.value.status == "EIO" and "open" in str(e.value)
    bad = tmp_path / "bad.parquet"
    bad.write_bytes(b"PAR1" + b"\x00" * 64 + b"XXXX")
    with pytest.raises(N.BowGpuError) as e:
        N.ParquetFile(str(bad))
    assert e.value.status == "EIO"
    short = tmp_path / "short.parquet"
    short.write_bytes(b"PAR1")
    with pytest.raises(N.BowGpuError):
        N.ParquetFile(str(short))
    # a footer length that points outside the file
    good = tmp_path / "g.parquet"
    pq.write_table(pa.table({"t": pa.array([1, 2, 3], type=pa.int64())}), str(good))
    raw = bytearray(good.read_bytes())
    raw[-8:-4] = (len(raw) * 2).to_bytes(4, "little")
    trunc = tmp_path / "t.parquet"
    trunc.write_bytes(bytes(raw))
    with pytest.raises(N.BowGpuError):
        N.ParquetFile(str(trunc))
    # garbage in place of the Thrift footer
    raw = bytearray(good.read_bytes())
    flen = int.from_bytes(raw[-8:-4], "little")
    raw[-8 - flen:-8] = b"\xff" * flen
    garb = tmp_path / "garb.parquet"
    garb.write_bytes(bytes(raw))
    with pytest.raises(N.BowGpuError):
        N.ParquetFile(str(garb))


def test_nested_schema_is_refused(tmp_path):
    t = pa.table({"l": pa.array([[1, 2], [3]], type=pa.list_(pa.int64()))})
    path = str(tmp_path / "n.parquet")
    pq.write_table(t, path)
    with pytest.raises(N.BowGpuError) as e:
        N.ParquetFile(path)
    assert e.value.status == "EUNSUPPORTED"


@pytest.mark.parametrize("opts", [dict(compression="SNAPPY", use_dictionary=False), dict(compression="NONE", use_dictionary=True),
                                  dict(compression="SNAPPY", use_dictionary=True, data_page_version="2.0", data_page_size=4096,
                                       row_group_size=7000)])
def test_page_walk_sizes_follow_the_metadata(tmp_path, opts):
    """the host-side page walk (bowgpu_parquet_plan: what sizes the device buffers) against pyarrow's column-chunk metadata"""
    rng = np.random.default_rng(2)
    n = 20_000
    t = pa.table({"t": pa.array(np.arange(n, dtype=np.int64) * 1000),
                  "v": pa.array(rng.normal(size=n), mask=rng.random(n) < 0.25),
                  "k": pa.array(rng.integers(0, 9, n).astype(np.int64), mask=rng.random(n) < 0.5),
                  "s": pa.array(["x"] * n)})
    path = str(tmp_path / "p.parquet")
    pq.write_table(t, path, **opts)
    meta = pq.ParquetFile(path).metadata
    comp = uncomp = 0
    for g in range(meta.num_row_groups):
        for c in range(3):   # t, v, k
            col = meta.row_group(g).column(c)
            comp += col.total_compressed_size
            uncomp += col.total_uncompressed_size
    with N.ParquetFile(path) as pf:
        p = pf.plan([0, 1, 2])
        chunks = 3 * meta.num_row_groups
        assert comp <= p["image_bytes"] <= comp + 32 * chunks + 64           # every chunk as stored, 16-byte padded
        if opts["compression"] == "SNAPPY":                                    # page bodies once uncompressed (headers excluded)
            assert 0 < p["scratch_bytes"] <= uncomp + 32 * p["pages"] + 64
        else:
            assert p["scratch_bytes"] == 64
        if opts["use_dictionary"]:
            assert 0 < p["aux_entries"] <= 3 * n                              # one index per value of a dictionary-encoded page
        else:
            assert p["aux_entries"] == 0
        assert p["pages"] >= chunks
        one = pf.plan([1])
        assert one["pages"] < p["pages"] and one["image_bytes"] < p["image_bytes"]
        with pytest.raises(N.BowGpuError) as e:
            pf.plan([3])
        assert e.value.status == "ETYPE"
