"""Generates tests/golden/config1_bow1_100000.npz from the reference's benchmark fixture
/root/reference/benchmarks/bow1-100000-rows.parquet (BASELINE.json configs[0]).  Only the numeric
columns the rolling path can consume are kept (Int64_ref = the sorted interval column, two int64 and
one float64 value column with ~30 % nulls); values and validity are stored as plain numpy arrays.
Run in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/make_config1_fixture.py
"""
import os

import numpy as np
import pyarrow.parquet as pq

SRC = "/root/reference/benchmarks/bow1-100000-rows.parquet"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1_bow1_100000.npz")
COLS = ["Int64_ref", "Int64_no_nils_bow1", "Int64_bow1", "Float64_bow1"]

if __name__ == "__main__":
    t = pq.read_table(SRC)
    out = {}
    for c in COLS:
        col = t.column(c).combine_chunks()
        valid = np.asarray(col.is_valid())
        vals = np.asarray(col.fill_null(0))
        out[c] = vals
        out[c + "__valid"] = np.packbits(valid, bitorder="little")
    np.savez_compressed(DST, **out)
    print(DST, os.path.getsize(DST), "bytes")
