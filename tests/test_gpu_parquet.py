"""Parquet ingest on the GPU (csrc/parquet.cu) against the oracle (oracle/parquet_ref.py, files up to 120 000 rows) and an
independent decoder (pyarrow, every file) — bit-exact values, validity and
null-slot zeros (bow.NewBuffer layout) — on the files the reference's own writer produced (tests/golden/parquet: SNAPPY,
PLAIN, data pages v1) and on files written here over the other encodings the reader supports.
Reference: bowparquet.go:44-155 (NewBowFromParquet) and its test bowparquet_test.go."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from bow_b200 import native as N

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "parquet")


@pytest.fixture(scope="module")
def ctx():
    c = N.Ctx(0)
    yield c
    c.close()


def arrow_cols(path, names):
    t = pq.read_table(path, columns=names)
    out = []
    for name in names:
        col = t.column(name).combine_chunks()
        valid = np.asarray(col.is_valid())
        vals = np.asarray(col.fill_null(0))
        out.append((vals, valid))
    return out


def check(ctx, path, names=None):
    with N.ParquetFile(path) as pf:
        idx = [j for j, d in enumerate(pf.dtypes) if d] if names is None else [pf.names.index(n) for n in names]
        names = [pf.names[j] for j in idx]
        fr = pf.read(ctx, idx)
        try:
            got = fr.download()
            nulls = [fr.null_count(j) for j in range(len(idx))]
            assert fr.num_rows == pf.num_rows
        finally:
            fr.close()
    want = arrow_cols(path, names)
    for name, (gv, gm), (wv, wm), nn in zip(names, got, want, nulls):
        assert gv.dtype == wv.dtype, name
        assert np.array_equal(gm, wm), f"{name}: validity differs at rows {np.nonzero(gm != wm)[0][:8]}"
        assert np.array_equal(gv.view(np.int64), wv.view(np.int64)), \
            f"{name}: values differ at rows {np.nonzero(gv.view(np.int64) != wv.view(np.int64))[0][:8]}"
        assert nn == int((~wm).sum()), name
    if len(want) and len(want[0][0]) <= 120_000:   # ... and against the repo's own restatement of the format (oracle/)
        from oracle import parquet_ref as P
        ora = P.read_parquet(path, names)
        for name, (gv, gm) in zip(names, got):
            ov, om = ora[name]
            assert np.array_equal(gm, om) and np.array_equal(gv.view(np.int64), ov.view(np.int64)), f"{name}: differs from the oracle"
    return got


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLD) if f.endswith(".parquet")))
def test_reference_written_files(ctx, name):
    check(ctx, os.path.join(GOLD, name))


def test_config0_fixture_matches_the_npz(ctx):
    """the same file as decoded when the round-1 fixture was made"""
    z = np.load(os.path.join(os.path.dirname(GOLD), "config1_bow1_100000.npz"))
    names = ["Int64_ref", "Int64_no_nils_bow1", "Int64_bow1", "Float64_bow1"]
    got = check(ctx, os.path.join(GOLD, "refwriter_bow1_100000_rows.parquet"), names)
    for name, (gv, gm) in zip(names, got):
        assert np.array_equal(gv.view(np.int64), z[name].view(np.int64))
        assert np.array_equal(np.packbits(gm, bitorder="little"), z[name + "__valid"])


def make_table(n, seed, null_frac, runs=False, nullable=True):
    rng = np.random.default_rng(seed)
    t = np.cumsum(rng.integers(0, 1000, size=n)).astype(np.int64) + 1_700_000_000_000_000_000
    v = rng.normal(size=n)
    lowcard = rng.integers(0, 17, size=n).astype(np.int64) * 1_000_003   # dictionary friendly
    const = np.full(n, 42.5)
    special = rng.choice(np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-310, 1.5]), size=n)

    def mask():
        if not nullable or null_frac == 0:
            return None
        if null_frac >= 1:
            return np.ones(n, dtype=bool)
        if runs:   # long runs of nulls / non-nulls: RLE runs in the definition levels
            m = np.zeros(n, dtype=bool)
            pos = 0
            while pos < n:
                ln = int(rng.integers(1, 5000))
                if rng.random() < null_frac:
                    m[pos:pos + ln] = True
                pos += ln
            return m
        return rng.random(n) < null_frac
    cols = {"t": (t, None), "v": (v, mask()), "lowcard": (lowcard, mask()), "const": (const, mask()), "special": (special, mask())}
    fields, arrays = [], []
    for name, (vals, m) in cols.items():
        arrays.append(pa.array(vals, mask=m))
        fields.append(pa.field(name, arrays[-1].type, nullable=nullable))
    return pa.Table.from_arrays(arrays, schema=pa.schema(fields))


WRITE_OPTS = [
    dict(compression="NONE", use_dictionary=False),
    dict(compression="SNAPPY", use_dictionary=False),
    dict(compression="SNAPPY", use_dictionary=True),
    dict(compression="NONE", use_dictionary=True, data_page_version="2.0"),
    dict(compression="SNAPPY", use_dictionary=False, data_page_version="2.0"),
    dict(compression="SNAPPY", use_dictionary=True, data_page_version="2.0", data_page_size=4096),
    dict(compression="SNAPPY", use_dictionary=False, data_page_size=1024),
]


@pytest.mark.parametrize("opts", range(len(WRITE_OPTS)))
@pytest.mark.parametrize("n,null_frac,runs", [(0, 0.3, False), (1, 0.0, False), (31, 0.5, False), (33, 0.5, False), (1000, 1.0, False),
                                              (20_000, 0.3, False), (20_000, 0.3, True), (300_000, 0.1, False), (300_000, 0.0, False)])
def test_written_files(ctx, tmp_path, opts, n, null_frac, runs):
    t = make_table(n, seed=n + opts, null_frac=null_frac, runs=runs)
    path = str(tmp_path / "f.parquet")
    pq.write_table(t, path, **WRITE_OPTS[opts])
    check(ctx, path)


def test_required_columns_and_row_groups(ctx, tmp_path):
    t = make_table(100_000, seed=5, null_frac=0.0, nullable=False)
    path = str(tmp_path / "r.parquet")
    pq.write_table(t, path, compression="SNAPPY", use_dictionary=False, row_group_size=7_777)
    got = check(ctx, path)
    assert all(m.all() for _, m in got)
    t = make_table(100_000, seed=6, null_frac=0.4)
    pq.write_table(t, path, compression="SNAPPY", use_dictionary=True, row_group_size=33_333, data_page_size=2048)
    check(ctx, path)


def test_column_selection_and_types(ctx, tmp_path):
    t = pa.table({"s": pa.array(["a", "b", None]), "t": pa.array([1, 2, 3], type=pa.int64()), "b": pa.array([True, None, False]),
                  "v": pa.array([1.5, None, 2.5])})
    path = str(tmp_path / "m.parquet")
    pq.write_table(t, path)
    got = check(ctx, path, ["v", "t"])          # any order
    assert got[0][0].dtype == np.float64 and got[1][0].dtype == np.int64
    with N.ParquetFile(path) as pf:
        with pytest.raises(N.BowGpuError) as e:
            pf.read(ctx, [0])
        assert e.value.status == "ETYPE" and "s" in str(e.value)
        with pytest.raises(N.BowGpuError) as e:
            pf.read(ctx, [9])
        assert e.value.status == "EINVAL"


def test_unsupported_codec_and_corrupt_pages(ctx, tmp_path):
    t = make_table(50_000, seed=9, null_frac=0.2)
    path = str(tmp_path / "z.parquet")
    pq.write_table(t, path, compression="GZIP")
    with N.ParquetFile(path) as pf:
        with pytest.raises(N.BowGpuError) as e:
            pf.read(ctx)
        assert e.value.status == "EUNSUPPORTED"
    # garbage in the middle of the SNAPPY pages: an error (or, if the bytes still parse, some frame) — never a crash
    pq.write_table(t, path, compression="SNAPPY", use_dictionary=False)
    raw = bytearray(open(path, "rb").read())
    rng = np.random.default_rng(1)
    for pos in rng.integers(1000, len(raw) // 2, size=200):
        raw[pos] = int(rng.integers(0, 256))
    bad = str(tmp_path / "bad.parquet")
    open(bad, "wb").write(bytes(raw))
    try:
        with N.ParquetFile(bad) as pf:
            pf.read(ctx).close()
    except N.BowGpuError as e:
        assert e.status in ("EIO", "EUNSUPPORTED")
    check(ctx, path)   # the context is still usable


def test_mirror_NewBowFromParquet(ctx, tmp_path):
    """the Python mirror of bow.NewBowFromParquet feeding the rolling path (bowparquet_test.go reads what it wrote)"""
    from bow_b200 import bow as B
    from bow_b200 import runtime
    runtime.set_default_ctx(ctx)
    try:
        path = os.path.join(GOLD, "refwriter_bow1_1000_rows.parquet")
        names = ["Int64_ref", "Int64_bow1", "Float64_bow1"]
        b = B.NewBowFromParquet(path, colNames=names)
        want = pq.read_table(path, columns=names)
        assert b.NumRows() == 1000 and [b.ColumnName(j) for j in range(3)] == names
        for j, name in enumerate(names):
            assert b.Column(j).to_pylist() == want.column(name).to_pylist()
        with pytest.raises(B.BowError):
            B.NewBowFromParquet(path)   # Boolean / String columns: no GPU type
        with pytest.raises(B.BowError):
            B.NewBowFromParquet(str(tmp_path / "nope.parquet"))
    finally:
        runtime.set_default_ctx(None)


@pytest.mark.parametrize("interval", [10, 1000])
def test_config0_from_the_parquet_file_stays_on_the_device(ctx, interval):
    """BASELINE.json configs[0] as the reference runs it (benchmarks/bow1-100000-rows.parquet: NewBowFromParquet ->
    IntervalRolling(Int64_ref, 10) -> ArithmeticMean / Count / Min / Max), file -> device frame -> Aggregate without a
    host round trip, against the oracle on the independently decoded columns."""
    from oracle import refc as R
    path = os.path.join(GOLD, "refwriter_bow1_100000_rows.parquet")
    names = ["Int64_ref", "Int64_bow1", "Float64_bow1"]
    with N.ParquetFile(path) as pf:
        fr = pf.read(ctx, [pf.names.index(n) for n in names])
    specs = [("WindowStart", 0)] + [(a, c) for c in (1, 2) for a in ("ArithmeticMean", "Count", "Min", "Max")]
    r = N.Rolling(fr, 0, interval)
    got = r.aggregate(specs)
    r.close()
    fr.close()
    cols = [(v, None if m.all() else m) for v, m in arrow_cols(path, names)]
    want = R.RefRolling(R.Frame(cols), 0, interval).aggregate(specs)
    for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
        assert np.array_equal(gm, wm), sp
        if sp[0] == "ArithmeticMean":
            assert np.allclose(gv[gm], wv[wm], rtol=1e-12, atol=0), sp
        else:
            assert np.array_equal(gv[gm].view(np.int64), wv[wm].view(np.int64)), sp
