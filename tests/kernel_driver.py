#!/usr/bin/env python
"""Drives every kernel family of libbowgpu.so once on small inputs, checked against the oracle — the workload of
scripts/sanitize.sh (compute-sanitizer memcheck / initcheck / racecheck / synccheck).  numpy + ctypes only (no torch:
its allocator and kernels would drown the report)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bow_b200 import native as N  # noqa: E402
from oracle import refc as R  # noqa: E402
from tests import helpers as H  # noqa: E402

ALL = ["WindowStart", "Count", "Sum", "ArithmeticMean", "Min", "Max", "First", "Last", "IntegralStep", "IntegralTrapezoid",
       "WeightedAverageStep", "WeightedAverageLinear"]


TOL_OPS = {"Sum", "ArithmeticMean", "IntegralStep", "IntegralTrapezoid", "WeightedAverageStep", "WeightedAverageLinear"}


def close(a, b, name, spec=None, scales=None):
    """bit-exact, or - float reductions - within the stated class |gpu - ref| <= 1e-12 * max(|ref|, sum|terms|)"""
    (gv, gm), (wv, wm) = a, b
    assert gv.dtype == wv.dtype and np.array_equal(gm, wm), name
    if spec is not None and spec[0] in TOL_OPS and not (spec[0] == "Sum" and spec[1] == 2):  # (Sum of the int64 column: exact)
        H.assert_in_tolerance_class(gv[gm], wv[wm], scales[spec[1]][spec[0]][gm], name)
    else:
        assert np.array_equal(gv[gm].view(np.int64), wv[wm].view(np.int64)), name


def main():
    n = int(os.environ.get("SAN_ROWS", "120000"))
    rng = np.random.default_rng(0)
    t = np.cumsum(rng.integers(0, 7, size=n)).astype(np.int64)
    v = rng.normal(size=n)
    m = rng.random(n) > 0.2
    iv = rng.integers(-1000, 1000, size=n).astype(np.int64)
    cols = [(t, None), (v, m), (iv, None)]
    ctx = N.Ctx(0)
    fr = N.Frame.from_numpy(ctx, cols)
    specs = [("WindowStart", 0)] + [(a, c) for c in (1, 2) for a in ALL[1:]]
    for interval, offset in ((50, 3), (5000, 0)):
        r = N.Rolling(fr, 0, interval, offset=offset)
        got = r.aggregate(specs)
        ref = R.RefRolling(R.Frame(cols), 0, interval, offset=offset)
        want = ref.aggregate(specs)
        # (the inclusive aggregations force inclusive windows for the whole call, aggregation.go:183-185)
        scales = {c: H.term_scales(cols, c, interval, offset=offset, inclusive=True) for c in (1, 2)}
        for s, g, w in zip(specs, got, want):
            close(g, w, f"aggregate {s} interval {interval}", s, scales)
        ops = ["WindowStart", "Linear", "StepPrevious"]
        got = r.interpolate_aggregate(ops, specs)                 # fused
        fi = r.interpolate(ops)                                   # materialised
        ri = N.Rolling(fi, 0, interval, offset=offset)
        got2 = ri.aggregate(specs)
        for s, g, w in zip(specs, got, got2):
            close(g, w, f"fused vs materialised {s}", s, scales)
        ri.close(); fi.close(); r.close()
    fr.aggregate_whole(0, specs)
    for meth in ("Previous", "Next", "Mean"):
        fr.fill(meth, 1).close()
    fr.fill_linear(0, 1).close()
    fr.drop_nils().close()
    assert fr.is_col_sorted(0)
    perm = rng.permutation(n)
    fs = N.Frame.from_numpy(ctx, [(t[perm], None), (v[perm], m[perm])])
    so = fs.sort_by_col(0)
    assert so.is_col_sorted(0)
    so.close(); fs.close(); fr.close()
    # pipelined one-shot host calls (worker threads, chunks of 16k rows)
    hs = [("WindowStart", 0), ("ArithmeticMean", 1), ("Min", 1), ("Max", 1), ("Count", 1), ("IntegralTrapezoid", 1)]
    got = N.aggregate_host(ctx, cols, 0, 50, hs, offset=3, chunk_rows=16384)
    want = R.RefRolling(R.Frame(cols), 0, 50, offset=3).aggregate(hs)
    scales = {1: H.term_scales(cols, 1, 50, offset=3, inclusive=True)}
    for s, g, w in zip(hs, got, want):
        close(g, w, f"aggregate_host {s}", s, scales)
    N.interpolate_aggregate_host(ctx, cols, 0, 50, ["WindowStart", "Linear", "StepPrevious"], hs, offset=3, chunk_rows=16384)
    # parquet ingest: the reference-written fixtures (Snappy, optional columns)
    gold = os.path.join(ROOT, "tests", "golden", "parquet")
    for f in sorted(os.listdir(gold)):
        if f.endswith(".parquet"):
            with N.ParquetFile(os.path.join(gold, f)) as pf:
                pf.read(ctx).close()
    ctx.synchronize()
    ctx.close()
    print("sanitize driver ok:", n, "rows")


if __name__ == "__main__":
    main()
