#!/usr/bin/env python
"""Device-resident timings of the BASELINE.json configs[1..4] pipelines on one B200 (informational; the
contract line is bench.py).  Prints one JSON line per config: rows/s, ms, algorithmic bytes and
the achieved fraction of the measured HBM peak for the whole chain."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bow_b200 import native as N  # noqa: E402

SEC = 1_000_000_000
SCALE = float(os.environ.get("BOW_BENCH_SCALE", "1.0"))
PEAK = 6548.5
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def dev_outs(W, nspecs):
    vals = [torch.empty(max(W, 1), dtype=torch.int64, device="cuda") for _ in range(nspecs)]
    bits = [torch.empty((W + 7) // 8 + 16, dtype=torch.uint8, device="cuda") for _ in range(nspecs)]
    outs = (N.OutCol * nspecs)()
    for j in range(nspecs):
        outs[j].values, outs[j].validity = vals[j].data_ptr(), bits[j].data_ptr()
    return outs, (vals, bits)


def timed(ctx, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ctx.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def report(name, n, secs, alg_bytes, extra=None):
    line = {"config": name, "rows": n, "ms": secs * 1e3, "rows_per_s": n / secs, "algorithmic_bytes": alg_bytes,
            "achieved_GBs": alg_bytes / secs / 1e9, "frac_of_measured_peak": alg_bytes / secs / 1e9 / PEAK}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def main():
    ctx = N.Ctx(0)
    which = sys.argv[1:] or ["1", "1b", "2", "3", "4", "whole", "fill", "sort", "parquet"]
    if "1" in which or "1b" in which:
        for tag, n in (("1", int(1e8 * SCALE)), ("1b", int(1e9 * SCALE))):
            if tag not in which:
                continue
            fr = N.Frame.generate(ctx, n, ncols=1, seed=42)
            r = N.Rolling(fr, 0, 60 * SEC)
            specs = [("WindowStart", 0), ("ArithmeticMean", 1), ("Sum", 1), ("Min", 1), ("Max", 1), ("Count", 1)]
            W = r.num_windows
            outs, keep = dev_outs(W, len(specs))
            arr = N.make_specs(specs)
            dt = timed(ctx, lambda: r.aggregate_device(arr, len(specs), outs), reps=20)
            report(f"configs[1] {n} rows mean/sum/min/max/count", n, dt, 16 * n + len(specs) * (8 * W + W // 8))
            r.close(); fr.close(); del keep
    if "2" in which:
        n = int(1e9 * SCALE)
        fr = N.Frame.generate(ctx, n, ncols=4, seed=7, null_mask=0xF, null_mod=10)
        r = N.Rolling(fr, 0, 900 * SEC, offset=420 * SEC)
        ops = ["WindowStart"] + ["Linear"] * 4
        specs = [("WindowStart", 0)]
        for c in range(1, 5):
            specs += [("WeightedAverageLinear", c), ("IntegralTrapezoid", c)]
        W = r.num_windows
        outs, keep = dev_outs(W, len(specs))
        arr = N.make_specs(specs)

        def chain():
            fi = r.interpolate(ops)
            r2 = N.Rolling(fi, 0, 900 * SEC, offset=420 * SEC)
            r2.aggregate_device(arr, len(specs), outs)
            ctx.synchronize()
            r2.close(); fi.close()
        dt = timed(ctx, chain, reps=3, warm=1)
        alg = 40 * n + 4 * n // 8 + len(specs) * (8 * W + W // 8)     # fused chain, inputs read once
        report("configs[2] Interpolate(WindowStart, Linear x4) -> WeightedAverageLinear + IntegralTrapezoid x4", n, dt, alg,
               {"note": "the interpolated frame is materialised: actual traffic ~3x algorithmic"})
        ctx.enable_timing(1)
        dtf = timed(ctx, lambda: r.interpolate_aggregate_device(ops, arr, len(specs), outs), reps=3, warm=1)
        report("configs[2] FUSED Interpolate -> Aggregate (bowgpu_rolling_interpolate_aggregate, no materialised frame)", n, dtf,
               alg, {"launches_last_call": ctx.last_timing().launches})
        ctx.enable_timing(0)
        fi = r.interpolate(ops)
        r2 = N.Rolling(fi, 0, 900 * SEC, offset=420 * SEC)
        dt2 = timed(ctx, lambda: r2.aggregate_device(arr, len(specs), outs), reps=5)
        report("configs[2] Aggregate only (on the interpolated frame)", n, dt2, alg)
        r2.close(); fi.close(); r.close(); fr.close(); del keep
    if "3" in which:
        n = int(1e9 * SCALE)
        fr = N.Frame.generate(ctx, n, ncols=1, seed=3, step=SEC, null_mask=1, null_mod=10, kind=1)
        r = N.Rolling(fr, 0, SEC)
        specs = [("WindowStart", 0), ("First", 1), ("Last", 1), ("Min", 1), ("Max", 1)]
        W = r.num_windows
        outs, keep = dev_outs(W, len(specs))
        arr = N.make_specs(specs)
        dt = timed(ctx, lambda: r.aggregate_device(arr, len(specs), outs), reps=10)
        report("configs[3] bursty Aggregate First/Last/Min/Max (no interpolation)", n, dt,
               16 * n + n // 8 + len(specs) * (8 * W + W // 8))

        def chain():
            fi = r.interpolate(["WindowStart", "StepPrevious"])
            r2 = N.Rolling(fi, 0, SEC)
            r2.aggregate_device(arr, len(specs), outs)
            ctx.synchronize()
            r2.close(); fi.close()
        dt = timed(ctx, chain, reps=3, warm=1)
        report("configs[3] bursty Interpolate(WindowStart, StepPrevious) -> First/Last/Min/Max", n, dt,
               16 * n + n // 8 + len(specs) * (8 * W + W // 8))
        ctx.enable_timing(1)
        dtf = timed(ctx, lambda: r.interpolate_aggregate_device(["WindowStart", "StepPrevious"], arr, len(specs), outs),
                    reps=5, warm=1)
        report("configs[3] FUSED Interpolate -> Aggregate", n, dtf, 16 * n + n // 8 + len(specs) * (8 * W + W // 8),
               {"launches_last_call": ctx.last_timing().launches})
        ctx.enable_timing(0)
        r.close(); fr.close(); del keep
    if "4" in which:
        n = int(5e8 * SCALE)
        fr = N.Frame.generate(ctx, n, ncols=16, seed=11, null_mask=0xAAAA, int_mask=0x00FF, null_mod=10)
        r = N.Rolling(fr, 0, 60 * SEC)
        opsl = ["Count", "Sum", "ArithmeticMean", "Min", "Max", "First", "Last", "IntegralStep", "IntegralTrapezoid",
                "WeightedAverageStep", "WeightedAverageLinear"]
        specs = [("WindowStart", 0)] + [(op, c) for c in range(1, 17) for op in opsl]
        W = r.num_windows
        outs, keep = dev_outs(W, len(specs))
        arr = N.make_specs(specs)
        dt = timed(ctx, lambda: r.aggregate_device(arr, len(specs), outs), reps=3, warm=1)
        report("configs[4] one GPU share: 16 columns x all aggregations (177 outputs)", n, dt,
               136 * n + 8 * n // 8 + len(specs) * (8 * W + W // 8))
        r.close(); fr.close(); del keep
    if "whole" in which:  # aggregation.Aggregate over the whole Bow = ONE window of 1e9 rows, and a rolling
        # whose single window holds every row (the huge-window end of the load-balance range: 122 070 tile records to join)
        n = int(1e9 * SCALE)
        fr = N.Frame.generate(ctx, n, ncols=1, seed=42)
        specs = [("WindowStart", 0), ("ArithmeticMean", 1), ("Sum", 1), ("Min", 1), ("Max", 1), ("Count", 1)]
        arr = N.make_specs(specs)
        outs, keep = dev_outs(1, len(specs))
        dt = timed(ctx, lambda: fr.aggregate_whole_device(0, arr, len(specs), outs), reps=10)
        report("whole-frame aggregation.Aggregate (one window) mean/sum/min/max/count", n, dt, 16 * n)
        r = N.Rolling(fr, 0, 2 * n * SEC)
        dt = timed(ctx, lambda: r.aggregate_device(arr, len(specs), outs), reps=10)
        report("IntervalRolling with ONE window holding every row, mean/sum/min/max/count", n, dt, 16 * n)
        specs2 = [("WindowStart", 0), ("IntegralStep", 1), ("IntegralTrapezoid", 1), ("WeightedAverageLinear", 1)]
        arr2 = N.make_specs(specs2)
        outs2, keep2 = dev_outs(1, len(specs2))
        dt = timed(ctx, lambda: fr.aggregate_whole_device(0, arr2, len(specs2), outs2), reps=10)
        report("whole-frame aggregation.Aggregate (one window) integrals", n, dt, 16 * n)
        r.close(); fr.close(); del keep, keep2
    if "fill" in which:   # next row of SURVEY 8(f): whole-column fills (bowfill.go)
        n = int(1e9 * SCALE)
        fr = N.Frame.generate(ctx, n, ncols=2, seed=5, null_mask=0x3, null_mod=10)
        for method in ("Previous", "Mean"):
            def run():
                out = fr.fill(method, 1)
                out.close()
            dt = timed(ctx, run, reps=3, warm=1)
            report(f"Fill{method} of one float64 column with 10 % nulls (2 untouched columns are copied)", n, dt,
                   16 * n + 2 * n // 8 + 2 * 16 * n)

        def run_lin():
            out = fr.fill_linear(0, 1)
            out.close()
        dt = timed(ctx, run_lin, reps=3, warm=1)
        report("FillLinear(time, value) incl. IsColSorted of the reference column", n, dt, 8 * n + 16 * n + 2 * n // 8 + 2 * 16 * n)
        fr.close()
    if "sort" in which:   # Bow.SortByCol (bowsort.go) — shuffled ns timestamps + 2 value columns
        n = int(1e8 * SCALE)
        perm = torch.randperm(n, device="cuda")
        t = 1_700_000_000_000_000_000 + perm * SEC
        v0 = torch.rand(n, dtype=torch.float64, device="cuda")
        v1 = torch.randint(0, 1 << 20, (n,), dtype=torch.int64, device="cuda")
        del perm
        arr = (N.Col * 3)()
        for j, (x, dt_) in enumerate(((t, N.INT64), (v0, N.FLOAT64), (v1, N.INT64))):
            arr[j].values, arr[j].validity, arr[j].offset, arr[j].length, arr[j].null_count, arr[j].dtype = \
                x.data_ptr(), None, 0, n, 0, dt_
        torch.cuda.synchronize()
        fr = N.Frame.from_col_descs(ctx, arr, 3, N.MEM_DEVICE, keep=[t, v0, v1])
        for col, what in ((0, "shuffled int64 ns timestamps (5 varying digits)"), (1, "uniform float64 keys (8 digits)"),
                          (2, "20-bit int64 keys (3 digits)")):
            def run():
                out = fr.sort_by_col(col)
                out.close()
            dt = timed(ctx, run, reps=3, warm=1)
            report(f"SortByCol, 3 columns, by {what}", n, dt, 2 * 3 * 8 * n)
        fr.close()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ks, order = torch.sort(t, stable=True)
            g0, g1 = v0[order], v1[order]
        torch.cuda.synchronize()
        report("(context) torch.sort(stable) + 2 gathers of the same timestamps", n, (time.perf_counter() - t0) / 3, 2 * 3 * 8 * n)
    if "parquet" in which:  # bow.NewBowFromParquet — file (page cache) -> device-resident frame
        import numpy as np
        import pyarrow as pa
        import pyarrow.parquet as pq
        import tempfile
        n = int(5e7 * SCALE)
        rng = np.random.default_rng(3)
        tcol = 1_700_000_000_000_000_000 + np.arange(n, dtype=np.int64) * SEC
        table = pa.table({"t": pa.array(tcol),
                          "a": pa.array(rng.random(n), mask=rng.random(n) < 0.1),
                          "b": pa.array(rng.random(n), mask=rng.random(n) < 0.1)})
        d = tempfile.mkdtemp()
        for tag, opts in (("SNAPPY, PLAIN, 1 MiB pages (pyarrow defaults without dictionary)", dict(compression="SNAPPY", use_dictionary=False)),
                          ("UNCOMPRESSED, PLAIN", dict(compression="NONE", use_dictionary=False)),
                          ("SNAPPY, PLAIN, 8 KiB pages (the page size of the reference's writer)",
                           dict(compression="SNAPPY", use_dictionary=False, data_page_size=8192))):
            path = os.path.join(d, "f.parquet")
            pq.write_table(table, path, **opts)
            fbytes = os.path.getsize(path)
            open(path, "rb").read()  # page cache
            pf = N.ParquetFile(path)

            def run():
                pf.read(ctx).close()
            dt = timed(ctx, run, reps=3, warm=1)
            t0 = time.perf_counter()
            pq.read_table(path, use_threads=True)
            t_arrow = time.perf_counter() - t0
            report(f"NewBowFromParquet {tag}: file -> device frame (3 columns, 10 % nulls)", n, dt, 3 * 8 * n,
                   {"file_bytes": fbytes, "pages": pf.plan()["pages"], "pyarrow_read_table_ms_all_threads": t_arrow * 1e3})
            pf.close()
            os.remove(path)


if __name__ == "__main__":
    main()
