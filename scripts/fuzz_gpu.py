#!/usr/bin/env python
"""Randomised differential campaign: every GPU entry point against the oracle on random shapes, biased towards the
edges of the kernels' geometry (16-row phases, 64-row thread runs, 8192-row tiles, 64-window shard cuts).
usage (on a B200): python scripts/fuzz_gpu.py [seconds] [seed]   -> exit status 1 and the failing seed on a mismatch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bow_b200 import native as N  # noqa: E402
from bow_b200 import parallel as PP  # noqa: E402
from oracle import refc as R  # noqa: E402
from tests import helpers as H  # noqa: E402

AGGS = ["Count", "Sum", "ArithmeticMean", "Min", "Max", "First", "Last", "IntegralStep", "IntegralTrapezoid",
        "WeightedAverageStep", "WeightedAverageLinear"]
TOL = {"Sum", "ArithmeticMean", "IntegralStep", "IntegralTrapezoid", "WeightedAverageStep", "WeightedAverageLinear"}


def pick_n(rng):
    if os.environ.get("FUZZ_BIG"):   # mostly multi-tile inputs
        base = int(rng.choice([8192, 16384, 24576, 65536, 131072]))
        return base + int(rng.integers(-70, 71)) if rng.random() < 0.7 else int(rng.integers(8000, 200000))
    base = int(rng.choice([0, 1, 16, 64, 1024, 8192, 16384, 24576, 65536]))
    return max(0, base + int(rng.integers(-3, 4)) if rng.random() < 0.6 else int(rng.integers(0, 70000)))


def same(sp, got, want, interval, what):
    (gv, gm), (wv, wm) = got, want
    assert gv.dtype == wv.dtype and np.array_equal(gm, wm), f"{what} {sp}: validity/dtype"
    a, b = gv[gm], wv[wm]
    if sp[0] in TOL:
        fin = np.isfinite(b)
        scale = 2e3 * (interval if "Integral" in sp[0] else 1.0)
        assert np.all(np.abs(a[fin] - b[fin]) <= 1e-12 * np.maximum(np.abs(b[fin]), scale) * 64), f"{what} {sp}: values"
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what} {sp}: nan pattern"
    else:
        eq = a.view(np.int64) == b.view(np.int64)
        if a.dtype == np.float64:
            eq |= np.isnan(a) & np.isnan(b)
        assert eq.all(), f"{what} {sp}: values {a[~eq][:3]} vs {b[~eq][:3]}"


def one(ctx, seed):
    rng = np.random.default_rng(seed)
    n = pick_n(rng)
    kind = str(rng.choice(["regular", "dense", "sparse", "bursty"]))
    t = H.random_times(rng, n, kind)
    if n:
        t = t - int(t[0]) + int(rng.integers(0, 5000))
    interval = int(rng.choice([1, 3, 17, 64, 500, 4000, 10 ** 6]))
    offset = int(rng.integers(-interval, 2 * interval))
    cols = [(t, None), H.random_values(rng, n, np.float64, float(rng.choice([0, 0.2, 0.9])), specials=rng.random() < 0.2),
            H.random_values(rng, n, np.int64, float(rng.choice([0, 0.5])))]
    specs = [("WindowStart", 0)] + [(a, int(rng.integers(1, 3))) for a in rng.choice(AGGS, size=int(rng.integers(1, 7)))]
    # sums over NaN / Inf / 1e300 are order dependent beyond any tolerance: exact ops only on such a column
    has_special = bool(np.isnan(cols[1][0]).any() or np.isinf(cols[1][0]).any() or (np.abs(cols[1][0]) > 1e200).any())
    if has_special:
        specs = [s for s in specs if s[0] not in TOL or s[1] != 1] or [("WindowStart", 0), ("Min", 1)]
    mode = str(rng.choice(["agg", "agg_inclusive", "sharded", "fused", "whole", "fill", "interp", "host", "sort"]))
    what = f"seed={seed} mode={mode} n={n} kind={kind} I={interval} off={offset}"
    fr = N.Frame.from_numpy(ctx, cols)
    try:
        if mode in ("agg", "agg_inclusive"):
            inc = mode == "agg_inclusive"
            got = N.Rolling(fr, 0, interval, offset=offset, inclusive=inc).aggregate(specs)
            want = R.RefRolling(R.Frame(cols), 0, interval, offset=offset, inclusive=inc).aggregate(specs)
            for sp, g, w in zip(specs, got, want):
                same(sp, g, w, interval, what)
        elif mode == "host" and n:
            got = N.aggregate_host(ctx, cols, 0, interval, specs, offset=offset)
            want = R.RefRolling(R.Frame(cols), 0, interval, offset=offset).aggregate(specs)
            for sp, g, w in zip(specs, got, want):
                same(sp, g, w, interval, what)
        elif mode == "sharded" and n:
            g = int(rng.integers(2, 6))
            shards, s0 = PP.plan_for_columns(t, interval, offset, g)
            per = [PP.aggregate_shard(cols, s, 0, interval, s0, False, specs) for s in shards]
            want = R.RefRolling(R.Frame(cols), 0, interval, offset=offset).aggregate(specs)
            for sp, gg, w in zip(specs, PP.concat_outputs(per), want):
                same(sp, gg, w, interval, what)
        elif mode in ("fused", "interp") and n:
            ops = ["WindowStart", str(rng.choice(["Linear", "StepPrevious", "None_", "StepNext"])), str(rng.choice(["Linear", "StepPrevious", "StepNext"]))]
            ref = R.RefRolling(R.Frame(cols), 0, interval, offset=offset)
            icols = ref.interpolate(ops)
            r = N.Rolling(fr, 0, interval, offset=offset)
            if mode == "interp":
                out = r.interpolate(ops)
                got = out.download()
                out.close()
                for j in range(3):
                    same(("col", j), got[j], icols[j], interval, what)
            else:
                ic = [(v, None if m.all() else m) for v, m in icols]
                want = R.RefRolling(R.Frame(ic), 0, interval, offset=offset).aggregate(specs)
                for sp, gg, w in zip(specs, r.interpolate_aggregate(ops, specs), want):
                    same(sp, gg, w, interval, what)
        elif mode == "whole":
            got = fr.aggregate_whole(0, specs)
            want = R.aggregate_whole(R.Frame(cols), 0, specs)
            span = float(t[-1] - t[0]) if n else 1.0
            for sp, gg, w in zip(specs, got, want):
                same(sp, gg, w, max(span, 1.0) * max(n, 1), what)
        elif mode == "sort" and n:
            key = rng.integers(-50, 50, size=n) if rng.random() < 0.5 else rng.permutation(n).astype(np.int64) - n // 2
            scols = [(key.astype(np.float64) / 2 if rng.random() < 0.3 else key, None), cols[1], cols[2]]
            sf = N.Frame.from_numpy(ctx, scols)
            try:
                out = sf.sort_by_col(0)
                want = R.sort_by_col(scols, 0)
                assert (out is None) == (want is None), what
                if out is not None:
                    gs = out.download()
                    out.close()
                    for c in range(3):
                        gv, gm = gs[c]
                        wv, wm = want[c]
                        assert np.array_equal(gm, wm) and np.array_equal(gv[gm].view(np.int64), wv[wm].view(np.int64)), f"{what} col {c}"
            finally:
                sf.close()
        elif mode == "fill":
            method = str(rng.choice(["Previous", "Next", "Mean"]))
            out = fr.fill(method, 1, 2)
            got = out.download()
            out.close()
            for c in (1, 2):
                same(("col", c), got[c], R.fill(R.Frame(cols), method, c), 1, what)
            dn = fr.drop_nils(1)
            gd = dn.download()
            dn.close()
            wd = R.drop_nils(cols, (1,))
            for c in range(3):
                same(("col", c), gd[c], wd[c], 1, what)
    finally:
        fr.close()


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    ctx = N.Ctx(0)
    t0, k = time.time(), 0
    while time.time() - t0 < seconds:
        try:
            one(ctx, seed0 + k)
        except AssertionError as e:
            print(f"MISMATCH {e}", flush=True)
            return 1
        k += 1
    print(f"fuzz ok: {k} cases in {time.time() - t0:.0f} s (seeds {seed0}..{seed0 + k - 1})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
