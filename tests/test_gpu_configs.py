"""The BASELINE.json configurations at their FULL sizes on one B200 (configs[1]..[4]; configs[0] is the fixture
test in test_gpu_rolling_api.py).  Inputs are generated on the device (csrc/generate.cu); parity is checked
  * against the oracle on sub-ranges regenerated bit-for-bit on the CPU (bow_b200/synth.py): the head of the
    series, its tail and chunks in the middle — every window that lies completely inside a chunk is compared;
  * through size-independent properties of the whole result: Σ Count == number of valid rows, every non-empty
    window's First/Last/Min/Max consistent (min <= first,last <= max), WindowStart == lattice, row counts of the
    interpolated frame == rows + missing window starts, sortedness of the interpolated time column.
BOW_TEST_SCALE (default 1.0) scales the row counts down for quick runs.
Bit-exact: bounds, WindowStart, Count, Min, Max, First, Last, interpolated rows.  Tolerance 1e-12 relative to
max(|ref|, Σ|terms|) for float64 sums, means, integrals and weighted averages (reduction order differs)."""
import os

import numpy as np
import pytest

from bow_b200 import synth
from oracle import refc as R

pytestmark = pytest.mark.gpu

SCALE = float(os.environ.get("BOW_TEST_SCALE", "1.0"))
T0 = synth.T0_DEFAULT
SEC = 1_000_000_000


@pytest.fixture(scope="module")
def ctx():
    from bow_b200 import native as N
    c = N.Ctx(0)
    yield c
    c.close()


def rows(n):
    return max(200_000, int(n * SCALE))


def device_outputs(ctx, rolling, specs):
    """aggregate with device-resident outputs -> list of (torch int64 values, torch uint8 bitmap)"""
    import torch
    from bow_b200 import native as N
    W = rolling.num_windows
    vals = [torch.empty(max(W, 1), dtype=torch.int64, device="cuda") for _ in specs]
    bits = [torch.zeros((W + 7) // 8 + 16, dtype=torch.uint8, device="cuda") for _ in specs]
    outs = (N.OutCol * len(specs))()
    for j in range(len(specs)):
        outs[j].values, outs[j].validity = vals[j].data_ptr(), bits[j].data_ptr()
    rolling.aggregate_device(N.make_specs(specs), len(specs), outs)
    ctx.synchronize()
    torch.cuda.synchronize()
    return vals, bits


def fused_outputs(ctx, rolling, ops, specs):
    """bowgpu_rolling_interpolate_aggregate with device-resident outputs"""
    import torch
    from bow_b200 import native as N
    W = rolling.num_windows
    vals = [torch.empty(max(W, 1), dtype=torch.int64, device="cuda") for _ in specs]
    bits = [torch.zeros((W + 7) // 8 + 16, dtype=torch.uint8, device="cuda") for _ in specs]
    outs = (N.OutCol * len(specs))()
    for j in range(len(specs)):
        outs[j].values, outs[j].validity = vals[j].data_ptr(), bits[j].data_ptr()
    rolling.interpolate_aggregate_device(ops, N.make_specs(specs), len(specs), outs)
    ctx.synchronize()
    torch.cuda.synchronize()
    return vals, bits


def assert_same_outputs(specs, a, b, W, in_dtypes, what, rel=1e-12, scale=1.0):
    """fused vs materialising chain on the device: identical bitmaps; values bit-exact, or within `rel` of
    max(|x|, scale) for the float64 sums"""
    import torch
    (va, ba), (vb, bb) = a, b
    nb = (W + 7) // 8
    for (op, col), x, y, p, q in zip(specs, va, vb, ba, bb):
        assert torch.equal(p[:nb], q[:nb]), f"{what} {op}({col}): validity differs"
        if op in TOL_OPS:
            fx, fy = x[:W].view(torch.float64), y[:W].view(torch.float64)
            tol = rel * torch.clamp(fy.abs(), min=scale)
            assert bool(((fx - fy).abs() <= tol).all()), f"{what} {op}({col}): max diff {float((fx - fy).abs().max())}"
        else:
            assert torch.equal(x[:W], y[:W]), f"{what} {op}({col}): values differ"


def host_window_range(vals, bits, k0, k1, is_float):
    """windows [k0, k1) of a device output -> (values ndarray, mask ndarray); k0 must be a multiple of 8"""
    from bow_b200 import native as N
    assert k0 % 8 == 0
    v = vals[k0:k1].cpu().numpy()
    b = bits[k0 // 8:(k1 + 7) // 8].cpu().numpy()
    return (v.view(np.float64) if is_float else v), N.unpack_bits(b, k1 - k0)


FLOAT_OPS = {"Sum", "ArithmeticMean", "Min", "Max", "IntegralStep", "IntegralTrapezoid", "WeightedAverageStep",
             "WeightedAverageLinear"}
TOL_OPS = {"Sum", "ArithmeticMean", "IntegralStep", "IntegralTrapezoid", "WeightedAverageStep", "WeightedAverageLinear"}


def out_is_float(op, in_dtype):
    return op in FLOAT_OPS or (op in ("First", "Last") and in_dtype == np.float64)


def compare_chunk(what, specs, vals, bits, k_glob0, ref_out, j0, j1, in_dtypes, scale_of):
    """oracle windows [j0, j1) of a chunk == device windows [k_glob0 + j0, k_glob0 + j1)"""
    ka = (k_glob0 + j0 + 7) // 8 * 8          # byte-aligned start inside the range
    ja = ka - k_glob0
    if ja >= j1:
        return 0
    for s, (op, col), v, b in zip(range(len(specs)), specs, vals, bits):
        isf = out_is_float(op, in_dtypes[col])
        gv, gm = host_window_range(v, b, ka, k_glob0 + j1, isf)
        wv, wm = ref_out[s][0][ja:j1], ref_out[s][1][ja:j1]
        assert np.array_equal(gm, wm), f"{what} {op}({col}): validity differs at windows {np.flatnonzero(gm != wm)[:5] + ka}"
        # Sum of an int64 column (values below 2^20 here): every partial sum is an exact integer -> bit-exact
        if op in TOL_OPS and not (op == "Sum" and in_dtypes[col] == np.int64):
            tol = 1e-12 * np.maximum(np.abs(wv[wm]), scale_of(op))
            bad = np.flatnonzero(np.abs(gv[gm] - wv[wm]) > tol)
            assert bad.size == 0, f"{what} {op}({col}): {gv[gm][bad[:3]]} vs {wv[wm][bad[:3]]}"
        else:
            a, c = gv[gm], wv[wm]
            same = (a.view(np.int64) == c.view(np.int64)) if isf else (a == c)
            assert same.all(), f"{what} {op}({col}): first mismatch at window {np.flatnonzero(~same)[:3]}"
    return j1 - ja


# ---------------------------------------------------------------------------------------------------------------
def test_config1_100M_regular_mean_sum_min_max_count(ctx):
    """configs[1]: 100M rows, 1 float64 column, 1-minute windows, WindowStart + mean/sum/min/max/count"""
    import torch
    from bow_b200 import native as N
    n, interval = rows(100_000_000), 60 * SEC
    fr = N.Frame.generate(ctx, n, ncols=1, seed=42)
    r = N.Rolling(fr, 0, interval)
    W = r.num_windows
    specs = [("WindowStart", 0), ("ArithmeticMean", 1), ("Sum", 1), ("Min", 1), ("Max", 1), ("Count", 1)]
    vals, bits = device_outputs(ctx, r, specs)
    s0 = r.first_window_start
    # whole-result properties
    assert int(vals[5][:W].sum().item()) == n                                    # Σ Count == rows (no nulls)
    assert torch.equal(vals[0][:W], s0 + torch.arange(W, device="cuda") * interval)
    mn, mx, mean = (vals[j][:W].view(torch.float64) for j in (3, 4, 1))
    assert bool(((mn <= mean) & (mean <= mx)).all())
    total = float(vals[2][:W].view(torch.float64).sum().item())
    assert abs(total - n * 0.5) < 8 * (n / 12) ** 0.5                            # uniform [0,1) values: 8 sigma
    # oracle on the head, a middle chunk and the tail (regenerated rows)
    m = min(n, 3_000_000)
    checked = 0
    for row0 in sorted({0, (n // 2) // 60 * 60, n - m}):
        m = min(m, n - row0)
        cols = synth.regular_frame(row0, m, 1, 42)
        ref = R.RefRolling(R.Frame(cols), 0, interval)
        out = ref.aggregate(specs)
        kg = (ref.first_window_start - s0) // interval
        lo = 0 if row0 == 0 else 1
        hi = ref.num_windows if row0 + m == n else ref.num_windows - 1
        checked += compare_chunk(f"config1 rows@{row0}", specs, vals, bits, kg, out, lo, hi,
                                 [np.int64, np.float64], lambda op: 60.0)
    assert checked > 0
    r.close()
    fr.close()


def test_config2_1B_interpolate_linear_then_weighted_average(ctx):
    """configs[2]: 1B rows, 4 float64 columns with 10 % nulls, 15-minute windows with Offset,
    Interpolate(WindowStart, Linear x4) then WindowStart + WeightedAverageLinear + IntegralTrapezoid per column"""
    import torch
    from bow_b200 import native as N
    n, interval, offset = rows(1_000_000_000), 900 * SEC, 420 * SEC
    fr = N.Frame.generate(ctx, n, ncols=4, seed=7, null_mask=0xF, null_mod=10)
    r = N.Rolling(fr, 0, interval, offset=offset)
    W, s0 = r.num_windows, r.first_window_start
    ops = ["WindowStart"] + ["Linear"] * 4
    fi = r.interpolate(ops)
    # every window start except the first one coincides with a row (1 s step): exactly one inserted row iff s0 < t[0]
    assert fi.num_rows == n + (1 if s0 < T0 else 0)
    r2 = N.Rolling(fi, 0, interval, offset=offset)
    assert r2.num_windows == W and r2.first_window_start == s0
    specs = [("WindowStart", 0)]
    for c in range(1, 5):
        specs += [("WeightedAverageLinear", c), ("IntegralTrapezoid", c)]
    vals, bits = device_outputs(ctx, r2, specs)
    assert torch.equal(vals[0][:W], s0 + torch.arange(W, device="cuda") * interval)
    # the fused chain (no materialised frame) must give the same windows
    assert_same_outputs(specs, fused_outputs(ctx, r, ops, specs), (vals, bits), W, None, "config2 fused",
                        scale=float(interval))
    # weighted average of values in [0,1) lies in [0,1]; integral = average * interval
    for j in range(1, len(specs), 2):
        wa = vals[j][:W - 1].view(torch.float64)
        it = vals[j + 1][:W - 1].view(torch.float64)
        assert bool(((wa >= 0) & (wa <= 1)).all())
        assert bool(((it - wa * float(interval)).abs() <= 1e-9 * float(interval)).all())
    m = min(n, 2_000_000)
    in_dtypes = [np.int64] + [np.float64] * 4
    checked = 0
    for row0 in sorted({0, (n // 3), n - m}):
        m = min(m, n - row0)
        cols = synth.regular_frame(row0, m, 4, 7, null_mask=0xF, null_mod=10)
        ref = R.RefRolling(R.Frame(cols), 0, interval, offset=offset)
        icols = ref.interpolate(ops)
        # interpolated rows of the chunk vs the device frame (skip the chunk's first window: its start row has no
        # previous valid value inside the chunk)
        if row0 == 0:
            got = fi.download(0, len(icols[0][0]))
            for j in range(5):
                assert np.array_equal(got[j][1], icols[j][1]), j
                assert np.array_equal(got[j][0][got[j][1]].view(np.int64), icols[j][0][icols[j][1]].view(np.int64)), j
        icols = [(v, None if mk.all() else mk) for v, mk in icols]
        ref2 = R.RefRolling(R.Frame(icols), 0, interval, offset=offset)
        out = ref2.aggregate(specs)
        kg = (ref2.first_window_start - s0) // interval
        lo = 0 if row0 == 0 else 2
        hi = ref2.num_windows if row0 + m == n else ref2.num_windows - 1
        checked += compare_chunk(f"config2 rows@{row0}", specs, vals, bits, kg, out, lo, hi, in_dtypes,
                                 lambda op: float(interval) if op.startswith("Integral") else 1.0)
    assert checked > 0
    for o in (r2, fi, r, fr):
        o.close()


def test_config3_1B_bursty_first_last_min_max_stepprevious(ctx):
    """configs[3]: ~1B rows of bursty timestamps (windows of 0 .. ~1e6 rows), 1 float64 column with 10 % nulls,
    Interpolate(WindowStart, StepPrevious) then WindowStart + First/Last/Min/Max (+ Count): load-balance stress"""
    import torch
    from bow_b200 import native as N
    n, interval, seed = rows(1_000_000_000), SEC, 3
    off = synth.bursty_offsets(seed, n)
    fr = N.Frame.generate(ctx, n, ncols=1, seed=seed, step=interval, null_mask=1, null_mod=10, kind=1)
    r = N.Rolling(fr, 0, interval)
    W, s0 = r.num_windows, r.first_window_start
    k_last = int(np.searchsorted(off, n - 1, side="right")) - 1          # window of the last row
    k_first = int(np.searchsorted(off, 0, side="right")) - 1             # window of row 0 (leading empty windows)
    assert s0 == T0 + k_first * interval and W == k_last - k_first + 1
    # window boundaries of the whole series == the generator's offsets
    first, _ = r.bounds()
    want_first = np.minimum(off[k_first:k_last + 2], n)
    want_first[0] = 0
    assert np.array_equal(first, want_first)
    fi = r.interpolate(["WindowStart", "StepPrevious"])
    # a start row is inserted for every window without a row exactly at S_k: empty windows and shifted ones
    cnt = np.diff(np.minimum(off[k_first:k_last + 2], n))
    shifted = ((synth.synth_u(seed, 0, np.arange(k_first, k_last + 1).astype(np.uint64)) >> np.uint64(40)) & np.uint64(1)) == 1
    missing = (cnt == 0) | shifted
    assert fi.num_rows == n + int(missing.sum())
    r2 = N.Rolling(fi, 0, interval)
    assert r2.num_windows == W
    specs = [("WindowStart", 0), ("First", 1), ("Last", 1), ("Min", 1), ("Max", 1), ("Count", 1)]
    vals, bits = device_outputs(ctx, r2, specs)
    assert_same_outputs(specs, fused_outputs(ctx, r, ["WindowStart", "StepPrevious"], specs), (vals, bits), W, None,
                        "config3 fused")
    # properties over all windows: Σ Count == valid rows of the interpolated frame; min <= first,last <= max
    _, vbits = fi.device_ptrs(1)
    assert vbits, "the interpolated value column must carry nulls"
    bm = device_bytes(vbits, (fi.num_rows + 7) // 8)       # trailing bits of the last byte are zero
    got_valid = sum(int(((bm >> k) & 1).sum(dtype=torch.int64).item()) for k in range(8))
    assert int(vals[5][:W].sum().item()) == got_valid
    ok = N.unpack_bits(bits[3][:(W + 7) // 8].cpu().numpy(), W)
    okt = torch.from_numpy(ok).cuda()
    f_, l_, mn, mx = (vals[j][:W].view(torch.float64)[okt] for j in (1, 2, 3, 4))
    assert bool(((mn <= f_) & (f_ <= mx) & (mn <= l_) & (l_ <= mx)).all())
    # oracle on chunks aligned to window starts: the head, the chunk holding the largest window, the tail
    big = k_first + int(np.argmax(cnt))
    chunks = [(k_first, min(k_last, k_first + 400)), (max(k_first, big - 40), min(k_last, big + 40)),
              (max(k_first, k_last - 400), k_last)]
    checked = 0
    for ka, kb in chunks:
        ra, rb = int(off[ka]), int(min(off[kb + 1], n))
        if rb - ra > 6_000_000 or rb <= ra:
            continue
        cols = synth.bursty_frame(ra, rb - ra, 1, seed, T0, interval, null_mask=1, null_mod=10, off=off)
        ref = R.RefRolling(R.Frame(cols), 0, interval)
        icols = ref.interpolate(["WindowStart", "StepPrevious"])
        icols = [(v, None if mk.all() else mk) for v, mk in icols]
        ref2 = R.RefRolling(R.Frame(icols), 0, interval)
        out = ref2.aggregate(specs)
        kg = (ref2.first_window_start - s0) // interval
        lo = 0 if ra == 0 else 2          # StepPrevious of the chunk's first windows needs rows before the chunk
        checked += compare_chunk(f"config3 windows {ka}..{kb}", specs, vals, bits, kg, out, lo, ref2.num_windows,
                                 [np.int64, np.float64], lambda op: 1.0)
    assert checked > 0
    for o in (r2, fi, r, fr):
        o.close()


class _DevMem:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_bytes(ptr, nbytes):
    """zero-copy torch view of library-owned device memory"""
    import torch
    return torch.as_tensor(_DevMem(ptr, nbytes), device="cuda")


def test_config4_500M_16_columns_all_aggregations(ctx):
    """configs[4], one GPU's share (4B rows / 8 GPUs): 500M rows, 8 int64 + 8 float64 columns (odd ones with 10 %
    nulls), every aggregation.* on every column in ONE Aggregate call (177 output columns)"""
    import torch
    from bow_b200 import native as N
    n, interval = rows(500_000_000), 60 * SEC
    int_mask, null_mask = 0x00FF, 0xAAAA
    fr = N.Frame.generate(ctx, n, ncols=16, seed=11, null_mask=null_mask, int_mask=int_mask, null_mod=10)
    r = N.Rolling(fr, 0, interval)
    W, s0 = r.num_windows, r.first_window_start
    ops = ["Count", "Sum", "ArithmeticMean", "Min", "Max", "First", "Last", "IntegralStep", "IntegralTrapezoid",
           "WeightedAverageStep", "WeightedAverageLinear"]
    specs = [("WindowStart", 0)] + [(op, c) for c in range(1, 17) for op in ops]
    assert len(specs) == 177
    vals, bits = device_outputs(ctx, r, specs)
    assert torch.equal(vals[0][:W], s0 + torch.arange(W, device="cuda") * interval)
    in_dtypes = [np.int64] + [np.int64 if (int_mask >> c) & 1 else np.float64 for c in range(16)]
    # Σ Count of a column without nulls == rows; int64 sums are exact (values < 2^20, partial sums < 2^53)
    for c in range(1, 17):
        j = 1 + (c - 1) * len(ops)
        if not (null_mask >> (c - 1)) & 1:
            assert int(vals[j][:W].sum().item()) == n, c
    m = min(n, 600_000)
    checked = 0
    for row0 in sorted({0, n - m}):
        m = min(m, n - row0)
        cols = synth.regular_frame(row0, m, 16, 11, null_mask=null_mask, int_mask=int_mask, null_mod=10)
        ref = R.RefRolling(R.Frame(cols), 0, interval)
        out = ref.aggregate(specs)
        kg = (ref.first_window_start - s0) // interval
        lo = 0 if row0 == 0 else 1
        hi = ref.num_windows if row0 + m == n else ref.num_windows - 1

        def scale_of(op):       # Σ|terms| of a window: values < 2^20, 60 rows, 60e9 ns
            return {"Sum": 2.0 ** 20 * 60, "IntegralStep": 2.0 ** 20 * 60e9, "IntegralTrapezoid": 2.0 ** 20 * 60e9}.get(op, 2.0 ** 20)
        checked += compare_chunk(f"config4 rows@{row0}", specs, vals, bits, kg, out, lo, hi, in_dtypes, scale_of)
    assert checked > 0
    r.close()
    fr.close()
