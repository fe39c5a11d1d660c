#!/usr/bin/env python
"""bench.py — input rows/s of IntervalRolling.Aggregate on B200 (BASELINE.json metric).

Default workload (`--config 1`, the north-star target): synthetic 1 000 000 000 rows — int64 ns time with a regular 1 s
step + one float64 column — 1-minute windows, WindowStart(time) + ArithmeticMean / Sum / Min / Max / Count(value).
That is BASELINE.json configs[1] at the row count the target is stated on (16 GB of input: it fits one GPU).
`--config 2`: configs[2], 1e9 rows x 4 float64 columns with 10 % nulls, 15-min windows with Offset,
Interpolate(WindowStart, Linear) -> WeightedAverageLinear + IntegralTrapezoid, through the fused call.

`--gpus N` is STRONG scaling: the same global frame is range-partitioned by window index over the N ranks
(bow_b200/partition.py; one process per GPU, cuts on multiples of 64 windows, one-row / one-window halos), no collective
on the data path; per-shard outputs are the concatenation.

One JSON line on stdout (rank 0):
  value     rows/s, inputs resident in HBM (CUDA events around K calls of the C-ABI entry point, max over ranks)
  e2e       rows/s through the one-shot reference-facing call with HOST (pinned) buffers: every rank runs the pipelined
            bowgpu_aggregate_host_ex / bowgpu_interpolate_aggregate_host on its shard (H2D + kernels + D2H inside)
  roofline  dominant kernel (segreduce main): algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json,
            plus the same fraction for the WHOLE step (SURVEY 8d definition of algorithmic bytes)
  cpu_baseline  the oracle (C port of the reference algorithm, 1 core) timed on this box's host CPU

`--impl reference` times the reference's CPU algorithm (oracle/ref.c port; the Go reference cannot be built in this
image — no Go toolchain) on bounded samples of the same workload, same `config`.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEC = 1_000_000_000
T0 = 1_700_000_000_000_000_000
STEP = SEC
METRIC = "input rows/s for IntervalRolling.Aggregate"

WORKLOADS = {
    1: dict(name="configs[1] at the north-star row count: synthetic {rows} rows, int64 ns time (regular 1 s step) + 1 float64 "
                 "col, 1-min windows, WindowStart + mean/sum/min/max/count",
            interval=60 * SEC, offset=0, ncols=1, seed=42, null_mask=0, null_mod=0, ops=None,
            aggs=[("WindowStart", 0), ("ArithmeticMean", 1), ("Sum", 1), ("Min", 1), ("Max", 1), ("Count", 1)]),
    2: dict(name="configs[2]: synthetic {rows} rows, 4 float64 cols with 10 % nulls, 15-min windows with Offset, "
                 "Interpolate(WindowStart, Linear x4) -> WeightedAverageLinear + IntegralTrapezoid x4 (fused call)",
            interval=900 * SEC, offset=420 * SEC, ncols=4, seed=7, null_mask=0xF, null_mod=10,
            ops=["WindowStart", "Linear", "Linear", "Linear", "Linear"],
            aggs=[("WindowStart", 0)] + [x for c in range(1, 5) for x in (("WeightedAverageLinear", c), ("IntegralTrapezoid", c))]),
}


def make_config(args, world):
    """identical for both arms (the driver compares them)"""
    from bow_b200 import partition as P
    wl = WORKLOADS[args.config]
    n = args.rows
    off = P.normalise_offset(wl["interval"], wl["offset"])
    W = P.num_windows(T0, T0 + (n - 1) * STEP, wl["interval"], off)
    return {"workload": wl["name"].format(rows=n), "rows": n, "windows": W, "interval_ns": wl["interval"],
            "offset_ns": wl["offset"], "value_columns": wl["ncols"], "aggregations": [a for a, _ in wl["aggs"]],
            "interpolations": wl["ops"], "parallelism": f"range-partition x{world} of the same global frame",
            "l2": "inputs (>= 2 GB per GPU) are larger than L2 (126 MB); no explicit flush"}


def algorithmic_bytes(wl, n, W):
    """every input buffer once + every output column once"""
    nulls = bin(wl["null_mask"]).count("1")
    return 8 * n + 8 * n * wl["ncols"] + nulls * ((n + 7) // 8) + len(wl["aggs"]) * (8 * W + (W + 7) // 8)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_near_gpu(local_rank):
    """pin this process to the CPUs next to its GPU (pinned host buffers are then NUMA-local to it)"""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
        return {"pci": bus, "numa_node": int(node), "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:80]}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def host_frame(wl, row0, n):
    from bow_b200 import synth
    return synth.regular_frame(row0, n, wl["ncols"], wl["seed"], T0, STEP, null_mask=wl["null_mask"], null_mod=wl["null_mod"] or 10)


def oracle_step(R, wl, fr):
    r = R.RefRolling(fr, 0, wl["interval"], offset=wl["offset"])
    if wl["ops"]:
        ops = ["None_" if o == "None" else o for o in wl["ops"]]
        icols = [(v, None if m.all() else m) for v, m in r.interpolate(ops)]
        r = R.RefRolling(R.Frame(icols), 0, wl["interval"], offset=wl["offset"])
    return r.aggregate(wl["aggs"])


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (C port, single thread like the single-goroutine Go path) on bounded
    samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import refc as R
    R.build()
    wl = WORKLOADS[args.config]
    sample = min(args.rows, args.ref_rows // max(1, wl["ncols"]))
    fr = R.Frame(host_frame(wl, 0, sample))
    for _ in range(args.warmup):
        oracle_step(R, wl, fr)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(R, wl, fr)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, world),
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": 1, "kind": "port",
                         "sample": f"each step = the first {sample} rows of the workload; Go reference not buildable here (no Go "
                                   "toolchain): C port of its algorithm, oracle/ref.c, one thread like the single-goroutine "
                                   "reference path (rolling.go:32,160,175)",
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bow_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_near_gpu(local_rank)
    from bow_b200 import native as N
    from bow_b200 import partition as P
    from bow_b200 import synth
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce(x, dist.ReduceOp.MAX) if world > 1 else x

    def sum_over_ranks(x):
        return reduce(x, dist.ReduceOp.SUM) if world > 1 else x

    wl = WORKLOADS[args.config]
    interval, offset, ncols = wl["interval"], wl["offset"], wl["ncols"]
    specs = wl["aggs"]
    fused = wl["ops"] is not None
    n_total = args.rows
    t_last = T0 + (n_total - 1) * STEP
    off_n = P.normalise_offset(interval, offset)
    s0 = P.first_window_start(T0, interval, off_n)
    lb = P.regular_lower_bound(T0, STEP, n_total)

    # ---- range partition of the global frame (no collective on the data path) --------------------------------------
    if fused:
        def valid_near(row, lo, hi):     # validity of rows [lo, hi) of every interpolated column (host mirror of the generator)
            m = np.ones(hi - lo, dtype=bool)
            for c in range(ncols):
                if (wl["null_mask"] >> c) & 1:
                    m &= synth.values(wl["seed"], c + 1, lo, hi - lo, False, wl["null_mod"])[1]
            return m

        def prev_valid_row(row):         # a row before `row` where EVERY interpolated column is valid (a superset halo)
            lo = max(0, row - 4096)
            idx = np.flatnonzero(valid_near(row, lo, row)) if row > lo else []
            return lo + int(idx[-1]) if len(idx) else 0

        def next_valid_row(row):
            hi = min(n_total, row + 4096)
            idx = np.flatnonzero(valid_near(row, row, hi)) if hi > row else []
            return row + int(idx[0]) + 1 if len(idx) else n_total
        shards = P.plan_interpolate(n_total, T0, t_last, interval, offset, world, lb, prev_valid_row, next_valid_row)
    else:
        shards = P.plan(n_total, T0, t_last, interval, offset, world, lb)
    sh = shards[rank]
    row_first = sh.first_row
    rows = sh.halo_hi - row_first                 # rows this rank holds (own rows + halos)
    own_rows = sh.row_hi - sh.row_lo
    W = sh.num_windows
    s0_shard = s0 + sh.k_lo * interval

    stream = torch.cuda.Stream()
    sarr = N.make_specs(specs)
    nspecs = len(specs)
    with torch.cuda.stream(stream):
        ctx = N.Ctx(local_rank, stream=stream.cuda_stream)
        frame = N.Frame.generate(ctx, rows, ncols=ncols, row0=row_first, t0=T0, step=STEP, seed=wl["seed"],
                                 null_mask=wl["null_mask"], null_mod=wl["null_mod"] or 10)
        rolling = N.Rolling(frame, 0, interval, shard=(s0_shard, W))
        out_vals = [torch.empty(max(W, 1), dtype=torch.int64, device="cuda") for _ in specs]
        out_bits = [torch.empty((W + 7) // 8 + 16, dtype=torch.uint8, device="cuda") for _ in specs]
        outs = (N.OutCol * nspecs)()
        for j in range(nspecs):
            outs[j].values, outs[j].validity = out_vals[j].data_ptr(), out_bits[j].data_ptr()

        def step():
            if fused:
                rolling.interpolate_aggregate_device(wl["ops"], sarr, nspecs, outs)
            else:
                rolling.aggregate_device(sarr, nspecs, outs)

        warm = max(args.warmup, 3)
        for _ in range(warm):
            step()
        ctx.synchronize()
        ctx.enable_timing(2)   # CUDA events around every 4th launch of the dominant kernel, on its own stream
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        sampler = ClockSampler(local_rank if vis is None else int(vis.split(",")[local_rank]))
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        torch.cuda.synchronize()
        barrier()
        clocks = sampler.stop()
        ms_local = ev0.elapsed_time(ev1) / args.steps
        ms_step = max_over_ranks(ms_local)
        tm = ctx.last_timing()
        main_ms = tm.main_ms / max(1, tm.main_launches)
        launches = tm.launches
        ctx.enable_timing(0)
        ctx.synchronize()      # surfaces EUNSORTED & co
        value = n_total / (ms_step * 1e-3)

        # ---- roofline ------------------------------------------------------------------------------------------------
        # dominant kernel = one segreduce launch: reads time 8n + one value column 8n (+ n/8 validity), writes what that
        # launch produces per window: config 1 cnt/sum/min/max = 4 x 8W; config 2 (per column) the trapezoid sum + its
        # point count = 2 x 8W.  The whole step is measured against SURVEY 8(d)'s algorithmic bytes.
        peak, peak_src = measured_peak()
        per_launch_cols = 1
        kernel_writes = (4 if not fused else 2) * 8 * W
        alg_kernel = 16 * rows * per_launch_cols + ((rows + 7) // 8 if wl["null_mask"] else 0) + kernel_writes
        achieved = alg_kernel / (main_ms * 1e-3) / 1e9 if main_ms > 0 else 0.0
        alg_step = algorithmic_bytes(wl, own_rows, W)
        step_gbs = alg_step / (ms_local * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                t_ = json.load(open(tp)).get(f"config{args.config}_main_kernel_bytes_per_launch", {})
                traffic = t_.get(str(rows))
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "segreduce_kernel (main streaming kernel, one launch per value column)",
                    "kernel_ms": main_ms, "kernel_launches_timed": int(tm.main_launches),
                    "algorithmic_bytes_per_launch": alg_kernel, "peak_source": peak_src,
                    "frac_of_nominal_8TBs": achieved / 8000.0,
                    "step": {"algorithmic_bytes": alg_step, "achieved": step_gbs, "frac": step_gbs / peak,
                             "frac_of_nominal_8TBs": step_gbs / 8000.0,
                             "note": "whole step of rank 0 (all launches of one call), SURVEY 8(d) algorithmic bytes"}}

        # ---- e2e: the one-shot reference-facing call with HOST (pinned) buffers, pipelined on every rank --------------
        e2e = None
        single = None
        if not args.no_e2e:
            if not fused:       # the rank's whole shard (own rows + the halo row), as a shard of the global lattice
                e_lead, e_rows, e_W, e_shard = 0, rows, W, (s0_shard, W)
            else:
                # config 2 holds 40 B/row: the host buffers take a prefix of the rank's OWN rows that ends on a window
                # boundary (plus the row after it), run as a frame of its own (a row sits on every window start)
                e_lead = sh.lead_rows
                e_W = min(W, max(64, (min(own_rows, args.e2e_rows_config2) * STEP // interval) // 64 * 64))
                e_rows = min(rows - e_lead, lb(s0_shard + e_W * interval) - sh.row_lo + 1)
                e_shard = None
            hcols = [torch.empty(e_rows, dtype=torch.int64).pin_memory()]
            hcols += [torch.empty(e_rows, dtype=torch.float64).pin_memory() for _ in range(ncols)]
            hbits = [None] + [torch.empty((e_rows + 7) // 8 + 8, dtype=torch.uint8).pin_memory() if (wl["null_mask"] >> c) & 1 else None
                              for c in range(ncols)]
            dl = (N.OutCol * (ncols + 1))()
            for j in range(ncols + 1):
                dl[j].values = hcols[j].data_ptr()
                dl[j].validity = hbits[j].data_ptr() if hbits[j] is not None else None
            ctx.check(N.lib().bowgpu_frame_download_range(frame.h, e_lead, e_rows, dl, ncols + 1))
            e_cap = e_W + 2      # (a frame of its own counts the window of its last row too)
            h_out_v = [torch.empty(e_cap, dtype=torch.int64).pin_memory() for _ in specs]
            h_out_b = [torch.empty((e_cap + 7) // 8 + 16, dtype=torch.uint8).pin_memory() for _ in specs]
            houts = (N.OutCol * nspecs)()
            for j in range(nspecs):
                houts[j].values, houts[j].validity = h_out_v[j].data_ptr(), h_out_b[j].data_ptr()
            harr = (N.Col * (ncols + 1))()
            for j in range(ncols + 1):
                harr[j].values, harr[j].offset, harr[j].length = hcols[j].data_ptr(), 0, e_rows
                harr[j].validity = hbits[j].data_ptr() if hbits[j] is not None else None
                harr[j].null_count = -1 if hbits[j] is not None else 0
                harr[j].dtype = N.INT64 if j == 0 else N.FLOAT64
            opts, okeep = N.make_host_opts(shard=e_shard)
            got_w = C.c_int64()
            if fused:
                codes = (C.c_int32 * (ncols + 1))(*[N.INTERP[o] for o in wl["ops"]])

            def e2e_step():
                if fused:
                    ctx.check(N.lib().bowgpu_interpolate_aggregate_host(ctx.h, harr, ncols + 1, 0, interval, offset, None, codes,
                                                                        ncols + 1, sarr, nspecs, houts, e_cap, C.byref(got_w),
                                                                        C.byref(opts)))
                else:
                    ctx.check(N.lib().bowgpu_aggregate_host_ex(ctx.h, harr, ncols + 1, 0, interval, offset, 0, sarr, nspecs, houts,
                                                               e_cap, C.byref(got_w), C.byref(opts)))

            e2e_steps = max(1, min(args.steps, args.e2e_steps))
            for _ in range(2):
                e2e_step()
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            dt_local = (time.perf_counter() - t0) / e2e_steps
            dt = max_over_ranks(dt_local)
            barrier()
            # the host results of the last e2e step must equal the device-resident ones
            # (sums are split at different row positions by the chunked call: equal within the 1e-12 tolerance class)
            exact_ops = ("WindowStart", "Min", "Max", "Count", "First", "Last")
            n_cmp = e_W if not fused else (e_W - 1 if sh.lead_rows == 0 and sh.k_lo == 0 else 0)   # (a prefix run on its own
            #   sees neither the rows before it nor its cut last window: compared on the first shard only)

            def scale_of(op):       # sum |terms| of a window: values in [0, 1)
                return {"Sum": interval / STEP, "IntegralStep": float(interval), "IntegralTrapezoid": float(interval)}.get(op, 1.0)
            for j, (op, _) in enumerate(specs):
                a, b = h_out_v[j][:n_cmp], out_vals[j][:n_cmp].cpu()
                vb = torch.from_numpy(N.unpack_bits(out_bits[j][:(n_cmp + 7) // 8].cpu().numpy(), n_cmp))
                va = torch.from_numpy(N.unpack_bits(h_out_b[j][:(n_cmp + 7) // 8].numpy(), n_cmp))
                assert torch.equal(va, vb), f"e2e validity of {op} differs"
                if op in exact_ops:
                    assert torch.equal(a[vb], b[vb]), f"e2e output {op} differs"
                else:
                    fa, fb = a.view(torch.float64), b.view(torch.float64)
                    ok = (fa - fb).abs() <= 1e-12 * torch.clamp(fb.abs(), min=scale_of(op))
                    assert bool(ok[vb].all()), f"e2e output {op} differs"
            e_rows_total = sum_over_ranks(min(own_rows, e_rows))
            h2d = sum(int(t.numel() * t.element_size()) for t in hcols) + sum(int((e_rows + 7) // 8) for b in hbits if b is not None)
            e2e = {"value": e_rows_total / dt, "unit": "rows/s", "ms_per_step": dt * 1e3, "steps": e2e_steps,
                   "rows_per_step": int(e_rows_total),
                   "h2d_bytes_per_step": int(sum_over_ranks(h2d)),
                   "d2h_bytes_per_step": int(sum_over_ranks(nspecs * (8 * e_W + (e_W + 7) // 8))),
                   "h2d_GBs_per_gpu": h2d / dt_local / 1e9,
                   "host_memory": "pinned, allocated after binding the process to the CPUs next to its GPU",
                   "numa": numa,
                   "call": ("bowgpu_interpolate_aggregate_host" if fused else "bowgpu_aggregate_host_ex") +
                           " (one pipelined call per rank on its shard)"}
            if fused:
                e2e["note"] = f"host buffers hold the first {e_rows} own rows of every shard (40 B/row of pinned memory), run as a frame of their own"

            # ---- the number a Go caller sees: PAGEABLE host memory (staged through pinned chunks), N = 1 ---------------
            if world == 1 and not fused and not args.no_pageable:
                p_rows = min(e_rows, args.pageable_rows)
                p_W = P.num_windows(T0, T0 + (p_rows - 1) * STEP, interval, off_n)
                pcols = [np.array(hcols[j][:p_rows].numpy(), copy=True) for j in range(ncols + 1)]    # plain malloc'ed memory
                parr = (N.Col * (ncols + 1))()
                for j in range(ncols + 1):
                    parr[j].values, parr[j].validity, parr[j].offset = pcols[j].ctypes.data, None, 0
                    parr[j].length, parr[j].null_count, parr[j].dtype = p_rows, 0, N.INT64 if j == 0 else N.FLOAT64
                pv = [np.empty(max(p_W, 1), dtype=np.int64) for _ in specs]
                pb = [np.empty((p_W + 7) // 8 + 16, dtype=np.uint8) for _ in specs]
                pouts = (N.OutCol * nspecs)()
                for j in range(nspecs):
                    pouts[j].values, pouts[j].validity = pv[j].ctypes.data, pb[j].ctypes.data

                def p_step():
                    ctx.check(N.lib().bowgpu_aggregate_host_ex(ctx.h, parr, ncols + 1, 0, interval, offset, 0, sarr, nspecs, pouts,
                                                               p_W, C.byref(got_w), None))
                p_step()
                t0 = time.perf_counter()
                for _ in range(3):
                    p_step()
                pdt = (time.perf_counter() - t0) / 3
                e2e["pageable"] = {"value": p_rows / pdt, "unit": "rows/s", "rows": p_rows, "ms_per_step": pdt * 1e3,
                                   "h2d_GBs": 16 * p_rows / pdt / 1e9,
                                   "note": "malloc'ed host buffers (what a Go caller hands over): staged through pinned chunks"}
                del pcols

            # ---- ONE process driving every GPU of the box through the C ABI (what a Go program would call) ---------------
            if world > 1 and not fused and not args.no_single_process:
                barrier()
                if rank == 0:
                    try:
                        sp_rows = min(e_rows, args.single_process_rows)
                        sp_W = P.num_windows(T0, T0 + (sp_rows - 1) * STEP, interval, off_n)
                        devs = list(range(world))
                        res = {}
                        for tag, dl_ in (("one_gpu", [local_rank]), ("all_gpus", devs)):
                            o2, k2 = N.make_host_opts(devices=dl_)
                            harr2 = (N.Col * (ncols + 1))()
                            for j in range(ncols + 1):
                                harr2[j] = harr[j]
                                harr2[j].length = sp_rows
                            for _ in range(2):
                                ctx.check(N.lib().bowgpu_aggregate_host_ex(ctx.h, harr2, ncols + 1, 0, interval, offset, 0, sarr, nspecs,
                                                                           houts, sp_W, C.byref(got_w), C.byref(o2)))
                            t0 = time.perf_counter()
                            ctx.check(N.lib().bowgpu_aggregate_host_ex(ctx.h, harr2, ncols + 1, 0, interval, offset, 0, sarr, nspecs,
                                                                       houts, sp_W, C.byref(got_w), C.byref(o2)))
                            res[tag] = time.perf_counter() - t0
                            res[tag + "_check"] = [int(h_out_v[j][:sp_W].sum().item()) for j in (0, 5)]
                        single = {"devices": world, "rows": sp_rows, "ms_one_gpu": res["one_gpu"] * 1e3,
                                  "ms_all_gpus": res["all_gpus"] * 1e3, "rows_per_s_all_gpus": sp_rows / res["all_gpus"],
                                  "same_results": res["one_gpu_check"] == res["all_gpus_check"],
                                  "call": "bowgpu_aggregate_host_ex(opts.devices = every GPU), one process; the other ranks idle"}
                    except Exception as e:  # noqa: BLE001
                        single = {"error": str(e)[:200]}
                barrier()

        # ---- CPU baseline: the oracle port on this box's host cores (rank 0, N == 1 only) -----------------------
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu:
            from oracle import refc as R
            R.build()
            sample = min(rows, args.cpu_rows // max(1, ncols))
            rf = R.Frame(host_frame(wl, row_first, sample))
            t0 = time.perf_counter()
            ref_out = oracle_step(R, wl, rf)
            cdt = time.perf_counter() - t0
            cpu = {"value": sample / cdt, "unit": "rows/s", "cores": 1, "kind": "port",
                   "sample": f"first {sample} rows of the workload, one pass ({cdt:.2f} s); C port of the reference algorithm "
                             "(oracle/ref.c): the Go reference cannot be built here (no Go toolchain), GOMAXPROCS n/a",
                   "host_cpus": os.cpu_count()}
            # cheap end-to-end parity check of the benchmarked configuration against the oracle
            Wc = len(ref_out[0][0]) - 2          # the last sampled windows are cut / lack their inclusive row
            for j, (op, _) in enumerate(specs):
                g = out_vals[j][:Wc].cpu().numpy()
                w, wm = ref_out[j][0][:Wc], ref_out[j][1][:Wc]
                if op in ("WindowStart", "Min", "Max", "Count", "First", "Last"):
                    assert np.array_equal(g[wm], w.view(np.int64)[wm]), op
                else:
                    scale = {"Sum": interval / STEP, "IntegralStep": float(interval), "IntegralTrapezoid": float(interval)}.get(op, 1.0)
                    assert np.all(np.abs(g.view(np.float64)[wm] - w[wm]) <= 1e-12 * np.maximum(np.abs(w[wm]), scale)), op

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "rows_per_gpu": rows, "windows_per_gpu": W,
        }
        if single is not None:
            line["single_process_multi_gpu"] = single
        print(json.dumps(line), flush=True)
    rolling.close()
    frame.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2])
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="rows of the GLOBAL frame (strong scaling)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-rows-config2", type=int, default=250_000_000)
    ap.add_argument("--pageable-rows", type=int, default=200_000_000)
    ap.add_argument("--single-process-rows", type=int, default=250_000_000)
    ap.add_argument("--cpu-rows", type=int, default=100_000_000)
    ap.add_argument("--ref-rows", type=int, default=20_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pageable", action="store_true")
    ap.add_argument("--no-single-process", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
