#!/usr/bin/env python
"""bench.py — input rows/s of IntervalRolling.Aggregate on B200 (BASELINE.json metric).

Workload at every N (per GPU, weak scaling): BASELINE.json configs[1] — synthetic 100M rows, int64 ns
time with a regular 1 s step + one float64 column, 1-minute windows, WindowStart(time) +
ArithmeticMean / Sum / Min / Max / Count (value).  For N > 1 the global N*100M-row series is
range-partitioned by window index (bow_b200/partition.py), one process per GPU, no collective on
the data path; per-shard outputs are the concatenation.

One JSON line on stdout (rank 0):
  value     rows/s, inputs resident in HBM (CUDA events around K calls of the C-ABI aggregate, max over ranks)
  e2e       rows/s through the C ABI with HOST (pinned) buffers: H2D + kernels + D2H inside the timed region
  roofline  dominant kernel (segreduce main): algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (C port of the reference algorithm, 1 core) timed on this box's host CPU

`--impl reference` times the reference's CPU algorithm (oracle/ref.c port; the Go reference cannot be
built in this image — no Go toolchain) on the same workload shape.
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

INTERVAL = 60_000_000_000          # 1-minute windows on ns timestamps
T0 = 1_700_000_000_000_000_000
STEP = 1_000_000_000
SEED = 42
AGGS = ["WindowStart", "ArithmeticMean", "Sum", "Min", "Max", "Count"]
METRIC = "input rows/s for IntervalRolling.Aggregate"


def specs_for():
    return [(a, 0 if a == "WindowStart" else 1) for a in AGGS]


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (C port, single thread like the single-goroutine Go
    path) on bounded samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bow_b200 import synth
    from oracle import refc as R
    R.build()
    sample = min(args.rows, args.ref_rows)
    cols = synth.regular_frame(0, sample, 1, SEED, T0, STEP)
    fr = R.Frame(cols)
    specs = specs_for()

    def step():
        r = R.RefRolling(fr, 0, INTERVAL)
        return r.aggregate(specs)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.rows), "sample_rows_per_step": sample,
                   "note": "Go reference not buildable here (no Go toolchain): C port of its algorithm, "
                           "oracle/ref.c, single thread like the single-goroutine reference path"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": 1, "kind": "port",
                         "sample": f"first {sample} rows of the workload per step"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(rows):
    return (f"configs[1]: synthetic {rows} rows/GPU, int64 ns time (regular 1 s step) + 1 float64 col, 1-min windows, "
            "WindowStart + mean/sum/min/max/count")


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bow_b200 import native as N
    from bow_b200 import partition as P

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bow_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- range partition of the global series (no collective on the data path) ----------------------------
    n_total = args.rows * world
    t_last = T0 + (n_total - 1) * STEP
    shards = P.plan(n_total, T0, t_last, INTERVAL, 0, world, P.regular_lower_bound(T0, STEP, n_total))
    sh = shards[rank]
    s0 = P.first_window_start(T0, INTERVAL, 0)
    rows = sh.halo_hi - sh.row_lo
    W = sh.num_windows

    stream = torch.cuda.Stream()
    specs = specs_for()
    sarr = N.make_specs(specs)
    with torch.cuda.stream(stream):
        ctx = N.Ctx(local_rank, stream=stream.cuda_stream)
        frame = N.Frame.generate(ctx, rows, ncols=1, row0=sh.row_lo, t0=T0, step=STEP, seed=SEED)
        rolling = N.Rolling(frame, 0, INTERVAL, shard=(s0 + sh.k_lo * INTERVAL, W))
        out_vals = [torch.empty(max(W, 1), dtype=torch.int64, device="cuda") for _ in specs]
        out_bits = [torch.empty((W + 7) // 8 + 16, dtype=torch.uint8, device="cuda") for _ in specs]
        outs = (N.OutCol * len(specs))()
        for j in range(len(specs)):
            outs[j].values, outs[j].validity = out_vals[j].data_ptr(), out_bits[j].data_ptr()

        def step():
            rolling.aggregate_device(sarr, len(specs), outs)

        for _ in range(max(args.warmup, 3)):
            step()
        ctx.synchronize()
        ctx.enable_timing(2)   # CUDA events around every 4th launch of the dominant kernel, on its own stream
        sampler = ClockSampler(local_rank if os.environ.get("CUDA_VISIBLE_DEVICES") is None else
                               int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
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
        ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
        tm = ctx.last_timing()
        main_ms = tm.main_ms / max(1, tm.main_launches)
        launches = tm.launches
        ctx.enable_timing(0)
        ctx.synchronize()      # surfaces EUNSORTED & co
        value = n_total / (ms_step * 1e-3)

        # ---- roofline of the dominant kernel (segreduce main): algorithmic bytes per launch ---------------
        # reads: time 8n + value 8n; writes: cnt, sum, mean, min, max = 5 * 8W   ("algorithmic bytes")
        alg_bytes = 16 * rows + 5 * 8 * W
        peak, peak_src = measured_peak()
        achieved = alg_bytes / (main_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("segreduce_basic_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "segreduce_basic_kernel", "kernel_ms": main_ms,
                    "kernel_launches_timed": int(tm.main_launches),
                    "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                    "frac_of_nominal_8TBs": achieved / 8000.0}

        # ---- e2e: the same call through the C ABI with HOST (pinned) buffers -----------------------------------
        e2e = None
        if not args.no_e2e:
            h_t = torch.empty(rows, dtype=torch.int64).pin_memory()
            h_v = torch.empty(rows, dtype=torch.float64).pin_memory()
            dl = (N.OutCol * 2)()
            dl[0].values, dl[1].values = h_t.data_ptr(), h_v.data_ptr()
            ctx.check(N.lib().bowgpu_frame_download_range(frame.h, 0, rows, dl, 2))
            h_out_v = [torch.empty(max(W, 1), dtype=torch.int64).pin_memory() for _ in specs]
            h_out_b = [torch.empty((W + 7) // 8 + 16, dtype=torch.uint8).pin_memory() for _ in specs]
            houts = (N.OutCol * len(specs))()
            for j in range(len(specs)):
                houts[j].values, houts[j].validity = h_out_v[j].data_ptr(), h_out_b[j].data_ptr()
            harr = (N.Col * 2)()
            for j, (h, dt) in enumerate(((h_t, N.INT64), (h_v, N.FLOAT64))):
                harr[j].values, harr[j].validity, harr[j].offset = h.data_ptr(), None, 0
                harr[j].length, harr[j].null_count, harr[j].dtype = rows, 0, dt

            got_w = C.c_int64()

            def e2e_step():
                if world == 1:
                    # the one-shot reference-facing call: IntervalRolling(b, ...).Aggregate(...) from host Arrow buffers to
                    # host result buffers (chunks of the window range pipelined over worker contexts)
                    ctx.check(N.lib().bowgpu_aggregate_host(ctx.h, harr, 2, 0, INTERVAL, 0, 0, sarr, len(specs), houts, W,
                                                            C.byref(got_w)))
                    return
                fr = N.Frame.from_col_descs(ctx, harr, 2, N.MEM_HOST)          # H2D inside
                r = N.Rolling(fr, 0, INTERVAL, shard=(s0 + sh.k_lo * INTERVAL, W))
                ctx.check(N.lib().bowgpu_rolling_aggregate(r.h, sarr, len(specs), houts, N.MEM_HOST))  # D2H inside
                r.close()
                fr.close()

            e2e_steps = max(1, min(args.steps, args.e2e_steps))
            for _ in range(2):
                e2e_step()
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            dt = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
            barrier()
            # the host results of the last e2e step must equal the device-resident ones
            # (sums are split at different row positions by the chunked call: equal within the 1e-12 tolerance class)
            for j in range(len(specs)):
                a, b = h_out_v[j][:W], out_vals[j][:W].cpu()
                if AGGS[j] in ("ArithmeticMean", "Sum"):
                    assert torch.allclose(a.view(torch.float64), b.view(torch.float64), rtol=1e-12, atol=0), \
                        f"e2e output {AGGS[j]} differs"
                else:
                    assert torch.equal(a, b), f"e2e output {AGGS[j]} differs"
            e2e = {"value": n_total / dt, "unit": "rows/s", "ms_per_step": dt * 1e3, "steps": e2e_steps,
                   "h2d_bytes_per_step": int(sum_over_ranks(16 * rows)),
                   "d2h_bytes_per_step": int(sum_over_ranks(len(specs) * (8 * W + (W + 7) // 8))),
                   "host_memory": "pinned",
                   "call": "bowgpu_aggregate_host" if world == 1 else
                           "bowgpu_frame_create + bowgpu_rolling_create_shard + bowgpu_rolling_aggregate"}

        # ---- CPU baseline: the oracle port on this box's host cores (rank 0, N == 1 only) -----------------------
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu:
            from oracle import refc as R
            R.build()
            sample = min(rows, args.cpu_rows)
            if not args.no_e2e:
                cols = [(h_t.numpy()[:sample], None), (h_v.numpy()[:sample], None)]
            else:
                from bow_b200 import synth
                cols = synth.regular_frame(0, sample, 1, SEED, T0, STEP)
            rf = R.Frame(cols)
            t0 = time.perf_counter()
            ref_out = R.RefRolling(rf, 0, INTERVAL).aggregate(specs)
            cdt = time.perf_counter() - t0
            cpu = {"value": sample / cdt, "unit": "rows/s", "cores": 1, "kind": "port",
                   "sample": f"first {sample} rows of the workload, one pass ({cdt:.2f} s)",
                   "host_cpus": os.cpu_count()}
            # cheap end-to-end parity check of the benchmarked configuration against the oracle
            Wc = len(ref_out[0][0]) - 1          # the last sampled window may be cut
            for j, name in enumerate(AGGS):
                g = out_vals[j][:Wc].cpu().numpy()
                w = ref_out[j][0][:Wc]
                if name in ("ArithmeticMean", "Sum"):
                    assert np.allclose(g.view(np.float64), w, rtol=1e-12, atol=0), name
                else:
                    assert np.array_equal(g, w.view(np.int64)), name

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.rows), "rows_per_gpu": args.rows, "windows_per_gpu": W,
                       "interval_ns": INTERVAL, "aggregations": AGGS, "parallelism": f"range-partition x{world}",
                       "l2": "inputs (1.6 GB/GPU) are larger than L2 (126 MB); no explicit flush"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    rolling.close()
    frame.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000, help="rows per GPU")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-rows", type=int, default=100_000_000)
    ap.add_argument("--ref-rows", type=int, default=20_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
