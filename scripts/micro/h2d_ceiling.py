#!/usr/bin/env python
"""Aggregate pinned host -> device bandwidth of the box with 1, 2, 4, 8 GPUs copying AT THE SAME TIME (one process, one
stream per GPU, plain cudaMemcpyAsync of 1 GiB pinned buffers; no library of this repo involved).  This is the ceiling of
every end-to-end number that starts in host memory (bench.py `e2e`).  Prints one JSON line per GPU subset."""
import json
import sys
import time

import torch

GIB = 1 << 30


def run(devs, reps=4):
    src = [torch.empty(GIB, dtype=torch.uint8).pin_memory() for _ in devs]
    dst = [torch.empty(GIB, dtype=torch.uint8, device=f"cuda:{d}") for d in devs]
    streams = [torch.cuda.Stream(device=f"cuda:{d}") for d in devs]

    def go():
        for s_, d_, st in zip(src, dst, streams):
            with torch.cuda.stream(st):
                d_.copy_(s_, non_blocking=True)
        for st in streams:
            st.synchronize()
    go()
    t0 = time.perf_counter()
    for _ in range(reps):
        go()
    dt = (time.perf_counter() - t0) / reps
    return len(devs) * GIB / dt / 1e9


def main():
    n = torch.cuda.device_count()
    subsets = [[0]]
    if n >= 2:
        subsets += [[0, 1]]
    if n >= 4:
        subsets += [[0, 1, 2, 3], [0, 2], [0, 4] if n >= 8 else [0, 3]]
    if n >= 8:
        subsets += [[4, 5, 6, 7], [0, 2, 4, 6], list(range(8))]
    for devs in subsets:
        gbs = run(devs)
        print(json.dumps({"gpus": devs, "aggregate_h2d_GBs": round(gbs, 1), "per_gpu_GBs": round(gbs / len(devs), 1)}), flush=True)


if __name__ == "__main__":
    sys.exit(main())
