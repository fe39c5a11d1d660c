#!/bin/bash
# Run under gpurun (one GPU).  Produces in gpurun_out/:
#   launches_<tag>.csv  every kernel launch of a short bench run with its device time
#   prof_<tag>.ncu-rep  one full capture of the dominant kernel (segreduce main)
tag=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:segreduce_kernel -s 3 -c 1 -o gpurun_out/prof_${tag} \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2_${tag}.log 2>&1
ls -la gpurun_out
