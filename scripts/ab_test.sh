#!/bin/bash
# A/B of two library builds on one box: bounds + segreduce kernel times from an ncu launch list, and the bench line
for tag in "" halo2; do
  lib=""; [ -n "$tag" ] && lib=$PWD/bow_b200/libbowgpu_$tag.so
  echo "== lib=${tag:-default}"
  BOWGPU_LIB=$lib BOW_BENCH_SCALE=0.2 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/ab_$tag.csv python scripts/bench_configs.py 2 > /dev/null 2>&1
  python scripts/ncu_list.py gpurun_out/ab_$tag.csv 9 | grep -E "bounds_kernel|segreduce|gather_kernel" | head -4
  BOWGPU_LIB=$lib python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.3e rows/s  ms_step %.4f  kernel_ms %.4f  GB/s %.0f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['roofline']['frac']))"
done
