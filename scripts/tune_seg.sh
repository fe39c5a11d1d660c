#!/bin/bash
# Sweep of the segreduce tile shape / residency on one B200 (run under gpurun).  Variants are separate builds
# (make BUILD=build_<tag> OUT=../libbowgpu_<tag>.so EXTRA="-DSEG_CFG_NT=.. -DSEG_CFG_R=.. -DSEG_CFG_CTAS=..").
# Each line of CFGS: "<lib tag or -> <stages> <ctas>"
CFGS=${CFGS:-$(cat scripts/tune_cfgs.txt)}
echo "$CFGS" | while read tag st ct; do
  [ -z "$tag" ] && continue
  lib=""; [ "$tag" != "-" ] && lib=$PWD/bow_b200/libbowgpu_$tag.so
  echo -n "lib=$tag stages=$st ctas=$ct: "
  BOWGPU_LIB=$lib BOWGPU_SEG_STAGES=$st BOWGPU_SEG_CTAS=$ct python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('value %.3e rows/s  ms_step %.4f  kernel_ms %.4f  GB/s %.0f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['roofline']['frac']))
except Exception as e:
    print('FAILED', e)"
done
