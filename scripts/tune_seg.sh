#!/bin/bash
# sweep of the segreduce launch knobs on one B200 (run under gpurun)
for cfg in "3 2" "2 3" "2 2" "1 4" "2 4"; do
  set -- $cfg
  echo "stages=$1 ctas=$2"
  BOWGPU_SEG_STAGES=$1 BOWGPU_SEG_CTAS=$2 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.3e rows/s  ms_step %.4f  kernel_ms %.4f  GB/s %.0f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['achieved'], d['roofline']['frac']))"
done
