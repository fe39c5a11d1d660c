#!/bin/bash
# A/B of sort.cu builds (make BUILD=build_s<items>_<minb> OUT=../libbowgpu_s<items>_<minb>.so EXTRA="-DSORT_CFG_ITEMS=.. -DSORT_CFG_MINB=..")
for lib in "" $(ls bow_b200/libbowgpu_s*.so 2>/dev/null); do
  echo "== ${lib:-default}"
  BOWGPU_LIB=${lib:+$PWD/$lib} BOW_BENCH_SCALE=${BOW_BENCH_SCALE:-1} python scripts/bench_configs.py sort 2>&1 | grep SortByCol | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   %-70s %.2f ms' % (d['config'][:70], d['ms']))"
done
