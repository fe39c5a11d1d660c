#!/bin/bash
# A/B of library builds (bow_b200/libbowgpu_<tag>.so next to the default one) on scripts/bench_configs.py sections
# usage: scripts/ab_configs.sh "<sections>" <tag> [<tag> ...]
secs=$1; shift
for tag in "" "$@"; do
  lib=""; [ -n "$tag" ] && lib=$PWD/bow_b200/libbowgpu_$tag.so
  echo "== ${tag:-default}"
  BOWGPU_LIB=$lib python scripts/bench_configs.py $secs 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try:
        d=json.loads(l); print('   %-88s %8.3f ms' % (d['config'][:88], d['ms']))
    except Exception: pass"
done
