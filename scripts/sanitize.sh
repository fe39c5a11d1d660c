#!/bin/bash
# compute-sanitizer over tests/kernel_driver.py (every kernel family, small inputs) -> gpurun_out/sanitizer_<tag>.txt
tag=${1:-r2}
out=gpurun_out/sanitizer_$tag.txt
mkdir -p gpurun_out
: > $out
run() {  # <label> <env...> -- <tool args...>
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "== $label: env ${envs[*]} compute-sanitizer $*" >> $out
  env "${envs[@]}" timeout 1500 compute-sanitizer "$@" python tests/kernel_driver.py 2>&1 | grep -E "sanitize driver|ERROR SUMMARY|RACECHECK SUMMARY|=========     at |========= Invalid|========= Uninit|========= Error|hazard|Traceback|Error|assert" | head -40 >> $out
}
run "memcheck (default path)" SAN_ROWS=120000 -- --tool memcheck
run "memcheck (side-by-side lanes)" SAN_ROWS=120000 BOWGPU_SEG_SIDE=3 -- --tool memcheck
run "memcheck (segmc experimental path)" SAN_ROWS=120000 BOWGPU_SEG_IMPL=mc -- --tool memcheck
run "initcheck" SAN_ROWS=60000 -- --tool initcheck
run "synccheck" SAN_ROWS=60000 -- --tool synccheck
run "racecheck" SAN_ROWS=30000 -- --tool racecheck
cat $out
