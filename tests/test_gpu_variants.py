"""The launch-shape knobs are read once per process, so each variant runs tests/kernel_driver.py (every kernel family on
small inputs, results checked against the oracle inside the driver) in a process of its own:
  BOWGPU_SEG_MERGE=0      basic and integral aggregations of a column as two launches (default: one merged launch)
  BOWGPU_BOUNDS_SEARCH=1  window boundaries of the fused path by binary search at every size / =0 never
  BOWGPU_SEG_SIDE=3       column launches side by side on three streams
  BOWGPU_SEG_IMPL=mc      the experimental multi-column kernels"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"BOWGPU_SEG_MERGE": "0"}, {"BOWGPU_BOUNDS_SEARCH": "1"}, {"BOWGPU_BOUNDS_SEARCH": "0"},
                                 {"BOWGPU_SEG_SIDE": "3"}, {"BOWGPU_SEG_IMPL": "mc"}, {}],
                         ids=["split-families", "search-always", "search-never", "side-by-side", "segmc", "default"])
def test_variant_matches_the_oracle(env):
    e = dict(os.environ)
    e.update(env)
    e["SAN_ROWS"] = "150000"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "kernel_driver.py")], env=e, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "sanitize driver ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
