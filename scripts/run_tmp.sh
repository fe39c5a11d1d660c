python - <<'PY'
import sys, time, os
sys.path.insert(0, '.')
from bow_b200 import native as N
ctx = N.Ctx(0)
n = 1_000_000_000
fr = N.Frame.generate(ctx, n, ncols=1, seed=7, null_mask=1, null_mod=10)
r = N.Rolling(fr, 0, 900_000_000_000, offset=420_000_000_000)
for _ in range(2): r.bounds()
ctx.synchronize()
ctx.enable_timing(1)
t0 = time.perf_counter()
for _ in range(5): r.bounds()
ctx.synchronize()
dt = (time.perf_counter() - t0) / 5
tm = ctx.last_timing()
print("bounds v1=%s: wall %.3f ms per call, kernel main_ms %.3f total_ms %.3f" % (os.environ.get("BOWGPU_BOUNDS_V1"), dt * 1e3, tm.main_ms, tm.total_ms))
PY
