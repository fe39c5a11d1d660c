"""Two contexts on two GPUs inside ONE process (the Go shim's multi-GPU fan-out): every entry point
selects its device itself, kernel attributes are configured per device.  Skipped on single-GPU boxes."""
import numpy as np
import pytest

from bow_b200 import parallel as PP
from oracle import refc as R
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_two_contexts_one_process():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from bow_b200 import native as N
    rng = np.random.default_rng(3)
    n, interval = 60000, 37
    t = (np.cumsum(rng.integers(0, 5, size=n)) + 10).astype(np.int64)
    v = (rng.normal(size=n), rng.random(n) > 0.2)
    cols = [(t, None), v]
    specs = [("WindowStart", 0), ("Count", 1), ("Sum", 1), ("Min", 1), ("Max", 1), ("IntegralTrapezoid", 1)]
    shards, s0 = PP.plan_for_columns(t, interval, 0, 2)
    ctxs = [N.Ctx(0), N.Ctx(1)]
    outs = []
    for g, sh in enumerate(shards):           # interleaved use of both devices from one thread
        fr = N.Frame.from_numpy(ctxs[g], PP.slice_cols(cols, sh.row_lo, sh.halo_hi))
        r = N.Rolling(fr, 0, interval, inclusive=True, shard=(s0 + sh.k_lo * interval, sh.num_windows))
        outs.append(r.aggregate(specs))
    got = PP.concat_outputs(outs)
    want = R.RefRolling(R.Frame(cols), 0, interval, inclusive=True).aggregate(specs)
    scales = H.term_scales(cols, 1, interval, inclusive=True)
    for sp, (gv, gm), (wv, wm) in zip(specs, got, want):
        assert np.array_equal(gm, wm), sp
        if sp[0] in ("Sum", "IntegralTrapezoid"):
            H.assert_in_tolerance_class(gv, wv, scales[sp[0]], str(sp))
        else:
            assert np.array_equal(gv[gm].view(np.int64), wv[wm].view(np.int64)), sp
    for c in ctxs:
        c.close()
