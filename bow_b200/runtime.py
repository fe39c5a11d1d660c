"""Process-wide default device context of the host-side mirror (one ctx per process, i.e. per GPU:
the deployment model is one process per GPU)."""
from __future__ import annotations

import os
from typing import Optional

from . import native as N

_ctx: Optional[N.Ctx] = None


def default_ctx() -> N.Ctx:
    """Lazily creates the context on device $BOWGPU_DEVICE (default: $LOCAL_RANK, else 0).
    Raises when no B200 / no library is available — there is no CPU fallback."""
    global _ctx
    if _ctx is None:
        dev = int(os.environ.get("BOWGPU_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _ctx = N.Ctx(dev)
    return _ctx


def set_default_ctx(ctx: Optional[N.Ctx]) -> None:
    global _ctx
    _ctx = ctx
