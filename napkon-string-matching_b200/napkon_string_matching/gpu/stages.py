"""Wall-clock stage timers of one ``compare()`` call (host side).  Off unless a caller opens
:func:`collect`; then every ``with stage(name)`` block adds its seconds to the collected dict.
Used by ``tools/cold_compare.py`` (the cold, one-shot timing through the drop-in API)."""
from __future__ import annotations

import contextlib
import time
from typing import Dict, Optional

_active: Optional[Dict[str, float]] = None


@contextlib.contextmanager
def collect():
    global _active
    previous, _active = _active, {}
    try:
        yield _active
    finally:
        _active = previous


@contextlib.contextmanager
def stage(name: str):
    if _active is None:
        yield
        return
    t0 = time.perf_counter()
    try:
        yield
    finally:
        _active[name] = _active.get(name, 0.0) + time.perf_counter() - t0
