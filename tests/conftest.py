import json
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
PKG = ROOT / "napkon-string-matching_b200"
for p in (str(PKG), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    """(meta, inputs, arrays) of one reference-generated fixture."""
    meta = json.loads((GOLDEN / "index.json").read_text())[name]
    inputs = json.loads((GOLDEN / f"{name}.inputs.json").read_text(encoding="utf-8"))
    arrays = dict(np.load(GOLDEN / f"{name}.npz", allow_pickle=False))
    return meta, inputs, arrays


def triples(left, right, score):
    """Canonical, order-free form of a result: rows sorted by (left, right)."""
    left = np.asarray(left, dtype=np.int64)
    right = np.asarray(right, dtype=np.int64)
    score = np.asarray(score, dtype=np.float64)
    order = np.lexsort((right, left))
    return left[order], right[order], score[order]


def assert_same_triples(got, want):
    gl, gr, gs = triples(*got)
    wl, wr, ws = triples(*want)
    assert len(gl) == len(wl), f"{len(gl)} pairs, expected {len(wl)}"
    assert np.array_equal(gl, wl) and np.array_equal(gr, wr), "different pair sets"
    # bit-exact float64 scores
    assert np.array_equal(gs.view(np.uint64), ws.view(np.uint64)), "scores differ bitwise"


@pytest.fixture(scope="session")
def engine():
    from napkon_string_matching.gpu.engine import Engine

    return Engine()
