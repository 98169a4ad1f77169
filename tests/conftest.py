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
    # the suite needs the built artefacts (they are git-ignored): the CUDA library for the C-ABI,
    # decoder and sort tests, the C oracle as the checker.  Build them when a fresh checkout lacks them.
    lib = PKG / "napkon_string_matching" / "gpu" / "libnsm_b200.so"
    oracle = ROOT / "oracle" / "_build" / "libnsm_oracle.so"
    if not lib.exists() or not oracle.exists():
        import importlib.util

        spec = importlib.util.spec_from_file_location("graft_entry_for_tests", ROOT / "__graft_entry__.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()


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


class OracleEngine:
    """Test double for gpu.engine.Engine in the CPU suite: same surface, but pairs are scored by
    the C oracle.  It lets the host logic around the kernels (filters, exceptions, frames,
    sharding) be tested without a GPU.  Never used by the product."""

    def __init__(self):
        self.last_info = {}
        self.launches = 0

    class _Cohort:
        def __init__(self, packed):
            import numpy as _np

            self.packed = packed
            self.kind = "sets" if hasattr(packed, "tok") else "strings"
            self.n_items = packed.n_items
            self.weights = _np.ones(packed.n_items)

    def upload(self, packed, pinned=None):
        return self._Cohort(packed)

    def upload_masks(self, masks):
        return masks

    def all_pairs(self, left, right, threshold, *, flat=False, rows=None, l_cat=None, r_cat=None,
                  cat_mode=0, **_):
        from oracle import c_oracle
        from napkon_string_matching.gpu.lib import PAIR_DTYPE

        begin, end = rows if rows is not None else (0, left.n_items)
        out, flags = c_oracle.all_pairs(left.packed, right.packed, threshold, flat=flat,
                                        l_begin=begin, l_end=end, l_cat=l_cat, r_cat=r_cat,
                                        cat_mode=cat_mode)
        self.last_info = {"count": len(out), "flags": flags}
        return out.astype(PAIR_DTYPE)


@pytest.fixture
def oracle_engine(monkeypatch):
    """Routes the host API through OracleEngine for the duration of one CPU test."""
    import napkon_string_matching.gpu.engine as eng_mod

    double = OracleEngine()
    monkeypatch.setattr(eng_mod, "default_engine", lambda: double)
    return double


@pytest.fixture
def cuda_engine(monkeypatch, engine):
    import napkon_string_matching.gpu.engine as eng_mod

    monkeypatch.setattr(eng_mod, "default_engine", lambda: engine)
    return engine
