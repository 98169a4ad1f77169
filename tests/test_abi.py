"""CPU suite: the C-ABI library is built, loads, and exports what include/nsm.h declares."""
import ctypes
import re

import pytest

from conftest import ROOT


def declared_functions():
    text = (ROOT / "include" / "nsm.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nsm_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_entry_points():
    names = declared_functions()
    for must in ("nsm_jaccard_allpairs", "nsm_qratio_allpairs", "nsm_version", "nsm_last_error"):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from napkon_string_matching.gpu import lib as nsmlib

    lib = nsmlib.load()
    assert lib.nsm_version() == 100
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert set(nsmlib.EXPORTS) == set(declared_functions())


def test_struct_layouts_match_the_header():
    from napkon_string_matching.gpu import lib as nsmlib

    assert ctypes.sizeof(nsmlib.NsmSets) == 12 * 8 + 8 * 4
    assert ctypes.sizeof(nsmlib.NsmStrings) == 5 * 8 + 6 * 4 + 8 * 4
    assert nsmlib.NsmJob.threshold.offset == 16
    assert nsmlib.NsmJob.out_pairs.offset == 40
    assert ctypes.sizeof(nsmlib.NsmJob) == 120 and nsmlib.NsmJob.out_dict.offset == 88
    assert nsmlib.CPACKET_DTYPE.itemsize == 256
    assert nsmlib.NsmJob.out_mode.offset == 80
    assert nsmlib.PACKET_DTYPE.itemsize == 496 and nsmlib.PACKET_DTYPE.fields["local"][1] == 400
    assert ctypes.sizeof(nsmlib.NsmRawSets) == 4 * 8 + 6 * 4
    assert ctypes.sizeof(nsmlib.NsmRawStrings) == 3 * 8 + 4 * 4 and nsmlib.NsmRawStrings.blank_sym.offset == 36
    assert nsmlib.PAIR_DTYPE.itemsize == 16
    assert nsmlib.PAIR_DTYPE.fields["score"][1] == 8


def test_no_gpu_means_loud_failure_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from napkon_string_matching.gpu import lib as nsmlib
    from napkon_string_matching.gpu.engine import Engine

    with pytest.raises(nsmlib.NsmError):
        Engine()
