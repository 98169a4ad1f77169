"""Backs the nltk / rapidfuzz shims.  Loaded by file path so that it does not depend on the
product package being importable under the same name as the reference's package."""
import ctypes
import functools
import importlib.util
import pathlib

_ROOT = pathlib.Path(__file__).resolve().parents[2]


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_port = _load("_nsm_oracle_port", _ROOT / "oracle" / "reference_port.py")
_tok_src = (_ROOT / "napkon-string-matching_b200" / "napkon_string_matching" / "text" /
            "tokenize.py").read_text(encoding="utf-8")
_ns: dict = {"__name__": "_nsm_tok"}
exec(compile(_tok_src, "tokenize.py", "exec"), _ns)  # noqa: S102 - our own file
GERMAN_STOP_WORDS = list(_ns["_GERMAN_STOP_WORDS"])
word_tokenize = _ns["word_tokenize"]

# rapidfuzz's LCS is compiled code: when the C oracle has been built, the shim counts the LCS with
# its bit-parallel routine (oracle/nsm_oracle.c:ora_lcs_utf32) so that a timing of the reference
# is not dominated by a Python dynamic programme; the float arithmetic below is the same either way
_c_lcs = None
_lib_path = _ROOT / "oracle" / "_build" / "libnsm_oracle.so"
if _lib_path.exists():
    try:
        _lib = ctypes.CDLL(str(_lib_path))
        _lib.ora_lcs_utf32.restype = ctypes.c_uint32
        _lib.ora_lcs_utf32.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32]
        _c_lcs = _lib.ora_lcs_utf32
    except (OSError, AttributeError):
        _c_lcs = None


def lcs_length(a: str, b: str) -> int:
    if _c_lcs is not None:
        return _c_lcs(a.encode("utf-32-le"), len(a), b.encode("utf-32-le"), len(b))
    return _port.lcs_length(a, b)


@functools.lru_cache(maxsize=1 << 16)
def _processed(s):
    """default_process once per distinct string (rapidfuzz does this in C++ on every call; a
    Python regex per call would make the shim, not the reference's loop, the thing timed)."""
    a = _port.default_process(s)
    return a, a.encode("utf-32-le"), len(a)


def qratio_percent(s1, s2):
    (a, a32, na), (b, b32, nb) = _processed(s1), _processed(s2)
    if not a or not b:
        return 0
    lcs = _c_lcs(a32, na, b32, nb) if _c_lcs is not None else _port.lcs_length(a, b)
    lensum = na + nb
    norm_dist = (lensum - 2 * lcs) / lensum if lensum else 0.0
    return (1.0 - norm_dist) * 100
