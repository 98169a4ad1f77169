"""TEST INFRASTRUCTURE — ctypes front end of oracle/nsm_oracle.c (see its header)."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = _HERE / "_build" / "libnsm_oracle.so"

PAIR_DTYPE = np.dtype([("left", np.uint32), ("right", np.uint32), ("score", np.float64)])
JACCARD, QRATIO = 0, 1
FLAG_ZERO_UNION, FLAG_INDEX_ERROR = 1, 2


class _Side(C.Structure):
    _fields_ = [("item_level_off", C.c_void_p), ("level_off", C.c_void_p),
                ("level_len", C.c_void_p), ("tok", C.c_void_p), ("chr", C.c_void_p),
                ("n_items", C.c_uint32)]


def _lib():
    if not _LIB.exists():
        subprocess.run(["sh", str(_HERE / "build.sh")], check=True, capture_output=True)
    lib = C.CDLL(str(_LIB))
    lib.ora_allpairs.restype = C.c_int64
    lib.ora_allpairs.argtypes = [C.c_int, C.c_int, C.POINTER(_Side), C.POINTER(_Side), C.c_uint32,
                                 C.c_uint32, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_int64, C.POINTER(C.c_uint32)]
    lib.ora_score_pairs.restype = None
    lib.ora_score_pairs.argtypes = [C.c_int, C.c_int, C.POINTER(_Side), C.POINTER(_Side),
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.POINTER(C.c_uint32)]
    return lib


def _side(p):
    """p is a PackedSets or PackedStrings (duck-typed; the oracle does not import the product)."""
    keep = []

    def ptr(a):
        a = np.ascontiguousarray(a)
        keep.append(a)
        return a.ctypes.data

    if hasattr(p, "tok"):
        s = _Side(ptr(p.item_level_off), ptr(p.level_tok_off), None, ptr(p.tok), None, p.n_items)
    else:
        s = _Side(ptr(p.item_level_off), ptr(p.level_chr_off), ptr(p.level_len), None, ptr(p.chr),
                  p.n_items)
    return s, keep


def all_pairs(left, right, threshold, flat=False, l_begin=0, l_end=None, l_cat=None, r_cat=None,
              cat_mode=0):
    """Returns (pairs[PAIR_DTYPE] in row-major order, flags)."""
    lib = _lib()
    func = JACCARD if hasattr(left, "tok") else QRATIO
    ls, k1 = _side(left)
    rs, k2 = _side(right)
    l_end = left.n_items if l_end is None else l_end
    lc = np.ascontiguousarray(l_cat, dtype=np.uint64) if l_cat is not None else None
    rc = np.ascontiguousarray(r_cat, dtype=np.uint64) if r_cat is not None else None
    flags = C.c_uint32(0)
    args = (func, int(flat), C.byref(ls), C.byref(rs), l_begin, l_end, float(threshold),
            lc.ctypes.data if lc is not None else None, rc.ctypes.data if rc is not None else None,
            cat_mode)
    n = lib.ora_allpairs(*args, None, 0, C.byref(flags))
    out = np.zeros(n, dtype=PAIR_DTYPE)
    if n:
        lib.ora_allpairs(*args, out.ctypes.data, n, C.byref(flags))
    # packs may store their items in another order (strings are grouped by length class):
    # report the caller's item indices
    if getattr(left, "perm", None) is not None:
        out["left"] = left.perm[out["left"]]
    if getattr(right, "perm", None) is not None:
        out["right"] = right.perm[out["right"]]
    return out, flags.value


def score_pairs(left, right, li, ri, flat=False):
    lib = _lib()
    func = JACCARD if hasattr(left, "tok") else QRATIO
    ls, k1 = _side(left)
    rs, k2 = _side(right)
    def stored(p, idx):  # caller's item index -> stored position
        idx = np.asarray(idx, dtype=np.int64)
        if getattr(p, "perm", None) is None:
            return idx
        inv = np.empty(len(p.perm), dtype=np.int64)
        inv[p.perm.astype(np.int64)] = np.arange(len(p.perm))
        return inv[idx]

    li = np.ascontiguousarray(stored(left, li), dtype=np.uint32)
    ri = np.ascontiguousarray(stored(right, ri), dtype=np.uint32)
    out = np.zeros(len(li), dtype=np.float64)
    flags = C.c_uint32(0)
    lib.ora_score_pairs(func, int(flat), C.byref(ls), C.byref(rs), li.ctypes.data, ri.ctypes.data,
                        len(li), out.ctypes.data, C.byref(flags))
    return out, flags.value
