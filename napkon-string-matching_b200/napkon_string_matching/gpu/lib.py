"""ctypes binding of libnsm_b200.so (include/nsm.h).  There is no CPU implementation behind it:
if the library or a CUDA device is missing, loading raises."""
from __future__ import annotations

import ctypes as C
import os
import pathlib

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
# NSM_B200_LIB points at another build of the same library (kernel A/B experiments)
LIB_PATH = pathlib.Path(os.environ.get("NSM_B200_LIB") or _HERE / "libnsm_b200.so")

NSM_OK = 0
FLAG_OVERFLOW, FLAG_ZERO_UNION, FLAG_EMPTY_ITEM = 1, 2, 4
CAT_OFF, CAT_LIST_LIST, CAT_MEMBER = 0, 1, 2
N_STATS = 5
STAT_NAMES = ("candidates", "level_evals", "level_merges", "bound_pairs", "kept")

PAIR_DTYPE = np.dtype([("left", np.uint32), ("right", np.uint32), ("score", np.float64)])
assert PAIR_DTYPE.itemsize == 16

# NSM_OUT_PACKETS (include/nsm.h:nsm_packet_t): up to 48 kept pairs of one 512 x 128 block
OUT_PAIRS, OUT_PACKETS, OUT_CODED = 0, 1, 2
PACKET_RECORDS = 48
CPACKET_RECORDS = 60
DICT_SLOTS = 65536
# NSM_OUT_CODED (nsm_cpacket_t): 16-bit position + 16-bit score code per kept pair
CPACKET_DTYPE = np.dtype([("left0", np.uint32), ("right0", np.uint32), ("count", np.uint32),
                          ("reserved_", np.uint32), ("rec", np.uint32, (CPACKET_RECORDS,))])
assert CPACKET_DTYPE.itemsize == 256
RECORD_BYTES = (16, 496, 256)                           # arena entry per out_mode
ENTRY_RECORDS = (1, PACKET_RECORDS, CPACKET_RECORDS)    # kept pairs one entry can hold
PACKET_DTYPE = np.dtype([("left0", np.uint32), ("right0", np.uint32), ("count", np.uint32),
                         ("reserved_", np.uint32), ("score", np.float64, (PACKET_RECORDS,)),
                         ("local", np.uint16, (PACKET_RECORDS,))])
assert PACKET_DTYPE.itemsize == 496
UNIT_LEFT, UNIT_RIGHT = 512, 128   # the block of the cross product one packet lies in


def decode_packets(packets: np.ndarray) -> np.ndarray:
    """``nsm_packet_t`` records -> ``PAIR_DTYPE`` records (host side of NSM_OUT_PACKETS)."""
    count = packets["count"].astype(np.int64)
    used = np.arange(PACKET_RECORDS, dtype=np.int64)[None, :] < count[:, None]
    out = np.empty(int(count.sum()), dtype=PAIR_DTYPE)
    local = packets["local"][used].astype(np.uint32)
    out["left"] = np.repeat(packets["left0"], count) + (local >> np.uint32(7))
    out["right"] = np.repeat(packets["right0"], count) + (local & np.uint32(127))
    out["score"] = packets["score"][used]
    return out


def decode_cpackets(packets: np.ndarray, dictionary: np.ndarray) -> np.ndarray:
    """``nsm_cpacket_t`` records + the score dictionary (uint64[DICT_SLOTS]) -> ``PAIR_DTYPE``."""
    count = packets["count"].astype(np.int64)
    used = np.arange(CPACKET_RECORDS, dtype=np.int64)[None, :] < count[:, None]
    rec = packets["rec"][used]
    out = np.empty(len(rec), dtype=PAIR_DTYPE)
    out["left"] = np.repeat(packets["left0"], count) + ((rec & np.uint32(0xffff)) >> np.uint32(7))
    out["right"] = np.repeat(packets["right0"], count) + (rec & np.uint32(127))
    out["score"] = np.ascontiguousarray(dictionary).view(np.float64)[rec >> np.uint32(16)]
    return out

def _perm_ptr(perm, keep):
    if perm is None:
        return None
    perm = np.ascontiguousarray(perm, dtype=np.uint32)
    keep.append(perm)
    return perm.ctypes.data


def decode_into(out: np.ndarray, mode: int, part: np.ndarray, dictionary=None, left_perm=None,
                right_perm=None) -> int:
    """Decodes one wire-format part into the head of ``out`` (contiguous ``PAIR_DTYPE``) with the
    library's host decoders (nsm_decode_*: one pass, the item permutations applied on the way);
    returns the number of records written."""
    lib, keep = load(), []
    lp, rp = _perm_ptr(left_perm, keep), _perm_ptr(right_perm, keep)
    part = np.ascontiguousarray(part)
    if mode == OUT_PACKETS:
        return int(lib.nsm_decode_packets(part.ctypes.data, len(part), lp, rp, out.ctypes.data))
    if mode == OUT_CODED:
        table = np.ascontiguousarray(dictionary, dtype=np.uint64)
        return int(lib.nsm_decode_cpackets(part.ctypes.data, len(part), table.ctypes.data, lp, rp, out.ctypes.data))
    n = len(part)
    out[:n] = part
    if left_perm is not None:
        out["left"][:n] = np.asarray(left_perm)[part["left"]]
    if right_perm is not None:
        out["right"][:n] = np.asarray(right_perm)[part["right"]]
    return n


def sort_pairs(records: np.ndarray, n_left: int) -> np.ndarray:
    """``records`` (``PAIR_DTYPE``) ordered by (left, right): the library's host-side counting sort
    (nsm_sort_pairs), 10x faster than ``np.lexsort`` on 10^8 records."""
    records = np.ascontiguousarray(records)
    out = np.empty_like(records)
    check(load().nsm_sort_pairs(records.ctypes.data, len(records), int(n_left), out.ctypes.data))
    return out


RAW_SUFFIX_PARTS, RAW_LEVELS = 0, 1
PACK_MAX_ITEM_IDS = 1024
PACK_FLAG_TOO_LARGE, PACK_FLAG_NOT_NESTED, PACK_FLAG_BAD_ID = 1, 2, 4

EXPORTS = ("nsm_version", "nsm_last_error", "nsm_last_launch_count", "nsm_jaccard_allpairs",
           "nsm_qratio_allpairs", "nsm_microbench", "nsm_pack_count_ids", "nsm_pack_scratch_bytes",
           "nsm_pack_sets_measure", "nsm_pack_sets_fill", "nsm_dict_reset", "nsm_publish",
           "nsm_pack_strings_measure", "nsm_pack_strings_fill", "nsm_decode_packets", "nsm_decode_cpackets", "nsm_sort_pairs")
STR_SYM_NONE, STR_FLAG_UNMAPPED = 0xffff, 1


class NsmSets(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in (
        "item_level_off", "level_tok_off", "tok", "tok_entry", "level_head", "level_tail", "level_tail2",
        "level_info", "item_any", "item_k", "slot_ht", "slot_info")] + [
        ("n_items", C.c_uint32), ("n_levels", C.c_uint32), ("max_levels", C.c_uint32),
        ("n_slots", C.c_uint32), ("exact_bits", C.c_uint32), ("slot_stride", C.c_uint32),
        ("nested", C.c_uint32), ("reserved_", C.c_uint32)]


class NsmStrings(C.Structure):
    _fields_ = [("item_level_off", C.c_void_p), ("level_chr_off", C.c_void_p),
                ("level_len", C.c_void_p), ("chr", C.c_void_p), ("level_hist", C.c_void_p),
                ("n_items", C.c_uint32),
                ("n_levels", C.c_uint32), ("max_levels", C.c_uint32), ("max_len", C.c_uint32),
                ("n_alphabet", C.c_uint32), ("reserved_", C.c_uint32),
                ("class_end", C.c_uint32 * 8)]


class NsmRawSets(C.Structure):
    _fields_ = [("item_grp_off", C.c_void_p), ("grp_id_off", C.c_void_p), ("ids", C.c_void_p),
                ("rank", C.c_void_p), ("n_items", C.c_uint32), ("n_groups", C.c_uint32),
                ("n_ids", C.c_uint32), ("n_vocab", C.c_uint32), ("mode", C.c_uint32),
                ("reserved_", C.c_uint32)]


class NsmRawStrings(C.Structure):
    _fields_ = [("level_off", C.c_void_p), ("cps", C.c_void_p), ("cp_sym", C.c_void_p),
                ("n_levels", C.c_uint32), ("n_cps", C.c_uint32), ("table_len", C.c_uint32),
                ("blank_sym", C.c_uint32)]


class NsmJob(C.Structure):
    _fields_ = [("l_row_begin", C.c_uint32), ("l_row_end", C.c_uint32), ("flat", C.c_uint32),
                ("cat_mode", C.c_uint32), ("threshold", C.c_double), ("l_cat", C.c_void_p),
                ("r_cat", C.c_void_p), ("out_pairs", C.c_void_p), ("out_capacity", C.c_uint64),
                ("out_count", C.c_void_p), ("out_flags", C.c_void_p), ("out_stats", C.c_void_p),
                ("out_mode", C.c_uint32), ("reserved_", C.c_uint32),
                ("out_dict", C.c_void_p), ("out_exc", C.c_void_p), ("out_exc_capacity", C.c_uint64),
                ("out_exc_count", C.c_void_p)]


class NsmError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Loads the CUDA library.  Raises if it has not been built (python build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NsmError(
            f"{LIB_PATH} is missing: build it with `python napkon-string-matching_b200/build.py` "
            "(there is no CPU fallback for the comparison path)")
    lib = C.CDLL(str(LIB_PATH))
    lib.nsm_version.restype = C.c_int
    lib.nsm_last_error.restype = C.c_char_p
    lib.nsm_last_launch_count.restype = C.c_int
    for name, first in (("nsm_jaccard_allpairs", NsmSets), ("nsm_qratio_allpairs", NsmStrings)):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(first), C.POINTER(first), C.POINTER(NsmJob), C.c_void_p]
    lib.nsm_publish.restype = C.c_int
    lib.nsm_publish.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.nsm_dict_reset.restype = C.c_int
    lib.nsm_dict_reset.argtypes = [C.c_void_p, C.c_void_p]
    lib.nsm_microbench.restype = C.c_int
    lib.nsm_microbench.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                   C.POINTER(C.c_uint64), C.c_void_p]
    lib.nsm_pack_count_ids.restype = C.c_int
    lib.nsm_pack_count_ids.argtypes = [C.POINTER(NsmRawSets), C.c_void_p, C.c_void_p]
    lib.nsm_pack_scratch_bytes.restype = C.c_uint64
    lib.nsm_pack_scratch_bytes.argtypes = [C.c_uint32]
    lib.nsm_pack_sets_measure.restype = C.c_int
    lib.nsm_pack_sets_measure.argtypes = [C.POINTER(NsmRawSets), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_uint64, C.c_void_p]
    lib.nsm_pack_sets_fill.restype = C.c_int
    lib.nsm_pack_sets_fill.argtypes = [C.POINTER(NsmRawSets), C.c_void_p, C.POINTER(NsmSets),
                                       C.c_void_p, C.c_void_p]
    lib.nsm_pack_strings_measure.restype = C.c_int
    lib.nsm_pack_strings_measure.argtypes = [C.POINTER(NsmRawStrings), C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
    lib.nsm_pack_strings_fill.restype = C.c_int
    lib.nsm_pack_strings_fill.argtypes = [C.POINTER(NsmRawStrings), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
    lib.nsm_decode_packets.restype = C.c_uint64
    lib.nsm_decode_packets.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nsm_decode_cpackets.restype = C.c_uint64
    lib.nsm_decode_cpackets.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nsm_sort_pairs.restype = C.c_int
    lib.nsm_sort_pairs.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != NSM_OK:
        raise NsmError(f"nsm error {rc}: {load().nsm_last_error().decode(errors='replace')}")
