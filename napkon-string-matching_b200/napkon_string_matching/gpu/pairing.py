"""
Host glue between ``ComparableData.gen_comparable`` and the CUDA engine: packs tokenised items,
reproduces the reference's exceptions for inputs its pair loop would have raised on, runs the
all-pairs job (sharded over ranks when ``torch.distributed`` is initialised) and gathers the
result frame from the kept ``(left, right, score)`` records.

Reference statements covered here (all /root/reference/napkon_string_matching/types/
comparable_data.py): category predicate :464-476, black-list exclusion :523-552, result columns
:236-240 + comparable.py:26-31, frame index = position in the cross product (:191).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import pandas as pd

from napkon_string_matching.compare.score_functions import KINDS
from napkon_string_matching.gpu import lib as nsmlib
from napkon_string_matching.gpu import pack
from napkon_string_matching.types.comparable import COLUMN_NAMES, Columns

PAIR_DTYPE = nsmlib.PAIR_DTYPE


# ------------------------------------------------------------------------------------------
# categories (Q9)
# ------------------------------------------------------------------------------------------
def _category_mode(first_left, first_right) -> str:
    if isinstance(first_left, list):
        if isinstance(first_right, list):
            return "list_list"
        # the reference evaluates `x in set(y)` with x the left *list*
        raise TypeError("unhashable type: 'list'")
    return "member" if isinstance(first_right, list) else "equal"


def _is_nan(v) -> bool:
    return isinstance(v, float) and v != v


def category_keep(mode: str, left, right) -> bool:
    if mode == "list_list":
        return (not set(left).isdisjoint(set(right))) or (not left and not right)
    if mode == "member":
        return left in set(right)
    return left == right


def category_keep_rows(col_left: Sequence, col_right: Sequence) -> List[bool]:
    col_left, col_right = list(col_left), list(col_right)
    mode = _category_mode(col_left[0], col_right[0])
    return [category_keep(mode, a, b) for a, b in zip(col_left, col_right)]


def category_masks(col_left: Sequence, col_right: Sequence):
    """Per-item category bit masks for the in-kernel predicate, or a host-side predicate when
    there are more than 64 distinct categories.  Returns a dict understood by
    :func:`score_all_pairs`."""
    col_left, col_right = list(col_left), list(col_right)
    mode = _category_mode(col_left[0], col_right[0])
    values = {}

    def bit(v):
        if _is_nan(v):
            return None  # NaN equals nothing
        return values.setdefault(v, len(values))

    def mask_of(cell, as_list):
        if as_list:
            bits = [bit(v) for v in (cell or [])]
        else:
            bits = [bit(cell)] if (mode == "equal" or cell is not None) else []
        return [b for b in bits if b is not None]

    left_bits = [mask_of(c, mode == "list_list") for c in col_left]
    right_bits = [mask_of(c, mode in ("list_list", "member")) for c in col_right]
    if len(values) > 64:
        return {"mode": mode, "host": (col_left, col_right)}
    to_mask = lambda bits: np.uint64(sum(1 << b for b in set(bits)))  # noqa: E731
    return {"mode": mode,
            "left": np.array([to_mask(b) for b in left_bits], dtype=np.uint64),
            "right": np.array([to_mask(b) for b in right_bits], dtype=np.uint64),
            "cat_mode": nsmlib.CAT_LIST_LIST if mode == "list_list" else nsmlib.CAT_MEMBER}


# ------------------------------------------------------------------------------------------
# exceptions the reference's pair loop raises (Q1, Q4)
# ------------------------------------------------------------------------------------------
def first_raising_pair(left_levels, right_levels, jaccard: bool,
                       skip: Optional[Callable[[int, int], bool]] = None):
    """The first pair, in the reference's row-major order, on which ``compare_terms`` raises:
    ``IndexError`` when exactly one item has no levels (``left[-1]`` on an empty list),
    ``ZeroDivisionError`` when ``intersection_vs_union`` meets two empty token sets.
    Returns ``(exception_type, left, right)`` or None.  O(N) unless such items exist."""
    kl = np.fromiter((len(v) for v in left_levels), dtype=np.int64, count=len(left_levels))
    kr = np.fromiter((len(v) for v in right_levels), dtype=np.int64, count=len(right_levels))

    def has_empty(levels):
        return np.fromiter((any(len(s) == 0 for s in v[1:] or v[:1]) for v in levels), dtype=bool,
                           count=len(levels)) if jaccard else np.zeros(len(levels), dtype=bool)

    el, er = has_empty(left_levels), has_empty(right_levels)
    special_l, special_r = (kl == 0) | el, (kr == 0) | er
    if not special_l.any() and not special_r.any():
        return None
    if len(kl) == 0 or len(kr) == 0:
        return None

    def raises(li, ri):
        a, b = left_levels[li], right_levels[ri]
        if not a and not b:
            return None
        if not a or not b:
            return IndexError
        if jaccard:
            for t in range(1, max(len(a), len(b)) + 1):
                if not a[min(t, len(a) - 1)] and not b[min(t, len(b) - 1)]:
                    return ZeroDivisionError
        return None

    right_special_idx = np.nonzero(special_r)[0]
    for li in range(len(kl)):
        candidates = range(len(kr)) if special_l[li] else right_special_idx
        for ri in candidates:
            ri = int(ri)
            exc = raises(li, ri)
            if exc is not None and not (skip is not None and skip(li, ri)):
                return exc, li, ri
    return None


# ------------------------------------------------------------------------------------------
# the all-pairs job
# ------------------------------------------------------------------------------------------
def pack_levels(left_levels, right_levels, score_func: str):
    if KINDS[score_func] == "sets":
        return pack.pack_sets(left_levels, right_levels)
    return pack.pack_strings(pack.fuzzy_level_strings(left_levels),
                             pack.fuzzy_level_strings(right_levels))


def upload_levels(engine, left_levels, right_levels, score_func: str):
    """Both sides in device memory, ready for ``engine.all_pairs``: ``(left, right, perms)``.

    Token sets: the host maps token strings to codes (exact string identity, Q3) and the GPU
    builds the packed arrays (gpu/device_pack.py).  An item beyond the device packer's per-item
    limit sends the comparison through the numpy packer instead.  Strings: the host joins each
    level's tokens (``join_sorted``), the GPU applies ``default_process`` and builds the packed
    arrays (csrc/pack_strings.cu)."""
    from napkon_string_matching.gpu.stages import stage

    packer = getattr(engine, "device_packer", None) if KINDS[score_func] == "sets" else None
    if packer is not None:
        from napkon_string_matching.gpu.device_pack import RawSets

        with stage("token codes (factorize)"):
            parts = [pack._csr_from_nested(s) for s in (left_levels, right_levels)]
            tokens = [t for _, _, flat in parts for t in flat]
            codes, uniques = pd.factorize(np.asarray(tokens, dtype=object)) if tokens else (np.zeros(0, np.int64), [])
            raws, pos = [], 0
            for item_level_off, level_off, flat in parts:
                raws.append(RawSets(item_level_off.astype(np.uint32), level_off.astype(np.uint32),
                                    codes[pos:pos + len(flat)].astype(np.uint32), nsmlib.RAW_LEVELS))
                pos += len(flat)
            # items of one level count next to each other, in chunks of the kernel's unit
            (raws[0], lperm), (raws[1], rperm) = (raws[0].ordered_by_levels(nsmlib.UNIT_LEFT),
                                                  raws[1].ordered_by_levels(nsmlib.UNIT_RIGHT))
        try:
            with stage("H2D + device packing"):
                dl, dr = packer.pack(raws, len(uniques), rank="frequency")
            dl.perm, dr.perm = lperm, rperm
            return dl, dr, (lperm, rperm)
        except pack.PackTooLarge:
            pass   # the numpy packer below has no per-item limit
    if KINDS[score_func] == "strings" and getattr(engine, "string_packer", None) is not None:
        from napkon_string_matching.gpu.device_pack import PackUnsupported
        from napkon_string_matching.text.process import join_sorted

        with stage("level strings (join_sorted)"):
            raw = [[[join_sorted(level) if isinstance(level, list) else level for level in lv] for lv in side]
                   for side in (left_levels, right_levels)]
        try:
            with stage("H2D + device packing"):
                dl, dr = engine.string_packer.pack(raw)
            return dl, dr, (dl.perm, dr.perm)
        except PackUnsupported:
            pass   # a code point whose lower-case form depends on its neighbours: host packer
    with stage("host packing (numpy)"):
        pl, pr = pack_levels(left_levels, right_levels, score_func)
    with stage("H2D"):
        return engine.upload(pl), engine.upload(pr), (getattr(pl, "perm", None), getattr(pr, "perm", None))


def score_all_pairs(left_levels, right_levels, score_func: str, score_threshold: float,
                    categories: Optional[dict] = None,
                    skip_pair: Optional[Callable[[int, int], bool]] = None,
                    engine=None) -> np.ndarray:
    """Kept records (``PAIR_DTYPE``) of ``compare_terms(l, r, score_func) >= score_threshold``
    over the whole cross product, sorted by (left, right)."""
    from napkon_string_matching.gpu import distributed
    from napkon_string_matching.gpu.engine import default_engine

    if score_func not in KINDS:
        raise AttributeError(f"no packed kernel for score function {score_func!r}")
    jaccard = KINDS[score_func] == "sets"

    host_cat = None
    if categories is not None and "host" in categories:
        host_cat = categories

    def excluded(li, ri):
        if skip_pair is not None and skip_pair(li, ri):
            return True
        if categories is not None:
            if host_cat is not None:
                return not category_keep(host_cat["mode"], host_cat["host"][0][li], host_cat["host"][1][ri])
            ml, mr = int(categories["left"][li]), int(categories["right"][ri])
            if categories["cat_mode"] == nsmlib.CAT_LIST_LIST:
                return not ((ml & mr) != 0 or (ml == 0 and mr == 0))
            return (ml & mr) == 0
        return False

    from napkon_string_matching.gpu.stages import stage

    with stage("scan for pairs the reference raises on"):
        bad = first_raising_pair(left_levels, right_levels, jaccard, excluded)
    if bad is not None:
        exc, li, ri = bad
        if exc is IndexError:
            raise IndexError("list index out of range")
        raise ZeroDivisionError("division by zero")

    if len(left_levels) == 0 or len(right_levels) == 0:
        return np.zeros(0, dtype=PAIR_DTYPE)
    engine = engine or default_engine()
    dl, dr, (lperm, rperm) = upload_levels(engine, left_levels, right_levels, score_func)
    kw = {}
    if categories is not None and host_cat is None:
        # masks are indexed by stored position (string packs group their items by length class)
        lmask, rmask = categories["left"], categories["right"]
        if lperm is not None:
            lmask, rmask = lmask[lperm], rmask[rperm]
        kw = dict(l_cat=engine.upload_masks(lmask), r_cat=engine.upload_masks(rmask),
                  cat_mode=categories["cat_mode"])
    with stage("kernels + D2H + decode (engine)"):
        records = distributed.sharded_all_pairs(
            lambda b, e: engine.all_pairs(dl, dr, score_threshold, rows=(b, e), **kw),
            dl.weights)
    # pairs the reference never scores (excluded before the loop) may carry the kernel's flags;
    # anything else was ruled out by first_raising_pair above
    if host_cat is not None and len(records):
        keep = np.fromiter((category_keep(host_cat["mode"], host_cat["host"][0][l], host_cat["host"][1][r])
                            for l, r in zip(records["left"], records["right"])), dtype=bool,
                           count=len(records))
        records = records[keep]
    with stage("sort records"):
        # the library's host-side counting sort: row-major order of the cross product
        return nsmlib.sort_pairs(records, len(left_levels))


def not_blocked(records: np.ndarray, left_ids: Sequence, right_ids: Sequence, blocked: set
                ) -> np.ndarray:
    """Boolean mask of records whose (left id, right id) is not black-listed (Q8)."""
    lids, rids = list(left_ids), list(right_ids)
    lpos = {v: i for i, v in enumerate(lids)}
    rpos = {v: i for i, v in enumerate(rids)}
    n_right = len(rids)
    keys = set()
    blocked_l, blocked_r = {l for l, _ in blocked}, {r for _, r in blocked}
    # identifiers are not necessarily unique: expand each black-listed id to all its positions
    lall, rall = {}, {}
    for i, v in enumerate(lids):
        if v in blocked_l:
            lall.setdefault(v, []).append(i)
    for i, v in enumerate(rids):
        if v in blocked_r:
            rall.setdefault(v, []).append(i)
    for l, r in blocked:
        for i in lall.get(l, ()):
            for j in rall.get(r, ()):
                keys.add(i * n_right + j)
    if not keys:
        return np.ones(len(records), dtype=bool)
    flat = records["left"].astype(np.int64) * n_right + records["right"].astype(np.int64)
    return ~np.isin(flat, np.fromiter(keys, dtype=np.int64, count=len(keys)))


def result_frame(records: np.ndarray, left_df: pd.DataFrame, right_df: pd.DataFrame,
                 left_prefix: str, right_prefix: str) -> pd.DataFrame:
    """``{Left}{Identifier, Sheet, Variable, Argument}``, ``{Right}...``, ``MatchScore`` for the
    kept pairs only; index = position of the pair in the cross product (Q10)."""
    lcols = [c for c in left_df.columns if c in COLUMN_NAMES]
    rcols = [c for c in right_df.columns if c in COLUMN_NAMES]
    li = records["left"].astype(np.int64)
    ri = records["right"].astype(np.int64)
    index = li * len(right_df) + ri
    # gather with the columns' own arrays (take keeps the dtype — Arrow-backed strings stay Arrow:
    # going through object arrays made pandas re-infer every gathered column, 3/4 of the time)
    data = {}
    for c in lcols:
        data[left_prefix + c] = left_df[c].array.take(li)
    for c in rcols:
        data[right_prefix + c] = right_df[c].array.take(ri)
    data[Columns.MATCH_SCORE.value] = records["score"].astype(np.float64)
    return pd.DataFrame(data, index=pd.Index(index))
