"""
Host-side packing of tokenised items into the arrays the CUDA kernels read (north_star "Token
packing").  numpy only — no per-pair work happens here.

Jaccard operands (``PackedSets``) — one cohort side, CSR over items -> levels -> tokens:

    item_level_off  uint32[n_items+1]   levels of item i are item_level_off[i]:item_level_off[i+1]
    level_tok_off   uint32[n_levels+1]  tokens of level g are tok[level_tok_off[g]:level_tok_off[g+1]]
    tok             uint32[n_tok]       token ids, sorted and unique inside a level
    tok_entry       uint8[n_tok]        first level of the item that holds this id (levels made by
                                        gen_comp_value are nested: level j is a subset of level
                                        j+1; ``nested`` says whether that holds for every item)
    level_head      uint64[n_levels]    exact bitset of the level's ids 0..63
    level_tail      uint64[n_levels]    signature of the ids >= 64 (bit = id - 64 if the whole
                                        vocabulary has <= 128 ids, else a multiplicative hash)
    level_tail2     uint64[n_levels]    a second, independent signature of the ids >= 64
    level_info      uint32[n_levels]    size | min(n_tail - popcount(tail), 255) << 16
    item_any        uint64[n_items, 2]  OR of (head, tail) over the levels compare_terms can use
                                        (levels 1..K-1, or level 0 when K == 1)
    item_k          uint32[n_items]     number of levels K of the item
    slot_ht         uint64[n_slots, stride, 2]   (head, tail) again, laid out by compare_terms step:
    slot_info       uint32[n_slots, stride]      slot t-1 holds level min(t, K-1) of each item
                                        (n_slots = clamp(max K - 1, 1, 10); slot-major, so a block of
                                        consecutive items is contiguous per step; stride = n_items
                                        rounded up to 128 and zero-filled, so that every block of
                                        128 items is one aligned, full-size bulk copy per step)

Token ids are assigned in order of falling frequency over all packed sides, so ids 0..63 (the
"head") are the 64 most frequent tokens: under the Zipf-like token statistics of questionnaire
text and MeSH ids most shared tokens are head tokens and their intersection is one popcount.

Levels are what ``ComparableData.gen_comp_value`` returns for an item
(/root/reference/napkon_string_matching/types/comparable_data.py:283-285): level j is the token
set of the last j+1 parts.  Token identity is the exact, case-sensitive string (Q3/Q4).

fuzzy_match operands (``PackedStrings``): per level the string ``fuzzy_match`` would hand to
``QRatio`` *after* ``join_sorted`` and rapidfuzz's ``default_process`` (Q5/Q6), as one byte per
code point through an alphabet shared by both sides:

    item_level_off  uint32[n_items+1]
    level_chr_off   uint32[n_levels]    start of the level's string in chr, a multiple of 8
    level_len       uint32[n_levels]    its length in code points
    chr             uint8[n_chr]        dense alphabet codes (0 .. n_alphabet-1), every string
                                        padded to a multiple of 8 bytes (the kernel reads 8 at a time)

Items are stored ordered by the number of 64-bit words their longest level string needs
(``class_end[w]`` = end of the items needing <= w+1 words), so that the kernel can run every class
with the narrowest bit-vectors; ``perm`` maps a stored position back to the caller's item index.
"""
from __future__ import annotations

from dataclasses import dataclass
from itertools import chain
from typing import Iterable, List, Sequence, Tuple

import numpy as np
import pandas as pd

from napkon_string_matching.text.process import default_process, join_sorted

SIG_HASH_MULT = np.uint32(0x9E3779B1)
MAX_LEVEL_SIZE = 0xFFFF
MAX_ALPHABET = 255


class PackError(ValueError):
    pass


class PackTooLarge(PackError):
    """An item exceeds what the device packer holds per item; the numpy packer has no such limit."""


@dataclass
class PackedSets:
    item_level_off: np.ndarray
    level_tok_off: np.ndarray
    tok: np.ndarray
    tok_entry: np.ndarray
    level_head: np.ndarray
    level_tail: np.ndarray
    level_tail2: np.ndarray
    level_info: np.ndarray
    item_any: np.ndarray
    item_k: np.ndarray
    slot_ht: np.ndarray
    slot_info: np.ndarray
    n_vocab: int
    exact_bits: bool
    max_levels: int = 0
    nested: bool = False

    @property
    def n_items(self) -> int:
        return len(self.item_level_off) - 1

    @property
    def n_levels(self) -> int:
        return len(self.level_tok_off) - 1

    def arrays(self):
        return [self.item_level_off, self.level_tok_off, self.tok, self.tok_entry, self.level_head,
                self.level_tail, self.level_tail2, self.level_info, self.item_any, self.item_k,
                self.slot_ht, self.slot_info]

    @property
    def n_slots(self) -> int:
        return self.slot_ht.shape[0]

    @property
    def slot_stride(self) -> int:
        return self.slot_ht.shape[1]

    def level_sizes(self) -> np.ndarray:
        return np.diff(self.level_tok_off.astype(np.int64))

    def levels_per_item(self) -> np.ndarray:
        return np.diff(self.item_level_off.astype(np.int64))

    def nbytes(self) -> int:
        return sum(a.nbytes for a in self.arrays())

    def rows(self, begin: int, end: int) -> "PackedSets":
        """Items begin:end as an independent pack (used by tests and CPU baselines)."""
        g0, g1 = int(self.item_level_off[begin]), int(self.item_level_off[end])
        t0, t1 = int(self.level_tok_off[g0]), int(self.level_tok_off[g1])
        return PackedSets(
            (self.item_level_off[begin : end + 1] - np.uint32(g0)).astype(np.uint32),
            (self.level_tok_off[g0 : g1 + 1] - np.uint32(t0)).astype(np.uint32),
            self.tok[t0:t1].copy(), self.tok_entry[t0:t1].copy(), self.level_head[g0:g1].copy(),
            self.level_tail[g0:g1].copy(),
            self.level_tail2[g0:g1].copy(), self.level_info[g0:g1].copy(),
            self.item_any[begin:end].copy(), self.item_k[begin:end].copy(),
            _pad_slots(self.slot_ht[:, begin:end]), _pad_slots(self.slot_info[:, begin:end]),
            self.n_vocab,
            self.exact_bits, self.max_levels, self.nested)


@dataclass
class PackedStrings:
    item_level_off: np.ndarray
    level_chr_off: np.ndarray
    level_len: np.ndarray
    chr: np.ndarray
    n_alphabet: int
    max_levels: int = 0
    max_len: int = 0
    perm: np.ndarray | None = None          # stored position -> original item index
    class_end: np.ndarray | None = None     # uint32[8]
    level_hist: np.ndarray | None = None    # uint32[n_levels, 8]: 32 saturating byte counters

    @property
    def n_items(self) -> int:
        return len(self.item_level_off) - 1

    @property
    def n_levels(self) -> int:
        return len(self.level_len)

    def classes(self) -> np.ndarray:
        if self.class_end is not None:
            return self.class_end
        return np.full(WORD_CLASSES, self.n_items, dtype=np.uint32)

    def arrays(self):
        if self.level_hist is None:
            self.level_hist = string_histograms(self.level_chr_off, self.level_len, self.chr)
        return [self.item_level_off, self.level_chr_off, self.level_len, self.chr, self.level_hist]

    def level_lengths(self) -> np.ndarray:
        return self.level_len.astype(np.int64)

    def levels_per_item(self) -> np.ndarray:
        return np.diff(self.item_level_off.astype(np.int64))

    def nbytes(self) -> int:
        return sum(a.nbytes for a in self.arrays())

    def level_string_codes(self, g: int) -> np.ndarray:
        o = int(self.level_chr_off[g])
        return self.chr[o : o + int(self.level_len[g])]

    def rows(self, begin: int, end: int) -> "PackedStrings":
        g0, g1 = int(self.item_level_off[begin]), int(self.item_level_off[end])
        c0 = int(self.level_chr_off[g0]) if g1 > g0 else 0
        c1 = int(self.level_chr_off[g1 - 1]) + _pad8(int(self.level_len[g1 - 1])) if g1 > g0 else 0
        cls = np.clip(self.classes().astype(np.int64) - begin, 0, end - begin).astype(np.uint32)
        return PackedStrings(
            (self.item_level_off[begin : end + 1] - np.uint32(g0)).astype(np.uint32),
            (self.level_chr_off[g0:g1] - np.uint32(c0)).astype(np.uint32),
            self.level_len[g0:g1].copy(), self.chr[c0:c1].copy(), self.n_alphabet, self.max_levels,
            self.max_len, None if self.perm is None else self.perm[begin:end].copy(), cls,
            None if self.level_hist is None else self.level_hist[g0:g1].copy())


def _pad_slots(a: np.ndarray) -> np.ndarray:
    """Slot rows cut to a sub-range of items, re-padded to the 128-item stride."""
    n = a.shape[1]
    stride = max(128, (n + 127) // 128 * 128)
    out = np.zeros((a.shape[0], stride) + a.shape[2:], dtype=a.dtype)
    out[:, :n] = a
    return out


def _pad8(n):
    return (n + 7) // 8 * 8


WORD_CLASSES = 8   # 64-bit words per pattern the kernel instantiates: up to 512 characters


# ------------------------------------------------------------------------------------------
# signatures
# ------------------------------------------------------------------------------------------
HEAD_IDS = 64
SLOT_CAP = 10
SLOT_BLOCK = 128   # items per right block of the kernel
SIG2_HASH_MULT = np.uint32(0x85EBCA77)


def tail_bits(ids: np.ndarray, exact_bits: bool, mult=SIG_HASH_MULT) -> np.ndarray:
    """Bit position (0..63) in the tail signature of every token id >= 64."""
    ids = ids.astype(np.uint32, copy=False)
    if exact_bits:
        return (ids - np.uint32(HEAD_IDS)).astype(np.uint64)
    with np.errstate(over="ignore"):
        return ((ids * mult) >> np.uint32(26)).astype(np.uint64)


def _segment_or(values: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """bitwise-or of values[offsets[g]:offsets[g+1]] for every g (empty segment -> 0)."""
    n_seg = len(offsets) - 1
    out = np.zeros(n_seg, dtype=np.uint64)
    sizes = np.diff(offsets.astype(np.int64))
    nonempty = np.nonzero(sizes > 0)[0]
    if len(nonempty):
        out[nonempty] = np.bitwise_or.reduceat(values, offsets[nonempty].astype(np.int64))
    return out


def finish_sets(item_level_off, level_tok_off, tok, n_vocab: int) -> PackedSets:
    """Adds signatures / info words to a CSR whose levels are already sorted and unique and
    whose ids are frequency-ranked."""
    item_level_off = np.ascontiguousarray(item_level_off, dtype=np.uint32)
    level_tok_off = np.ascontiguousarray(level_tok_off, dtype=np.uint32)
    tok = np.ascontiguousarray(tok, dtype=np.uint32)
    sizes = np.diff(level_tok_off.astype(np.int64))
    if len(sizes) and sizes.max() > MAX_LEVEL_SIZE:
        raise PackError(f"a level holds {sizes.max()} tokens; the packed format allows 65535")
    exact_bits = n_vocab <= 2 * HEAD_IDS
    is_head = tok < HEAD_IDS
    one = np.uint64(1)
    head_bit = np.where(is_head, np.left_shift(one, np.where(is_head, tok, 0).astype(np.uint64)),
                        np.uint64(0))
    tail_ids = np.where(is_head, HEAD_IDS, tok)
    tail_bit = np.where(is_head, np.uint64(0), np.left_shift(one, tail_bits(tail_ids, exact_bits)))
    tail2_bit = np.where(is_head, np.uint64(0),
                         np.left_shift(one, tail_bits(tail_ids, exact_bits, SIG2_HASH_MULT)))
    head = _segment_or(head_bit, level_tok_off)
    tail = _segment_or(tail_bit, level_tok_off)
    tail2 = _segment_or(tail2_bit, level_tok_off)
    n_head = np.bitwise_count(head).astype(np.int64)
    extra = np.minimum(sizes - n_head - np.bitwise_count(tail).astype(np.int64), 255)
    info = (sizes.astype(np.uint32) | (extra.astype(np.uint32) << np.uint32(16))).astype(np.uint32)
    k = np.diff(item_level_off.astype(np.int64))
    n_items, n_levels = len(k), len(sizes)
    # union over the levels compare_terms can touch: 1..K-1, or 0 when K == 1
    item_any = np.zeros((n_items, 2), dtype=np.uint64)
    if n_levels:
        level_item = np.repeat(np.arange(n_items, dtype=np.int64), k)
        level_j = np.arange(n_levels, dtype=np.int64) - item_level_off[:-1].astype(np.int64)[level_item]
        used = (level_j >= 1) | (k[level_item] == 1)
        np.bitwise_or.at(item_any[:, 0], level_item[used], head[used])
        np.bitwise_or.at(item_any[:, 1], level_item[used], tail[used])
    # compare_terms' schedule, materialised: slot t-1 = level min(t, K-1)
    max_k = int(k.max()) if n_items else 0
    n_slots = min(max(max_k - 1, 1), SLOT_CAP)
    stride = max(SLOT_BLOCK, (n_items + SLOT_BLOCK - 1) // SLOT_BLOCK * SLOT_BLOCK)
    slot_ht = np.zeros((n_slots, stride, 2), dtype=np.uint64)
    slot_info = np.zeros((n_slots, stride), dtype=np.uint32)
    has = k > 0
    if n_levels and has.any():
        base = item_level_off[:-1].astype(np.int64)
        for t in range(1, n_slots + 1):
            g = (base + np.minimum(t, np.maximum(k - 1, 0)))[has]
            slot_ht[t - 1, :n_items][has, 0] = head[g]
            slot_ht[t - 1, :n_items][has, 1] = tail[g]
            slot_info[t - 1, :n_items][has] = info[g]
    tok_entry, nested = _entry_levels(item_level_off, level_tok_off, tok, k)
    return PackedSets(item_level_off, level_tok_off, tok, tok_entry, head, tail, tail2, info,
                      item_any, np.minimum(k, 0xFFFFFFFF).astype(np.uint32), slot_ht, slot_info,
                      int(n_vocab), bool(exact_bits), max_k, nested)


def _entry_levels(item_level_off, level_tok_off, tok, k) -> Tuple[np.ndarray, bool]:
    """Per token row the first level of its item that holds the id, and whether every item's
    levels are nested (each id is held by all levels from its entry level to the deepest)."""
    n_levels, n_tok = len(level_tok_off) - 1, len(tok)
    if n_tok == 0:
        return np.zeros(0, dtype=np.uint8), True
    sizes = np.diff(level_tok_off.astype(np.int64))
    level_item = np.repeat(np.arange(len(k), dtype=np.int64), k)
    level_j = np.arange(n_levels, dtype=np.int64) - item_level_off[:-1].astype(np.int64)[level_item]
    row_item = np.repeat(level_item, sizes)
    row_j = np.repeat(level_j, sizes)
    span, kspan = int(tok.max()) + 1, int(k.max()) + 1
    # one sort of the combined (item, id, level) key; rows of one (item, id) end up adjacent with
    # their levels ascending
    if len(k) * span * kspan < 2 ** 62:
        key = (row_item * span + tok.astype(np.int64)) * kspan + row_j
        order = np.argsort(key, kind="stable")
        skey = key[order] // kspan
    else:  # the combined key would overflow 64 bits: three-key sort
        order = np.lexsort((row_j, tok, row_item))
        skey = np.cumsum(np.concatenate([[0], (np.diff(row_item[order]) != 0) | (np.diff(tok[order].astype(np.int64)) != 0)]))
    sj = row_j[order]
    first = np.ones(n_tok, dtype=bool)
    first[1:] = skey[1:] != skey[:-1]
    group = np.cumsum(first) - 1
    entry_of_group = sj[first]
    count_of_group = np.bincount(group)
    k_of_group = k[row_item[order][first]]
    nested = bool(np.all(count_of_group == k_of_group - entry_of_group)) and int(k.max(initial=0)) <= 255
    entry = np.empty(n_tok, dtype=np.int64)
    entry[order] = entry_of_group[group]
    return np.minimum(entry, 255).astype(np.uint8), nested


def chunked_level_order(k: np.ndarray, chunk: int) -> np.ndarray:
    """Storage order for the items of a token-set cohort: ``perm[stored position] = item index``.

    Items are sorted by their level count K and cut into chunks of ``chunk`` items (the kernel's
    unit: 512 left rows, 128 right columns), so that the items a warp scores together walk
    ``compare_terms``' schedule for the same number of steps (stage C runs its step loop to the
    deepest schedule of the 32 pairs of a round; in a mixed cohort a third of those lane-steps are
    idle).  The chunks are then laid out in bit-reversed order, so that every contiguous row range
    — a rank's row block, the engine's probe block — holds the cohort's mix of level counts.  A
    trailing partial chunk stays last (the chunks before it stay aligned)."""
    k = np.asarray(k, dtype=np.int64)
    order = np.argsort(k, kind="stable")
    full = len(k) // chunk
    if full < 2:
        return order
    bits = (full - 1).bit_length()
    idx = np.arange(1 << bits, dtype=np.int64)
    rev = np.zeros_like(idx)
    for b in range(bits):
        rev |= ((idx >> b) & 1) << (bits - 1 - b)
    seq = rev[rev < full]
    head = order[: full * chunk].reshape(full, chunk)[seq].reshape(-1)
    return np.concatenate([head, order[full * chunk:]])


def rank_by_frequency(codes_per_side: List[np.ndarray], n_vocab: int) -> List[np.ndarray]:
    """Renumbers ids so that id 0 is the most frequent token over all sides."""
    if n_vocab == 0:
        return codes_per_side
    counts = np.zeros(n_vocab, dtype=np.int64)
    for c in codes_per_side:
        counts += np.bincount(c, minlength=n_vocab)
    order = np.argsort(-counts, kind="stable")
    rank = np.empty(n_vocab, dtype=np.int64)
    rank[order] = np.arange(n_vocab, dtype=np.int64)
    return [rank[c] for c in codes_per_side]


def _csr_from_nested(items_levels: Sequence[Sequence[Sequence]]) -> Tuple[np.ndarray, np.ndarray, list]:
    k = np.fromiter(map(len, items_levels), dtype=np.int64, count=len(items_levels))
    item_level_off = np.zeros(len(k) + 1, dtype=np.int64)
    np.cumsum(k, out=item_level_off[1:])
    levels = list(chain.from_iterable(items_levels))
    if any(isinstance(level, str) for level in levels):  # intersection_vs_union splits a str operand
        levels = [level.split() if isinstance(level, str) else level for level in levels]
    sizes = np.fromiter(map(len, levels), dtype=np.int64, count=len(levels))
    flat = list(chain.from_iterable(levels))
    level_off = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes, out=level_off[1:])
    return item_level_off, level_off, flat


def _sort_unique_levels(level_off: np.ndarray, codes: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Sort ids inside every level and drop duplicates (the score function works on sets).
    One sort of a combined (level, id) key instead of a two-key lexsort."""
    n_levels = len(level_off) - 1
    sizes = np.diff(level_off)
    if len(codes) == 0:
        return np.zeros(n_levels + 1, dtype=np.int64), codes.astype(np.int64)
    span = int(codes.max()) + 1
    level_id = np.repeat(np.arange(n_levels, dtype=np.int64), sizes)
    if n_levels * span < 2 ** 62:
        key = np.sort(level_id * span + codes)
        keep = np.ones(len(key), dtype=bool)
        keep[1:] = key[1:] != key[:-1]
        key = key[keep]
        level_id, codes = key // span, key % span
    else:  # cannot happen with 32-bit ids and < 2^30 levels; kept for safety
        order = np.lexsort((codes, level_id))
        codes, level_id = codes[order], level_id[order]
        keep = np.ones(len(codes), dtype=bool)
        keep[1:] = (codes[1:] != codes[:-1]) | (level_id[1:] != level_id[:-1])
        codes, level_id = codes[keep], level_id[keep]
    new_sizes = np.bincount(level_id, minlength=n_levels).astype(np.int64)
    new_off = np.zeros(n_levels + 1, dtype=np.int64)
    np.cumsum(new_sizes, out=new_off[1:])
    return new_off, codes


def pack_sets(*sides: Sequence[Sequence[Sequence[str]]]) -> List[PackedSets]:
    """Dictionary-encodes the token strings of all ``sides`` together (exact string identity)
    and packs each side.  ``sides[s][i][j]`` is the token list of level j of item i."""
    parts = [_csr_from_nested(s) for s in sides]
    all_tokens = [t for _, _, flat in parts for t in flat]
    if all_tokens:
        codes_all, uniques = pd.factorize(np.asarray(all_tokens, dtype=object))
        n_vocab = len(uniques)
    else:
        codes_all, n_vocab = np.zeros(0, dtype=np.int64), 0
    sides_codes, pos = [], 0
    for _, level_off, flat in parts:
        codes = codes_all[pos : pos + len(flat)].astype(np.int64)
        pos += len(flat)
        sides_codes.append(_sort_unique_levels(level_off, codes))
    ranked = rank_by_frequency([c for _, c in sides_codes], n_vocab)
    out = []
    for (item_level_off, _, _), (new_off, _), codes in zip(parts, sides_codes, ranked):
        new_off, codes = _sort_unique_levels(new_off, codes)  # ids changed: sort again
        out.append(finish_sets(item_level_off, new_off, codes, n_vocab))
    return out


def frequency_rank(id_arrays: Sequence[np.ndarray], n_vocab: int) -> np.ndarray:
    """rank[id] = position of ``id`` when ids are ordered by falling count over all arrays."""
    counts = np.zeros(n_vocab, dtype=np.int64)
    for a in id_arrays:
        counts += np.bincount(np.asarray(a, dtype=np.int64), minlength=n_vocab)
    order = np.argsort(-counts, kind="stable")
    rank = np.empty(n_vocab, dtype=np.int64)
    rank[order] = np.arange(n_vocab, dtype=np.int64)
    return rank


def pack_suffix_id_sets(lens: np.ndarray, flat_ids: np.ndarray, n_vocab: int,
                        rank: np.ndarray | None = None) -> PackedSets:
    """Integer fast path for list-valued columns whose parts are single tokens (``TokenIds``):
    item i has the id list ``v = flat_ids[o_i : o_i + lens[i]]`` and level j is ``set(v[-(j+1):])``
    (Q2), built without Python loops.  ``rank`` (see :func:`frequency_rank`) renumbers the ids;
    all sides of one comparison must be packed with the same ``rank``."""
    lens = np.asarray(lens, dtype=np.int64)
    if rank is not None:
        flat_ids = np.asarray(rank)[np.asarray(flat_ids, dtype=np.int64)]
    n = len(lens)
    item_end = np.cumsum(lens)
    item_level_off = np.zeros(n + 1, dtype=np.int64)
    item_level_off[1:] = item_end  # one level per list element
    n_levels = int(item_end[-1]) if n else 0
    level_item = np.repeat(np.arange(n, dtype=np.int64), lens)
    level_j = np.arange(n_levels, dtype=np.int64) - np.repeat(item_end - lens, lens)
    raw_sizes = level_j + 1
    raw_off = np.zeros(n_levels + 1, dtype=np.int64)
    np.cumsum(raw_sizes, out=raw_off[1:])
    # element e of level (i, j) is v[len_i - 1 - j + e]
    within = np.arange(int(raw_off[-1]), dtype=np.int64) - np.repeat(raw_off[:-1], raw_sizes)
    src = np.repeat(item_end[level_item] - 1 - level_j, raw_sizes) + within
    codes = np.asarray(flat_ids, dtype=np.int64)[src]
    new_off, codes = _sort_unique_levels(raw_off, codes)
    return finish_sets(item_level_off, new_off, codes, n_vocab)


def pack_part_id_sets(part_lens: np.ndarray, flat_ids: np.ndarray, n_vocab: int,
                      rank: np.ndarray | None = None) -> PackedSets:
    """Integer fast path for list-valued columns whose parts hold several tokens (``Term``):
    ``part_lens[i, q]`` ids of part q of item i follow one another in ``flat_ids`` (absent parts
    have length 0); level j is the id set of the last j+1 present parts (Q2)."""
    part_lens = np.asarray(part_lens, dtype=np.int64)
    n, n_parts = part_lens.shape
    flat_ids = np.asarray(flat_ids, dtype=np.int64)
    if rank is not None:
        flat_ids = np.asarray(rank)[flat_ids]
    present = part_lens > 0
    k = present.sum(axis=1)
    item_level_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(k, out=item_level_off[1:])
    item_len = part_lens.sum(axis=1)
    item_end = np.cumsum(item_len)
    # suffix sizes: ids in parts q..last, for every present part q, deepest suffix last
    suffix = np.cumsum(part_lens[:, ::-1], axis=1)[:, ::-1]          # ids from part q to the end
    # levels of item i, j = 0..k-1: the present parts taken from the back
    order = np.argsort(~present[:, ::-1], axis=1, kind="stable")     # present parts from the back, first
    back_idx = n_parts - 1 - order                                   # part index of the j-th present part from the back
    level_sizes_raw = np.take_along_axis(suffix, back_idx, axis=1)   # [n, n_parts]; valid for j < k
    valid = np.arange(n_parts)[None, :] < k[:, None]
    raw_sizes = level_sizes_raw[valid]
    level_item = np.repeat(np.arange(n, dtype=np.int64), k)
    raw_off = np.zeros(len(raw_sizes) + 1, dtype=np.int64)
    np.cumsum(raw_sizes, out=raw_off[1:])
    within = np.arange(int(raw_off[-1]), dtype=np.int64) - np.repeat(raw_off[:-1], raw_sizes)
    src = np.repeat(item_end[level_item] - raw_sizes, raw_sizes) + within
    codes = flat_ids[src]
    new_off, codes = _sort_unique_levels(raw_off, codes)
    return finish_sets(item_level_off, new_off, codes, n_vocab)


# ------------------------------------------------------------------------------------------
# strings
# ------------------------------------------------------------------------------------------
def fuzzy_level_strings(items_levels: Sequence[Sequence]) -> List[List[str]]:
    """Per level: ``default_process(join_sorted(tokens))`` (or of the str itself)."""
    return [[default_process(join_sorted(level) if isinstance(level, list) else level)
             for level in lv] for lv in items_levels]


def string_histograms(level_chr_off: np.ndarray, level_len: np.ndarray, chr_: np.ndarray) -> np.ndarray:
    """uint32[n_levels, 8]: per level string 32 byte counters, saturating at 255, of its codes;
    bucket = code & 31, counter of bucket b = byte (b & 3) of word b >> 2.  One insertion or deletion
    changes one counter by one, so the L1 distance of two strings' counters is a lower bound of
    their Indel distance (merging codes into buckets and saturating only lower it)."""
    n_levels = len(level_len)
    lens = level_len.astype(np.int64)
    counts = np.zeros((n_levels, 32), dtype=np.int64)
    if n_levels and lens.sum():
        level_id = np.repeat(np.arange(n_levels, dtype=np.int64), lens)
        src = np.repeat(level_chr_off.astype(np.int64), lens) + (
            np.arange(int(lens.sum()), dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens))
        np.add.at(counts, (level_id, chr_[src].astype(np.int64) & 31), 1)
    packed = np.minimum(counts, 255).astype(np.uint8).reshape(n_levels, 8, 4)
    return np.ascontiguousarray(packed).view("<u4").reshape(n_levels, 8)


def pack_strings(*sides: Sequence[Sequence[str]]) -> List[PackedStrings]:
    """``sides[s][i][j]`` is the already processed string of level j of item i.

    Alphabet: one code per code point while there are at most 255 of them.  Beyond that (two
    sides): code points that occur on BOTH sides keep a code each; a code point that occurs on one
    side only can never match anything of the other side, so all such code points of a side share
    one code (the left side's and the right side's differ).  255 codes therefore cover any pair of
    sides with at most 253 common code points."""
    per_side = []
    for s in sides:
        k = np.fromiter((len(lv) for lv in s), dtype=np.int64, count=len(s))
        # order items by the length of their longest level (stable): lanes of a warp then hold
        # patterns of nearly one length and a tile texts of nearly one length; the order is also
        # the order by 64-bit words, which the kernels' word classes need
        longest = np.fromiter((max((len(x) for x in lv), default=0) for lv in s), dtype=np.int64,
                              count=len(s))
        words = np.minimum(np.maximum((longest + 63) // 64, 1), WORD_CLASSES + 1)
        perm = np.argsort(longest, kind="stable")
        class_end = np.searchsorted(words[perm], np.arange(1, WORD_CLASSES + 1), side="right")
        strs = [x for i in perm for x in s[i]]
        lens = np.fromiter((len(x) for x in strs), dtype=np.int64, count=len(strs))
        cps = np.frombuffer("".join(strs).encode("utf-32-le", "surrogatepass"), dtype=np.uint32)
        per_side.append((k[perm], lens, cps, perm, class_end))
    sets = [np.unique(c[2]) for c in per_side]
    common = np.unique(np.concatenate(sets)) if sets else np.zeros(0, dtype=np.uint32)
    one_sided = [False] * len(sets)
    if len(common) > MAX_ALPHABET and len(sets) == 2:
        # too many code points for one byte each: only those COMMON to both sides need a code of
        # their own; all others of a side collapse onto one code of that side
        common = np.intersect1d(sets[0], sets[1], assume_unique=True)
        one_sided = [len(u) > len(common) for u in sets]
    n_alphabet = len(common) + sum(one_sided)
    if n_alphabet > MAX_ALPHABET:
        raise PackError(f"{len(common)} code points are common to both sides; the packed format allows "
                        f"{MAX_ALPHABET - 2}")
    out, extra = [], len(common)
    for (k, lens, cps, perm, class_end), own in zip(per_side, one_sided):
        item_level_off = np.zeros(len(k) + 1, dtype=np.int64)
        np.cumsum(k, out=item_level_off[1:])
        padded = _pad8(lens)
        level_chr_off = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(padded, out=level_chr_off[1:])
        if level_chr_off[-1] >= 2 ** 32:
            raise PackError("more than 4 GiB of level strings on one side")
        pos = np.searchsorted(common, cps)
        hit = (pos < len(common)) & (common[np.minimum(pos, max(len(common) - 1, 0))] == cps) \
            if len(common) else np.zeros(len(cps), dtype=bool)
        codes = np.where(hit, pos, extra).astype(np.uint8)
        if own:
            extra += 1
        chr_ = np.zeros(int(level_chr_off[-1]), dtype=np.uint8)
        if len(codes):
            src_off = np.zeros(len(lens) + 1, dtype=np.int64)
            np.cumsum(lens, out=src_off[1:])
            dest = np.repeat(level_chr_off[:-1] - src_off[:-1], lens) + np.arange(len(codes), dtype=np.int64)
            chr_[dest] = codes
        off32, len32 = level_chr_off[:-1].astype(np.uint32), lens.astype(np.uint32)
        out.append(PackedStrings(item_level_off.astype(np.uint32), off32, len32, chr_, int(n_alphabet),
                                 int(k.max()) if len(k) else 0,
                                 int(lens.max()) if len(lens) else 0,
                                 perm.astype(np.uint32), class_end.astype(np.uint32),
                                 string_histograms(off32, len32, chr_)))
    return out


def flat_items(values: Iterable) -> List[list]:
    """K = 1 items for the flat (scalar ``score_func``) entry points."""
    return [[v] for v in values]
