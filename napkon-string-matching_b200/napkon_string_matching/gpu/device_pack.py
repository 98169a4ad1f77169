"""
Device-side token packing: the CUDA counterpart of ``pack.finish_sets`` and of
the integer fast paths ``pack.pack_part_id_sets`` / ``pack.pack_suffix_id_sets``.

The host keeps what north_star assigns to it — tokenising and mapping token strings to integer
codes — and hands over a raw CSR of codes (``RawSets``: items -> groups -> ids, a group being one
*part* of the item's value or one explicit *level*).  ``nsm_pack_sets_measure`` /
``nsm_pack_sets_fill`` (csrc/pack.cu) then build every array of ``nsm_sets_t`` in device memory,
bit-identical to the numpy packer, and the cohort is ready for ``Engine.all_pairs`` without a
host round trip of the packed arrays.  No GPU -> raises (``Engine`` does).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from napkon_string_matching.gpu import lib as nsmlib
from napkon_string_matching.gpu.engine import DeviceCohort, Engine
from napkon_string_matching.gpu.pack import (HEAD_IDS, SLOT_BLOCK, SLOT_CAP, PackedSets, PackError,
                                             PackTooLarge)


@dataclass
class RawSets:
    """Dictionary codes of one cohort side: ids of group g are ``ids[grp_id_off[g]:grp_id_off[g+1]]``,
    groups of item i are ``item_grp_off[i]:item_grp_off[i+1]``.  ``mode`` says what a group is:
    ``RAW_SUFFIX_PARTS`` — one part of the item's value, level j = id set of the last j+1 parts
    (``items[-i:]``, /root/reference/napkon_string_matching/types/comparable_data.py:283-285);
    ``RAW_LEVELS`` — level j itself."""
    item_grp_off: np.ndarray
    grp_id_off: np.ndarray
    ids: np.ndarray
    mode: int = nsmlib.RAW_SUFFIX_PARTS
    pinned: Optional[List[torch.Tensor]] = None   # page-locked copies of the three arrays (pin())

    def pin(self) -> "RawSets":
        """Page-locks copies of the arrays so that the upload is one asynchronous DMA each."""
        self.pinned = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.uint8).reshape(-1)
                                        .copy()).pin_memory() if a.size else None
                       for a in (self.item_grp_off, self.grp_id_off, self.ids)]
        return self

    def nbytes(self) -> int:
        return 4 * (len(self.item_grp_off) + len(self.grp_id_off) + len(self.ids))

    @property
    def n_items(self) -> int:
        return len(self.item_grp_off) - 1

    @property
    def n_groups(self) -> int:
        return len(self.grp_id_off) - 1

    def max_levels(self) -> int:
        return int(np.diff(self.item_grp_off.astype(np.int64)).max(initial=0))

    def reordered(self, perm: np.ndarray) -> "RawSets":
        """The same items stored in the order ``perm`` (stored position -> item index)."""
        perm = np.asarray(perm, dtype=np.int64)
        igo, gio = self.item_grp_off.astype(np.int64), self.grp_id_off.astype(np.int64)
        n_grp = (igo[1:] - igo[:-1])[perm]
        new_igo = np.zeros(len(perm) + 1, dtype=np.int64)
        np.cumsum(n_grp, out=new_igo[1:])
        grp = np.repeat(igo[:-1][perm] - new_igo[:-1], n_grp) + np.arange(int(new_igo[-1]), dtype=np.int64)
        n_ids = (gio[1:] - gio[:-1])[grp]
        new_gio = np.zeros(len(grp) + 1, dtype=np.int64)
        np.cumsum(n_ids, out=new_gio[1:])
        ids = np.repeat(gio[:-1][grp] - new_gio[:-1], n_ids) + np.arange(int(new_gio[-1]), dtype=np.int64)
        return RawSets(new_igo.astype(np.uint32), new_gio.astype(np.uint32),
                       np.ascontiguousarray(self.ids[ids], dtype=np.uint32), self.mode)

    def ordered_by_levels(self, chunk: int):
        """``(raw, perm)``: the items in :func:`pack.chunked_level_order` (level-count chunks of
        the kernel's unit size); ``perm`` goes to ``DeviceCohort.perm`` so that the engine returns
        the caller's item indices."""
        from napkon_string_matching.gpu.pack import chunked_level_order

        perm = chunked_level_order(np.diff(self.item_grp_off.astype(np.int64)), chunk)
        return self.reordered(perm), perm


def _offsets(lens: np.ndarray) -> np.ndarray:
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    if off[-1] >= 2 ** 32:
        raise PackError("more than 2^32 rows on one side")
    return off.astype(np.uint32)


def raw_from_parts(part_lens: np.ndarray, flat_ids: np.ndarray) -> RawSets:
    """Same input convention as :func:`pack.pack_part_id_sets`: ``part_lens[i, q]`` ids of part q
    of item i follow one another in ``flat_ids``; parts of length 0 are absent."""
    part_lens = np.asarray(part_lens, dtype=np.int64)
    present = part_lens > 0
    return RawSets(_offsets(present.sum(axis=1)), _offsets(part_lens[present]),
                   np.ascontiguousarray(flat_ids, dtype=np.uint32), nsmlib.RAW_SUFFIX_PARTS)


def raw_from_id_lists(lens: np.ndarray, flat_ids: np.ndarray) -> RawSets:
    """Same input convention as :func:`pack.pack_suffix_id_sets`: item i is a list of ``lens[i]``
    single-token parts (the ``TokenIds`` column)."""
    lens = np.asarray(lens, dtype=np.int64)
    n_ids = int(lens.sum())
    return RawSets(_offsets(lens), np.arange(n_ids + 1, dtype=np.uint32),
                   np.ascontiguousarray(flat_ids, dtype=np.uint32), nsmlib.RAW_SUFFIX_PARTS)


def raw_from_levels(items_levels: Sequence[Sequence[Sequence[int]]]) -> RawSets:
    """Explicit levels: ``items_levels[i][j]`` is the code list of level j of item i."""
    k = [len(lv) for lv in items_levels]
    sizes = [len(level) for lv in items_levels for level in lv]
    flat = [c for lv in items_levels for level in lv for c in level]
    return RawSets(_offsets(np.asarray(k, dtype=np.int64)), _offsets(np.asarray(sizes, dtype=np.int64)),
                   np.asarray(flat, dtype=np.uint32), nsmlib.RAW_LEVELS)


_SETS_FIELDS = ("item_level_off", "level_tok_off", "tok", "tok_entry", "level_head", "level_tail",
                "level_tail2", "level_info", "item_any", "item_k", "slot_ht", "slot_info")


class DevicePacker:
    """Packs raw code CSRs on the engine's device.  All sides of one comparison must go through
    one call of :meth:`pack` (they share the frequency ranking, like ``pack.pack_sets``)."""

    def __init__(self, engine: Engine):
        self.engine = engine
        self.lib = engine.lib
        self.device = engine.device
        self.last_rank: Optional[np.ndarray] = None

    def _dev(self, arr: np.ndarray, dtype) -> torch.Tensor:
        arr = np.ascontiguousarray(arr, dtype=dtype)
        if arr.size == 0:
            arr = np.zeros(1, dtype=dtype)
        return torch.from_numpy(arr.view(np.uint8).reshape(-1)).to(self.device)

    def _empty(self, n_bytes: int) -> torch.Tensor:
        return torch.empty(max(int(n_bytes), 16), dtype=torch.uint8, device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def frequency_rank(self, raws: List[C.Structure], n_vocab: int) -> np.ndarray:
        """``pack.frequency_rank`` with the histogram taken on the device: rank[id] = position of
        id when ids are ordered by falling count over all sides (ties: smaller id first)."""
        counts = torch.zeros(max(n_vocab, 1), dtype=torch.int32, device=self.device)
        for st in raws:
            nsmlib.check(self.lib.nsm_pack_count_ids(C.byref(st), counts.data_ptr(), self._stream()))
            self.engine.launches += self.lib.nsm_last_launch_count()
        host = counts.cpu().numpy().view(np.uint32)[:n_vocab].astype(np.int64)
        order = np.argsort(-host, kind="stable")
        rank = np.empty(n_vocab, dtype=np.int64)
        rank[order] = np.arange(n_vocab, dtype=np.int64)
        return rank

    def pack(self, sides: Sequence[RawSets], n_vocab: int,
             rank: Union[None, str, np.ndarray] = "frequency") -> List[DeviceCohort]:
        """``rank``: "frequency" (ids renumbered by falling frequency over all sides, what
        ``pack.pack_sets`` does), an explicit renumbering array, or None (ids used as they are)."""
        with torch.cuda.device(self.device):
            return self._pack(sides, n_vocab, rank)

    def _pack(self, sides, n_vocab, rank):
        staged = []
        for raw in sides:
            if raw.n_items >= 2 ** 32 - 1 or len(raw.ids) >= 2 ** 32:
                raise PackError("more than 2^32 rows on one side")
            arrays = (raw.item_grp_off, raw.grp_id_off, raw.ids)
            pinned = raw.pinned or (None, None, None)
            tensors = [pin.to(self.device, non_blocking=True) if pin is not None else self._dev(a, np.uint32)
                       for a, pin in zip(arrays, pinned)]
            st = nsmlib.NsmRawSets(tensors[0].data_ptr(), tensors[1].data_ptr(), tensors[2].data_ptr(),
                                   None, raw.n_items, raw.n_groups, len(raw.ids), n_vocab, raw.mode, 0)
            staged.append((raw, st, tensors))
        rank_dev = None
        if isinstance(rank, str):
            if rank != "frequency":
                raise ValueError(rank)
            rank = self.frequency_rank([st for _, st, _ in staged], n_vocab)
        if rank is not None:
            self.last_rank = np.asarray(rank)
            rank_dev = self._dev(rank, np.uint32)
        out = []
        for raw, st, tensors in staged:
            if rank_dev is not None:
                st.rank = rank_dev.data_ptr()
            out.append(self._pack_one(raw, st, n_vocab))
        return out

    def _pack_one(self, raw: RawSets, st, n_vocab: int) -> DeviceCohort:
        n, n_levels = raw.n_items, raw.n_groups
        item_tok_off = self._empty((n + 1) * 4)
        totals = self._empty(16)
        scratch_bytes = int(self.lib.nsm_pack_scratch_bytes(n))
        scratch = self._empty(scratch_bytes)
        nsmlib.check(self.lib.nsm_pack_sets_measure(C.byref(st), item_tok_off.data_ptr(), totals.data_ptr(),
                                                    scratch.data_ptr(), scratch_bytes, self._stream()))
        self.engine.launches += self.lib.nsm_last_launch_count()
        n_tok, flags = (int(x) for x in totals.cpu().numpy().view(np.uint64)[:2])
        self._raise_for(flags)
        if n_tok >= 2 ** 32:
            raise PackError("more than 2^32 token rows on one side")
        max_k = raw.max_levels()
        n_slots = min(max(max_k - 1, 1), SLOT_CAP)
        stride = max(SLOT_BLOCK, (n + SLOT_BLOCK - 1) // SLOT_BLOCK * SLOT_BLOCK)
        exact_bits = n_vocab <= 2 * HEAD_IDS
        sizes = {"item_level_off": (n + 1) * 4, "level_tok_off": (n_levels + 1) * 4, "tok": n_tok * 4,
                 "tok_entry": n_tok, "level_head": n_levels * 8, "level_tail": n_levels * 8,
                 "level_tail2": n_levels * 8, "level_info": n_levels * 4, "item_any": n * 16,
                 "item_k": n * 4, "slot_ht": n_slots * stride * 16, "slot_info": n_slots * stride * 4}
        tensors = [self._empty(sizes[f]) for f in _SETS_FIELDS]
        sets = nsmlib.NsmSets(*[t.data_ptr() for t in tensors], n, n_levels, max_k, n_slots,
                              int(exact_bits), stride, 0, 0)
        flags_dev = self._empty(4)
        nsmlib.check(self.lib.nsm_pack_sets_fill(C.byref(st), item_tok_off.data_ptr(), C.byref(sets),
                                                 flags_dev.data_ptr(), self._stream()))
        self.engine.launches += self.lib.nsm_last_launch_count()
        off = item_tok_off.cpu().numpy().view(np.uint32)[: n + 1].astype(np.int64)
        flags = int(flags_dev.cpu().numpy().view(np.uint32)[0])
        self._raise_for(flags)
        sets.nested = int(n_tok == 0 or (not (flags & nsmlib.PACK_FLAG_NOT_NESTED) and max_k <= 255))
        weights = np.diff(off).astype(np.float64) + 1.0
        return DeviceCohort("sets", sets, tensors, n, max_k, sum(sizes.values()), weights, None,
                            n_vocab, sizes)

    @staticmethod
    def _raise_for(flags: int) -> None:
        if flags & nsmlib.PACK_FLAG_BAD_ID:
            raise PackError("a token code is >= n_vocab")
        if flags & nsmlib.PACK_FLAG_TOO_LARGE:
            raise PackTooLarge(f"an item holds more than {nsmlib.PACK_MAX_ITEM_IDS} ids (or 65535 levels); "
                               "pack this cohort with gpu.pack on the host")


_DTYPES = {"item_level_off": np.uint32, "level_tok_off": np.uint32, "tok": np.uint32,
           "tok_entry": np.uint8, "level_head": np.uint64, "level_tail": np.uint64,
           "level_tail2": np.uint64, "level_info": np.uint32, "item_any": np.uint64,
           "item_k": np.uint32, "slot_ht": np.uint64, "slot_info": np.uint32}


def to_host(cohort: DeviceCohort) -> PackedSets:
    """Copies a device-packed cohort back as a ``PackedSets`` (tests, inspection)."""
    st = cohort.struct
    arrays = {}
    for f, t in zip(_SETS_FIELDS, cohort.tensors):
        n_bytes = cohort.sizes[f]
        arrays[f] = t[:n_bytes].cpu().numpy().view(_DTYPES[f]).copy() if n_bytes else \
            np.zeros(0, dtype=_DTYPES[f])
    arrays["item_any"] = arrays["item_any"].reshape(-1, 2)
    arrays["slot_ht"] = arrays["slot_ht"].reshape(st.n_slots, st.slot_stride, 2)
    arrays["slot_info"] = arrays["slot_info"].reshape(st.n_slots, st.slot_stride)
    return PackedSets(*[arrays[f] for f in _SETS_FIELDS], cohort.n_vocab, bool(st.exact_bits),
                      int(st.max_levels), bool(st.nested))


# ------------------------------------------------------------------------------------------
# strings (fuzzy_match): csrc/pack_strings.cu
# ------------------------------------------------------------------------------------------
class PackUnsupported(PackError):
    """The device string packer cannot reproduce ``default_process`` for this input (a code point
    whose lower-case form is not one code point, or a capital sigma, whose lower-case form depends
    on its position in the word); the host packer can."""


def processed_code_points(cps: np.ndarray):
    """``(processed, ok)``: per distinct code point what ``text.process.default_process`` turns it
    into — a blank when it is not alphanumeric (regex ``\\W``), else its lower-case form."""
    from napkon_string_matching.text.process import _NON_ALNUM

    out = np.empty(len(cps), dtype=np.uint32)
    for i, c in enumerate(cps.tolist()):
        ch = chr(c)
        if c == 0x3A3:   # final-sigma rule: str.lower() looks at the neighbours
            return out, False
        low = " " if _NON_ALNUM.match(ch) else ch.lower()
        if len(low) != 1:
            return out, False
        out[i] = ord(low)
    return out, True


class DeviceStringPacker:
    """``pack.pack_strings(fuzzy_level_strings(...))`` with the per-character work on the GPU.

    Input per side: the UNPROCESSED level strings (``join_sorted`` of a token list, or the str
    itself), ``sides[s][i][j]`` = level j of item i.  All sides of a comparison go through one call:
    they share the alphabet."""

    def __init__(self, engine: Engine):
        self.engine, self.lib, self.device = engine, engine.lib, engine.device

    def _dev(self, arr: np.ndarray, dtype) -> torch.Tensor:
        arr = np.ascontiguousarray(arr, dtype=dtype)
        if arr.size == 0:
            arr = np.zeros(1, dtype=dtype)
        if not arr.flags.writeable:   # np.frombuffer over the encoded text: torch wants a writable source
            arr = arr.copy()
        return torch.from_numpy(arr.view(np.uint8).reshape(-1)).to(self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def pack(self, sides: Sequence[Sequence[Sequence[str]]]) -> List[DeviceCohort]:
        with torch.cuda.device(self.device):
            return self._pack(sides)

    def _pack(self, sides):
        from napkon_string_matching.gpu.pack import MAX_ALPHABET, WORD_CLASSES, _pad8

        staged = []
        for s in sides:
            k = np.fromiter((len(lv) for lv in s), dtype=np.int64, count=len(s))
            strs = [x for lv in s for x in lv]
            raw_len = np.fromiter((len(x) for x in strs), dtype=np.int64, count=len(strs))
            cps = np.frombuffer("".join(strs).encode("utf-32-le", "surrogatepass"), dtype=np.uint32)
            if len(cps) >= 2 ** 32:
                raise PackError("more than 2^32 characters on one side")
            counts = np.bincount(cps) if len(cps) else np.zeros(0, dtype=np.int64)
            staged.append((k, raw_len, cps, counts))
        # what is per distinct code point stays with Python's Unicode tables
        table_len = max((len(c[3]) for c in staged), default=0)
        present = np.zeros(table_len, dtype=bool)
        for _, _, _, counts in staged:
            present[: len(counts)] |= counts > 0
        uniq = np.nonzero(present)[0].astype(np.uint32)
        processed, ok = processed_code_points(uniq)
        if not ok:
            raise PackUnsupported("a code point whose lower-case form is position dependent or longer "
                                  "than one code point")
        symbols = np.unique(np.concatenate([processed, np.array([0x20], dtype=np.uint32)]))
        if len(symbols) >= nsmlib.STR_SYM_NONE:
            raise PackError("more than 65534 distinct processed code points")
        cp_sym = np.full(max(table_len, 1), nsmlib.STR_SYM_NONE, dtype=np.uint16)
        cp_sym[uniq] = np.searchsorted(symbols, processed).astype(np.uint16)
        blank = int(np.searchsorted(symbols, 0x20))
        cp_sym_dev = self._dev(cp_sym, np.uint16)

        measured = []
        for k, raw_len, cps, counts in staged:
            level_off = np.zeros(len(raw_len) + 1, dtype=np.int64)
            np.cumsum(raw_len, out=level_off[1:])
            t_off, t_cps = self._dev(level_off, np.uint32), self._dev(cps, np.uint32)
            n_levels = len(raw_len)
            first = torch.empty(max(n_levels, 1), dtype=torch.int32, device=self.device)
            length = torch.empty(max(n_levels, 1), dtype=torch.int32, device=self.device)
            flags = torch.zeros(1, dtype=torch.int32, device=self.device)
            st = nsmlib.NsmRawStrings(t_off.data_ptr(), t_cps.data_ptr(), cp_sym_dev.data_ptr(), n_levels,
                                      len(cps), len(cp_sym), blank)
            nsmlib.check(self.lib.nsm_pack_strings_measure(C.byref(st), first.data_ptr(), length.data_ptr(),
                                                           flags.data_ptr(), self._stream()))
            self.engine.launches += self.lib.nsm_last_launch_count()
            measured.append((st, (t_off, t_cps), first, length, flags))
        sides_out, sets = [], []
        for (k, raw_len, cps, counts), (st, keep, first, length, flags) in zip(staged, measured):
            lens = length.cpu().numpy().view(np.uint32)[: len(raw_len)].astype(np.int64)
            if int(flags.cpu().numpy()[0]) & nsmlib.STR_FLAG_UNMAPPED:
                raise PackError("a code point without a processed form (internal error)")
            # the processed code points this side still holds after the trim (only blanks are trimmed)
            sym_count = np.zeros(len(symbols), dtype=np.int64)
            np.add.at(sym_count, cp_sym[: len(counts)][counts > 0].astype(np.int64), counts[counts > 0])
            sym_count[blank] -= int((raw_len - lens).sum())
            sets.append(symbols[sym_count > 0])
            sides_out.append(lens)
        common = np.unique(np.concatenate(sets)) if sets else np.zeros(0, dtype=np.uint32)
        one_sided = [False] * len(sets)
        if len(common) > MAX_ALPHABET and len(sets) == 2:
            common = np.intersect1d(sets[0], sets[1], assume_unique=True)
            one_sided = [len(u) > len(common) for u in sets]
        n_alphabet = len(common) + sum(one_sided)
        if n_alphabet > MAX_ALPHABET:
            raise PackError(f"{len(common)} code points are common to both sides; the packed format allows "
                            f"{MAX_ALPHABET - 2}")
        out, extra = [], len(common)
        for (k, raw_len, cps, counts), (st, keep, first, length, flags), lens, own in zip(
                staged, measured, sides_out, one_sided):
            n = len(k)
            item_off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(k, out=item_off[1:])
            longest = np.zeros(n, dtype=np.int64)
            has = k > 0
            if len(lens) and has.any():
                longest[has] = np.maximum.reduceat(lens, item_off[:-1][has])
            words = np.minimum(np.maximum((longest + 63) // 64, 1), WORD_CLASSES + 1)
            perm = np.argsort(longest, kind="stable")
            class_end = np.searchsorted(words[perm], np.arange(1, WORD_CLASSES + 1), side="right")
            kp = k[perm]
            new_item_off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(kp, out=new_item_off[1:])
            src_level = np.repeat(item_off[:-1][perm] - new_item_off[:-1], kp) + np.arange(int(new_item_off[-1]),
                                                                                         dtype=np.int64)
            stored_len = lens[src_level] if len(lens) else np.zeros(0, dtype=np.int64)
            level_chr_off = np.zeros(len(stored_len) + 1, dtype=np.int64)
            np.cumsum(_pad8(stored_len), out=level_chr_off[1:])
            if level_chr_off[-1] >= 2 ** 32:
                raise PackError("more than 4 GiB of level strings on one side")
            pos = np.searchsorted(common, symbols)
            hit = (pos < len(common)) & (common[np.minimum(pos, max(len(common) - 1, 0))] == symbols) \
                if len(common) else np.zeros(len(symbols), dtype=bool)
            sym_code = np.where(hit, pos, extra).astype(np.uint8)
            if own:
                extra += 1
            n_levels = len(stored_len)
            t_item = self._dev(new_item_off, np.uint32)
            t_off = self._dev(level_chr_off[:-1], np.uint32)
            t_len = self._dev(stored_len, np.uint32)
            t_src = self._dev(src_level, np.uint32)
            t_code = self._dev(sym_code, np.uint8)
            chr_ = torch.empty(max(int(level_chr_off[-1]), 16), dtype=torch.uint8, device=self.device)
            hist = torch.empty(max(n_levels, 1) * 32, dtype=torch.uint8, device=self.device)
            nsmlib.check(self.lib.nsm_pack_strings_fill(C.byref(st), first.data_ptr(), t_src.data_ptr(),
                                                        t_off.data_ptr(), t_len.data_ptr(), t_code.data_ptr(),
                                                        n_levels, chr_.data_ptr(), hist.data_ptr(), self._stream()))
            self.engine.launches += self.lib.nsm_last_launch_count()
            struct = nsmlib.NsmStrings(t_item.data_ptr(), t_off.data_ptr(), t_len.data_ptr(), chr_.data_ptr(),
                                       hist.data_ptr(), n, n_levels, int(k.max()) if n else 0,
                                       int(stored_len.max()) if n_levels else 0, int(n_alphabet), 0,
                                       (C.c_uint32 * 8)(*[int(x) for x in class_end]))
            csum = np.concatenate([[0], np.cumsum(stored_len)])
            weights = (csum[new_item_off[1:]] - csum[new_item_off[:-1]]).astype(np.float64) + 1.0
            sizes = {"item_level_off": (n + 1) * 4, "level_chr_off": n_levels * 4, "level_len": n_levels * 4,
                     "chr": int(level_chr_off[-1]), "level_hist": n_levels * 32}
            out.append(DeviceCohort("strings", struct, [t_item, t_off, t_len, chr_, hist], n,
                                    int(k.max()) if n else 0, 4 * (len(cps) + len(raw_len) + 1), weights,
                                    perm.astype(np.uint32), 0, sizes))
        # the raw inputs must outlive the fill kernels (same stream: they do once these are queued)
        torch.cuda.current_stream(self.device).synchronize()
        return out


def strings_to_host(cohort: DeviceCohort):
    """Copies a device-packed string cohort back as a ``PackedStrings`` (tests, inspection)."""
    from napkon_string_matching.gpu.pack import PackedStrings

    st = cohort.struct
    dt = {"item_level_off": np.uint32, "level_chr_off": np.uint32, "level_len": np.uint32, "chr": np.uint8,
          "level_hist": np.uint32}
    arr = {}
    for f, t in zip(dt, cohort.tensors):
        n_bytes = cohort.sizes[f]
        arr[f] = t[:n_bytes].cpu().numpy().view(dt[f]).copy() if n_bytes else np.zeros(0, dtype=dt[f])
    return PackedStrings(arr["item_level_off"], arr["level_chr_off"], arr["level_len"], arr["chr"],
                         int(st.n_alphabet), int(st.max_levels), int(st.max_len), cohort.perm,
                         np.array(list(st.class_end), dtype=np.uint32), arr["level_hist"].reshape(-1, 8))
