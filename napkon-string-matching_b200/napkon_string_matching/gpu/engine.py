"""
Drives the CUDA comparison kernels: uploads packed cohorts, runs the all-pairs job for a left row
block, and returns the kept ``(left, right, score)`` records as one numpy array.

PyTorch is used for device memory, pinned host memory and streams only.  No GPU -> raises.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from napkon_string_matching.gpu import lib as nsmlib
from napkon_string_matching.gpu.pack import PackedSets, PackedStrings

PAIR_DTYPE = nsmlib.PAIR_DTYPE


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise nsmlib.NsmError(
            "no CUDA device: the comparison path runs on the GPU only (no CPU fallback)")


@dataclass
class DeviceCohort:
    kind: str                       # "sets" | "strings"
    struct: C.Structure
    tensors: List[torch.Tensor]
    n_items: int
    max_levels: int
    h2d_bytes: int
    weights: np.ndarray             # per-item work estimate, for row-block balancing
    perm: Optional[np.ndarray] = None  # stored position -> caller's item index (strings)
    n_vocab: int = 0                   # device-packed cohorts (gpu/device_pack.py) ...
    sizes: Optional[Dict[str, int]] = None  # ... and the byte size of each of their arrays


@dataclass
class Job:
    """One all-pairs comparison: left[rows] x right, keep score >= threshold."""
    left: DeviceCohort
    right: DeviceCohort
    threshold: float
    flat: bool = False
    rows: Optional[Tuple[int, int]] = None
    l_cat: Optional[torch.Tensor] = None
    r_cat: Optional[torch.Tensor] = None
    cat_mode: int = nsmlib.CAT_OFF


class Records:
    """Kept pairs of one job as they arrived from the device: views of the engine's pinned arena
    (valid until its next call), part by part in the wire format of the part — ``PAIR_DTYPE``
    records or ``nsm_packet_t`` packets (include/nsm.h)."""

    def __init__(self, parts, count: int, left_perm=None, right_perm=None, dictionary=None):
        self.parts, self.count = parts, int(count)
        self.left_perm, self.right_perm = left_perm, right_perm
        self.dictionary = dictionary   # uint64[DICT_SLOTS]: the scores behind NSM_OUT_CODED parts

    def __len__(self) -> int:
        return self.count

    @property
    def nbytes(self) -> int:
        return sum(a.nbytes for _, a in self.parts) + (self.dictionary.nbytes if self.dictionary is not None else 0)

    def decode(self, copy: bool = True) -> np.ndarray:
        """All parts as one ``PAIR_DTYPE`` array with the callers' item indices."""
        if not self.parts:
            return np.zeros(0, dtype=PAIR_DTYPE)
        if len(self.parts) == 1 and self.parts[0][0] == nsmlib.OUT_PAIRS and self.left_perm is None \
                and self.right_perm is None:
            return self.parts[0][1].copy() if copy else self.parts[0][1]
        # one pass per part through the library's host decoders (packets -> 16-byte records, the
        # stored-position -> item-index maps applied on the way)
        total = sum(len(a) * nsmlib.ENTRY_RECORDS[mode] for mode, a in self.parts)
        out = np.empty(total, dtype=PAIR_DTYPE)
        pos = 0
        for mode, a in self.parts:
            pos += nsmlib.decode_into(out[pos:], mode, a, self.dictionary, self.left_perm, self.right_perm)
        return out[:pos]


class Engine:
    PIPELINE_BLOCK_BYTES = 160 << 20   # record bytes per row block: the last block's copy-out is exposed
    PIPELINE_MIN_PAIRS = 1 << 26   # smaller jobs are one launch: a probe costs more than it saves

    def __init__(self, device: Optional[int] = None, max_pairs_per_block: int = 1 << 30):
        _require_cuda()
        self.lib = nsmlib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.max_pairs_per_block = int(max_pairs_per_block)
        self.pipeline_d2h = True    # probe + row blocks for results that go to the host
        # record format of token-set results that go to the host (include/nsm.h NSM_OUT_*):
        # "auto": coded packets (4.3 B per kept pair) when the probe block finds the result dense
        # (>= 2 packets per warp and unit), else 16-byte pairs; "coded" / True or "packets": that
        # format for every token-set job; False: always 16-byte pairs
        self.compact = "auto"
        # results of large jobs that go to the host: the kernels store the records STRAIGHT into the
        # engine's page-locked host arena (zero-copy stores over PCIe/C2C) — no device arena, no
        # device->host copy to wait for, no host round trip per row block; see _run_jobs_direct
        self.direct_host = True
        self._buffers: Dict[str, torch.Tensor] = {}
        self.launches = 0           # kernels of ours launched so far
        self.time_kernels = False   # bracket every comparison kernel with CUDA events
        self.kernel_ms = 0.0
        self.kernel_launches_timed = 0
        self._timed: List[Tuple[torch.cuda.Event, torch.cuda.Event]] = []
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self.last_infos: List[Dict] = []
        self.last_info: Dict = {}
        self._packer = None
        self._string_packer = None
        self.trace: Optional[List[Tuple[float, str]]] = None   # set to [] to collect (host time, label)

    @property
    def device_packer(self):
        """Device-side token packing (gpu/device_pack.py), bound to this engine's device."""
        if self._packer is None:
            from napkon_string_matching.gpu.device_pack import DevicePacker

            self._packer = DevicePacker(self)
        return self._packer

    @property
    def string_packer(self):
        """Device-side string packing (gpu/device_pack.py:DeviceStringPacker) on this engine's device."""
        if self._string_packer is None:
            from napkon_string_matching.gpu.device_pack import DeviceStringPacker

            self._string_packer = DeviceStringPacker(self)
        return self._string_packer

    # ------------------------------------------------------------------ uploads
    def _to_device(self, arr) -> torch.Tensor:
        if isinstance(arr, torch.Tensor):  # already a (pinned) byte tensor
            return arr.to(self.device, non_blocking=True)
        arr = np.ascontiguousarray(arr)
        if arr.size == 0:  # keep a valid pointer
            arr = np.zeros(1, dtype=arr.dtype)
        host = torch.from_numpy(arr.view(np.uint8).reshape(-1))
        return host.to(self.device, non_blocking=False)

    @staticmethod
    def _arrays(packed) -> List[np.ndarray]:
        if isinstance(packed, PackedSets):
            return packed.arrays()
        if isinstance(packed, PackedStrings):
            return packed.arrays()
        raise TypeError(type(packed))

    def pin(self, packed) -> List[torch.Tensor]:
        """Page-locked copies of a pack's arrays, so that :meth:`upload` is one async DMA each."""
        out = []
        for a in self._arrays(packed):
            a = np.ascontiguousarray(a)
            if a.size == 0:
                a = np.zeros(1, dtype=a.dtype)
            out.append(torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).pin_memory())
        return out

    def upload(self, packed, pinned: Optional[List[torch.Tensor]] = None) -> DeviceCohort:
        arrays = self._arrays(packed)
        tensors = [self._to_device(a) for a in (pinned if pinned is not None else arrays)]
        if isinstance(packed, PackedSets):
            st = nsmlib.NsmSets(*[t.data_ptr() for t in tensors], packed.n_items, packed.n_levels,
                                packed.max_levels, packed.n_slots, int(packed.exact_bits), packed.slot_stride,
                                int(packed.nested), 0)
            per_level = packed.level_sizes()
            kind = "sets"
        else:
            st = nsmlib.NsmStrings(*[t.data_ptr() for t in tensors[:5]], packed.n_items,
                                   packed.n_levels, packed.max_levels, packed.max_len,
                                   packed.n_alphabet, 0,
                                   (C.c_uint32 * 8)(*[int(x) for x in packed.classes()]))
            per_level = packed.level_lengths()
            kind = "strings"
        # per-item work estimate (sum of level sizes + 1), used to balance row blocks over GPUs
        csum = np.concatenate([[0], np.cumsum(per_level, dtype=np.int64)])
        off = packed.item_level_off.astype(np.int64)
        per_item = (csum[off[1:]] - csum[off[:-1]]).astype(np.float64) + 1.0
        return DeviceCohort(kind, st, tensors, packed.n_items, packed.max_levels,
                            sum(a.nbytes for a in arrays), per_item, getattr(packed, "perm", None))

    def upload_masks(self, masks: Optional[np.ndarray]) -> Optional[torch.Tensor]:
        if masks is None:
            return None
        return self._to_device(np.ascontiguousarray(masks, dtype=np.uint64))

    # ------------------------------------------------------------------ buffers
    def _arena(self, name: str, n_bytes: int, pinned: bool) -> torch.Tensor:
        """Grow-only byte arenas, kept across calls so that steady-state calls neither allocate
        device memory nor page-lock host memory."""
        cur = self._buffers.get(name)
        if cur is None or cur.numel() < n_bytes:
            self._buffers.pop(name, None)
            if pinned:
                cur = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
            else:
                cur = torch.empty(n_bytes, dtype=torch.uint8, device=self.device)
            self._buffers[name] = cur
        return cur

    # ------------------------------------------------------------------ the job
    def all_pairs(self, left: DeviceCohort, right: DeviceCohort, threshold: float, *,
                  flat: bool = False, rows: Optional[Tuple[int, int]] = None,
                  l_cat: Optional[torch.Tensor] = None, r_cat: Optional[torch.Tensor] = None,
                  cat_mode: int = nsmlib.CAT_OFF, capacity: Optional[int] = None,
                  to_host: bool = True, copy: bool = True) -> np.ndarray:
        """Scores left[rows] x right and returns the kept records (``PAIR_DTYPE``), in no
        particular order.  See :meth:`run_jobs`; counters are in ``self.last_info``."""
        job = Job(left, right, threshold, flat=flat, rows=rows, l_cat=l_cat, r_cat=r_cat,
                  cat_mode=cat_mode)
        out = self.run_jobs([job], capacity=capacity, to_host=to_host, copy=copy)[0]
        self.last_info = self.last_infos[0]
        return out

    def run_jobs(self, jobs: List["Job"], *, capacity: Optional[int] = None, to_host: bool = True,
                 copy: bool = True, decode: bool = True):
        """Runs several all-pairs jobs back to back.

        One kernel launch covers a whole job; kept records are compacted into one of two device
        arenas, so the device->host copy of job k (on a second stream, into the engine's pinned
        arena) overlaps the kernel of job k+1.  If an arena was too small the kernel still counts
        exactly: the arena is grown to the exact need and the launch repeated once.  A job that
        keeps more than ``max_pairs_per_block`` records is split into left row blocks.
        ``to_host=False`` leaves the records on the device (kernel-only timing) and returns empty
        arrays; ``copy=False`` returns views of the pinned arena (valid until the next call);
        ``decode=False`` returns :class:`Records` (the parts in their wire format) instead of
        ``PAIR_DTYPE`` arrays."""
        # the C ABI launches on the calling thread's current device: make it this engine's
        with torch.cuda.device(self.device):
            if to_host and self.direct_host and self.pipeline_d2h and capacity is None:
                return self._run_jobs_direct(jobs, copy, decode)
            return self._run_jobs(jobs, capacity, to_host, copy, decode)

    def _mark(self, label: str) -> None:
        if self.trace is not None:
            import time

            self.trace.append((time.perf_counter(), label))

    def _run_jobs(self, jobs: List["Job"], capacity: Optional[int], to_host: bool, copy: bool,
                  decode: bool = True):
        stream = torch.cuda.current_stream(self.device)
        self._mark("run_jobs begin")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        copy_stream = self._copy_stream
        ctl = [self._arena(f"ctl{i}", 64, pinned=False) for i in range(2)]
        ctl_pin = [self._arena(f"ctl_pin{i}", 64, pinned=True) for i in range(2)]
        slot_free: List[Optional[torch.cuda.Event]] = [None, None]   # D2H out of the arena done
        rec_bytes, entry_records = nsmlib.RECORD_BYTES, nsmlib.ENTRY_RECORDS
        dicts: Dict[int, torch.Tensor] = {}      # job -> device score dictionary (NSM_OUT_CODED)
        dict_host: Dict[int, int] = {}           # job -> offset of its copy in the pinned arena
        dict_bytes = nsmlib.DICT_SLOTS * 8

        def n_units(rows: int, n_right: int) -> int:
            return -(-rows // nsmlib.UNIT_LEFT) * -(-n_right // nsmlib.UNIT_RIGHT)

        def capacity_for(mode: int, records: float, rows: int, n_right: int) -> int:
            """Arena entries (pairs or packets) that hold `records` kept pairs of a row block."""
            records = min(float(self.max_pairs_per_block), records)
            if mode != nsmlib.OUT_PAIRS:   # full packets + at most one partial per warp and unit
                return int(records / entry_records[mode]) + 4 * n_units(rows, n_right) + 64
            return max(int(records) + 4096, 1 << 18)

        # Work items: row blocks of the jobs.  A job whose records go to the host starts with a
        # small PROBE block: its kept-pair density sizes the arenas of the rest exactly (no
        # overflow re-run), picks the record format (16-byte pairs, or packets when the result is
        # dense enough to be bound by the device->host link) and cuts the rest into blocks so that
        # the copy-out of one block runs beside the kernel of the next.
        work: List[dict] = []
        infos = []
        for j, job in enumerate(jobs):
            if job.left.kind != job.right.kind:
                raise TypeError("left and right must be packed for the same score function")
            begin, end = job.rows if job.rows is not None else (0, job.left.n_items)
            infos.append({"count": 0, "flags": 0, "reruns": 0, "blocks": 0, "d2h_bytes": 0,
                          "packets": 0, "uncoded": 0, "stats": dict.fromkeys(nsmlib.STAT_NAMES, 0),
                          "item_pairs": max(0, end - begin) * job.right.n_items, "parts": []})
            if end <= begin or not job.right.n_items:
                continue
            rows = end - begin
            forced = nsmlib.OUT_PAIRS
            if job.left.kind == "sets" and self.compact in (True, "coded", "packets"):
                forced = nsmlib.OUT_PACKETS if self.compact == "packets" else nsmlib.OUT_CODED
            if (to_host and self.pipeline_d2h and capacity is None and rows >= 4 * nsmlib.UNIT_LEFT
                    and rows * job.right.n_items >= self.PIPELINE_MIN_PAIRS):
                cut = begin + max(nsmlib.UNIT_LEFT, rows // 16 // nsmlib.UNIT_LEFT * nsmlib.UNIT_LEFT)
                work.append({"j": j, "rb": begin, "re": cut, "mode": nsmlib.OUT_PAIRS, "probe": True})
                work.append({"j": j, "rb": cut, "re": end, "mode": forced, "rest": True})
            else:
                work.append({"j": j, "rb": begin, "re": end, "mode": forced})
        self.last_infos = infos

        def launch(slot: int, item: dict):
            job = jobs[item["j"]]
            mode = item["mode"]
            fn = self.lib.nsm_jaccard_allpairs if job.left.kind == "sets" else self.lib.nsm_qratio_allpairs
            if slot_free[slot] is not None:
                stream.wait_event(slot_free[slot])
            dev = self._arena(f"out{slot}", item["cap"] * rec_bytes[mode], pinned=False)
            cap = dev.numel() // rec_bytes[mode]
            c = ctl[slot]
            coded = (None, None, 0, None)
            if mode == nsmlib.OUT_CODED:
                if item["j"] not in dicts:   # one dictionary per result, shared by its row blocks
                    dicts[item["j"]] = torch.empty(dict_bytes, dtype=torch.uint8, device=self.device)
                    nsmlib.check(self.lib.nsm_dict_reset(dicts[item["j"]].data_ptr(),
                                                         C.c_void_p(stream.cuda_stream)))
                item.setdefault("exc_cap", max(1 << 16, cap * entry_records[mode] // 16))
                exc = self._arena(f"exc{slot}", item["exc_cap"] * 16, pinned=False)
                coded = (dicts[item["j"]].data_ptr(), exc.data_ptr(), exc.numel() // 16, c.data_ptr() + 56)
            cjob = nsmlib.NsmJob(item["rb"], item["re"], int(job.flat), int(job.cat_mode), float(job.threshold),
                                 job.l_cat.data_ptr() if job.l_cat is not None else None,
                                 job.r_cat.data_ptr() if job.r_cat is not None else None,
                                 dev.data_ptr(), cap, c.data_ptr(), c.data_ptr() + 8, c.data_ptr() + 16,
                                 mode, 0, *coded)
            if self.time_kernels:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            nsmlib.check(fn(C.byref(job.left.struct), C.byref(job.right.struct), C.byref(cjob),
                            C.c_void_p(stream.cuda_stream)))
            if self.time_kernels:
                e1.record(stream)
                self._timed.append((e0, e1))
            self.launches += self.lib.nsm_last_launch_count()
            # the counters go to the host through a one-warp kernel, not the copy engine: a 64-byte
            # copy would queue behind the record copy of the previous block (milliseconds)
            nsmlib.check(self.lib.nsm_publish(c.data_ptr(), ctl_pin[slot].data_ptr(), 64,
                                              C.c_void_p(stream.cuda_stream)))
            self.launches += 1
            done = torch.cuda.Event()
            done.record(stream)
            return done

        def default_capacity(item: dict) -> int:
            job = jobs[item["j"]]
            rows, n_right = item["re"] - item["rb"], job.right.n_items
            if capacity is not None:
                return capacity
            have = max((self._buffers[k].numel() for k in ("out0", "out1") if k in self._buffers),
                       default=0) // rec_bytes[item["mode"]]
            want = capacity_for(item["mode"], min(1 << 24, max(1 << 16, rows * n_right // 8)), rows, n_right)
            return max(have, want)

        def reserve_host(n_bytes: int, fill: int):
            """The pinned arena holds `fill` bytes already and gets room for n_bytes more."""
            pin = self._buffers.get("pin")
            if pin is None or pin.numel() < fill + n_bytes:
                copy_stream.synchronize()
                grown = torch.empty(max((fill + n_bytes) * 5 // 4, (pin.numel() * 3 // 2) if pin is not None else 0),
                                    dtype=torch.uint8).pin_memory()
                if fill:
                    grown[:fill].copy_(pin[:fill])
                self._buffers["pin"] = pin = grown
            return pin

        host_fill = 0
        pending = None      # (slot, item, event)
        queue = list(work)
        slot = 0
        while queue or pending:
            nxt = None
            if pending is None:
                item = queue.pop(0)
                item.setdefault("cap", default_capacity(item))
                pending = (slot, item, launch(slot, item))
                slot ^= 1
            p_slot, p_item, p_done = pending
            p_done.synchronize()
            self._mark(f"block done j{p_item['j']} rows {p_item['rb']}:{p_item['re']} mode {p_item['mode']}")
            words = ctl_pin[p_slot].numpy().view(np.uint64)
            count, flags = int(words[0]), int(words[1]) & 0xffffffff
            stats = [int(x) for x in words[2:2 + nsmlib.N_STATS]]
            j, rb, re_, mode = p_item["j"], p_item["rb"], p_item["re"], p_item["mode"]
            job, info = jobs[j], infos[j]
            n_exc = int(words[7]) if mode == nsmlib.OUT_CODED else 0
            i_kept = nsmlib.STAT_NAMES.index("kept")
            if mode == nsmlib.OUT_PAIRS:
                stats[i_kept] = count    # the fuzzy kernel does not keep this counter
            kept = stats[i_kept]
            if flags & nsmlib.FLAG_OVERFLOW:
                # `count` is exact (pairs or packets): repeat with that size, or split the block
                info["reruns"] += 1
                limit = self.max_pairs_per_block if mode == nsmlib.OUT_PAIRS else \
                    max(1, self.max_pairs_per_block * 16 // rec_bytes[mode])
                if count > limit and re_ - rb > 1:
                    n_parts = min(re_ - rb, -(-count // max(1, limit // 2)))
                    cuts = np.linspace(rb, re_, n_parts + 1).astype(np.int64)
                    queue[:0] = [{"j": j, "rb": int(a), "re": int(b), "mode": mode, "cap": limit,
                                  "exc_cap": int(n_exc * 1.1) + 1024}
                                 for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
                else:
                    queue.insert(0, {**p_item, "cap": int(count * 1.02) + 1024,
                                     "exc_cap": int(n_exc * 1.1) + 1024})
                pending = None
                slot = p_slot
                continue
            if p_item.get("probe") and queue and queue[0].get("rest") and queue[0]["j"] == j:
                rest = queue.pop(0)
                rest_rows = rest["re"] - rest["rb"]
                density = kept / max(1, (re_ - rb) * job.right.n_items)
                expect = density * rest_rows * job.right.n_items        # kept pairs still to come
                r_mode = rest["mode"]
                if (self.compact == "auto" and job.left.kind == "sets" and
                        density * nsmlib.UNIT_LEFT * 32 >= 2 * nsmlib.CPACKET_RECORDS):
                    r_mode = nsmlib.OUT_CODED   # >= two packets per warp and unit
                per_rec = rec_bytes[r_mode] / entry_records[r_mode]
                n_parts = int(min(16, max(1, -(-(expect * per_rec) // self.PIPELINE_BLOCK_BYTES))))
                step = -(-rest_rows // n_parts)
                step = -(-step // nsmlib.UNIT_LEFT) * nsmlib.UNIT_LEFT   # whole 512-row chunks
                cuts = list(range(rest["rb"], rest["re"], step)) + [rest["re"]]
                margin = 1.15 if expect > 1e6 else 2.0
                queue[:0] = [{"j": j, "rb": a, "re": b, "mode": r_mode,
                              "cap": capacity_for(r_mode, density * (b - a) * job.right.n_items * margin + 4096,
                                                  b - a, job.right.n_items)}
                             for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
                if to_host:   # one allocation of page-locked memory for everything still to come
                    ahead = sum(q["cap"] * rec_bytes[q["mode"]] for q in queue if q["j"] == j and "cap" in q)
                    if r_mode == nsmlib.OUT_CODED:   # uncoded pairs (a few per cent) + the dictionary
                        ahead += int(expect * margin * 0.1) * 16 + dict_bytes
                    reserve_host(count * 16 + ahead, host_fill)
            # launch the next block before copying this one out, so that the two overlap
            if queue:
                item = queue.pop(0)
                item.setdefault("cap", default_capacity(item))
                nxt = (slot, item, launch(slot, item))
                slot ^= 1
            info["count"] += kept
            info["flags"] |= flags
            info["blocks"] += 1
            for name, v in zip(nsmlib.STAT_NAMES, stats):
                info["stats"][name] += v
            if mode != nsmlib.OUT_PAIRS:
                info["packets"] += count
                info["uncoded"] += n_exc
            if to_host and (count or n_exc):
                n_bytes, x_bytes = count * rec_bytes[mode], n_exc * 16
                extra = dict_bytes if (mode == nsmlib.OUT_CODED and j not in dict_host) else 0
                pin = reserve_host(n_bytes + x_bytes + extra, host_fill)
                if extra:
                    dict_host[j] = host_fill
                    host_fill += dict_bytes
                    info["d2h_bytes"] += dict_bytes
                copy_stream.wait_event(p_done)
                with torch.cuda.stream(copy_stream):
                    pin[host_fill:host_fill + n_bytes].copy_(self._buffers[f"out{p_slot}"][:n_bytes],
                                                             non_blocking=True)
                    if x_bytes:
                        pin[host_fill + n_bytes:host_fill + n_bytes + x_bytes].copy_(
                            self._buffers[f"exc{p_slot}"][:x_bytes], non_blocking=True)
                    if mode == nsmlib.OUT_CODED:   # the dictionary as of this block (it only grows)
                        pin[dict_host[j]:dict_host[j] + dict_bytes].copy_(dicts[j], non_blocking=True)
                    freed = torch.cuda.Event()
                    freed.record(copy_stream)
                slot_free[p_slot] = freed
                if n_bytes:
                    info["parts"].append((host_fill, n_bytes, mode))
                if x_bytes:
                    info["parts"].append((host_fill + n_bytes, x_bytes, nsmlib.OUT_PAIRS))
                info["d2h_bytes"] += n_bytes + x_bytes
                host_fill += n_bytes + x_bytes
            pending = nxt
        self._mark("last block launched and counted")
        copy_stream.synchronize()
        self._mark("copies done")
        for ev in slot_free:
            if ev is not None:
                stream.wait_event(ev)
        if self.time_kernels:
            for e0, e1 in self._timed:
                self.kernel_ms += e0.elapsed_time(e1)
                self.kernel_launches_timed += 1
            self._timed.clear()

        outs = []
        wire = (PAIR_DTYPE, nsmlib.PACKET_DTYPE, nsmlib.CPACKET_DTYPE)
        for j, (job, info) in enumerate(zip(jobs, infos)):
            parts = info.pop("parts")
            pin = self._buffers.get("pin")
            views = [(mode, pin[lo:lo + n].numpy().view(wire[mode])) for lo, n, mode in parts] if to_host else []
            dictionary = pin[dict_host[j]:dict_host[j] + dict_bytes].numpy().view(np.uint64) \
                if (to_host and j in dict_host) else None
            rec = Records(views, info["count"] if to_host else 0, job.left.perm, job.right.perm, dictionary)
            outs.append(rec.decode(copy=copy) if decode else rec)
        return outs

    # ------------------------------------------------------------------ the job, records stored by the kernels
    def _run_jobs_direct(self, jobs: List["Job"], copy: bool, decode: bool = True):
        """``run_jobs(to_host=True)`` without device arenas for the large jobs.

        A job of PIPELINE_MIN_PAIRS item pairs or more runs as a PROBE (its first 512 left rows,
        16-byte pairs into a small device arena) and a REST launch whose ``out_pairs`` is the
        engine's page-locked host arena itself: the probe's kept-pair density sizes that region and
        picks the record format; the kernel's 16-byte vector stores of whole packets cross the
        link as they are produced.  All probes are launched first (one wait), then all rests back
        to back (one wait): the GPU never idles on a host decision, nothing queues behind a bulk
        copy, and the step ends when the last kernel does.  Smaller jobs take one launch into a
        device arena and one copy.  An arena that turns out too small is counted exactly by the
        kernel and that block is run again."""
        stream = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        copy_stream = self._copy_stream
        self._mark("run_jobs begin")
        rec_bytes, entry_records = nsmlib.RECORD_BYTES, nsmlib.ENTRY_RECORDS
        dict_bytes = nsmlib.DICT_SLOTS * 8
        n_ctl = 2 * max(1, len(jobs))
        ctl = self._arena("ctl_direct", 64 * n_ctl, pinned=False)
        ctl_pin = self._arena("ctl_direct_pin", 64 * n_ctl, pinned=True)
        i_kept = nsmlib.STAT_NAMES.index("kept")

        def n_units(rows: int, n_right: int) -> int:
            return -(-rows // nsmlib.UNIT_LEFT) * -(-n_right // nsmlib.UNIT_RIGHT)

        def entries_for(mode: int, records: float, rows: int, n_right: int) -> int:
            if mode != nsmlib.OUT_PAIRS:   # full packets + at most one partial per warp and unit
                return int(records / entry_records[mode]) + 4 * n_units(rows, n_right) + 64
            return int(records) + 4096

        def launch(job, rb, re_, mode, out_ptr, cap, slot, dict_ptr=None, exc_ptr=None, exc_cap=0):
            fn = self.lib.nsm_jaccard_allpairs if job.left.kind == "sets" else self.lib.nsm_qratio_allpairs
            c = ctl.data_ptr() + 64 * slot
            coded = (dict_ptr, exc_ptr, exc_cap, c + 56) if mode == nsmlib.OUT_CODED else (None, None, 0, None)
            cjob = nsmlib.NsmJob(rb, re_, int(job.flat), int(job.cat_mode), float(job.threshold),
                                 job.l_cat.data_ptr() if job.l_cat is not None else None,
                                 job.r_cat.data_ptr() if job.r_cat is not None else None,
                                 out_ptr, cap, c, c + 8, c + 16, mode, 0, *coded)
            if self.time_kernels:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            nsmlib.check(fn(C.byref(job.left.struct), C.byref(job.right.struct), C.byref(cjob),
                            C.c_void_p(stream.cuda_stream)))
            if self.time_kernels:
                e1.record(stream)
                self._timed.append((e0, e1))
            self.launches += self.lib.nsm_last_launch_count()
            nsmlib.check(self.lib.nsm_publish(c, ctl_pin.data_ptr() + 64 * slot, 64,
                                              C.c_void_p(stream.cuda_stream)))
            self.launches += 1

        def counters(slot: int, mode: int):
            words = ctl_pin.numpy().view(np.uint64)[8 * slot: 8 * slot + 8]
            count, flags = int(words[0]), int(words[1]) & 0xffffffff
            stats = [int(x) for x in words[2:2 + nsmlib.N_STATS]]
            if mode == nsmlib.OUT_PAIRS:
                stats[i_kept] = count
            return count, flags, stats, (int(words[7]) if mode == nsmlib.OUT_CODED else 0)

        def account(info, count, flags, stats, n_exc, mode):
            info["count"] += stats[i_kept]
            info["flags"] |= flags & ~nsmlib.FLAG_OVERFLOW
            info["blocks"] += 1
            for name, v in zip(nsmlib.STAT_NAMES, stats):
                info["stats"][name] += v
            if mode != nsmlib.OUT_PAIRS:
                info["packets"] += count
                info["uncoded"] += n_exc

        infos, plans = [], []
        for j, job in enumerate(jobs):
            if job.left.kind != job.right.kind:
                raise TypeError("left and right must be packed for the same score function")
            begin, end = job.rows if job.rows is not None else (0, job.left.n_items)
            infos.append({"count": 0, "flags": 0, "reruns": 0, "blocks": 0, "d2h_bytes": 0,
                          "packets": 0, "uncoded": 0, "stats": dict.fromkeys(nsmlib.STAT_NAMES, 0),
                          "item_pairs": max(0, end - begin) * job.right.n_items, "parts": []})
            rows = end - begin
            if rows <= 0 or not job.right.n_items:
                plans.append(None)
                continue
            forced = nsmlib.OUT_PAIRS
            if job.left.kind == "sets" and self.compact in (True, "coded", "packets"):
                forced = nsmlib.OUT_PACKETS if self.compact == "packets" else nsmlib.OUT_CODED
            # the probe: four 512-row chunks.  Token-set cohorts are stored in level-count chunks laid
            # out in bit-reversed order (pack.chunked_level_order), so four consecutive chunks hold
            # the cohort's mix of level counts.  It already runs in the compact format (a sparse
            # result costs it at most one packet per warp and unit).
            big = rows >= 8 * nsmlib.UNIT_LEFT and rows * job.right.n_items >= self.PIPELINE_MIN_PAIRS
            probe_mode = forced
            if big and job.left.kind == "sets" and self.compact == "auto":
                probe_mode = nsmlib.OUT_CODED
            plans.append({"begin": begin, "end": end, "forced": forced, "big": big,
                          "first_mode": probe_mode if big else forced,
                          "cut": begin + 4 * nsmlib.UNIT_LEFT if big else end})
        self.last_infos = infos

        # ---- phase 1: probes of the big jobs, whole small jobs; 16-byte pairs into device arenas ----
        def first_capacity(rows, n_right, mode):
            return max(entries_for(mode, min(1 << 24, max(1 << 16, rows * n_right // 8)), rows, n_right), 1 << 16)

        def run_first(js, caps):
            for j in js:
                job, pl = jobs[j], plans[j]
                mode = pl["first_mode"]
                dev = self._arena(f"first{j}", caps[j] * rec_bytes[mode], pinned=False)
                pl["first_cap"] = dev.numel() // rec_bytes[mode]
                extra = {}
                if mode == nsmlib.OUT_CODED:
                    if "dict" not in pl:   # one dictionary per result, shared by all its launches
                        pl["dict"] = torch.empty(dict_bytes, dtype=torch.uint8, device=self.device)
                        nsmlib.check(self.lib.nsm_dict_reset(pl["dict"].data_ptr(), C.c_void_p(stream.cuda_stream)))
                    exc = self._arena(f"first_exc{j}", max(1 << 16, pl["first_cap"] * entry_records[mode] // 16) * 16,
                                      pinned=False)
                    extra = dict(dict_ptr=pl["dict"].data_ptr(), exc_ptr=exc.data_ptr(), exc_cap=exc.numel() // 16)
                launch(job, pl["begin"], pl["cut"], mode, dev.data_ptr(), pl["first_cap"], 2 * j, **extra)
            stream.synchronize()

        live = [j for j, pl in enumerate(plans) if pl is not None]
        caps = {j: first_capacity(plans[j]["cut"] - plans[j]["begin"], jobs[j].right.n_items,
                                  plans[j]["first_mode"]) for j in live}
        todo = list(live)
        while todo:
            run_first(todo, caps)
            again = []
            for j in todo:
                count, flags, stats, n_exc = counters(2 * j, plans[j]["first_mode"])
                if flags & nsmlib.FLAG_OVERFLOW:   # `count` is exact: run it again with that size
                    infos[j]["reruns"] += 1
                    caps[j] = int(count * 1.02) + 1024
                    if plans[j]["first_mode"] == nsmlib.OUT_CODED:
                        self._buffers.pop(f"first_exc{j}", None)
                        self._arena(f"first_exc{j}", (int(n_exc * 1.1) + 1024) * 16, pinned=False)
                    again.append(j)
                else:
                    plans[j]["first"] = (count, flags, stats, n_exc)
            todo = again
        self._mark("probes and small jobs done")

        # ---- phase 2: host regions for everything, then the rests straight into them ----
        align = lambda x: (x + 255) & ~255   # noqa: E731
        need = 0
        for j in live:
            job, pl = jobs[j], plans[j]
            count, flags, stats, n_exc = pl["first"]
            mode = pl["first_mode"]
            pl["first_off"] = need
            need = align(need + count * rec_bytes[mode])
            if n_exc:
                pl["first_exc_off"] = need
                need = align(need + n_exc * 16)
            if not pl["big"]:
                if mode == nsmlib.OUT_CODED:
                    pl["dict_off"] = need
                    need = align(need + dict_bytes)
                continue
            kept = stats[i_kept]
            n_right = job.right.n_items
            density = kept / max(1, (pl["cut"] - pl["begin"]) * n_right)
            rest_rows = pl["end"] - pl["cut"]
            expect = density * rest_rows * n_right
            r_mode = pl["forced"]
            if (self.compact == "auto" and job.left.kind == "sets" and
                    density * nsmlib.UNIT_LEFT * 32 >= 2 * nsmlib.CPACKET_RECORDS):
                r_mode = nsmlib.OUT_CODED   # >= two packets per warp and unit
            limit = self.max_pairs_per_block
            # strings are stored by rising length: the probe rows are the shortest ones
            margin = (1.25 if expect > 1e6 else 2.0) if job.left.kind == "sets" else 3.0
            n_parts = int(max(1, -(-(expect * margin) // limit)))
            step = -(-rest_rows // n_parts)
            step = -(-step // nsmlib.UNIT_LEFT) * nsmlib.UNIT_LEFT
            cuts = list(range(pl["cut"], pl["end"], step)) + [pl["end"]]
            pl["rest_mode"], pl["rest"] = r_mode, []
            if r_mode == nsmlib.OUT_CODED and "dict" not in pl:
                pl["dict"] = torch.empty(dict_bytes, dtype=torch.uint8, device=self.device)
                nsmlib.check(self.lib.nsm_dict_reset(pl["dict"].data_ptr(), C.c_void_p(stream.cuda_stream)))
            if "dict" in pl:
                pl["dict_off"] = need
                need = align(need + dict_bytes)
            for a, b in zip(cuts[:-1], cuts[1:]):
                if b <= a:
                    continue
                cap = entries_for(r_mode, density * (b - a) * n_right * margin + 4096, b - a, n_right)
                blk = {"rb": a, "re": b, "cap": cap, "off": need}
                need = align(need + cap * rec_bytes[r_mode])
                if r_mode == nsmlib.OUT_CODED:
                    blk["exc_cap"] = int(density * (b - a) * n_right * 0.1) + (1 << 14)
                    blk["exc_off"] = need
                    need = align(need + blk["exc_cap"] * 16)
                pl["rest"].append(blk)
        pin = self._buffers.get("pin")
        if pin is None or pin.numel() < need:
            # page-locking is slow (hundreds of ms per GB, more inside a VM) and the need varies a
            # little from call to call (which scores find a dictionary slot is a race): headroom
            self._buffers.pop("pin", None)
            pin = self._buffers["pin"] = torch.empty(max(need + need // 4, 1 << 20), dtype=torch.uint8).pin_memory()
        base = pin.data_ptr()
        self._mark("host arena ready")

        # the first launches' records: exact-size copies on the copy stream (they are small)
        with torch.cuda.stream(copy_stream):
            for j in live:
                pl = plans[j]
                count, flags, stats, n_exc = pl["first"]
                mode = pl["first_mode"]
                n_bytes = count * rec_bytes[mode]
                if n_bytes:
                    pin[pl["first_off"]:pl["first_off"] + n_bytes].copy_(self._buffers[f"first{j}"][:n_bytes],
                                                                         non_blocking=True)
                    infos[j]["parts"].append((pl["first_off"], n_bytes, mode))
                if n_exc:
                    pin[pl["first_exc_off"]:pl["first_exc_off"] + n_exc * 16].copy_(
                        self._buffers[f"first_exc{j}"][:n_exc * 16], non_blocking=True)
                    infos[j]["parts"].append((pl["first_exc_off"], n_exc * 16, nsmlib.OUT_PAIRS))
                infos[j]["d2h_bytes"] += n_bytes + n_exc * 16 + (dict_bytes if "dict" in pl else 0)
                account(infos[j], count, flags, stats, n_exc, mode)

        pending = [(j, blk) for j in live if plans[j]["big"] for blk in plans[j]["rest"]]
        while pending:
            ctl_rest = self._arena("ctl_rest", 64 * len(pending), pinned=False)
            ctl_rest_pin = self._arena("ctl_rest_pin", 64 * len(pending), pinned=True)
            ctl, ctl_pin = ctl_rest, ctl_rest_pin   # launch() / counters() address these
            for slot, (j, blk) in enumerate(pending):
                job, pl = jobs[j], plans[j]
                mode = pl["rest_mode"]
                extra = {}
                if mode == nsmlib.OUT_CODED:
                    extra = dict(dict_ptr=pl["dict"].data_ptr(), exc_ptr=base + blk["exc_off"], exc_cap=blk["exc_cap"])
                launch(job, blk["rb"], blk["re"], mode, base + blk["off"], blk["cap"], slot, **extra)
            self._mark("rests launched")
            stream.synchronize()
            self._mark("rests done")
            again = []
            for slot, (j, blk) in enumerate(pending):
                pl = plans[j]
                mode = pl["rest_mode"]
                count, flags, stats, n_exc = counters(slot, mode)
                if flags & nsmlib.FLAG_OVERFLOW:
                    infos[j]["reruns"] += 1
                    again.append((j, dict(blk, cap=int(count * 1.02) + 1024, exc_cap=int(n_exc * 1.1) + 1024)))
                    continue
                n_bytes = count * rec_bytes[mode]
                if n_bytes:
                    infos[j]["parts"].append((blk["off"], n_bytes, mode))
                if n_exc:
                    infos[j]["parts"].append((blk["exc_off"], n_exc * 16, nsmlib.OUT_PAIRS))
                infos[j]["d2h_bytes"] += n_bytes + n_exc * 16
                account(infos[j], count, flags, stats, n_exc, mode)
            if again:
                # regions for the repeated blocks behind everything else; the arena may have to grow,
                # which is safe now: no kernel is writing to it
                copy_stream.synchronize()
                fill = need
                for j, blk in again:
                    mode = plans[j]["rest_mode"]
                    blk["off"] = need
                    need = align(need + blk["cap"] * rec_bytes[mode])
                    if mode == nsmlib.OUT_CODED:
                        blk["exc_off"] = need
                        need = align(need + blk["exc_cap"] * 16)
                if pin.numel() < need:
                    grown = torch.empty(need + need // 4, dtype=torch.uint8).pin_memory()
                    grown[:fill].copy_(pin[:fill])
                    self._buffers["pin"] = pin = grown
                    base = pin.data_ptr()
            pending = again
        for j in live:   # the dictionaries as the last launch left them (queued behind the kernels)
            pl = plans[j]
            if "dict" in pl:
                pin[pl["dict_off"]:pl["dict_off"] + dict_bytes].copy_(pl["dict"], non_blocking=True)
        copy_stream.synchronize()
        stream.synchronize()
        self._mark("copies done")
        if self.time_kernels:
            for e0, e1 in self._timed:
                self.kernel_ms += e0.elapsed_time(e1)
                self.kernel_launches_timed += 1
            self._timed.clear()

        outs = []
        wire = (PAIR_DTYPE, nsmlib.PACKET_DTYPE, nsmlib.CPACKET_DTYPE)
        for j, (job, info) in enumerate(zip(jobs, infos)):
            parts = info.pop("parts")
            views = [(mode, pin[lo:lo + n].numpy().view(wire[mode])) for lo, n, mode in parts]
            pl = plans[j]
            dictionary = None
            if pl is not None and "dict_off" in pl:
                dictionary = pin[pl["dict_off"]:pl["dict_off"] + dict_bytes].numpy().view(np.uint64)
            rec = Records(views, info["count"], job.left.perm, job.right.perm, dictionary)
            outs.append(rec.decode(copy=copy) if decode else rec)
        return outs

    # ------------------------------------------------------------------ roofline denominators
    def microbench(self, kind: int, iters: int = 4096, blocks_per_sm: int = 8,
                   threads: int = 256) -> float:
        """Measured 32-bit integer ops/s of one instruction kind (see nsm_microbench)."""
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        torch.cuda.set_device(self.device)
        blocks = sms * blocks_per_sm
        sink = torch.zeros(4, dtype=torch.int32, device=self.device)
        ops = C.c_uint64(0)
        stream = torch.cuda.current_stream(self.device)
        best = 0.0
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            nsmlib.check(self.lib.nsm_microbench(kind, blocks, threads, iters, sink.data_ptr(),
                                                 C.byref(ops), C.c_void_p(stream.cuda_stream)))
            e1.record(stream)
            e1.synchronize()
            self.launches += 1
            if rep:
                best = max(best, ops.value * blocks * threads / (e0.elapsed_time(e1) * 1e-3))
        return best


_default_engine: Optional[Engine] = None
_thread = threading.local()


def default_engine() -> Engine:
    """The engine the host API scores with: the one bound to this thread by :func:`use_engine`
    (gpu/scheduler.py runs one worker thread per GPU), else a process-wide engine on the current
    CUDA device."""
    global _default_engine
    bound = getattr(_thread, "engine", None)
    if bound is not None:
        return bound
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


@contextlib.contextmanager
def use_engine(engine):
    """Binds ``engine`` to the calling thread (and makes its CUDA device the thread's current
    device, which the C ABI's kernel launches need) for the duration of the block."""
    previous = getattr(_thread, "engine", None)
    _thread.engine = engine
    device = getattr(engine, "device", None)
    try:
        if device is not None:
            with torch.cuda.device(device):
                yield engine
        else:
            yield engine
    finally:
        _thread.engine = previous
