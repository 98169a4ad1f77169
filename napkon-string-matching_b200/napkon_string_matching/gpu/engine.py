"""
Drives the CUDA comparison kernels: uploads packed cohorts, runs the all-pairs job for a left row
block, and returns the kept ``(left, right, score)`` records as one numpy array.

PyTorch is used for device memory, pinned host memory and streams only.  No GPU -> raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from napkon_string_matching.gpu import lib as nsmlib
from napkon_string_matching.gpu.pack import PackedSets, PackedStrings

PAIR_DTYPE = nsmlib.PAIR_DTYPE


class ScoreError(Exception):
    """Carries the reference's exception type for inputs its pair loop would have raised on."""


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


class Engine:
    def __init__(self, device: Optional[int] = None, max_pairs_per_block: int = 1 << 30):
        _require_cuda()
        self.lib = nsmlib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.max_pairs_per_block = int(max_pairs_per_block)
        self._buffers: Dict[str, torch.Tensor] = {}
        self.launches = 0           # kernels of ours launched so far
        self.time_kernels = False   # bracket every comparison kernel with CUDA events
        self.kernel_ms = 0.0
        self.kernel_launches_timed = 0
        self.last_info: Dict = {}

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
            return [packed.item_level_off, packed.level_chr_off, packed.chr]
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
                                packed.max_levels, packed.n_slots, int(packed.exact_bits), 0)
            per_level = packed.level_sizes()
            kind = "sets"
        else:
            st = nsmlib.NsmStrings(*[t.data_ptr() for t in tensors], packed.n_items,
                                   packed.n_levels, packed.max_levels, packed.max_len,
                                   packed.n_alphabet)
            per_level = packed.level_lengths()
            kind = "strings"
        # per-item work estimate (sum of level sizes + 1), used to balance row blocks over GPUs
        csum = np.concatenate([[0], np.cumsum(per_level, dtype=np.int64)])
        off = packed.item_level_off.astype(np.int64)
        per_item = (csum[off[1:]] - csum[off[:-1]]).astype(np.float64) + 1.0
        return DeviceCohort(kind, st, tensors, packed.n_items, packed.max_levels,
                            sum(a.nbytes for a in arrays), per_item)

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
        particular order.

        One kernel launch covers the whole row block; kept records are compacted into a device
        arena.  If the arena was too small the kernel still counts exactly, so the arena is grown
        to the exact need and the launch repeated once.  Results larger than
        ``max_pairs_per_block`` records are produced in several row blocks.
        ``to_host=False`` leaves the records on the device (kernel-only timing) and returns an
        empty array; ``copy=False`` returns a view of the engine's pinned arena (valid until the
        next call).  Counters are in ``self.last_info``."""
        if left.kind != right.kind:
            raise TypeError("left and right must be packed for the same score function")
        fn = self.lib.nsm_jaccard_allpairs if left.kind == "sets" else self.lib.nsm_qratio_allpairs
        begin, end = rows if rows is not None else (0, left.n_items)
        n_right = right.n_items
        info = {"count": 0, "flags": 0, "reruns": 0, "blocks": 0, "d2h_bytes": 0,
                "stats": dict.fromkeys(nsmlib.STAT_NAMES, 0),
                "item_pairs": max(0, end - begin) * n_right}
        self.last_info = info
        if end <= begin or n_right == 0:
            return np.zeros(0, dtype=PAIR_DTYPE)
        stream = torch.cuda.current_stream(self.device)
        ctl = self._arena("ctl", 64, pinned=False)
        ctl_pin = self._arena("ctl_pin", 64, pinned=True)

        def run(rb: int, re_: int, cap: int):
            dev = self._arena("out", cap * 16, pinned=False)
            cap = dev.numel() // 16
            job = nsmlib.NsmJob(rb, re_, int(flat), int(cat_mode), float(threshold),
                                l_cat.data_ptr() if l_cat is not None else None,
                                r_cat.data_ptr() if r_cat is not None else None,
                                dev.data_ptr(), cap, ctl.data_ptr(), ctl.data_ptr() + 8,
                                ctl.data_ptr() + 16)
            if self.time_kernels:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            nsmlib.check(fn(C.byref(left.struct), C.byref(right.struct), C.byref(job),
                            C.c_void_p(stream.cuda_stream)))
            if self.time_kernels:
                e1.record(stream)
            self.launches += 1
            ctl_pin.copy_(ctl, non_blocking=True)
            stream.synchronize()
            if self.time_kernels:
                self.kernel_ms += e0.elapsed_time(e1)
                self.kernel_launches_timed += 1
            words = ctl_pin.numpy().view(np.uint64)
            return int(words[0]), int(words[1]) & 0xffffffff, [int(x) for x in words[2:2 + nsmlib.N_STATS]]

        blocks = [(begin, end)]
        parts: List[Tuple[int, int]] = []   # (offset, count) inside the pinned arena
        host_fill = 0
        if capacity is None:
            have = self._buffers["out"].numel() // 16 if "out" in self._buffers else 0
            capacity = max(have, min(self.max_pairs_per_block, 1 << 24,
                                     max(1 << 16, info["item_pairs"] // 8)))
        while blocks:
            rb, re_ = blocks.pop(0)
            count, flags, stats = run(rb, re_, capacity)
            if flags & nsmlib.FLAG_OVERFLOW:
                info["reruns"] += 1
                if count > self.max_pairs_per_block and re_ - rb > 1:
                    # split by the observed density so that every part should fit
                    n_parts = min(re_ - rb, -(-count // max(1, self.max_pairs_per_block // 2)))
                    cuts = np.linspace(rb, re_, n_parts + 1).astype(np.int64)
                    blocks[:0] = [(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
                    capacity = self.max_pairs_per_block
                else:
                    capacity = int(count * 1.02) + 1024
                    blocks.insert(0, (rb, re_))
                continue
            info["count"] += count
            info["flags"] |= flags
            info["blocks"] += 1
            for name, v in zip(nsmlib.STAT_NAMES, stats):
                info["stats"][name] += v
            if to_host and count:
                n_bytes = count * 16
                pin = self._buffers.get("pin")
                if pin is None or pin.numel() < host_fill + n_bytes:
                    grown = torch.empty(max(host_fill + n_bytes, 2 * (pin.numel() if pin is not None else 0)),
                                        dtype=torch.uint8).pin_memory()
                    if host_fill:
                        grown[:host_fill].copy_(pin[:host_fill])
                    self._buffers["pin"] = pin = grown
                pin[host_fill:host_fill + n_bytes].copy_(self._buffers["out"][:n_bytes],
                                                         non_blocking=True)
                stream.synchronize()
                host_fill += n_bytes
                info["d2h_bytes"] += n_bytes
        if not to_host or host_fill == 0:
            return np.zeros(0, dtype=PAIR_DTYPE)
        out = self._buffers["pin"][:host_fill].numpy().view(PAIR_DTYPE)
        return out.copy() if copy else out

    # ------------------------------------------------------------------ roofline denominators
    def microbench(self, kind: int, iters: int = 4096, blocks_per_sm: int = 8,
                   threads: int = 256) -> float:
        """Measured 32-bit integer ops/s of one instruction kind (see nsm_microbench)."""
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
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


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine
