"""
Multi-GPU partitioning (north_star "Multi-GPU partitioning"): one process per
GPU, the left cohort split into contiguous row blocks balanced by work, the right cohort
replicated.  Pairs are independent, so there is no data-path collective: every rank scores its
block, the ranks all-gather their kept-pair COUNTS (NCCL when the process group is NCCL) and the
records themselves are concatenated on the host through a gloo group.
"""
from __future__ import annotations

import contextlib
import threading
from typing import Callable, List, Tuple

import numpy as np
import torch
import torch.distributed as dist

from napkon_string_matching.gpu.lib import PAIR_DTYPE

_host_group = None
last_counts: List[int] = []


def partition_rows(weights: np.ndarray, parts: int) -> List[Tuple[int, int]]:
    """``parts`` contiguous row blocks [begin, end) with near-equal weight sums (blocks may be
    empty when there are fewer rows than parts)."""
    n = len(weights)
    if parts <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (parts - 1)
    csum = np.cumsum(np.asarray(weights, dtype=np.float64))
    targets = csum[-1] * np.arange(1, parts) / parts
    cuts = np.searchsorted(csum, targets, side="left") + 1
    cuts = np.minimum(np.maximum.accumulate(np.concatenate([[0], cuts, [n]])), n)
    return [(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:])]


_local = threading.local()


@contextlib.contextmanager
def whole_comparisons():
    """Inside the block this rank scores every comparison completely by itself: the scheduler
    (gpu/scheduler.py) has dealt out whole comparisons, so rows are not sharded again."""
    previous = getattr(_local, "whole", False)
    _local.whole = True
    try:
        yield
    finally:
        _local.whole = previous


def _world() -> Tuple[int, int]:
    if getattr(_local, "whole", False):
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def host_group():
    """The gloo group host-side exchanges go through (None: the default group is gloo)."""
    return _get_host_group()


def _get_host_group():
    global _host_group
    if dist.get_backend() == "gloo":
        return None  # default group already lives on the host
    if _host_group is None:
        _host_group = dist.new_group(backend="gloo")
    return _host_group


def allgather_counts(count: int) -> List[int]:
    """The path's one collective."""
    rank, world = _world()
    if world == 1:
        return [int(count)]
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" \
        else torch.device("cpu")
    mine = torch.tensor([int(count)], dtype=torch.int64, device=device)
    everyone = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine)
    return [int(t.item()) for t in everyone]


def sharded_all_pairs(score_block: Callable[[int, int], np.ndarray], weights: np.ndarray,
                      gather: bool = True) -> np.ndarray:
    """Runs ``score_block(begin, end)`` on this rank's row block and returns the records of all
    ranks (``gather=True``, identical on every rank) or only this rank's."""
    global last_counts
    rank, world = _world()
    n = len(weights)
    if world == 1:
        out = score_block(0, n)
        last_counts = [len(out)]
        return out
    begin, end = partition_rows(weights, world)[rank]
    failure = None
    try:
        mine = score_block(begin, end) if end > begin else np.zeros(0, dtype=PAIR_DTYPE)
    except Exception as exc:  # noqa: BLE001 - the other ranks wait in the collective below
        failure, mine = exc, np.zeros(0, dtype=PAIR_DTYPE)
    # a count of -1 tells every rank that a block failed, so that nobody is left in a gather
    counts = allgather_counts(-1 if failure is not None else len(mine))
    if failure is not None:
        raise failure
    if min(counts) < 0:
        raise RuntimeError(f"scoring the row block of rank {counts.index(-1)} failed")
    last_counts = counts
    if not gather:
        return mine
    group = _get_host_group()
    cap = max(counts)
    if cap == 0:
        return np.zeros(0, dtype=PAIR_DTYPE)
    buf = np.zeros(cap, dtype=PAIR_DTYPE)
    buf[: len(mine)] = mine
    send = torch.from_numpy(buf.view(np.uint8))
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    parts = [r.numpy().view(PAIR_DTYPE)[:c] for r, c in zip(recv, counts)]
    return np.concatenate(parts)
