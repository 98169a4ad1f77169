"""
Multi-GPU partitioning (north_star "Multi-GPU partitioning"): one process per
GPU, the left cohort split into contiguous row blocks balanced by work, the right cohort
replicated.  Pairs are independent, so there is no data-path collective: every rank scores its
block, the ranks all-gather their kept-pair COUNTS (NCCL when the process group is NCCL) and the
records themselves are concatenated on the host through a gloo group.
"""
from __future__ import annotations

import contextlib
import os
import threading
from typing import Callable, List, Sequence, Tuple

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


def allgather_count_vectors(counts: Sequence[int]) -> List[List[int]]:
    """:func:`allgather_counts` for several counters at once: ``result[rank][k]``."""
    rank, world = _world()
    if world == 1:
        return [[int(c) for c in counts]]
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" \
        else torch.device("cpu")
    mine = torch.tensor([int(c) for c in counts], dtype=torch.int64, device=device)
    everyone = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine)
    return [[int(v) for v in t.tolist()] for t in everyone]


def sharded_run_jobs(engine, jobs, *, to_host: bool = True, decode: bool = True, copy: bool = True):
    """Several comparisons, each split over the ranks by left row blocks, through ONE
    ``engine.run_jobs`` call per rank (so the copy-out of one comparison runs beside the kernel of
    the next), followed by the path's one collective: the all-gather of the kept-pair counts.

    ``jobs`` are ``engine.Job``s over whole cohorts (or a ``rows`` range); every rank must pass the
    same list.  Returns ``(outs, counts)``: this rank's records per job (they stay sharded — each
    rank holds the pairs of its own rows) and ``counts[rank][job]``."""
    import dataclasses

    rank, world = _world()
    mine = []
    for job in jobs:
        begin, end = job.rows if job.rows is not None else (0, job.left.n_items)
        b, e = partition_rows(job.left.weights[begin:end], world)[rank]
        mine.append(dataclasses.replace(job, rows=(begin + b, begin + e)))
    failure = None
    try:
        outs = engine.run_jobs(mine, to_host=to_host, decode=decode, copy=copy)
        kept = [info["count"] for info in engine.last_infos]
    except Exception as exc:  # noqa: BLE001 - the other ranks wait in the collective below
        failure, outs, kept = exc, [], [-1] * len(jobs)
    counts = allgather_count_vectors(kept)
    if failure is not None:
        raise failure
    for r, row in enumerate(counts):
        if min(row, default=0) < 0:
            raise RuntimeError(f"scoring the row blocks of rank {r} failed")
    return outs, counts


def default_gather() -> str:
    """Where the records of a row-sharded comparison end up: "rank0" (rank 0 holds the complete
    result, every other rank the pairs of its own rows; exact-size point-to-point transfers, no
    redundant traffic), "all" (identical on every rank) or "none".  ``NSM_GATHER`` overrides."""
    return os.environ.get("NSM_GATHER", "rank0")


def sharded_all_pairs(score_block: Callable[[int, int], np.ndarray], weights: np.ndarray,
                      gather=None) -> np.ndarray:
    """Runs ``score_block(begin, end)`` on this rank's row block, all-gathers the kept-pair counts
    and returns records according to ``gather`` (see :func:`default_gather`; True means "all",
    False "none")."""
    global last_counts
    rank, world = _world()
    n = len(weights)
    if world == 1:
        out = score_block(0, n)
        last_counts = [len(out)]
        return out
    if gather is None:
        gather = default_gather()
    gather = {True: "all", False: "none"}.get(gather, gather)
    if gather not in ("all", "rank0", "none"):
        raise ValueError(f"gather={gather!r}")
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
    if gather == "none" or max(counts) == 0:
        return mine
    group = _get_host_group()
    mine = np.ascontiguousarray(mine)
    if gather == "rank0":
        # exact-size transfers to rank 0 only: every record crosses the host once
        if rank != 0:
            if len(mine):
                dist.send(torch.from_numpy(mine.view(np.uint8)), dst=0, group=group)
            return mine
        out = np.empty(sum(counts), dtype=PAIR_DTYPE)
        out[: counts[0]] = mine
        pos = counts[0]
        for src in range(1, world):
            if counts[src]:
                dist.recv(torch.from_numpy(out[pos: pos + counts[src]].view(np.uint8)), src=src, group=group)
                pos += counts[src]
        return out
    cap = max(counts)
    buf = np.zeros(cap, dtype=PAIR_DTYPE)
    buf[: len(mine)] = mine
    send = torch.from_numpy(buf.view(np.uint8))
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    parts = [r.numpy().view(PAIR_DTYPE)[:c] for r, c in zip(recv, counts)]
    return np.concatenate(parts)
