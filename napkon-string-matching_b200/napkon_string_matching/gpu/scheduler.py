"""
Concurrent scheduling of the comparisons of one matching run.

The reference runs its comparison steps one after the other
(/root/reference/napkon_string_matching/matcher.py:228-284: gecco vs every cohort, every unordered
cohort pair, the same pairs again on the ``Variable`` column).  They are independent of one
another, so here whole comparisons are spread over the GPUs of the box instead:

* one process, several visible GPUs: one worker thread per GPU, each with its own ``Engine``
  (its own stream, arenas and pinned buffers); ctypes releases the GIL during kernel launches
  and event waits, so the GPUs run concurrently;
* one process per GPU (``torch.distributed`` initialised): every rank takes its share of whole
  comparisons, runs them unsharded on its GPU, and the (small) result frames are exchanged with
  one ``all_gather_object`` on the host group.

Comparisons are dealt out longest-first by estimated work (|left| x |right| items) onto the least
loaded worker (LPT), deterministically, so every rank computes the same plan without talking.
When there are fewer comparisons than workers the row-block sharding of
``gpu/distributed.py`` (every comparison split over all ranks) uses the box better and is kept.
Results are always delivered in submission order, so ``Matcher.results`` is filled exactly as by
the sequential loop.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence

from napkon_string_matching.gpu import distributed
from napkon_string_matching.gpu import engine as engine_mod


@dataclass
class ComparisonTask:
    """One ``left.compare(right, **kwargs)`` call; ``name`` is the key in ``Matcher.results``."""
    name: str
    left: Any
    right: Any
    kwargs: Dict[str, Any] = field(default_factory=dict)

    @property
    def work(self) -> int:
        return max(1, len(self.left)) * max(1, len(self.right))

    def run(self):
        return self.left.compare(self.right, **self.kwargs)


def plan(work: Sequence[int], n_workers: int) -> List[List[int]]:
    """LPT: task indices per worker, largest first onto the least loaded worker (ties: lowest
    worker index, lowest task index).  Every worker's list is in submission order."""
    n_workers = max(1, n_workers)
    order = sorted(range(len(work)), key=lambda i: (-work[i], i))
    load = [0] * n_workers
    out: List[List[int]] = [[] for _ in range(n_workers)]
    for i in order:
        w = min(range(n_workers), key=lambda k: (load[k], k))
        out[w].append(i)
        load[w] += work[i]
    return [sorted(ix) for ix in out]


def _run_threads(tasks: List[ComparisonTask], engines: List[Any]) -> List[Any]:
    """One thread per engine; engine k runs the tasks ``plan`` gives worker k."""
    assignment = plan([t.work for t in tasks], len(engines))
    results: List[Any] = [None] * len(tasks)
    errors: List[BaseException] = []

    def worker(k: int):
        try:
            device = getattr(engines[k], "device", None)
            if device is not None and len(engines) > 1:   # this thread's pinned arenas: local node
                from napkon_string_matching.gpu import affinity

                affinity.bind_to_gpu(device.index)
            with engine_mod.use_engine(engines[k]):
                for i in assignment[k]:
                    results[i] = tasks[i].run()
        except BaseException as exc:  # re-raised in the caller's thread
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(k,), name=f"nsm-gpu{k}") for k in range(len(engines))
               if assignment[k]]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def _run_ranks(tasks: List[ComparisonTask]) -> List[Any]:
    """torch.distributed: my share of whole comparisons, then one exchange of the results."""
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    assignment = plan([t.work for t in tasks], world)
    mine: Dict[int, Any] = {}
    failure: Optional[BaseException] = None
    with distributed.whole_comparisons():
        try:
            for i in assignment[rank]:
                mine[i] = tasks[i].run()
        except Exception as exc:  # noqa: BLE001 - must still reach the exchange: the others wait there
            failure = exc
    everyone: List[Optional[tuple]] = [None] * world
    dist.all_gather_object(everyone, (mine, failure), group=distributed.host_group())
    for r, (_, exc) in enumerate(everyone):
        if exc is not None:   # every rank raises the first failure (lowest rank), like the
            raise exc         # sequential loop would have raised it on its only process
    merged: Dict[int, Any] = {}
    for part, _ in everyone:
        merged.update(part)
    return [merged[i] for i in range(len(tasks))]


def run_comparisons(tasks: Sequence[ComparisonTask], engines: Optional[List[Any]] = None,
                    engine_factory: Optional[Callable[[int], Any]] = None) -> List[Any]:
    """Results of ``task.run()`` for every task, in submission order.

    ``engines`` (or ``engine_factory(device_index)``) override the per-GPU engines of the threaded
    mode; by default one ``Engine`` per visible CUDA device is created (and kept)."""
    tasks = list(tasks)
    if not tasks:
        return []
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if len(tasks) >= dist.get_world_size():
            return _run_ranks(tasks)
        return [t.run() for t in tasks]  # every comparison row-sharded over all ranks
    if engines is None:
        n_dev = torch.cuda.device_count() if torch.cuda.is_available() else 0
        if engine_factory is not None:
            engines = [engine_factory(k) for k in range(max(1, n_dev))]
        elif n_dev > 1:
            engines = device_engines(n_dev)
    if not engines or len(engines) == 1 or len(tasks) == 1:
        if engines:
            with engine_mod.use_engine(engines[0]):
                return [t.run() for t in tasks]
        return [t.run() for t in tasks]
    return _run_threads(tasks, engines)


_device_engines: Dict[int, Any] = {}


def device_engines(n_dev: int) -> List[Any]:
    """One ``Engine`` per CUDA device, created once."""
    for k in range(n_dev):
        if k not in _device_engines:
            _device_engines[k] = engine_mod.Engine(k)
    return [_device_engines[k] for k in range(n_dev)]
