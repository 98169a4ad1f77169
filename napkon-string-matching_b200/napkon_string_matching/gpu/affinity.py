"""
CPU / NUMA affinity of a GPU worker.  On a multi-socket box the kept records of every GPU land in
pinned host memory; if that memory sits on the other socket, every copy crosses the inter-socket
link.  Binding the worker (process or thread) to the CPUs NVML reports as local to its GPU *before*
the pinned arenas are allocated puts them on the local node (first touch).  Best effort: any
failure (no NVML, restricted cpuset, non-Linux) leaves the affinity as it was.
"""
from __future__ import annotations

import os
from typing import List, Optional


def gpu_local_cpus(device_index: int) -> Optional[List[int]]:
    """CPUs NVML considers local to CUDA device ``device_index`` (matched by PCI bus id, so
    CUDA_VISIBLE_DEVICES re-numbering does not matter), or None."""
    try:
        import pynvml
        import torch

        props = torch.cuda.get_device_properties(device_index)
        bus_id = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        pynvml.nvmlInit()
        try:
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
            n_cpus = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
        return cpus or None
    except Exception:  # noqa: BLE001 - best effort by design
        return None


def bind_to_gpu(device_index: int) -> Optional[List[int]]:
    """Restricts the calling thread to the CPUs local to the GPU (intersected with what it may
    use already).  Returns the CPU list that is now in effect, or None if nothing was changed."""
    if not hasattr(os, "sched_setaffinity"):
        return None
    local = gpu_local_cpus(device_index)
    if not local:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        target = sorted(allowed.intersection(local))
        if not target or len(target) == len(allowed):
            return None
        os.sched_setaffinity(0, target)
        return target
    except OSError:
        return None
