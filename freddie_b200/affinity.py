"""NUMA placement of a GPU's host side.  Pinned host buffers are first-touch allocations: a rank whose threads run
on the far socket pins its batches there and every DMA of the end-to-end path crosses the socket interconnect
(measured on a 4-GPU box: 2.97 ms per step for the GPUs next to the socket the ranks ran on, 3.43 ms for the
others).  ``bind_to_gpu(device)`` restricts the calling process (all its future threads and their first-touch
allocations) to the CPUs of the GPU's NUMA node.  Best effort: any failure leaves the affinity untouched."""
from __future__ import annotations

import os
from typing import Optional


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device: int) -> Optional[int]:
    """NUMA node of CUDA device ``device`` (index inside CUDA_VISIBLE_DEVICES), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device
        if vis:
            ent = [v.strip() for v in vis.split(",") if v.strip()]
            if device < len(ent) and ent[device].isdigit():
                idx = int(ent[device])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        # NVML prints an 8-digit domain, sysfs uses 4
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())
        with open(path) as fh:
            node = int(fh.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu(device: int) -> Optional[dict]:
    """Pins the process to the CPUs of the GPU's NUMA node.  Returns {'node', 'cpus'} or None when nothing was done
    (single-node host, unknown topology, FRS_NO_NUMA_BIND set)."""
    if os.environ.get("FRS_NO_NUMA_BIND"):
        return None
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if len(nodes) < 2:
            return None
        node = gpu_numa_node(device)
        if node is None:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = _parse_cpulist(fh.read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return dict(node=node, cpus=len(cpus))
    except Exception:
        return None
