"""Bind a rank's host threads (and with them its first-touch host allocations, pinned buffers included) to the NUMA
node its GPU hangs off.

With one process per GPU and eight GPUs on a two-socket host, pinned buffers that all land on node 0 make every
device<->host copy of the far socket's GPUs cross the inter-socket link, and all eight ranks compete for one socket's
memory controllers (round 1: 11.4 GB/s per GPU at N = 8 against 52.8 GB/s at N = 1).  Linux allocates on the node of
the touching thread, so restricting the process to the GPU's node before any pinned allocation is enough.
"""

from __future__ import annotations

import os
from typing import List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of a CUDA device from sysfs, or None when the platform does not say (single node, container)."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def node_cpus(node: int) -> List[int]:
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            return _parse_cpulist(f.read())
    except Exception:
        return []


def bind_to_gpu_node(device_index: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node (intersected with the CPUs it may already use).
    Returns what was done, for the bench report.  Never raises: without NUMA information nothing changes."""
    info = {"node": None, "cpus": None, "bound": False}
    node = gpu_numa_node(device_index)
    if node is None:
        return info
    info["node"] = node
    try:
        allowed = os.sched_getaffinity(0)
        want = sorted(set(node_cpus(node)) & set(allowed))
        if want and len(want) < len(allowed):
            os.sched_setaffinity(0, want)
            info["bound"] = True
        info["cpus"] = len(want) if want else len(allowed)
    except Exception:
        pass
    return info
