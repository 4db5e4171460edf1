"""Host-side placement for the one-process-per-GPU runtime.

Each rank stages its inputs in pinned host memory and copies them over its own PCIe link.  On a two-socket box the
pages of a pinned buffer live on the NUMA node of the CPU that first touched them, so a rank whose process floats over
both sockets ends up pulling half of its H2D traffic across the inter-socket link, and several ranks doing so at once
saturate it.  `bind_to_gpu_numa_node` pins the calling process to the CPUs of the NUMA node its GPU hangs off (read
from sysfs) BEFORE the pinned buffers are allocated and filled.
"""
from __future__ import annotations

import os
import subprocess
from typing import Optional


def _parse_cpulist(text: str) -> set:
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


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of GPU `device_index` (as numbered by the driver), or None when the platform does not expose it."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout
        for line in out.splitlines():
            idx, bus = [t.strip() for t in line.split(",")]
            if int(idx) == device_index:
                bus = bus.lower()
                if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:   # 00000000:3B:00.0 -> 0000:3b:00.0
                    bus = bus[4:]
                with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                    node = int(f.read().strip())
                return node if node >= 0 else None
    except Exception:
        return None
    return None


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Restrict this process to the CPUs of the GPU's NUMA node (no-op when unknown).  Returns the node or None."""
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target:
            os.sched_setaffinity(0, target)
            return node
    except Exception:
        return None
    return None
