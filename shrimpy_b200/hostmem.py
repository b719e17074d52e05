"""Host-side placement for the PCIe-bound legs of the path.

With one process per GPU, pinned staging buffers should live on the NUMA node the GPU hangs off; otherwise
every H2D/D2H crosses the socket interconnect and the eight ranks of a box contend for it (measured: the
end-to-end rate of 8 ranks was barely above that of one).  ``bind_to_gpu`` pins the calling process to the
CPUs NVML reports as local to the device *before* the pinned buffers are allocated (first touch decides the
node).  Best effort: silently does nothing when NVML or the affinity call is unavailable.
"""

from __future__ import annotations

import os

__all__ = ["bind_to_gpu"]


def bind_to_gpu(device_index: int) -> bool:
    """Restrict this process to the CPUs local to CUDA device ``device_index``; returns True when applied."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
        phys = int(vis[device_index]) if device_index < len(vis) else device_index
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False
