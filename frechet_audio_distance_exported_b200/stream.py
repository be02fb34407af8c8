"""Host -> device streaming of PCM for the FAD path (the reference moves one clip at a time
with a blocking `.to(device)` / `.cpu()` pair per clip, fad.py:389-396).

`HostRing` owns a ring of device staging buffers fed by ONE dedicated copy stream.  Chunk i is copied into buffer
i % depth as soon as the kernels that last read that buffer have finished — never earlier, never later — so with a
depth of 4 the copy engine runs up to four chunks ahead of the compute stream and is never idle while work is left
(a two-buffer scheme stalls the copy engine whenever a chunk's kernels outlast the next chunk's copy).  The host
thread only enqueues: no synchronisation inside the loop.

`bind_to_gpu_numa_node` pins the calling process to the CPUs of the NUMA node the GPU hangs off, so that pinned host
buffers allocated afterwards are local to the GPU's PCIe root (first-touch policy).  With eight ranks pulling PCM
at the same time, remote-socket buffers are what limits the host link.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple

import torch


def chunk_bounds(n: int, chunk: int, first: Optional[int] = None) -> List[Tuple[int, int]]:
    """[(start, count)] covering range(n) in chunks of `chunk`; the first chunk may be shorter (`first`) so that the
    kernels start after a fraction of a chunk's copy time."""
    out, c0 = [], 0
    while c0 < n:
        nc = min(first if (c0 == 0 and first) else chunk, n - c0)
        out.append((c0, nc))
        c0 += nc
    return out


class HostRing:
    def __init__(self, device: torch.device, depth: int = 4):
        self.device = device
        self.depth = int(depth)
        self.stream = torch.cuda.Stream(device=device)
        self.bufs: List[torch.Tensor] = []
        self.copied = [torch.cuda.Event() for _ in range(self.depth)]
        self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self.next = 0
        self._key = None

    def _ensure(self, chunk: int, row_shape: Tuple[int, ...], dtype: torch.dtype) -> None:
        key = (tuple(row_shape), dtype)
        if key == self._key and self.bufs and self.bufs[0].shape[0] >= chunk:
            return                                                   # buffers only ever grow
        torch.cuda.current_stream(self.device).synchronize()       # nobody may still read the old buffers
        self.stream.synchronize()
        self.bufs = [torch.empty((chunk,) + tuple(row_shape), dtype=dtype, device=self.device) for _ in range(self.depth)]
        self.copied = [torch.cuda.Event() for _ in range(self.depth)]
        self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self.next = 0
        self._key = key

    def run(self, src: torch.Tensor, chunk: int, consume: Callable[[torch.Tensor, int, int], None],
            first: Optional[int] = None) -> None:
        """src: HOST tensor [n, ...] (pinned for asynchronous copies).  For every chunk: copy on the ring's stream,
        then `consume(device_view, start, count)` on the CURRENT stream.  Events persist across calls, so the first
        copy of a call overlaps the tail of the previous call's kernels."""
        n = src.shape[0]
        if n == 0:
            return
        chunk = max(1, min(int(chunk), n))
        self._ensure(chunk, tuple(src.shape[1:]), src.dtype)
        cur = torch.cuda.current_stream(self.device)
        for c0, nc in chunk_bounds(n, chunk, first if n > chunk else None):
            b = self.next
            self.next = (self.next + 1) % self.depth
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(self.done[b])               # no-op until the event has been recorded once
                self.bufs[b][:nc].copy_(src[c0:c0 + nc], non_blocking=True)
                self.copied[b].record(self.stream)
            cur.wait_event(self.copied[b])
            consume(self.bufs[b][:nc], c0, nc)
            self.done[b].record(cur)


def gpu_numa_cpus(device_index: int) -> Optional[List[int]]:
    """CPUs local to the GPU's PCIe root complex, from sysfs; None when the topology cannot be read."""
    try:
        try:
            pr = torch.cuda.get_device_properties(device_index)
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        except AttributeError:
            import pynvml
            pynvml.nvmlInit()
            bus_id = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
            bus_id = bus_id.decode() if isinstance(bus_id, bytes) else bus_id
            bdf = bus_id.lower()[-12:]                               # "00000000:1b:00.0" -> "0000:1b:00.0"
        path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
        with open(path) as fh:
            txt = fh.read().strip()
        cpus: List[int] = []
        for part in txt.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Restrict this process to the CPUs next to `device_index` (and therefore its first-touch allocations to that
    NUMA node).  Returns the number of CPUs bound to, or None if nothing was changed."""
    cpus = gpu_numa_cpus(device_index)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        target = set(cpus) & set(allowed)
        if not target or target == set(allowed):
            return None
        os.sched_setaffinity(0, target)
        return len(target)
    except Exception:
        return None
