"""Multi-GPU work distribution for the colour stage: frames are independent (no filter in the reference chain has
temporal state), so one clip is split into contiguous frame ranges (BASELINE cfg 4) and a batch of clips is dealt out
clip by clip (cfg 5).  There is no collective on the data path; ``torch.distributed`` is used by the harness only to
agree on the plan and to reduce timings."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple


def frame_range(num_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of rank ``rank``; ranges differ by at most one frame and tile [0, num_frames)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(max(0, num_frames), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gop_aligned_ranges(num_frames: int, world: int, gop: int) -> List[Tuple[int, int]]:
    """Frame ranges whose boundaries fall on multiples of ``gop`` (so every rank's encoder starts on a keyframe)."""
    gop = max(1, gop)
    gops = (num_frames + gop - 1) // gop
    out = []
    for r in range(world):
        a, b = frame_range(gops, r, world)
        out.append((min(a * gop, num_frames), min(b * gop, num_frames)))
    return out


@dataclass(frozen=True)
class Clip:
    name: str
    width: int
    height: int
    pix_fmt: str
    frames: int
    lut: str

    @property
    def cost(self) -> float:
        """Relative cost: pixels to process (the stage is per-pixel work)."""
        return float(self.width) * self.height * self.frames


def assign_clips(clips: Sequence[Clip], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of whole clips to ranks; deterministic (ties by index).
    Returns, per rank, the indices of its clips in processing order."""
    if world <= 0:
        raise ValueError("world must be positive")
    load = [0.0] * world
    plan: List[List[int]] = [[] for _ in range(world)]
    for idx in sorted(range(len(clips)), key=lambda i: (-clips[i].cost, i)):
        r = min(range(world), key=lambda k: (load[k], k))
        plan[r].append(idx)
        load[r] += clips[idx].cost
    return plan


def luts_per_rank(clips: Sequence[Clip], plan: Sequence[Sequence[int]]) -> List[Dict[str, int]]:
    """How many clips of each rank share each LUT (every distinct LUT is uploaded once per device)."""
    out = []
    for mine in plan:
        d: Dict[str, int] = {}
        for i in mine:
            d[clips[i].lut] = d.get(clips[i].lut, 0) + 1
        out.append(d)
    return out


def bind_to_device_cpus(device: int):
    """Pin the calling process to the CPU cores local to GPU ``device`` (NVML's ideal-affinity mask) so that the pinned host buffers
    it allocates afterwards land on that GPU's NUMA node: with one process per GPU, host frames otherwise cross the socket
    interconnect on their way to PCIe.  Returns the previous affinity (restore it with ``os.sched_setaffinity(0, prev)``), or None
    when NVML or the mask is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return prev
    except Exception:
        return None
