"""Frame sharding across the GPUs of one box.

Frames are independent in every kernel of the path, so rank r simply owns a contiguous block
of frames — the same partition the reference gets from DistributedSampler
(model/MvRoPose_FR3.py:946) — and there is NO collective inside decode -> triangulate -> FK.
The only communication is the final result gather (NCCL all_gather over NVLink; gloo in the
CPU tests), < 1 KB per frame.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def frame_range(n_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of rank `rank`; the remainder goes to the LAST ranks."""
    if world_size < 1 or not (0 <= rank < world_size) or n_frames < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_frames, world_size)
    first_big = world_size - rem  # ranks >= first_big get base + 1 frames
    begin = rank * base + max(0, rank - first_big)
    return begin, begin + base + (1 if rank >= first_big else 0)


def gpu_local_cpus(device_index: int) -> List[int]:
    """CPUs on the NUMA node / socket the GPU hangs off (NVML's ideal CPU affinity); [] if unknown."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        return [c for c in cpus if c < n_cpu]
    except Exception:
        return []


def bind_to_gpu_numa(device_index: int) -> Optional[List[int]]:
    """Pin the calling thread (and the threads it creates afterwards) to the CPUs next to GPU
    `device_index`, so that pinned host buffers allocated from now on are first-touched on that
    socket and H2D copies do not cross the inter-socket link. One process per GPU (torchrun does not
    bind ranks): call it right after torch.cuda.set_device, BEFORE allocating pinned memory.
    Returns the CPU list it bound to, or None when the topology is unknown or the binding is refused
    (cgroup cpusets)."""
    cpus = gpu_local_cpus(device_index)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        target = sorted(set(cpus) & set(allowed))
        if not target:
            return None
        os.sched_setaffinity(0, target)
        return target
    except Exception:
        return None


def gather_frames(local: Dict[str, torch.Tensor], n_frames: int, group=None) -> Dict[str, torch.Tensor]:
    """All-gather per-frame result tensors (leading dimension = this rank's frames, in
    frame_range order) into full tensors of leading dimension n_frames on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [frame_range(n_frames, r, ws)[1] - frame_range(n_frames, r, ws)[0] for r in range(ws)]
    cap = max(counts)
    out = {}
    for name, t in local.items():
        if t.dim() == 0:
            continue
        if t.shape[0] != counts[rank]:
            raise ValueError(f"{name}: leading dimension {t.shape[0]} != this rank's {counts[rank]} frames")
        pad = t
        if counts[rank] < cap:  # equal-sized contributions: one all_gather_into_tensor
            pad = torch.cat([t, t.new_zeros((cap - counts[rank],) + tuple(t.shape[1:]))], dim=0)
        buf = pad.new_empty((ws * cap,) + tuple(t.shape[1:]))
        dist.all_gather_into_tensor(buf, pad.contiguous(), group=group)
        parts = [buf[r * cap: r * cap + counts[r]] for r in range(ws)]
        out[name] = torch.cat(parts, dim=0) if any(c != cap for c in counts) else buf
    return out


class PackedResults:
    """Per-frame result tensors laid out back to back in ONE flat device buffer, so the final
    gather is a single all_gather_into_tensor however many result arrays there are (launch
    latency, not bytes, is what a ~1 MB gather costs on NVSwitch). Every entry is a dense,
    contiguous view the kernels can write straight into.

        pr = PackedResults({"X_tri": ((B, K, 3), torch.float32), "score": ((B, V, K), torch.float32)}, device)
        kernel(..., pr["X_tri"].data_ptr(), ...)
        full = pr.all_gather()          # {"X_tri": (world, B, K, 3), ...}: views, no copy
    Requires the same frame count on every rank (weak-scaling batches)."""

    def __init__(self, spec: Dict[str, tuple], device, storage: torch.Tensor = None):
        self._spec, self._views, off = {}, {}, 0
        for name, (shape, dtype) in spec.items():
            if torch.empty((), dtype=dtype).element_size() != 4:
                raise ValueError("PackedResults holds 32-bit types only")
            n = 1
            for d in shape:
                n *= int(d)
            self._spec[name] = (off, n, tuple(shape), dtype)
            off += n
        if storage is None:
            storage = torch.empty((off,), dtype=torch.float32, device=device)
        elif storage.dtype != torch.float32 or storage.numel() != off or not storage.is_contiguous():
            raise ValueError(f"storage must be a contiguous float32 tensor of {off} elements")
        self.flat = storage
        for name, (o, n, shape, dtype) in self._spec.items():
            self._views[name] = self.flat[o:o + n].view(dtype).view(shape)
        self._gathered = None

    @staticmethod
    def numel_of(spec: Dict[str, tuple]) -> int:
        total = 0
        for shape, _ in spec.values():
            n = 1
            for d in shape:
                n *= int(d)
            total += n
        return total

    def __getitem__(self, name: str) -> torch.Tensor:
        return self._views[name]

    def keys(self):
        return self._views.keys()

    def all_gather(self, group=None, async_op: bool = False):
        """One collective for every result array. With async_op=True returns (views, work): the
        gather runs on the communicator's own stream, overlapping whatever the caller launches
        next; call work.wait() before reading the views or overwriting this buffer again."""
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            views = {k: v.unsqueeze(0) for k, v in self._views.items()}
            return (views, None) if async_op else views
        ws = dist.get_world_size(group)
        if self._gathered is None or self._gathered.numel() != ws * self.flat.numel():
            self._gathered = self.flat.new_empty((ws * self.flat.numel(),))  # concatenated form (gloo and nccl)
        work = dist.all_gather_into_tensor(self._gathered, self.flat, group=group, async_op=async_op)
        g2 = self._gathered.view(ws, self.flat.numel())
        views = {name: g2[:, o:o + n].view(dtype).view((ws,) + shape)
                 for name, (o, n, shape, dtype) in self._spec.items()}
        return (views, work) if async_op else views


class ResultRing:
    """`slots` PackedResults laid out in ONE device buffer: batch k of a job writes its results
    into slot k, and the whole job's results are exchanged with a single final
    all_gather_into_tensor (the only collective of the path; nothing is exchanged per batch)."""

    def __init__(self, spec: Dict[str, tuple], slots: int, device):
        self.spec, self.slots = spec, int(slots)
        self.record = PackedResults.numel_of(spec)
        self.flat = torch.empty((self.slots * self.record,), dtype=torch.float32, device=device)
        self.slot = [PackedResults(spec, device, self.flat[i * self.record:(i + 1) * self.record])
                     for i in range(self.slots)]
        self._gathered = None

    def final_gather(self, used_slots: int = None, group=None) -> torch.Tensor:
        """Gather slots [0, used_slots) of every rank: returns (world, used_slots, record) float32."""
        n = self.slots if used_slots is None else int(used_slots)
        src = self.flat[: n * self.record]
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return src.view(1, n, self.record)
        ws = dist.get_world_size(group)
        if self._gathered is None or self._gathered.numel() < ws * src.numel():
            self._gathered = self.flat.new_empty((ws * self.slots * self.record,))
        dst = self._gathered[: ws * src.numel()]
        dist.all_gather_into_tensor(dst, src, group=group)
        return dst.view(ws, n, self.record)
