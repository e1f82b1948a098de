"""Site sharding over ranks and the single collective of the path.

Sites are independent, so a job is cut into contiguous global site ranges balanced by cumulative read
count; rank r scores sites [bounds[r], bounds[r+1]) with site_id_base = bounds[r], which makes the
result independent of the number of ranks.  The only exchange is one all-gather of the per-site outputs
(site_prob float32 + mod_count int32 = 8 bytes per site, packed into one tensor).
Backend: NCCL for CUDA tensors, gloo for the CPU tests of this logic.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np


def env_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process per GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def bind_host_to_device(device_index: int, min_cpus: int = 4) -> List[int]:
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index` (its NUMA node / PCIe root), so
    that the pinned staging buffers a rank allocates afterwards are first-touched next to its GPU and its ingest / CSV
    threads stay on that socket.  One process per GPU on a two-socket 8-GPU box otherwise sends half of the ranks' H2D
    traffic across the inter-socket link.  Best effort: returns the CPU list it bound to, or [] (nothing changed) when
    NVML is unavailable, CUDA_VISIBLE_DEVICES remaps devices in a way NVML cannot see, or the set is implausibly small."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if device_index >= len(ids) or not ids[device_index].isdigit():
                    return []
                physical = int(ids[device_index])
            else:
                physical = device_index
            handle = pynvml.nvmlDeviceGetHandleByIndex(physical)
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
            cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        finally:
            pynvml.nvmlShutdown()
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in cpus if c in allowed)
        if len(cpus) < min_cpus or len(cpus) >= len(allowed):
            return []
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:      # noqa: BLE001 - purely an optimisation
        return []


def shard_bounds(n_reads: np.ndarray, world_size: int) -> List[int]:
    """Contiguous site ranges with (nearly) equal total reads.  Returns world_size + 1 site indices."""
    n_reads = np.asarray(n_reads, dtype=np.int64)
    n_sites = len(n_reads)
    if world_size <= 1 or n_sites == 0:
        return [0] + [n_sites] * max(1, world_size)
    cum = np.cumsum(n_reads)
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        b = int(np.searchsorted(cum, target, side="left")) + 1     # first prefix reaching the target
        if b > 0 and b <= n_sites and abs(cum[b - 1] - target) > abs((cum[b - 2] if b >= 2 else 0) - target):
            b -= 1
        bounds.append(min(max(b, bounds[-1]), n_sites))
    bounds.append(n_sites)
    return bounds


def all_gather_rows(rows, bounds: List[int], group=None):
    """One all-gather of every rank's per-site rows (float32 [n_local, C]) -> [n_sites, C] on every rank, shards padded
    to the longest one.  Used for the per-pass predictions of validate() (C = n_iterations)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [bounds[r + 1] - bounds[r] for r in range(world)]
    assert rows.dim() == 2 and rows.shape[0] == sizes[rank] and rows.dtype == torch.float32
    m, c = max(max(sizes), 1), rows.shape[1]
    pack = torch.zeros((m, c), dtype=torch.float32, device=rows.device)
    pack[: sizes[rank]] = rows
    out = torch.empty((world * m, c), dtype=torch.float32, device=rows.device)
    dist.all_gather_into_tensor(out, pack, group=group)
    out = out.view(world, m, c)
    return torch.cat([out[r, : sizes[r]] for r in range(world)])


def all_gather_site_outputs(site_prob, mod_count, bounds: List[int], group=None):
    """One all-gather of every rank's (site_prob, mod_count) -> full-length tensors on every rank.

    site_prob float32 [n_local], mod_count int32 [n_local] (torch tensors, CUDA for NCCL / CPU for gloo).
    Shards are padded to the longest shard and packed as [max_shard, 2] float32 words so that a single
    collective moves both outputs; mod_count travels bit-cast, not converted."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [bounds[r + 1] - bounds[r] for r in range(world)]
    assert site_prob.numel() == sizes[rank] == mod_count.numel()
    m = max(max(sizes), 1)
    pack = torch.zeros((m, 2), dtype=torch.float32, device=site_prob.device)
    pack[: sizes[rank], 0] = site_prob
    pack[: sizes[rank], 1] = mod_count.view(torch.float32)
    out = torch.empty((world * m, 2), dtype=torch.float32, device=site_prob.device)
    dist.all_gather_into_tensor(out, pack, group=group)
    out = out.view(world, m, 2)
    sp = torch.cat([out[r, : sizes[r], 0] for r in range(world)])
    mc = torch.cat([out[r, : sizes[r], 1].contiguous().view(torch.int32) for r in range(world)])
    return sp, mc
