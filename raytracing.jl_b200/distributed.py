"""Multi-GPU plumbing: one process per GPU, tracks sharded by contiguous uid range (equal total track length
per rank), mesh replicated, ONE collective -- the NCCL all-reduce of per-element sum(delta*len) inside
rt_volumes (reference src/trackgenerator.jl:378-386 is the serial loop it replaces).  torch.distributed is only
used to ship the 128-byte NCCL unique id; the communicator lives inside librt_b200.so."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib


def track_costs(tg, cost_w: float = 0.0) -> np.ndarray:
    """Planning cost of all tracks in uid order: the length (vectorised restatement of src/trackgenerator.jl:188-226; only used to
    balance shards, not for results) plus ``cost_w`` times the reciprocal sines of the angles at which the track leaves and
    reaches the bounding box -- the cells it spends in the boundary band, where every transition takes the slow side of the
    walk.  Needs tg.azimuthal_quadrature.phis, i.e. the tables of trace_.  (Twin of k_trace's planning branch, trace.cuh.)"""
    from .api import nazim2, nazim4

    aq = tg.azimuthal_quadrature
    n2, n4 = nazim2(aq), nazim4(aq)
    dx, dy = tg.mesh.width, tg.mesh.height
    out = []
    for i in range(1, n2 + 1):
        nx, ny = int(tg.n_tracks_x[i - 1]), int(tg.n_tracks_y[i - 1])
        right = i <= n4
        dxe, dye = dx / nx, dy / ny
        j = np.arange(1, nx + ny + 1, dtype=np.float64)
        onx = j <= nx
        px = np.where(onx, np.where(right, dxe * (nx - j + 0.5), dxe * (j - 0.5)), 0.0 if right else dx)
        py = np.where(onx, 0.0, dye * (j - nx - 0.5))
        m = np.tan(aq.phis[i - 1])
        qx = px - (py - dy) / m
        qy = np.full_like(qx, dy)
        bad = ~((0 <= qx) & (qx <= dx))
        if right:
            qy = np.where(bad, py + m * (dx - px), qy)
            qx = np.where(bad, dx, qx)
        else:
            qy = np.where(bad, py - m * px, qy)
            qx = np.where(bad, 0.0, qx)
        cost = np.hypot(px - qx, py - qy)
        if cost_w > 0.0:
            t = abs(m)
            c = 1.0 / np.sqrt(1.0 + t * t)
            sn = t * c
            s_in = np.where(onx, sn, c)
            s_out = np.where(bad, c, sn)
            cost = cost + cost_w * (1.0 / np.maximum(s_in, 1e-6) + 1.0 / np.maximum(s_out, 1e-6))
        out.append(cost)
    return np.concatenate(out)


def track_lengths(tg) -> np.ndarray:
    return track_costs(tg, 0.0)


def plan_shards(tg, n_parts: int) -> np.ndarray:
    """bounds[r] .. bounds[r+1] (1-based uids, end exclusive) with equal total planning cost per part (track_costs; the weight of
    the boundary band comes from the context when the TrackGenerator has one: band_cost fast transitions per band cell)."""
    cost_w = 0.0
    if getattr(tg, "_ctx", None):
        rho = tg.info("rho")
        cost_w = tg.info("band_cost") / rho if rho > 0 else 0.0
    lens = track_costs(tg, cost_w)
    cum = np.concatenate([[0.0], np.cumsum(lens)])
    targets = cum[-1] * (np.arange(1, n_parts) / n_parts)
    inner = np.searchsorted(cum, targets, side="left") + 1
    bounds = np.concatenate([[1], inner, [lens.size + 1]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def bind_to_gpu_numa(device_index: int) -> bool:
    """Restrict this process to the CPUs NVML reports as local to the GPU (its NUMA node), so that the page-locked host buffers
    it allocates afterwards -- the destination of the Segment downloads -- are placed next to that GPU's PCIe root.  With one
    process per GPU this keeps N concurrent device->host streams from funnelling into one socket's memory.  Returns False (and
    changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [int(v) for v in vis.split(",") if v.strip().isdigit()]
        phys = ids[device_index] if ids and device_index < len(ids) else device_index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1 and 64 * w + b < n_cpu}
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload: bytes | None, n: int, group=None, device=None) -> bytes:
    """Ship ``n`` bytes from rank 0 to every rank of the torch.distributed group (the NCCL unique id)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    t = torch.zeros(n, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    if dist.get_backend(group) == "nccl":
        t = t.cuda() if device is None else t.to(device)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().numpy().tobytes())


def init_comm(tg, group=None) -> None:
    """Create the NCCL communicator of this TrackGenerator's context across the torch.distributed group."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    L = _lib.lib()
    buf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(tg._ctx, L.rt_comm_unique_id(tg._ctx, buf))
    ident = broadcast_bytes(buf.raw if rank == 0 else None, 128, group)
    _lib.check(tg._ctx, L.rt_comm_init(tg._ctx, world, rank, ident))
    tg._has_comm = True


def shard_range(tg, rank: int, n_ranks: int):
    """(uid_begin, uid_end) of ``rank``: 1-based, end exclusive."""
    b = plan_shards(tg, n_ranks)
    return int(b[rank]), int(b[rank + 1])
