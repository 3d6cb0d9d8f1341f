"""Ray sharding across the GPUs of one node: one process per GPU, rays partitioned by
`iray % world == rank` (interleaved, because ray length varies smoothly along the launch loops), no
traffic during integration, then one reduce of the deposition profile and one gather of the fixed-size
per-ray summaries over torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
Device-side sharding of a fan that is already in HBM is rays_b200_fan_shard.  This is the one-process-per-GPU
plumbing bench.py uses under torchrun; a single host process (the Fortran `program rays`) uses the library's own
rays_b200_init_multi / rays_b200_trace_multi(_binned), which runs the same two collectives with NCCL inside the library."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(nray: int, rank: int, world: int) -> np.ndarray:
    return np.arange(rank, nray, world, dtype=np.int64)


def reduce_profile(profile_and_qsum: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """Sum of the per-GPU partial profiles (n_bins doubles + Q_sum) on rank `dst`
    (calculate_deposition_profiles sums over rays: deposition_profiles_m.f90:244-257)."""
    t = profile_and_qsum.clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t


def gather_summaries(local: torch.Tensor, nray_total: int, rank: int, world: int) -> torch.Tensor:
    """All-gather one per-ray summary array of the shards and put it back into fan order."""
    if not dist.is_initialized() or world == 1:
        return local
    n_max = (nray_total + world - 1) // world
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.empty((nray_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        cnt = len(range(r, nray_total, world))
        out[r::world] = parts[r][:cnt]
    return out


# ---- fixed-point deposition bins (rays_b200.h: rays_b200_deposition_fixed) -------------------------------------------------------
def fixed_point_unit(total_weight: float) -> float:
    """Value of one count of the device's 64-bit deposition bins: 2^(e-62) with 2^e > sum |ray_pwr_wt| of the WHOLE fan (the library
    forms it at fan upload / launch and keeps it through rays_b200_fan_shard, so every shard uses the same unit)."""
    import math
    e = math.frexp(total_weight)[1] if total_weight > 0.0 and math.isfinite(total_weight) else 0
    return math.ldexp(1.0, e - 62)


def reduce_bins(acc: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """THE collective of a sharded deposition run: integer sum of the per-GPU bins on rank `dst`.  Integer addition is associative,
    so the reduced profile is bitwise independent of the number of shards and of the reduction order."""
    assert acc.dtype == torch.int64
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM)
    return acc


def gather_packed_summaries(local: torch.Tensor, gathered: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather of the packed per-ray summaries (rays_b200_summaries_pack rows, every rank padded to the same row count):
    returns [world * rows, row_doubles], block r = the rays r, r + world, ... in order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if gathered is None:
        gathered = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local)
    return gathered


def unshard_rows(gathered: torch.Tensor, nray_total: int, world: int) -> torch.Tensor:
    """rows of gather_packed_summaries back in fan order (ray i sits in block i % world at row i // world)"""
    n_max = gathered.shape[0] // world
    out = torch.empty((nray_total,) + tuple(gathered.shape[1:]), dtype=gathered.dtype, device=gathered.device)
    for r in range(world):
        cnt = len(range(r, nray_total, world))
        out[r::world] = gathered[r * n_max: r * n_max + cnt]
    return out
