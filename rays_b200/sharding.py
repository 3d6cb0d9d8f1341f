"""Ray sharding across the GPUs of one node: one process per GPU, rays partitioned by
`iray % world == rank` (interleaved, because ray length varies smoothly along the launch loops), no
traffic during integration, then one reduce of the deposition profile and one gather of the fixed-size
per-ray summaries over torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
Device-side sharding of a fan that is already in HBM is rays_b200_fan_shard."""
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
