"""Multi-GPU host logic on CPU: two gloo ranks shard a fan by iray % world, trace their
shards with the oracle standing in for the device, reduce the deposition profile and gather the per-ray
summaries; the result must equal the single-process run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _oracle as orc
    from _cases import init_case, oracle_fan
    from rays_b200.sharding import shard_indices, reduce_profile, gather_summaries
    cfg = init_case("axisym_deposition_fan.in", nstep_max=300)
    r, n, w, _, _ = oracle_fan(cfg, n_rindex_theta=4, delta_rindex_theta=0.1, n_rindex_phi=5, delta_rindex_phi=0.07)
    idx = shard_indices(r.shape[0], rank, world)
    o, st, _ = orc.trace(cfg, r[idx], n[idx], w[idx])
    prof, q = orc.deposition(cfg, o, 64, 0.0, 1.0)
    tot = reduce_profile(torch.from_numpy(np.append(prof, q)), dst=0)
    npts = gather_summaries(torch.from_numpy(o.npoints.astype(np.int64)), r.shape[0], rank, world)
    # the path bench.py's config 5 takes: fixed-point bins (one contribution per ray here, quantised like the device quantises every
    # segment), integer reduce, packed summaries all-gathered and put back in fan order
    from rays_b200.sharding import fixed_point_unit, reduce_bins, gather_packed_summaries, unshard_rows
    from rays_b200 import ResultArrays
    unit = fixed_point_unit(float(np.sum(np.abs(w))))
    acc = np.zeros(64, dtype=np.int64)
    for j in range(len(idx)):
        one = ResultArrays(0, int(cfg.nv), int(cfg.nstep_max) + 1, store=False)
        one.nray = 1
        one.ray_vec, one.residual, one.npoints, one.initial_ray_power = o.ray_vec[j:j + 1], o.residual[j:j + 1], o.npoints[j:j + 1], o.initial_ray_power[j:j + 1]
        one.c.nray = 1
        import ctypes as C
        from rays_b200 import _abi
        one.c.ray_vec = one.ray_vec.ctypes.data_as(_abi.c_double_p); one.c.npoints = one.npoints.ctypes.data_as(_abi.c_int32_p)
        one.c.initial_ray_power = one.initial_ray_power.ctypes.data_as(_abi.c_double_p)
        pj, _ = orc.deposition(cfg, one, 64, 0.0, 1.0)
        acc += np.rint(pj / unit).astype(np.int64)
    acc_t = reduce_bins(torch.from_numpy(acc.copy()), dst=0)
    n_max = (r.shape[0] + world - 1) // world
    packed = np.zeros((n_max, 3))
    packed[: len(idx), 0] = o.npoints; packed[: len(idx), 1] = o.ray_stop_code; packed[: len(idx), 2] = idx
    gath = unshard_rows(gather_packed_summaries(torch.from_numpy(packed)), r.shape[0], world)
    if rank == 0:
        np.save(os.path.join(out_dir, "prof.npy"), tot.numpy())
        np.save(os.path.join(out_dir, "npts.npy"), npts.numpy())
        np.save(os.path.join(out_dir, "acc.npy"), acc_t.numpy())
        np.save(os.path.join(out_dir, "gath.npy"), gath.numpy())
    np.save(os.path.join(out_dir, f"acc_rank{rank}.npy"), acc)
    dist.destroy_process_group()


def test_two_rank_shard_reduce_gather(built, tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    import _oracle as orc
    from _cases import init_case, oracle_fan
    cfg = init_case("axisym_deposition_fan.in", nstep_max=300)
    r, n, w, _, _ = oracle_fan(cfg, n_rindex_theta=4, delta_rindex_theta=0.1, n_rindex_phi=5, delta_rindex_phi=0.07)
    o, st, _ = orc.trace(cfg, r, n, w)
    prof, q = orc.deposition(cfg, o, 64, 0.0, 1.0)
    got = np.load(tmp_path / "prof.npy")
    assert np.max(np.abs(got[:-1] - prof)) <= 1e-12 * np.max(np.abs(prof))
    assert abs(got[-1] - q) <= 1e-12 * abs(q)
    assert np.array_equal(np.load(tmp_path / "npts.npy"), o.npoints.astype(np.int64))
    # fixed-point path: the reduced bins are the exact integer sum of the shards' bins, whatever the sharding
    acc = np.load(tmp_path / "acc.npy")
    assert np.array_equal(acc, np.load(tmp_path / "acc_rank0.npy") + np.load(tmp_path / "acc_rank1.npy"))
    from rays_b200.sharding import fixed_point_unit
    unit = fixed_point_unit(float(np.sum(np.abs(w))))
    assert np.max(np.abs(acc * unit - prof)) <= 1e-12 * np.max(np.abs(prof))
    gath = np.load(tmp_path / "gath.npy")
    assert np.array_equal(gath[:, 2], np.arange(r.shape[0])) and np.array_equal(gath[:, 0], o.npoints)


def test_shard_indices_partition():
    from rays_b200.sharding import shard_indices
    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if n else np.zeros(0, dtype=np.int64)
            assert np.array_equal(allidx, np.arange(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
