#!/usr/bin/env python
"""Config 5: an axisym_toroid + solovev_magnetics fan with fundamental ECH damping, sharded
over the GPUs of one node, deposition profile vs psi_N binned on the device while tracing (no trajectory
storage) and summed with ONE NCCL reduce.

  python examples/deposition_fan.py --grid 1024                      # 1 GPU, 1024 x 1024 = 1M candidates
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/deposition_fan.py --grid 2896   # 8.4M rays

Every rank builds the launch fan on its GPU (1 ms), keeps rays iray % world == rank (rays_b200_fan_shard),
traces them with fused binning (rays_b200_trace_device_binned) and contributes n_bins + 1 doubles.
--check re-traces the whole fan on rank 0 and compares the reduced profile with the single-GPU one."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=1024, help="n_rindex_theta = n_rindex_phi")
    ap.add_argument("--bins", type=int, default=501)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import rays_b200 as rb
    from rays_b200.sharding import reduce_profile

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    text = open(rb.config_path("axisym_deposition_fan.in")).read()
    n = args.grid
    text = text.replace("n_rindex_theta = 1024", f"n_rindex_theta = {n}").replace("n_rindex_phi = 1024", f"n_rindex_phi = {n}")
    text = text.replace("delta_rindex_theta = 0.0003910068426197458", f"delta_rindex_theta = {0.4 / max(n - 1, 1)!r}")
    text = text.replace("delta_rindex_phi = 0.0003421309872922776", f"delta_rindex_phi = {0.35 / max(n - 1, 1)!r}")
    text = text.replace("nray_max = 8388608", f"nray_max = {max(n * n, 1)}")
    d = tempfile.mkdtemp(prefix="rays_dep_")
    path = os.path.join(d, "rays.in")
    open(path, "w").write(text)

    rb.initialize(path, ray_init=True, device=local_rank)       # launch fan on the device
    cfg = rb.host_cfg()
    rvec0, nvec0, wt = rb.get_fan()
    nray_total = rvec0.shape[0]
    rb.set_config(cfg)
    rb.fan_upload(rvec0, nvec0, wt)
    rb.fan_shard(rank, world)
    part = torch.zeros(args.bins + 1, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = rb.trace_device(store=False, bins=(args.bins, 0.0, 1.0))
    prof_local, q_local = rb.deposition(args.bins, 0.0, 1.0, d_profile_out=part.data_ptr())
    total = reduce_profile(part, dst=0)                          # the single collective of the run
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = torch.tensor([float(st["ray_steps"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(steps)
    if rank == 0:
        prof = total[:-1].cpu().numpy()
        out = {"n_gpus": world, "rays": nray_total, "ray_steps": int(steps.item()), "wall_s": dt, "kernel_ms_rank0": st["kernel_ms"],
               "ray_steps_per_sec": steps.item() / dt, "Q_sum": float(total[-1].item()), "profile_peak_bin": int(np.argmax(prof)),
               "profile_peak": float(prof.max()), "kernel": st["kernel"]}
        if args.check:
            rb.fan_upload(rvec0, nvec0, wt)
            rb.trace_device(store=False, bins=(args.bins, 0.0, 1.0))
            p1, q1 = rb.deposition(args.bins, 0.0, 1.0)
            out["max_abs_diff_vs_single_gpu"] = float(np.max(np.abs(p1 - prof)))
            out["rel_diff_Q_sum"] = abs(q1 - out["Q_sum"]) / abs(q1)
            assert out["max_abs_diff_vs_single_gpu"] <= 1e-12 * max(float(np.max(np.abs(p1))), 1e-300) * 1e3
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
