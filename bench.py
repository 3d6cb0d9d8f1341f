#!/usr/bin/env python
"""bench.py — ray-steps/s and wall time per fan for the hot path, on N GPUs of one node.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one trace of the whole launch fan.  Workload (BASELINE.json configs[3]):
1 048 576-candidate Solov'ev fan, ray_deriv_name='numerical', RK4, ds = 5e-11, nstep_max = 1000, per GPU
(weak scaling: every rank traces its own fan of that size; rank r shifts rindex_phi0 by r*1e-4).
  value  = ray-steps/s with the fan resident in HBM, trajectories written to HBM (device time, CUDA events)
  e2e    = the same through rays_b200_trace with HOST buffers: H2D of the fan, kernels, D2H of the
           trajectories + summaries into the reference's ray_results_m layout, inside the timed region
  roofline = algorithmic fp64 flops of the trace kernel / its duration, against the DFMA peak measured on
           this GPU in the same run (MEASURED_PEAKS.json carries no fp64 figure)
  cpu_baseline = the C++ oracle (restatement of the Fortran; the Fortran itself cannot be built here) on
           a bounded sample of the same fan, all host cores
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "solovev_fan_1M"
# exact algorithmic flops per ray-step of this workload from the oracle's counting scalar type
# (oracle_count_flops; +,-,*,/,sqrt and libm calls = 1 each, reference code as written): see DESIGN.md
FLOPS_PER_RAY_STEP_FALLBACK = 20789.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=0, help="debug: shrink the fan to about this many candidates")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-rays", type=int, default=98304, help="rays of the bounded CPU-baseline sample")
    ap.add_argument("--ode", default="", help="measurement aid: override ode_solver_name (RK4_ODE | SG_ODE)")
    ap.add_argument("--deriv", default="", help="measurement aid: override ray_deriv_name (cold | numerical)")
    return ap.parse_args()


def workload_namelist(rank: int, rays: int) -> str:
    """The config-4 namelist, rank-shifted; written to a temp dir (the Brz/ray_init side files are not needed)."""
    import rays_b200 as rb
    text = open(rb.config_path("solovev_fan_1M.in")).read()
    text = text.replace("rindex_phi0 = 0.05", f"rindex_phi0 = {0.05 + 1e-4 * rank!r}")
    if rays:
        # shrink the two outer (position) loops, keep the 32 x 32 direction grid
        npos = max(1, rays // 1024)
        nr = max(1, min(16, npos // 64))
        nt = max(1, min(64, npos // nr))
        text = text.replace("n_r_launch = 16", f"n_r_launch = {nr}").replace("n_theta_launch = 64", f"n_theta_launch = {nt}")
    d = tempfile.mkdtemp(prefix="rays_bench_")
    p = os.path.join(d, "rays.in")
    open(p, "w").write(text)
    return p


def host_threads() -> int:
    """threads for the CPU arms: every core this process may run on (torchrun exports OMP_NUM_THREADS=1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def host_memory_budget(world_local: int) -> float:
    """bytes of host memory one rank may lock for result arrays: half of min(cgroup limit, MemAvailable) / ranks"""
    avail = None
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = float(ln.split()[1]) * 1024.0
    except Exception:
        pass
    for f in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(f).read().strip()
            if v != "max":
                avail = min(avail, float(v)) if avail else float(v)
        except Exception:
            pass
    if not avail:
        avail = 64e9
    return 0.5 * avail / max(world_local, 1)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled during the timed region"""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.maxmhz, self._halt = gpu_index, [], set(), 0.0, threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.maxmhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.maxmhz or None, "reasons": sorted(self.reasons)}


def run_reference(args, rank, world):
    """The reference arm: the CPU implementation of the path (C++ oracle; kind 'port') on all host cores,
    each step a bounded sample of the workload's fan.  Rank 0 only."""
    if rank != 0:
        return
    import numpy as np
    import _oracle as orc
    from _cases import init_case, oracle_fan
    orc.build()
    import rays_b200 as rb
    from rays_b200 import _abi
    L = _abi.load()
    path = workload_namelist(0, args.rays)
    assert L.rays_host_initialize(path.encode(), 0) == 0, L.rays_host_last_error()
    cfg = rb.host_cfg()
    r, n, w, _, _ = oracle_fan(cfg, cap=1 << 21)
    stride = max(1, r.shape[0] // args.cpu_rays)
    idx = np.arange(0, r.shape[0], stride)
    rs, ns_, ws = r[idx].copy(), n[idx].copy(), w[idx].copy()
    cores = host_threads()
    for _ in range(max(args.warmup, 1) if args.steps > 1 else 1):
        orc.trace(cfg, rs[:256], ns_[:256], ws[:256], store=False, nthreads=cores)
    t0 = time.perf_counter()
    steps = 0
    for _ in range(args.steps):
        o, st, _ = orc.trace(cfg, rs, ns_, ws, store=True, nthreads=cores)
        steps += o.total_ray_steps
    dt = time.perf_counter() - t0
    val = steps / dt
    sample = f"{len(idx)} rays (every {stride}th of the {r.shape[0]}-ray fan), {steps // args.steps} ray-steps per step"
    print(json.dumps({
        "impl": "reference", "metric": "ray_steps_per_sec", "value": val, "unit": "ray-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "rays_per_gpu": int(r.shape[0]), "ode": "RK4", "ray_deriv": "numerical",
                                                         "ds": 5e-11, "nstep_max": 1000, "sharding": "fan per GPU, no collective"},
        "cpu_baseline": {"value": val, "unit": "ray-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C++ restatement of the reference's Fortran path (oracle/); the Fortran itself cannot be compiled in this image"}))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import numpy as np
    import torch
    import rays_b200 as rb
    from rays_b200 import _abi
    import ctypes as C

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- initialize(read_input): namelist -> module state -> launch fan built on the device ------------
    path = workload_namelist(rank, args.rays)
    rb.initialize(path, ray_init=True, device=local_rank)
    if args.ode or args.deriv:
        rb.set_ode(ode_solver_name=args.ode, ray_deriv_name=args.deriv, rel_err0=1e-6, abs_err0=1e-6, SG_error_limit=0.1)
    cfg = rb.host_cfg()
    rvec0, nvec0, wt = rb.get_fan()
    nray = rvec0.shape[0]
    nv, npa = int(cfg.nv), int(cfg.nstep_max) + 1
    L = _abi.load()
    stream = torch.cuda.ExternalStream(L.rays_b200_stream(), device=torch.device("cuda", local_rank))
    peak_tf, _ = rb.fp64_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    # ---- device-resident arm: fan in HBM, trajectories to HBM -------------------------------------------
    rb.set_config(cfg)
    rb.fan_upload(rvec0, nvec0, wt)
    for _ in range(args.warmup):
        st = rb.trace_device(store=True)
    barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_ms, steps_total, launches, trace_launches, pilot_ms = 0.0, 0, 0, 0, 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0.record(stream)
        st = rb.trace_device(store=True)
        e1.record(stream)
        torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e1)
        kernel_ms += st["trace_kernel_ms"]      # the trace kernel proper (the roofline's kernel)
        pilot_ms += st["resume_pass_ms"]         # of which: the resume pass over the suspended (long) rays
        steps_total += st["ray_steps"]
        launches += st["n_launches"]
        trace_launches += 1
    barrier(); torch.cuda.synchronize()
    clocks = sampler.finish()
    dev_ms_max = max_over_ranks(dev_ms)
    steps_all = sum_over_ranks(float(steps_total))
    value = steps_all / (dev_ms_max * 1e-3)
    ray_steps_per_fan = steps_total // max(args.steps, 1)
    kinfo = st

    # ---- roofline of the trace kernel --------------------------------------------------------------------
    flops_per_step = FLOPS_PER_RAY_STEP_FALLBACK
    try:
        import _oracle as orc
        orc.load()
        idx = np.arange(0, nray, max(1, nray // 64))
        fl, stp, _ = orc.count_flops(cfg, rvec0[idx], nvec0[idx], 0, len(idx))
        if stp > 0:
            flops_per_step = fl / stp
    except Exception:
        pass
    avg_kernel_s = kernel_ms * 1e-3 / max(trace_launches, 1)
    # DRAM bytes of the trace kernel from the committed `ncu --set full` capture of this very workload
    # (profiles/r1_trace_rk4_1M_ncu_full_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum of the first pass, which
    # takes 79 % of the fan's kernel time); null for the measurement-aid workloads that have no capture
    ncu_traffic = None
    if not (args.rays or args.ode or args.deriv):
        try:
            txt = open(os.path.join(ROOT, "profiles", "r1_trace_rk4_1M_ncu_full_summary.txt")).read()
            import re
            rd = float(re.search(r"dram__bytes_read.sum \[Gbyte\] = ([0-9.]+)", txt).group(1))
            wr = float(re.search(r"dram__bytes_write.sum \[Gbyte\] = ([0-9.]+)", txt).group(1))
            ncu_traffic = (rd + wr) * 1e9
        except Exception:
            ncu_traffic = None
    achieved_tf = flops_per_step * (steps_total / max(trace_launches, 1)) / avg_kernel_s / 1e12
    wb_bytes = (nv + 1) * 8.0 * (steps_total / max(trace_launches, 1))
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                "traffic": ncu_traffic, "peak_source": "DFMA microbenchmark on this GPU in this run (rays_b200_fp64_peak); MEASURED_PEAKS.json has no fp64 figure",
                "flops_per_ray_step": flops_per_step, "kernel": kinfo["kernel"], "grid": kinfo["grid"], "ctas_per_sm": kinfo["blocks_per_sm"],
                "hbm_writeback_gbs": wb_bytes / avg_kernel_s / 1e9, "avg_kernel_ms": avg_kernel_s * 1e3,
                "resume_pass_ms": pilot_ms / max(trace_launches, 1), "launches_per_fan": kinfo["n_passes"],
                "note": "achieved = algorithmic flops of one fan / summed CUDA-event duration of the trace kernel's launches for that fan (first pass + the resume pass over the time-sliced long rays)"}

    # ---- end-to-end arm: host buffers through rays_b200_trace ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        pinned = []

        def host_array(shape, dtype=np.float64):
            """page-locked array from the library's allocator (allocate_ray_results); pageable if that fails"""
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            p = C.c_void_p()
            if L.rays_b200_host_alloc(C.byref(p), max(nbytes, 8)) != 0:
                return np.zeros(shape, dtype=dtype)
            pinned.append(p)
            buf = (C.c_char * max(nbytes, 8)).from_address(p.value)
            a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
            a[...] = 0
            return a
        # The reference's result arrays are dense: ray_vec(nv, nstep_max+1, nray) = 58.8 GB + residual 8.4 GB for
        # this fan.  When that exceeds this rank's share of host memory the host traces the fan in equal batches
        # of rays, re-using the same arrays (as a host that writes each batch out before the next would).
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        budget = host_memory_budget(local_world)
        per_ray = npa * (nv + 1) * 8.0
        n_batches = max(1, int(np.ceil(nray * per_ray / budget)))
        nb_rays = (nray + n_batches - 1) // n_batches
        out = rb.ResultArrays(0, nv, npa, store=False)
        out.nray = nb_rays
        out.ray_vec = host_array((nb_rays, npa, nv))
        out.residual = host_array((nb_rays, npa))
        out.npoints = host_array((nb_rays,), np.int32)
        out.ray_stop_code = host_array((nb_rays,), np.int32)
        out._flags = C.create_string_buffer(nb_rays * _abi.FLAG_LEN)
        for nm in ("initial_ray_power", "ray_trace_time", "end_residuals", "max_residuals", "end_ray_parameter"):
            setattr(out, nm, host_array((nb_rays,)))
        out.start_ray_vec, out.end_ray_vec = host_array((nb_rays, nv)), host_array((nb_rays, nv))
        c = out.c
        c.nray, c.nv, c.npoints_alloc = nb_rays, nv, npa
        dp, ip = (lambda a: a.ctypes.data_as(_abi.c_double_p)), (lambda a: a.ctypes.data_as(_abi.c_int32_p))
        c.ray_vec, c.residual, c.npoints, c.ray_stop_code = dp(out.ray_vec), dp(out.residual), ip(out.npoints), ip(out.ray_stop_code)
        c.ray_stop_flag = C.cast(out._flags, C.c_char_p)
        c.initial_ray_power, c.ray_trace_time, c.end_residuals = dp(out.initial_ray_power), dp(out.ray_trace_time), dp(out.end_residuals)
        c.max_residuals, c.end_ray_parameter, c.start_ray_vec, c.end_ray_vec = dp(out.max_residuals), dp(out.end_ray_parameter), dp(out.start_ray_vec), dp(out.end_ray_vec)
        h_r, h_n, h_w = host_array((nray, 3)), host_array((nray, 3)), host_array((nray,))
        h_r[...], h_n[...], h_w[...] = rvec0, nvec0, wt
        fans = []
        for b in range(n_batches):
            lo, hi = b * nb_rays, min(nray, (b + 1) * nb_rays)
            fans.append(rb.make_fan(h_r[lo:hi], h_n[lo:hi], h_w[lo:hi]))

        def e2e_step():
            steps_, ms_, nl_, pts_ = 0, 0.0, 0, 0
            for fan, _keep in fans:
                assert L.rays_b200_trace(C.byref(cfg), C.byref(fan), C.byref(c)) == 0, L.rays_b200_last_error()
                steps_ += int(c.total_ray_steps)
                stt = rb.last_trace_stats()
                nl_ += stt["n_launches"]
                ms_ += stt["kernel_ms"]
                pts_ += int(np.sum(out.npoints[: int(fan.nray)], dtype=np.int64))
            return steps_, ms_, nl_, pts_
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_steps, e2e_dev_ms, npts_sum = 0, 0.0, 0
        for _ in range(args.steps):
            a_, b_, c_, d_ = e2e_step()
            e2e_steps += a_; e2e_dev_ms += b_; launches += c_; npts_sum = d_
        torch.cuda.synchronize(); barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_all = sum_over_ranks(float(e2e_steps))
        d2h = float(npts_sum * (nv + 1) * 8 + nray * (2 * 4 + 4 * 8 + 2 * nv * 8))
        e2e = {"value": e2e_all / dt, "unit": "ray-steps/s", "h2d_bytes_per_step": float(nray * 7 * 8), "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * dt / args.steps, "device_ms_per_step": e2e_dev_ms / args.steps,
               "note": "host fan in (pinned), trajectories + summaries out in the reference layout (pinned); finished rays are copied out by the trace kernel while others integrate; d2h counts saved points + summaries; device_ms = H2D + kernel + summary D2H by CUDA events"}
        # spot check: the end-to-end run saved exactly the points the device-resident run counted
        assert npts_sum - nray == ray_steps_per_fan, (npts_sum, nray, ray_steps_per_fan)
        e2e["host_batches"] = n_batches
        for p in pinned:
            L.rays_b200_host_free(p)

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import _oracle as orc
        stride = max(1, nray // args.cpu_rays)
        idx = np.arange(0, nray, stride)
        t0 = time.perf_counter()
        o, stt, _ = orc.trace(cfg, rvec0[idx], nvec0[idx], wt[idx], store=True, nthreads=host_threads())
        dt = time.perf_counter() - t0
        cpu = {"value": o.total_ray_steps / dt, "unit": "ray-steps/s", "cores": host_threads(), "kind": "port",
               "sample": f"{len(idx)} rays (every {stride}th of the fan), {o.total_ray_steps} ray-steps, {dt:.1f} s; C++ restatement of the Fortran path, OpenMP over rays"}

    if rank == 0:
        print(json.dumps({
            "metric": "ray_steps_per_sec", "value": value, "unit": "ray-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_gpu": nray, "ray_steps_per_fan": int(ray_steps_per_fan),
                       "ode": "SG" if cfg.ode_solver == 2 else "RK4", "ray_deriv": "numerical" if cfg.ray_deriv == 2 else "cold",
                       "ds": 5e-11, "nstep_max": 1000, "nv": nv, "sharding": "one fan per GPU, no collective during integration",
                       "l2": "256 MiB buffer written between timed iterations; trajectory output (>= 14 GB per fan) exceeds L2"},
            "wall_time_per_fan_s": dev_ms_max * 1e-3 / args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
