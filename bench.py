#!/usr/bin/env python
"""bench.py — ray-steps/s and wall time per fan for the hot path, on N GPUs of one node.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one trace of the whole launch fan.  Headline workload (BASELINE.json configs[3]):
1 048 576-candidate Solov'ev fan, ray_deriv_name='numerical', RK4, ds = 5e-11, nstep_max = 1000, per GPU (weak scaling:
every rank traces its own fan of that size; rank r shifts rindex_phi0 by r*1e-4).
  value    = ray-steps/s with the fan resident in HBM, trajectories written to HBM (device time, CUDA events)
  e2e      = the same through rays_b200_trace with HOST buffers: H2D of the fan, kernels, D2H of the trajectories +
             summaries into the reference's ray_results_m layout, inside the timed region (page-locked result arrays from
             rays_b200_host_alloc, which is what the Fortran binding's allocate_ray_results uses)
  parity   = rays of that very end-to-end run compared BITWISE with the CPU oracle's trace of the same rays (rank 0)
  e2e_summaries = the same call without trajectory arrays (ray_vec = residual = NULL): summaries only, the scalable mode
  e2e_pageable  = the same call with ordinary pageable arrays (an unmodified host's `allocate`), one step
  roofline = algorithmic fp64 flops of the trace kernel / its duration (frac, frac_nominal) and the executed-instruction
             view (frac_pipe, from the committed ncu capture of the same kernel), see DESIGN.md section 4
  also     = the other BASELINE configs on this GPU (N = 1): Solov'ev RK4 + deriv_cold, Solov'ev Shampine-Gordon, the 1M-ray
             mirror fan on the MPEX field; each with its own roofline fraction and an oracle parity check on a sample
  config5  = the sharded 8.39M-ray axisym deposition fan: rays iray % N, fused fixed-point binning, ONE reduce of the
             profile + ONE all-gather of the per-ray summaries inside the timed region (strong scaling)
  cpu_baseline = the C++ oracle (restatement of the Fortran; the Fortran itself cannot be built here) on a bounded sample
             of the same fan, all host cores
"""
import argparse
import hashlib
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
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12   # 148 SMs x 64 DFMA lanes x 2 flop x 1.965 GHz = 37.2
# exact algorithmic flops per ray-step of the headline workload from the oracle's counting scalar type
# (oracle_count_flops; +,-,*,/,sqrt and libm calls = 1 each, reference code as written): see DESIGN.md
FLOPS_PER_RAY_STEP_FALLBACK = 20789.0
BENCH_CFG_JSON = os.path.join(ROOT, "tests", "golden", "bench_cfg_solovev_fan_1M.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=0, help="debug: shrink the fan to about this many candidates")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the other BASELINE configs (N = 1 extras)")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--e2e-pinned-only", action="store_true", help="measurement aid: of the end-to-end modes run the page-locked one only")
    ap.add_argument("--config5-grid", type=int, default=2896, help="n_rindex_theta = n_rindex_phi of the deposition fan (2896^2 = 8.39M)")
    ap.add_argument("--cpu-rays", type=int, default=98304, help="rays of the bounded CPU-baseline sample")
    ap.add_argument("--ode", default="", help="measurement aid: override ode_solver_name (RK4_ODE | SG_ODE)")
    ap.add_argument("--deriv", default="", help="measurement aid: override ray_deriv_name (cold | numerical)")
    ap.add_argument("--workload", default=WORKLOAD, choices=[WORKLOAD, "mirror_fan_1M", "axisym_deposition_fan"],
                    help="measurement aid: run the device-resident arm on another workload (profiling)")
    return ap.parse_args()


def shrink_solovev(text: str, rays: int) -> str:
    if rays:
        # shrink the two outer (position) loops, keep the 32 x 32 direction grid
        npos = max(1, rays // 1024)
        nr = max(1, min(16, npos // 64))
        nt = max(1, min(64, npos // nr))
        text = text.replace("n_r_launch = 16", f"n_r_launch = {nr}").replace("n_theta_launch = 64", f"n_theta_launch = {nt}")
    return text


def workload_namelist(rank: int, rays: int) -> str:
    """The config-4 namelist, rank-shifted; written to a temp dir (the Brz/ray_init side files are not needed)."""
    import rays_b200 as rb
    text = open(rb.config_path("solovev_fan_1M.in")).read()
    text = text.replace("rindex_phi0 = 0.05", f"rindex_phi0 = {0.05 + 1e-4 * rank!r}")
    text = shrink_solovev(text, rays)
    d = tempfile.mkdtemp(prefix="rays_bench_")
    p = os.path.join(d, "rays.in")
    open(p, "w").write(text)
    return p


def deposition_namelist(grid: int) -> str:
    """axisym_toroid + solovev_magnetics, damp_fund_ECH, (n_theta, n_phi) launch grid"""
    import rays_b200 as rb
    text = open(rb.config_path("axisym_deposition_fan.in")).read()
    n = grid
    text = text.replace("n_rindex_theta = 1024", f"n_rindex_theta = {n}").replace("n_rindex_phi = 1024", f"n_rindex_phi = {n}")
    text = text.replace("delta_rindex_theta = 0.0003910068426197458", f"delta_rindex_theta = {0.4 / max(n - 1, 1)!r}")
    text = text.replace("delta_rindex_phi = 0.0003421309872922776", f"delta_rindex_phi = {0.35 / max(n - 1, 1)!r}")
    text = text.replace("nray_max = 8388608", f"nray_max = {max(n * n, 1)}")
    d = tempfile.mkdtemp(prefix="rays_dep_")
    p = os.path.join(d, "rays.in")
    open(p, "w").write(text)
    return p


def mirror_fan_inputs(rays: int):
    """Synthetic fan on the MPEX mirror field (SURVEY.md 8d config 3 scaled up): a grid over launch position (x, z) on the
    MPEX launcher's line y = -0.1 m and over two launch angles around the shipped 11-ray scans; the device launcher solves
    n(theta) for every candidate (one_ray_init_XYZ_k_direction).  32^4 = 1 048 576 candidates by default."""
    import numpy as np
    m = 32
    if rays:
        m = max(2, int(round(rays ** 0.25)))
    xs, zs = np.linspace(-0.03, 0.03, m), np.linspace(3.2, 3.4, m)
    ax, az = np.deg2rad(np.linspace(-5.0, 5.0, m)), np.deg2rad(np.linspace(25.0, 45.0, m))
    X, Z, AX, AZ = np.meshgrid(xs, zs, ax, az, indexing="ij")
    r = np.stack([X.ravel(), np.full(X.size, -0.1), Z.ravel()], 1)
    d = np.stack([np.sin(AX.ravel()), np.cos(AX.ravel()) * np.cos(AZ.ravel()), -np.cos(AX.ravel()) * np.sin(AZ.ravel())], 1)
    return np.ascontiguousarray(r), np.ascontiguousarray(d)


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


def headline_config(nray, ray_steps_per_fan, ode, deriv, nv):
    """the `config` object both arms print (same keys, so that the driver's same_config check compares like with like)"""
    return {"workload": WORKLOAD, "rays_per_gpu": int(nray), "ray_steps_per_fan": int(ray_steps_per_fan), "ode": ode, "ray_deriv": deriv,
            "ds": 5e-11, "nstep_max": 1000, "nv": int(nv), "sharding": "one fan per GPU, no collective during integration",
            "l2": "256 MiB buffer written between timed iterations; trajectory output (>= 14 GB per fan) exceeds L2"}


# ============================ reference arm =====================================================================
def run_reference(args, rank, world):
    """The reference arm: the CPU implementation of the path (C++ oracle; kind 'port') on all host cores, each step a bounded
    sample of the workload's fan.  Rank 0 only.  It does NOT load rays_b200/lib/librays_b200.so: the marshalled module state
    comes from tests/golden/bench_cfg_solovev_fan_1M.json (written by tests/golden/make_bench_cfg.py from the same namelist)."""
    if rank != 0:
        return
    import ctypes as C
    import numpy as np
    import _oracle as orc
    from rays_b200 import _abi
    orc.build()
    snap = json.load(open(BENCH_CFG_JSON))
    assert snap["sizeof_rays_cfg"] == C.sizeof(_abi.Cfg), "bench cfg snapshot is stale: run tests/golden/make_bench_cfg.py"
    cfg = _abi.Cfg.from_buffer_copy(bytes.fromhex(snap["rays_cfg_hex"]))
    launch = _abi.SolovevLaunch.from_buffer_copy(bytes.fromhex(snap["solovev_launch_hex"]))
    if args.rays:
        npos = max(1, args.rays // 1024)
        launch.n_r_launch = max(1, min(16, npos // 64))
        launch.n_theta_launch = max(1, min(64, npos // launch.n_r_launch))
    r, n, w = orc.launch_fan(cfg, "solovev", launch, 1 << 21)
    # ~5 s of CPU work per step (so that --steps 20 stays within minutes), plus ONE trace of the whole fan that validates
    # the sampling (full_fan below)
    stride = max(1, r.shape[0] // max(1, args.cpu_rays // 2))
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
    del o
    full = None
    if not args.no_cpu:
        t1 = time.perf_counter()
        of, _, _ = orc.trace(cfg, r, n, w, store=False, nthreads=cores)
        d1 = time.perf_counter() - t1
        full = {"rays": int(r.shape[0]), "ray_steps": int(of.total_ray_steps), "seconds": d1, "value": of.total_ray_steps / d1,
                "note": "one untimed-loop trace of the WHOLE fan (summaries only): validates the sampled per-step value"}
    sample = f"{len(idx)} rays (every {stride}th of the {r.shape[0]}-ray fan), {steps // args.steps} ray-steps per step"
    print(json.dumps({
        "impl": "reference", "metric": "ray_steps_per_sec", "value": val, "unit": "ray-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": headline_config(r.shape[0], (of.total_ray_steps if full else 0), "RK4", "numerical", cfg.nv),
        "cpu_baseline": {"value": val, "unit": "ray-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "full_fan": full,
        "note": "C++ restatement of the reference's Fortran path (oracle/); the Fortran itself cannot be compiled in this image"}))


# ============================ helpers of our arm ================================================================
def compare_with_oracle(np, g_vec, g_res, g_np, g_code, o, rk4: bool, tol: float = 0.0, rel: bool = False):
    """GPU rows vs oracle ResultArrays of the same rays.  rk4: bitwise on every saved point; else within tol: absolute for SG (the run's
    own ODE tolerance), relative (rel=True) for RK4 configurations that call libm (tanh/cosh/pow profiles: CUDA's libm vs glibc's)."""
    n = len(g_np)
    np_equal = bool(np.array_equal(g_np, o.npoints))
    stop_equal = bool(np.array_equal(g_code, o.ray_stop_code))
    bitwise, max_rel, max_abs, pts = True, 0.0, 0.0, 0
    for i in range(n):
        m = int(min(g_np[i], o.npoints[i]))
        a, b = g_vec[i, :m], o.ray_vec[i, :m]
        pts += m
        if not np.array_equal(a, b, equal_nan=True):
            bitwise = False
            fin = np.isfinite(b) & np.isfinite(a)
            if fin.any():
                d = np.abs(a - b)[fin]
                max_abs = max(max_abs, float(d.max()))
                # position / wave-vector triples against their norms (north_star: final (x, k, power) to 1e-10 relative)
                for sl in (slice(0, 3), slice(3, 6)):
                    nb = np.linalg.norm(b[:, sl], axis=1)
                    da = np.linalg.norm(a[:, sl] - b[:, sl], axis=1)
                    ok = np.isfinite(nb) & np.isfinite(da) & (nb > 0)
                    if ok.any():
                        max_rel = max(max_rel, float((da[ok] / nb[ok]).max()))
                # every other slot (s, integrated gradients, power) against the largest magnitude that slot reaches on this ray
                for l in range(6, b.shape[1]):
                    fl = fin[:, l]
                    sc = float(np.max(np.abs(b[fl, l]))) if fl.any() else 0.0
                    if sc > 0:
                        max_rel = max(max_rel, float(np.max(np.abs(a[fl, l] - b[fl, l]))) / sc)
        if rk4 and g_res is not None and o.residual is not None:
            if not np.array_equal(g_res[i, :m], o.residual[i, :m], equal_nan=True):
                bitwise = False
    out = {"rays": int(n), "points": int(pts), "npoints_equal": np_equal, "stop_reasons_equal": stop_equal, "bitwise": bool(bitwise),
           "max_rel": max_rel, "max_abs": max_abs, "against": "oracle/ (C++ restatement of the Fortran path) on the same rays"}
    out["ok"] = bool(np_equal and stop_equal and (bitwise if rk4 else (bitwise or (max_rel <= tol if rel else max_abs <= tol))))
    if not rk4:
        out["tolerance"] = tol
        out["tolerance_kind"] = "relative (max_rel: |dx|/|x|, |dk|/|k|, other slots against their largest magnitude on the ray)" if rel else "absolute (max_abs)"
    return out


def flops_per_ray_step(np, cfg, rvec0, nvec0, fallback):
    try:
        import _oracle as orc
        orc.load()
        nray = rvec0.shape[0]
        idx = np.arange(0, nray, max(1, nray // 64))
        fl, stp, _ = orc.count_flops(cfg, rvec0[idx], nvec0[idx], 0, len(idx))
        if stp > 0:
            return fl / stp
    except Exception:
        pass
    return fallback


def ncu_profile_numbers(kernel_tag: str):
    """executed-instruction view of a kernel from its committed `ncu --set full` summary (profiles/r2_<tag>_ncu_summary.txt):
    FP64 pipe utilisation and DRAM bytes.  These are PROFILE numbers (one capture of the same kernel on the same workload),
    not measured in this run; None when no summary is committed."""
    import re
    p = os.path.join(ROOT, "profiles", f"r2_{kernel_tag}_ncu_summary.txt")
    if not os.path.exists(p):
        return None
    txt = open(p).read()

    def grab(name):
        m = re.search(re.escape(name) + r"[^=\n]*=\s*([0-9.eE+-]+)", txt)
        return float(m.group(1)) if m else None
    return {"file": os.path.relpath(p, ROOT), "fp64_pipe_pct": grab("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "fp64_inst_pct": grab("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "dram_read_gb": grab("dram__bytes_read.sum [Gbyte]"), "dram_write_gb": grab("dram__bytes_write.sum [Gbyte]"),
            "duration_ms": grab("gpu__time_duration.sum [ms]")}


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

    L = _abi.load()
    rb.init(local_rank)
    stream = torch.cuda.ExternalStream(L.rays_b200_stream(), device=torch.device("cuda", local_rank))
    peak_tf, _ = rb.fp64_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_device(steps, warmup, store=True, bins=None):
        """device-resident arm on the fan that is in HBM: CUDA events on the library's stream"""
        st = None
        for _ in range(warmup):
            st = rb.trace_device(store=store, bins=bins)
        barrier(); torch.cuda.synchronize()
        dev_ms = kernel_ms = resume_ms = 0.0
        steps_total = launches = 0
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0.record(stream)
            st = rb.trace_device(store=store, bins=bins)
            e1.record(stream)
            torch.cuda.synchronize()
            dev_ms += e0.elapsed_time(e1)
            kernel_ms += st["trace_kernel_ms"]      # the trace kernel proper (the roofline's kernel)
            resume_ms += st["resume_pass_ms"]        # of which: the resume pass over the suspended (long) rays
            steps_total += st["ray_steps"]
            launches += st["n_launches"]
        barrier(); torch.cuda.synchronize()
        return dict(dev_ms=dev_ms, kernel_ms=kernel_ms, resume_ms=resume_ms, ray_steps=steps_total, launches=launches, info=st)

    def roofline_of(cfg, rvec0, nvec0, t, steps, fallback_flops, tag):
        nv = int(cfg.nv)
        fps = flops_per_ray_step(np, cfg, rvec0, nvec0, fallback_flops)
        avg_s = t["kernel_ms"] * 1e-3 / max(steps, 1)
        per_launch = t["ray_steps"] / max(steps, 1)
        ach = fps * per_launch / avg_s / 1e12 if avg_s > 0 else 0.0
        prof = ncu_profile_numbers(tag)
        r = {"bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
             "frac_algorithmic": ach / peak_tf if peak_tf else None, "frac_nominal": ach / FP64_NOMINAL_TFLOPS,
             "frac_pipe": (prof["fp64_pipe_pct"] / 100.0) if prof and prof.get("fp64_pipe_pct") else None,
             "traffic": ((prof["dram_read_gb"] or 0) + (prof["dram_write_gb"] or 0)) * 1e9 if prof and prof.get("dram_write_gb") is not None else None,
             "traffic_from_profile": ((prof["dram_read_gb"] or 0) + (prof["dram_write_gb"] or 0)) * 1e9 if prof and prof.get("dram_write_gb") is not None else None,
             "profile": prof["file"] if prof else None,
             "peak_source": "DFMA microbenchmark on this GPU in this run (rays_b200_fp64_peak); MEASURED_PEAKS.json has no fp64 figure; "
                            f"frac_nominal is against 148 SMs x 64 lanes x 2 x 1.965 GHz = {FP64_NOMINAL_TFLOPS:.1f} TFLOP/s",
             "flops_per_ray_step": fps, "kernel": t["info"]["kernel"], "grid": t["info"]["grid"], "ctas_per_sm": t["info"]["blocks_per_sm"],
             "hbm_writeback_gbs": (nv + 1) * 8.0 * per_launch / avg_s / 1e9 if avg_s > 0 else 0.0, "avg_kernel_ms": avg_s * 1e3,
             "resume_pass_ms": t["resume_ms"] / max(steps, 1), "launches_per_fan": t["info"]["n_passes"],
             "note": "achieved = algorithmic flops of one fan (the reference's arithmetic as written, counted by the oracle) / summed CUDA-event "
                     "duration of the trace kernel's launches for that fan; frac_pipe = FP64 pipe-active share from the committed ncu capture "
                     "(traffic_from_profile likewise: DRAM read + write bytes of that capture, not measured in this run)"}
        return r

    # ---- measurement aid: another workload through the device-resident arm only (profiling) -------------
    if args.workload != WORKLOAD:
        if args.workload == "mirror_fan_1M":
            rb.initialize(rb.config_path("mpex/rays.in"), ray_init=False, device=local_rank)
            if args.ode or args.deriv:
                rb.set_ode(ode_solver_name=args.ode, ray_deriv_name=args.deriv, rel_err0=1e-6, abs_err0=1e-6, SG_error_limit=0.1)
            cfg = rb.host_cfg()
            rb.set_config(cfg)
            pos, dirs = mirror_fan_inputs(args.rays)
            nray = rb.launch_fan_directions(pos, dirs)
            bins = None
        else:
            rb.initialize(deposition_namelist(args.config5_grid if not args.rays else max(2, int(args.rays ** 0.5))), ray_init=True, device=local_rank)
            cfg = rb.host_cfg()
            rb.set_config(cfg)
            rvec0, nvec0, wt = rb.get_fan()
            nray = rb.fan_upload(rvec0, nvec0, wt)
            bins = (501, 0.0, 1.0)
        rvec0, nvec0, wt = rb.fan_download(nray)
        t = time_device(args.steps, args.warmup, store=bins is None, bins=bins)
        rl = roofline_of(cfg, rvec0, nvec0, t, args.steps, 5000.0, args.workload)
        print(json.dumps({"metric": "ray_steps_per_sec", "value": t["ray_steps"] / (t["dev_ms"] * 1e-3), "unit": "ray-steps/s", "n_gpus": 1,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": t["dev_ms"] / args.steps, "higher_is_better": True,
                          "config": {"workload": args.workload, "rays_per_gpu": int(nray), "nv": int(cfg.nv)}, "roofline": rl}))
        return

    # ---- initialize(read_input): namelist -> module state -> launch fan built on the device ------------
    path = workload_namelist(rank, args.rays)
    rb.initialize(path, ray_init=True, device=local_rank)
    if args.ode or args.deriv:
        rb.set_ode(ode_solver_name=args.ode, ray_deriv_name=args.deriv, rel_err0=1e-6, abs_err0=1e-6, SG_error_limit=0.1)
    cfg = _abi.Cfg.from_buffer_copy(bytes(rb.host_cfg()))   # a snapshot: the extras below re-initialise the host mirror (no tables in this config)
    rvec0, nvec0, wt = rb.get_fan()
    nray = rvec0.shape[0]
    nv, npa = int(cfg.nv), int(cfg.nstep_max) + 1

    # ---- device-resident arm: fan in HBM, trajectories to HBM -------------------------------------------
    rb.set_config(cfg)
    rb.fan_upload(rvec0, nvec0, wt)
    sampler = ClockSampler(local_rank)
    # warm-up is untimed; the sampler covers the timed region (and a little of the warm-up, same load)
    for _ in range(args.warmup):
        rb.trace_device(store=True)
    sampler.start()
    t_main = time_device(args.steps, 0)
    clocks = sampler.finish()
    launches = t_main["launches"]
    dev_ms_max = max_over_ranks(t_main["dev_ms"])
    steps_all = sum_over_ranks(float(t_main["ray_steps"]))
    value = steps_all / (dev_ms_max * 1e-3)
    ray_steps_per_fan = t_main["ray_steps"] // max(args.steps, 1)
    is_headline = not (args.rays or args.ode or args.deriv)
    tag = "trace_rk4_num" if is_headline else "none"
    roofline = roofline_of(cfg, rvec0, nvec0, t_main, args.steps, FLOPS_PER_RAY_STEP_FALLBACK, tag)

    # ---- CPU oracle on a bounded sample (rank 0): the cpu_baseline AND the reference of the parity block --
    oracle_sample = None
    cpu = None
    if rank == 0 and not args.no_cpu:
        import _oracle as orc
        stride = max(1, nray // args.cpu_rays)
        idx = np.arange(0, nray, stride)
        t0 = time.perf_counter()
        o, stt, _ = orc.trace(cfg, rvec0[idx], nvec0[idx], wt[idx], store=True, nthreads=host_threads())
        dt = time.perf_counter() - t0
        oracle_sample = (idx, o)
        if world == 1:
            cpu = {"value": o.total_ray_steps / dt, "unit": "ray-steps/s", "cores": host_threads(), "kind": "port",
                   "sample": f"{len(idx)} rays (every {stride}th of the fan), {o.total_ray_steps} ray-steps, {dt:.1f} s; C++ restatement of the Fortran path, OpenMP over rays"}

    # ---- end-to-end arms: host buffers through rays_b200_trace ---------------------------------------------
    e2e = e2e_summaries = e2e_pageable = parity = None
    if not args.no_e2e:
        pinned = []

        def host_array(shape, dtype=np.float64, pin=True):
            """page-locked array from the library's allocator (allocate_ray_results); pageable if asked or if that fails"""
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            p = C.c_void_p()
            if not pin or L.rays_b200_host_alloc(C.byref(p), max(nbytes, 8)) != 0:
                return np.zeros(shape, dtype=dtype)
            pinned.append(p)
            buf = (C.c_char * max(nbytes, 8)).from_address(p.value)
            a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
            a[...] = 0
            return a

        def free_pinned():
            for p in pinned:
                L.rays_b200_host_free(p)
            pinned.clear()
        # The reference's result arrays are dense: ray_vec(nv, nstep_max+1, nray) = 58.8 GB + residual 8.4 GB for
        # this fan.  When that exceeds this rank's share of host memory the host traces the fan in equal batches
        # of rays, re-using the same arrays (as a host that writes each batch out before the next would).
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        budget = host_memory_budget(local_world)
        per_ray = npa * (nv + 1) * 8.0
        h_r, h_n, h_w = host_array((nray, 3)), host_array((nray, 3)), host_array((nray,))
        h_r[...], h_n[...], h_w[...] = rvec0, nvec0, wt

        def make_out(nb_rays, store, pin):
            out = rb.ResultArrays(0, nv, npa, store=False)
            out.nray = nb_rays
            out.ray_vec = host_array((nb_rays, npa, nv), pin=pin) if store else None
            out.residual = host_array((nb_rays, npa), pin=pin) if store else None
            out.npoints = host_array((nb_rays,), np.int32, pin=pin)
            out.ray_stop_code = host_array((nb_rays,), np.int32, pin=pin)
            out._flags = C.create_string_buffer(nb_rays * _abi.FLAG_LEN)
            for nm in ("initial_ray_power", "ray_trace_time", "end_residuals", "max_residuals", "end_ray_parameter"):
                setattr(out, nm, host_array((nb_rays,), pin=pin))
            out.start_ray_vec, out.end_ray_vec = host_array((nb_rays, nv), pin=pin), host_array((nb_rays, nv), pin=pin)
            c = out.c
            c.nray, c.nv, c.npoints_alloc = nb_rays, nv, npa
            dp, ip = (lambda a: None if a is None else a.ctypes.data_as(_abi.c_double_p)), (lambda a: a.ctypes.data_as(_abi.c_int32_p))
            c.ray_vec, c.residual, c.npoints, c.ray_stop_code = dp(out.ray_vec), dp(out.residual), ip(out.npoints), ip(out.ray_stop_code)
            c.ray_stop_flag = C.cast(out._flags, C.c_char_p)
            c.initial_ray_power, c.ray_trace_time, c.end_residuals = dp(out.initial_ray_power), dp(out.ray_trace_time), dp(out.end_residuals)
            c.max_residuals, c.end_ray_parameter, c.start_ray_vec, c.end_ray_vec = dp(out.max_residuals), dp(out.end_ray_parameter), dp(out.start_ray_vec), dp(out.end_ray_vec)
            return out

        def run_e2e(store, pin, steps, warmup, check=None):
            """steps traces of the whole fan through rays_b200_trace; `check(batch_lo, batch_hi, out)` is called after every
            batch of ONE extra, untimed pass"""
            n_batches = max(1, int(np.ceil(nray * per_ray / budget))) if store else 1
            nb_rays = (nray + n_batches - 1) // n_batches
            out = make_out(nb_rays, store, pin)
            c = out.c
            fans = []
            for b in range(n_batches):
                lo, hi = b * nb_rays, min(nray, (b + 1) * nb_rays)
                fans.append((lo, hi) + rb.make_fan(h_r[lo:hi], h_n[lo:hi], h_w[lo:hi]))

            def one(check_fn=None):
                steps_, ms_, nl_, pts_ = 0, 0.0, 0, 0
                for lo, hi, fan, _keep in fans:
                    assert L.rays_b200_trace(C.byref(cfg), C.byref(fan), C.byref(c)) == 0, L.rays_b200_last_error()
                    steps_ += int(c.total_ray_steps)
                    stt = rb.last_trace_stats()
                    nl_ += stt["n_launches"]
                    ms_ += stt["kernel_ms"]
                    pts_ += int(np.sum(out.npoints[: int(fan.nray)], dtype=np.int64))
                    if check_fn:
                        check_fn(lo, hi, out)
                return steps_, ms_, nl_, pts_
            for _ in range(warmup):
                one()
            barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            tot_steps, dev_ms, nl, npts = 0, 0.0, 0, 0
            for _ in range(steps):
                a_, b_, c_, d_ = one()
                tot_steps += a_; dev_ms += b_; nl += c_; npts = d_
            torch.cuda.synchronize(); barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            all_steps = sum_over_ranks(float(tot_steps))
            if check is not None:
                one(check)
            d2h = float((npts * (nv + 1) * 8 if store else 0) + nray * (2 * 4 + 4 * 8 + 2 * nv * 8))
            rec = {"value": all_steps / dt, "unit": "ray-steps/s", "h2d_bytes_per_step": float(nray * 7 * 8), "d2h_bytes_per_step": d2h,
                   "ms_per_step": 1e3 * dt / steps, "device_ms_per_step": dev_ms / steps, "host_batches": n_batches}
            return rec, nl, npts

        # (1) the headline end-to-end number: page-locked arrays, trajectories + summaries out
        cmp_parts = []

        def check_batch(lo, hi, out):
            if oracle_sample is None:
                return
            idx, o = oracle_sample
            sel = np.nonzero((idx >= lo) & (idx < hi))[0]
            if len(sel) == 0:
                return
            loc = idx[sel] - lo
            sub = rb.ResultArrays(0, nv, npa, store=False)   # view of the oracle rows of this batch
            sub.ray_vec, sub.residual, sub.npoints, sub.ray_stop_code = o.ray_vec[sel], o.residual[sel], o.npoints[sel], o.ray_stop_code[sel]
            cmp_parts.append(compare_with_oracle(np, out.ray_vec[loc], out.residual[loc], out.npoints[loc], out.ray_stop_code[loc], sub,
                                                 rk4=int(cfg.ode_solver) == 1, tol=float(cfg.abs_err0) * 10))
        e2e, nl, npts_sum = run_e2e(True, True, args.steps, max(1, min(args.warmup, 2)), check=check_batch)
        launches += nl
        e2e["copy_out"] = "copy_out_kernel (concurrent copier)" if "copy_out_kernel" in rb.last_trace_stats()["kernel"] else "in-kernel (flush_finished_rays)"
        e2e["note"] = ("host fan in (pinned), trajectories + summaries out in the reference layout (pinned); the trace kernels write the trajectories to "
                       "HBM and append every ended ray to a list that a concurrent copier kernel (second stream, 16 CTAs) drains into the caller's arrays "
                       "while the other rays integrate (RAYS_B200_COPIER=0, or a fan whose trajectories do not fit in HBM: the trace kernel's warps copy "
                       "their finished rays themselves); d2h counts saved points + summaries; device_ms = H2D + kernels + summary D2H by CUDA events")
        # spot check: the end-to-end run saved exactly the points the device-resident run counted
        assert npts_sum - nray == ray_steps_per_fan, (npts_sum, nray, ray_steps_per_fan)
        if cmp_parts:
            parity = {"rays": sum(p["rays"] for p in cmp_parts), "points": sum(p["points"] for p in cmp_parts),
                      "npoints_equal": all(p["npoints_equal"] for p in cmp_parts), "stop_reasons_equal": all(p["stop_reasons_equal"] for p in cmp_parts),
                      "bitwise": all(p["bitwise"] for p in cmp_parts), "max_rel": max(p["max_rel"] for p in cmp_parts),
                      "max_abs": max(p["max_abs"] for p in cmp_parts), "ok": all(p["ok"] for p in cmp_parts),
                      "what": "trajectories, residuals, npoints and stop codes that the timed end-to-end path (rays_b200_trace, host arrays) "
                              "delivered for every sampled ray vs oracle/ (C++ restatement of the Fortran path; the Fortran was never executed)"}
        free_pinned()
        h_r, h_n, h_w = rvec0, nvec0, wt
        # (2) the scalable mode: no trajectory arrays, summaries only
        if not args.e2e_pinned_only:
            e2e_summaries, nl, _ = run_e2e(False, False, args.steps, 1)
            launches += nl
            e2e_summaries["note"] = "rays_b200_trace with ray_vec = residual = NULL: H2D of the fan, trace without trajectory storage, D2H of the per-ray summaries"
        # (3) an unmodified host: pageable result arrays (batched cudaMemcpy2DAsync copy-out); one warm-up, one step; N = 1 only
        if world == 1 and is_headline and not args.e2e_pinned_only:
            try:
                e2e_pageable, nl, _ = run_e2e(True, False, 1, 1)
                launches += nl
                e2e_pageable["note"] = ("pageable numpy arrays (what `allocate(ray_vec(...))` of an unmodified Fortran host gives): arrays above 1 GiB are not "
                                        "registered, trajectories go through HBM-sized batches + trimmed cudaMemcpy2DAsync; 1 warm-up + 1 timed step")
            except MemoryError:
                e2e_pageable = {"unavailable": "not enough host memory for pageable result arrays"}

    # ---- what the host can take: every rank copies 1 GiB device -> pinned host with the copy engine at the same time.  The
    # aggregate is the ceiling of any end-to-end number that delivers full trajectories to host memory (PCIe + host DRAM).
    if e2e is not None:
        nb = 1 << 30
        dsrc = torch.empty(nb, dtype=torch.uint8, device="cuda")
        hdst = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
        hdst.copy_(dsrc); torch.cuda.synchronize(); barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(4):
            hdst.copy_(dsrc, non_blocking=True)
        c1.record(); torch.cuda.synchronize()
        my_gbs = 4 * nb / (c0.elapsed_time(c1) * 1e-3) / 1e9
        barrier()
        agg = sum_over_ranks(my_gbs)
        e2e["d2h_dma_gbs_this_rank"] = my_gbs
        e2e["d2h_dma_gbs_all_ranks"] = agg
        e2e["d2h_gbs_achieved_all_ranks"] = sum_over_ranks(e2e["d2h_bytes_per_step"]) / (e2e["ms_per_step"] * 1e-3) / 1e9
        e2e["d2h_ceiling_note"] = ("all ranks copy 1 GiB device -> pinned host concurrently with cudaMemcpyAsync (copy engine): the host-side ceiling "
                                   "(PCIe links + host DRAM ingest) for delivering full trajectories; d2h_gbs_achieved is what the trace delivered")
        del dsrc, hdst

    # ---- the other BASELINE configs on this GPU (N = 1) ----------------------------------------------------
    also = []
    if world == 1 and is_headline and not args.no_also:
        import _oracle as orc

        def sample_parity(cfg_x, r_, n_, w_, rk4, nsample=4096, tol=0.0, rel=False):
            idx = np.arange(0, r_.shape[0], max(1, r_.shape[0] // nsample))
            g = rb.trace(cfg_x, r_[idx], n_[idx], w_[idx])
            o, _, _ = orc.trace(cfg_x, r_[idx], n_[idx], w_[idx], store=True, nthreads=host_threads())
            return compare_with_oracle(np, g.ray_vec, g.residual, g.npoints, g.ray_stop_code, o, rk4=rk4, tol=tol, rel=rel)

        def extra(name, desc, cfg_x, r_, n_, w_, rk4, tag, tol=0.0, fallback=4000.0, rel=False):
            nonlocal launches
            rb.set_config(cfg_x)
            rb.fan_upload(r_, n_, w_)
            t = time_device(2, 1)
            launches += t["launches"]
            rl = roofline_of(cfg_x, r_, n_, t, 2, fallback, tag)
            rec = {"workload": name, "what": desc, "rays": int(r_.shape[0]), "ray_steps_per_fan": int(t["ray_steps"] // 2),
                   "value": t["ray_steps"] / (t["dev_ms"] * 1e-3), "unit": "ray-steps/s", "ms_per_fan": t["dev_ms"] / 2, "steps": 2, "warmup": 1,
                   "roofline": {k: rl[k] for k in ("achieved", "peak", "frac", "frac_nominal", "frac_pipe", "flops_per_ray_step", "kernel", "ctas_per_sm",
                                                   "avg_kernel_ms", "launches_per_fan", "profile", "traffic_from_profile")},
                   "parity": sample_parity(cfg_x, r_, n_, w_, rk4, tol=tol, rel=rel)}
            also.append(rec)
        # Solov'ev fan, RK4 + deriv_cold (config 2's equilibrium with the analytic derivatives)
        rb.set_ode(ode_solver_name="RK4_ODE", ray_deriv_name="cold")
        extra("solovev_fan_1M/RK4/cold", "the headline fan with ray_deriv_name='cold'", rb.host_cfg(), rvec0, nvec0, wt, True, "trace_rk4_cold")
        # Solov'ev fan, Shampine-Gordon, tol 1e-6 (config 2 at scale)
        rb.set_ode(ode_solver_name="SG_ODE", ray_deriv_name="cold", rel_err0=1e-6, abs_err0=1e-6, SG_error_limit=0.1)
        extra("solovev_fan_1M/SG/cold", "the headline fan with ode_solver_name='SG_ODE', rel_err0 = abs_err0 = 1e-6", rb.host_cfg(), rvec0, nvec0, wt, False,
              "trace_sg2", tol=1e-5, fallback=19000.0)
        # mirror fan on the MPEX field (config 3 at scale)
        rb.initialize(rb.config_path("mpex/rays.in"), ray_init=False, device=local_rank)
        cfg_m = rb.host_cfg()
        rb.set_config(cfg_m)
        pos, dirs = mirror_fan_inputs(0)
        n_m = rb.launch_fan_directions(pos, dirs)
        r_m, n_mm, w_m = rb.fan_download(n_m)
        extra("mirror_fan_1M/RK4/cold", "32^4 grid over launch position (x, z) and two launch angles on the MPEX mirror field (bicubic Br, Bz, Aphi tables), "
              "RK4, ds = 1e-12, nstep_max = 500, nv = 12", cfg_m, r_m, n_mm, w_m, False, "trace_rk4_mirror", tol=1e-10, fallback=5400.0, rel=True)

    # ---- config 5: sharded deposition fan with the north-star collective inside the timed region ------------
    config5 = None
    if is_headline and not args.no_config5:
        from rays_b200.sharding import reduce_bins, gather_packed_summaries
        rb.initialize(deposition_namelist(args.config5_grid), ray_init=True, device=local_rank)   # every rank launches the whole fan on its GPU
        cfg5 = rb.host_cfg()
        rb.set_config(cfg5)
        r5, n5, w5 = rb.get_fan()
        n_total = r5.shape[0]
        nbins = 501
        rb.fan_upload(r5, n5, w5)
        rb.fan_shard(rank, world)                      # rays iray % world == rank, order preserved
        n_local = len(range(rank, n_total, world))
        n_max = (n_total + world - 1) // world
        rows_d = 6 + 2 * int(cfg5.nv)
        acc = torch.zeros(nbins, dtype=torch.int64, device="cuda")
        summ = torch.zeros((n_max, rows_d), dtype=torch.float64, device="cuda")
        gathered = torch.empty((world * n_max, rows_d), dtype=torch.float64, device="cuda") if world > 1 else summ

        def step5():
            st = rb.trace_device(store=False, bins=(nbins, 0.0, 1.0))
            rb.deposition_fixed(nbins, 0.0, 1.0, d_acc_out=acc.data_ptr())   # this GPU's fixed-point bins
            rb.summaries_pack(summ.data_ptr(), n_max)
            if dist is not None:
                reduce_bins(acc, dst=0)                                     # THE collective of the run: 501 int64 (exact sum)
                gather_packed_summaries(summ, gathered)                     # per-ray summaries, 6 + 2 nv doubles per ray
            return st
        for _ in range(2):
            acc.zero_()
            st5 = step5()
        barrier(); torch.cuda.synchronize()
        k5 = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        steps5 = 0
        for _ in range(k5):
            st5 = step5()
            steps5 += st5["ray_steps"]
        torch.cuda.synchronize(); barrier()
        dt5 = max_over_ranks(time.perf_counter() - t0)
        steps5_all = sum_over_ranks(float(steps5))
        launches += (k5 + 2) * (st5["n_launches"] + 1)
        _, unit, _, _ = rb.deposition_fixed(nbins, 0.0, 1.0)
        acc_h = acc.cpu().numpy()
        # a fresh reduce for the checksum (the timed loop reduced into rank 0's buffer k5 times)
        rb.deposition_fixed(nbins, 0.0, 1.0, d_acc_out=acc.data_ptr())
        reduce_bins(acc, dst=0)
        torch.cuda.synchronize()
        acc_h = acc.cpu().numpy()
        prof = acc_h.astype(np.float64) * unit
        if rank == 0:
            config5 = {"workload": "axisym_deposition_fan", "what": "SURVEY.md 8d config 5: axisym_toroid + solovev_magnetics, damp_fund_ECH, nv = 8, "
                       f"{args.config5_grid} x {args.config5_grid} launch grid, rays sharded iray % N, deposition binned while tracing (no trajectory storage), "
                       "one reduce of the 501-bin fixed-point profile + one all-gather of the per-ray summaries inside the timed region",
                       "rays": int(n_total), "ray_steps_per_fan": int(steps5_all // k5), "n_gpus": world, "scaling": "strong", "steps": k5, "warmup": 2,
                       "ms_per_fan": 1e3 * dt5 / k5, "value": steps5_all / dt5, "unit": "ray-steps/s", "kernel": st5["kernel"],
                       "kernel_ms_rank0": st5["kernel_ms"], "collective": ("NCCL reduce (int64 x 501) + all_gather_into_tensor (%d doubles per ray)" % rows_d) if world > 1 else "none (1 GPU)",
                       "gathered_bytes": int(world * n_max * rows_d * 8) if world > 1 else 0,
                       "Q_sum": float(prof.sum()), "profile_peak_bin": int(np.argmax(prof)), "profile_peak": float(prof.max()),
                       "profile_sha256": hashlib.sha256(acc_h.tobytes()).hexdigest(),
                       "note": "profile_sha256 hashes the reduced int64 bins: it is the same string at N = 1, 2, 4, 8 (integer sums do not depend on the sharding)"}
            if world > 1:   # the whole fan once more on rank 0 alone: the sharded profile must equal it bit for bit
                rb.fan_upload(r5, n5, w5)
                rb.trace_device(store=False, bins=(nbins, 0.0, 1.0))
                acc1, _, _, _ = rb.deposition_fixed(nbins, 0.0, 1.0)
                config5["equals_single_gpu_profile_bitwise"] = bool(np.array_equal(acc1, acc_h))
                config5["max_abs_dev_from_single_gpu"] = float(np.max(np.abs(acc1.astype(np.float64) - acc_h.astype(np.float64))) * unit)
        barrier()

    if rank == 0:
        print(json.dumps({
            "metric": "ray_steps_per_sec", "value": value, "unit": "ray-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": headline_config(nray, ray_steps_per_fan, "SG" if cfg.ode_solver == 2 else "RK4", "numerical" if cfg.ray_deriv == 2 else "cold", nv),
            "wall_time_per_fan_s": dev_ms_max * 1e-3 / args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "e2e_summaries": e2e_summaries, "e2e_pageable": e2e_pageable,
            "also": also or None, "config5": config5, "gpu_launches": int(launches), "clocks": clocks}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
