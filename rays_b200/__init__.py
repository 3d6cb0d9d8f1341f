"""rays_b200 — B200-native ray-integration engine for the hot path of ORNL-Fusion/RAYS.

Host-side mirror of the reference's program flow (RAYS_project/RAYS_code/RAYS.f90:9-15):

    import rays_b200 as rb
    rb.initialize("rays.in")        # initialize(read_input)   RAYS_lib/intialize.f90
    rb.trace_rays()                 # trace_rays               RAYS_lib/ray_tracing.f90   <- sm_100a kernels
    rb.finalize_run("outdir")       # finalize_run             RAYS_lib/finalize_run.f90  -> run_results.<label>.nc

Everything numerical runs in `rays_b200/lib/librays_b200.so` (C ABI in include/rays_b200.h); this
module is only ctypes plumbing plus numpy views of the result arrays.  There is no CPU fallback: the
calls fail if the library is missing or no B200 is present.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._abi import Cfg, Deposition, Fan, Results  # noqa: F401


class RaysError(RuntimeError):
    pass


def _lib():
    return _abi.load()


def _ck(rc: int, host: bool = False) -> None:
    if rc != 0:
        L = _lib()
        msg = (L.rays_host_last_error() if host else L.rays_b200_last_error()) or b""
        if not msg:
            msg = L.rays_b200_last_error() or b""
        raise RaysError(f"rays_b200 error {rc}: {msg.decode(errors='replace')}")


def _dp(a):
    return None if a is None else a.ctypes.data_as(_abi.c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_abi.c_int32_p)


# ---- life cycle ------------------------------------------------------------------------------------
def init(device: int = 0) -> None:
    """Select the CUDA device of this process (replaces initialize_openmp_m, openmp_m.f90:39-71)."""
    _ck(_lib().rays_b200_init(int(device)))


def finalize() -> None:
    _ck(_lib().rays_b200_finalize())


def stop_string(code: int) -> str:
    buf = C.create_string_buffer(_abi.FLAG_LEN)
    _lib().rays_b200_stop_string(int(code), buf, _abi.FLAG_LEN)
    return buf.raw.decode()


# ---- the reference's program flow -------------------------------------------------------------------
def initialize(namelist_path: str, ray_init: bool = True, device: int | None = None) -> None:
    """initialize(read_input) (intialize.f90:1-94).  With ray_init=True the launch fan named by
    /ray_init_list/ is built on the device; with False the caller supplies one through set_fan()."""
    if ray_init or device is not None:
        init(0 if device is None else device)
    _ck(_lib().rays_host_initialize(str(namelist_path).encode(), 1 if ray_init else 0), host=True)


def host_cfg() -> Cfg:
    """The marshalled module state (rays_cfg) of the initialised host; pointers stay owned by the host."""
    p = _lib().rays_host_cfg()
    if not p:
        raise RaysError("initialize() has not been called")
    return p.contents


def set_ode(ode_solver_name: str = "", ray_deriv_name: str = "", nstep_max: int = 0, s_max: float = 0.0, ds: float = 0.0,
            rel_err0: float = 0.0, abs_err0: float = 0.0, SG_error_limit: float = 0.0) -> None:
    """Poke /ode_list/ module variables between runs, as ray_scan does (ray_scan.f90:33-49)."""
    _ck(_lib().rays_host_set_ode(ode_solver_name.encode(), ray_deriv_name.encode(), int(nstep_max), float(s_max), float(ds),
                                 float(rel_err0), float(abs_err0), float(SG_error_limit)), host=True)


def set_fan(rvec0: np.ndarray, rindex_vec0: np.ndarray, ray_pwr_wt: np.ndarray | None = None) -> None:
    rvec0 = np.ascontiguousarray(rvec0, dtype=np.float64).reshape(-1, 3)
    rindex_vec0 = np.ascontiguousarray(rindex_vec0, dtype=np.float64).reshape(-1, 3)
    w = None if ray_pwr_wt is None else np.ascontiguousarray(ray_pwr_wt, dtype=np.float64)
    _ck(_lib().rays_host_set_fan(rvec0.shape[0], _dp(rvec0), _dp(rindex_vec0), _dp(w)), host=True)


def get_fan():
    """(rvec0[nray,3], rindex_vec0[nray,3], ray_pwr_wt[nray]) of ray_init_m (copies)."""
    r, n, w = _abi.c_double_p(), _abi.c_double_p(), _abi.c_double_p()
    nray = _lib().rays_host_get_fan(C.byref(r), C.byref(n), C.byref(w))
    if nray == 0:
        return np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0)
    return (np.ctypeslib.as_array(r, (nray, 3)).copy(), np.ctypeslib.as_array(n, (nray, 3)).copy(),
            np.ctypeslib.as_array(w, (nray,)).copy())


def trace_rays() -> None:
    """trace_rays (ray_tracing.f90:1-290): fills the ray_results_m arrays."""
    _ck(_lib().rays_host_trace_rays(), host=True)


def finalize_run(outdir: str = ".") -> None:
    """finalize_run (finalize_run.f90:1-51): writes run_results.<run_label>.nc (netCDF classic)."""
    _ck(_lib().rays_host_finalize_run(str(outdir).encode()), host=True)


def write_deposition_profiles(outdir: str, profiles) -> None:
    """write_deposition_profiles_NC (deposition_profiles_m.f90:336-420): deposition_profiles.<run_label>.nc from a list of
    dicts with profile_name, grid_name, grid_min, grid_max, profile (n_bins values), Q_sum"""
    n = len(profiles)
    nb = len(profiles[0]["profile"])
    assert all(len(p["profile"]) == nb for p in profiles)
    names = "".join(p["profile_name"].ljust(20)[:20] for p in profiles).encode()
    gnames = "".join(p["grid_name"].ljust(20)[:20] for p in profiles).encode()
    gmin = np.array([p["grid_min"] for p in profiles], dtype=np.float64)
    gmax = np.array([p["grid_max"] for p in profiles], dtype=np.float64)
    prof = np.ascontiguousarray(np.stack([np.asarray(p["profile"], dtype=np.float64) for p in profiles]))
    q = np.array([p["Q_sum"] for p in profiles], dtype=np.float64)
    _ck(_lib().rays_host_write_deposition_profiles(str(outdir).encode(), n, names, gnames, nb, _dp(gmin), _dp(gmax), _dp(prof), _dp(q)), host=True)


def results() -> dict:
    """numpy copies of the ray_results_m arrays after trace_rays()."""
    r = Results()
    _ck(_lib().rays_host_results(C.byref(r)), host=True)
    nray, nv, npa = int(r.nray), int(r.nv), int(r.npoints_alloc)
    arr = np.ctypeslib.as_array

    def get(p, shape):
        return arr(p, shape).copy() if nray else np.zeros(shape)
    flags = C.string_at(r.ray_stop_flag, nray * _abi.FLAG_LEN) if nray else b""
    return dict(
        nray=nray, nv=nv, npoints_alloc=npa,
        ray_vec=get(r.ray_vec, (nray, npa, nv)), residual=get(r.residual, (nray, npa)),
        npoints=get(r.npoints, (nray,)), ray_stop_code=get(r.ray_stop_code, (nray,)),
        ray_stop_flag=[flags[i * _abi.FLAG_LEN:(i + 1) * _abi.FLAG_LEN].decode() for i in range(nray)],
        initial_ray_power=get(r.initial_ray_power, (nray,)), ray_trace_time=get(r.ray_trace_time, (nray,)),
        end_residuals=get(r.end_residuals, (nray,)), max_residuals=get(r.max_residuals, (nray,)),
        end_ray_parameter=get(r.end_ray_parameter, (nray,)), start_ray_vec=get(r.start_ray_vec, (nray, nv)),
        end_ray_vec=get(r.end_ray_vec, (nray, nv)), total_trace_time=float(r.total_trace_time), total_ray_steps=int(r.total_ray_steps))


# ---- direct use of the C ABI -------------------------------------------------------------------------
class ResultArrays:
    """Caller-owned result arrays in the reference layout (ray_results_m.f90:44-58), zero-filled like
    initialize_ray_results_m.  store=False leaves ray_vec/residual NULL (summaries only)."""

    def __init__(self, nray: int, nv: int, npoints_alloc: int, store: bool = True):
        self.nray, self.nv, self.npoints_alloc = int(nray), int(nv), int(npoints_alloc)
        self.ray_vec = np.zeros((nray, npoints_alloc, nv)) if store else None
        self.residual = np.zeros((nray, npoints_alloc)) if store else None
        self.npoints = np.zeros(nray, dtype=np.int32)
        self.ray_stop_code = np.zeros(nray, dtype=np.int32)
        self._flags = C.create_string_buffer(max(nray, 1) * _abi.FLAG_LEN)
        self.initial_ray_power = np.zeros(nray)
        self.ray_trace_time = np.zeros(nray)
        self.end_residuals = np.zeros(nray)
        self.max_residuals = np.zeros(nray)
        self.end_ray_parameter = np.zeros(nray)
        self.start_ray_vec = np.zeros((nray, nv))
        self.end_ray_vec = np.zeros((nray, nv))
        self.c = Results()
        c = self.c
        c.nray, c.nv, c.npoints_alloc = self.nray, self.nv, self.npoints_alloc
        c.ray_vec, c.residual = _dp(self.ray_vec), _dp(self.residual)
        c.npoints, c.ray_stop_code = _ip(self.npoints), _ip(self.ray_stop_code)
        c.ray_stop_flag = C.cast(self._flags, C.c_char_p)
        c.initial_ray_power, c.ray_trace_time = _dp(self.initial_ray_power), _dp(self.ray_trace_time)
        c.end_residuals, c.max_residuals, c.end_ray_parameter = _dp(self.end_residuals), _dp(self.max_residuals), _dp(self.end_ray_parameter)
        c.start_ray_vec, c.end_ray_vec = _dp(self.start_ray_vec), _dp(self.end_ray_vec)

    @property
    def ray_stop_flag(self):
        raw = self._flags.raw
        return [raw[i * _abi.FLAG_LEN:(i + 1) * _abi.FLAG_LEN].decode() for i in range(self.nray)]

    @property
    def total_ray_steps(self) -> int:
        return int(self.c.total_ray_steps)

    @property
    def total_trace_time(self) -> float:
        return float(self.c.total_trace_time)


def make_fan(rvec0: np.ndarray, rindex_vec0: np.ndarray, ray_pwr_wt: np.ndarray | None = None):
    """Returns (Fan struct, keep-alive tuple)."""
    r = np.ascontiguousarray(rvec0, dtype=np.float64).reshape(-1, 3)
    n = np.ascontiguousarray(rindex_vec0, dtype=np.float64).reshape(-1, 3)
    w = np.full(r.shape[0], 1.0 / max(r.shape[0], 1)) if ray_pwr_wt is None else np.ascontiguousarray(ray_pwr_wt, dtype=np.float64)
    f = Fan()
    f.nray = r.shape[0]
    f.rvec0, f.rindex_vec0, f.ray_pwr_wt = _dp(r), _dp(n), _dp(w)
    return f, (r, n, w)


def set_config(cfg: Cfg) -> None:
    _ck(_lib().rays_b200_set_config(C.byref(cfg)))


def trace(cfg: Cfg | None, rvec0, rindex_vec0, ray_pwr_wt=None, store: bool = True, out: ResultArrays | None = None) -> ResultArrays:
    """rays_b200_trace: host buffers in, host buffers out (H2D + kernels + D2H)."""
    cfg = cfg if cfg is not None else host_cfg()
    fan, keep = make_fan(rvec0, rindex_vec0, ray_pwr_wt)
    res = out if out is not None else ResultArrays(int(fan.nray), int(cfg.nv), int(cfg.nstep_max) + 1, store)
    _ck(_lib().rays_b200_trace(C.byref(cfg), C.byref(fan), C.byref(res.c)))
    del keep
    return res


def init_multi(ngpu: int = 0) -> int:
    """One host process, every GPU of the box (rays_b200.h): contexts + NCCL communicator; returns the GPU count."""
    _ck(_lib().rays_b200_init_multi(int(ngpu)))
    return int(_lib().rays_b200_ngpu())


def finalize_multi() -> None:
    _ck(_lib().rays_b200_finalize_multi())


def trace_multi(cfg: Cfg | None, rvec0, rindex_vec0, ray_pwr_wt=None, store: bool = True, out: ResultArrays | None = None,
                bins: tuple | None = None):
    """rays_b200_trace_multi / _binned: the fan sharded iray % ngpu over the GPUs of init_multi, results in fan order.
    bins=(n_bins, grid_min, grid_max): fused deposition + NCCL reduce / all-gather; returns (results, profile, Q_sum)."""
    cfg = cfg if cfg is not None else host_cfg()
    fan, keep = make_fan(rvec0, rindex_vec0, ray_pwr_wt)
    res = out if out is not None else ResultArrays(int(fan.nray), int(cfg.nv), int(cfg.nstep_max) + 1, store and bins is None)
    if bins is None:
        _ck(_lib().rays_b200_trace_multi(C.byref(cfg), C.byref(fan), C.byref(res.c)))
        del keep
        return res
    prof = np.zeros(int(bins[0]))
    d = Deposition()
    d.n_bins, d.grid_min, d.grid_max, d.profile = int(bins[0]), float(bins[1]), float(bins[2]), _dp(prof)
    _ck(_lib().rays_b200_trace_multi_binned(C.byref(cfg), C.byref(fan), C.byref(res.c), C.byref(d)))
    del keep
    return res, prof, float(d.Q_sum)


def fan_upload(rvec0, rindex_vec0, ray_pwr_wt=None) -> int:
    fan, keep = make_fan(rvec0, rindex_vec0, ray_pwr_wt)
    _ck(_lib().rays_b200_fan_upload(C.byref(fan)))
    del keep
    return int(fan.nray)


def fan_shard(rank: int, world: int) -> None:
    _ck(_lib().rays_b200_fan_shard(int(rank), int(world)))


def fan_download(nray: int):
    r, n, w = np.zeros((nray, 3)), np.zeros((nray, 3)), np.zeros(nray)
    _ck(_lib().rays_b200_fan_download(_dp(r), _dp(n), _dp(w)))
    return r, n, w


def launch_fan(kind: str, params) -> int:
    """Build the launch fan on the device; kind in {'slab','solovev','axisym'}; returns surviving rays."""
    L = _lib()
    n = C.c_int64(0)
    fn = {"slab": L.rays_b200_launch_fan_slab, "solovev": L.rays_b200_launch_fan_solovev, "axisym": L.rays_b200_launch_fan_axisym}[kind]
    _ck(fn(C.byref(params), C.byref(n)))
    return int(n.value)


def launch_fan_directions(rvec_in, nvec_in) -> int:
    r = np.ascontiguousarray(rvec_in, dtype=np.float64).reshape(-1, 3)
    d = np.ascontiguousarray(nvec_in, dtype=np.float64).reshape(-1, 3)
    n = C.c_int64(0)
    _ck(_lib().rays_b200_launch_fan_directions(r.shape[0], _dp(r), _dp(d), C.byref(n)))
    return int(n.value)


def trace_device(store: bool = True, bins: tuple | None = None) -> dict:
    """Trace the device-resident fan; bins=(n_bins, grid_min, grid_max) fuses deposition binning."""
    L = _lib()
    if bins is None:
        _ck(L.rays_b200_trace_device(1 if store else 0))
    else:
        _ck(L.rays_b200_trace_device_binned(int(bins[0]), float(bins[1]), float(bins[2]), 1 if store else 0))
    return last_trace_stats()


def last_trace_stats() -> dict:
    L = _lib()
    ms, steps, nl = C.c_double(0), C.c_int64(0), C.c_int32(0)
    L.rays_b200_last_trace_stats(C.byref(ms), C.byref(steps), C.byref(nl))
    rhs, grid, bps = C.c_int64(0), C.c_int32(0), C.c_int32(0)
    name = C.create_string_buffer(96)
    L.rays_b200_last_trace_info(C.byref(rhs), name, 96, C.byref(grid), C.byref(bps))
    first, resume, npass = C.c_double(0), C.c_double(0), C.c_int32(0)
    L.rays_b200_last_trace_breakdown(C.byref(first), C.byref(resume), C.byref(npass))
    return dict(kernel_ms=ms.value, ray_steps=steps.value, n_launches=nl.value, rhs_evals=rhs.value, kernel=name.value.decode(),
                grid=grid.value, blocks_per_sm=bps.value, first_pass_ms=first.value, resume_pass_ms=resume.value, n_passes=npass.value,
                trace_kernel_ms=first.value + resume.value)


def results_download(nray: int, nv: int, npoints_alloc: int, store: bool = True) -> ResultArrays:
    res = ResultArrays(nray, nv, npoints_alloc, store)
    _ck(_lib().rays_b200_results_download(C.byref(res.c)))
    return res


def deposition(n_bins: int, grid_min: float, grid_max: float, d_profile_out: int | None = None):
    """calculate_deposition_profiles (deposition_profiles_m.f90:228-260): (profile[n_bins], Q_sum).
    d_profile_out: optional DEVICE address of n_bins+1 doubles that receives this GPU's partial
    profile and Q_sum (for an NCCL reduce by the caller)."""
    prof = np.zeros(n_bins)
    d = Deposition()
    d.n_bins, d.grid_min, d.grid_max, d.profile = int(n_bins), float(grid_min), float(grid_max), _dp(prof)
    _ck(_lib().rays_b200_deposition(C.byref(d), C.c_void_p(d_profile_out) if d_profile_out else None))
    return prof, float(d.Q_sum)


def deposition_fixed(n_bins: int, grid_min: float, grid_max: float, d_acc_out: int | None = None):
    """This GPU's raw fixed-point bins: (acc[n_bins] int64, unit, profile, Q_sum); profile = acc * unit.  Summing `acc` over
    GPUs in integer arithmetic gives a profile that does not depend on the sharding (rays_b200.h).  d_acc_out: optional
    DEVICE address of n_bins int64 that receives the same bins (for an NCCL reduce)."""
    prof = np.zeros(n_bins)
    acc = np.zeros(n_bins, dtype=np.int64)
    unit = C.c_double(0)
    d = Deposition()
    d.n_bins, d.grid_min, d.grid_max, d.profile = int(n_bins), float(grid_min), float(grid_max), _dp(prof)
    _ck(_lib().rays_b200_deposition_fixed(C.byref(d), acc.ctypes.data_as(C.POINTER(C.c_int64)), C.c_void_p(d_acc_out) if d_acc_out else None, C.byref(unit)))
    return acc, unit.value, prof, float(d.Q_sum)


def summaries_pack(d_out: int | None = None, rows_capacity: int = 0):
    """Pack the per-ray summaries of the last trace into DEVICE memory at d_out (rows of 6 + 2 nv doubles); returns (rows, row_doubles)."""
    rows, rd = C.c_int64(0), C.c_int32(0)
    _ck(_lib().rays_b200_summaries_pack(C.c_void_p(d_out) if d_out else None, int(rows_capacity), C.byref(rows), C.byref(rd)))
    return int(rows.value), int(rd.value)


def deposition_set_total_weight(total_weight: float) -> None:
    _ck(_lib().rays_b200_deposition_set_total_weight(float(total_weight)))


def probe_equilibrium(rvec: np.ndarray):
    r = np.ascontiguousarray(rvec, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((r.shape[0], _abi.EQ_OUT))
    err = np.zeros(r.shape[0], dtype=np.int32)
    _ck(_lib().rays_b200_probe_equilibrium(r.shape[0], _dp(r), _dp(out), _ip(err)))
    return out, err


def probe_rhs(v: np.ndarray):
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros_like(v)
    st = np.zeros(v.shape[0], dtype=np.int32)
    _ck(_lib().rays_b200_probe_rhs(v.shape[0], _dp(v), _dp(out), _ip(st)))
    return out, st


def probe_check_save(v: np.ndarray):
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(v.shape[0])
    st = np.zeros(v.shape[0], dtype=np.int32)
    _ck(_lib().rays_b200_probe_check_save(v.shape[0], _dp(v), _dp(out), _ip(st)))
    return out, st


def ox_conv_analysis(nray: int):
    """analyze_OX_conv (OX_conv_analysis_m.f90:91-198) on the trajectories of the last stored device trace:
    (records[nray] as a numpy structured array, number of converted rays)"""
    from ._abi import OX_DTYPE
    out = np.zeros(nray, dtype=np.dtype(OX_DTYPE, align=True))
    assert out.dtype.itemsize == C.sizeof(_abi.OxConv)
    n = C.c_int64(0)
    _ck(_lib().rays_b200_ox_conv_analysis(out.ctypes.data_as(C.c_void_p), C.byref(n)))
    return out, int(n.value)


def make_coils(coils) -> "C.Array":
    """list of dicts (inner_radius, outer_radius, z_center, z_width, I_coil, n_turns, n_r_layers, n_z_slices) -> rays_coil[]"""
    from ._abi import Coil
    arr = (Coil * len(coils))()
    for c, d in zip(arr, coils):
        for k, v in d.items():
            setattr(c, k, v)
    return arr


def mirror_Brz_grid(coils, n_r: int, r_min: float, r_max: float, n_z: int, z_min: float, z_max: float):
    """calculate_B_on_rz_grid (mirror_magnetics_m.f90:324-368) on the GPU: r_grid, z_grid, Br, Bz, Aphi; fields [n_z][n_r]"""
    arr = coils if not isinstance(coils, (list, tuple)) else make_coils(coils)
    rg, zg = np.zeros(n_r), np.zeros(n_z)
    Br, Bz, Aphi = (np.zeros((n_z, n_r)) for _ in range(3))
    _ck(_lib().rays_b200_mirror_brz_grid(arr, len(arr), n_r, r_min, r_max, n_z, z_min, z_max, _dp(rg), _dp(zg), _dp(Br), _dp(Bz), _dp(Aphi)))
    return rg, zg, Br, Bz, Aphi


def mirror_magnetics(namelist_path: str, outdir: str = ".") -> str:
    """program mirror_magnetics (mirror_magnetics_lib/mirror_magnetics.f90): namelists -> Brz_fields.<name>.nc; returns its path"""
    buf = C.create_string_buffer(512)
    _ck(_lib().rays_host_mirror_magnetics(str(namelist_path).encode(), str(outdir).encode(), buf, 512), host=True)
    return buf.value.decode()


def fp64_peak() -> tuple[float, float]:
    """(measured DFMA TFLOP/s, max SM clock MHz): the FP64 roofline denominator."""
    t, m = C.c_double(0), C.c_double(0)
    _ck(_lib().rays_b200_fp64_peak(C.byref(t), C.byref(m)))
    return t.value, m.value


def config_path(name: str) -> str:
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", name)
