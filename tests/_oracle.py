"""ctypes binding of the CPU oracle (oracle/_build/librays_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rays_b200 import _abi
from rays_b200 import ResultArrays, make_fan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "librays_oracle.so")
_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_SO):
        build()
    L = C.CDLL(ORACLE_SO)
    P, i, l, d = C.POINTER, C.c_int, C.c_long, C.c_double
    dp, ip = _abi.c_double_p, _abi.c_int32_p
    L.oracle_trace.argtypes = [P(_abi.Cfg), P(_abi.Fan), P(_abi.Results), i, P(l)]
    L.oracle_count_flops.argtypes = [P(_abi.Cfg), P(_abi.Fan), l, l, P(d), P(l), P(l)]
    L.oracle_probe_equilibrium.argtypes = [P(_abi.Cfg), l, dp, dp, ip]
    L.oracle_probe_rhs.argtypes = [P(_abi.Cfg), l, dp, dp, ip]
    L.oracle_probe_check_save.argtypes = [P(_abi.Cfg), l, dp, dp, ip]
    L.oracle_launch_fan_solovev.argtypes = [P(_abi.Cfg), P(_abi.SolovevLaunch), l, dp, dp, dp]
    L.oracle_launch_fan_solovev.restype = l
    L.oracle_launch_fan_axisym.argtypes = [P(_abi.Cfg), P(_abi.AxisymLaunch), l, dp, dp, dp]
    L.oracle_launch_fan_axisym.restype = l
    L.oracle_launch_fan_slab.argtypes = [P(_abi.Cfg), P(_abi.SlabLaunch), l, dp, dp, dp]
    L.oracle_launch_fan_slab.restype = l
    L.oracle_launch_fan_directions.argtypes = [P(_abi.Cfg), l, dp, dp, i, l, dp, dp, dp]
    L.oracle_launch_fan_directions.restype = l
    L.oracle_deposition.argtypes = [P(_abi.Cfg), P(_abi.Results), P(_abi.Deposition)]
    L.oracle_binner.argtypes = [dp, dp, i, d, d, dp, i]
    L.oracle_zfun.argtypes = [P(_abi.Cfg), l, dp, dp, dp, dp]
    L.oracle_cspeval.argtypes = [P(_abi.Spline1D), l, dp, dp, dp]
    L.oracle_bcspeval.argtypes = [P(_abi.Spline2D), l, dp, dp, dp, dp, dp]
    L.oracle_solve_n1.argtypes = [P(_abi.Cfg), dp, d, d, P(d), P(d)]
    L.oracle_solve_nsq_theta.argtypes = [P(_abi.Cfg), dp, d, dp]
    L.oracle_stop_string.argtypes = [i, C.c_char_p, i]
    _lib = L
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(_abi.c_double_p)


def _ip(a):
    return a.ctypes.data_as(_abi.c_int32_p)


def trace(cfg, rvec0, rindex_vec0, ray_pwr_wt=None, store=True, nthreads=0):
    """oracle trace_rays over a fan -> (ResultArrays, status, nrhs)"""
    L = load()
    fan, keep = make_fan(rvec0, rindex_vec0, ray_pwr_wt)
    res = ResultArrays(int(fan.nray), int(cfg.nv), int(cfg.nstep_max) + 1, store)
    nrhs = C.c_long(0)
    st = L.oracle_trace(C.byref(cfg), C.byref(fan), C.byref(res.c), int(nthreads), C.byref(nrhs))
    del keep
    return res, st, int(nrhs.value)


def count_flops(cfg, rvec0, rindex_vec0, first=0, count=8):
    L = load()
    fan, keep = make_fan(rvec0, rindex_vec0, None)
    fl, st, nr = C.c_double(0), C.c_long(0), C.c_long(0)
    L.oracle_count_flops(C.byref(cfg), C.byref(fan), first, count, C.byref(fl), C.byref(st), C.byref(nr))
    del keep
    return fl.value, int(st.value), int(nr.value)


def launch_fan(cfg, kind, params, cap):
    L = load()
    r, n, w = np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros(cap)
    fn = {"slab": L.oracle_launch_fan_slab, "solovev": L.oracle_launch_fan_solovev, "axisym": L.oracle_launch_fan_axisym}[kind]
    k = fn(C.byref(cfg), C.byref(params), cap, _dp(r), _dp(n), _dp(w))
    if k < 0:
        raise RuntimeError(f"fan needs capacity {-k}")
    return r[:k].copy(), n[:k].copy(), w[:k].copy()


def launch_fan_directions(cfg, rvec_in, nvec_in, all_weights_zero=False):
    L = load()
    rin = np.ascontiguousarray(rvec_in, dtype=np.float64).reshape(-1, 3)
    nin = np.ascontiguousarray(nvec_in, dtype=np.float64).reshape(-1, 3)
    cap = rin.shape[0]
    r, n, w = np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros(cap)
    k = L.oracle_launch_fan_directions(C.byref(cfg), cap, _dp(rin), _dp(nin), 1 if all_weights_zero else 0, cap, _dp(r), _dp(n), _dp(w))
    return r[:k].copy(), n[:k].copy(), w[:k].copy()


def probe_equilibrium(cfg, rvec):
    L = load()
    r = np.ascontiguousarray(rvec, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((r.shape[0], _abi.EQ_OUT))
    err = np.zeros(r.shape[0], dtype=np.int32)
    L.oracle_probe_equilibrium(C.byref(cfg), r.shape[0], _dp(r), _dp(out), _ip(err))
    return out, err


def probe_rhs(cfg, v):
    L = load()
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros_like(v)
    st = np.zeros(v.shape[0], dtype=np.int32)
    L.oracle_probe_rhs(C.byref(cfg), v.shape[0], _dp(v), _dp(out), _ip(st))
    return out, st


def probe_check_save(cfg, v):
    L = load()
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(v.shape[0])
    st = np.zeros(v.shape[0], dtype=np.int32)
    L.oracle_probe_check_save(C.byref(cfg), v.shape[0], _dp(v), _dp(out), _ip(st))
    return out, st


def deposition(cfg, res: ResultArrays, n_bins, grid_min, grid_max):
    L = load()
    prof = np.zeros(n_bins)
    d = _abi.Deposition()
    d.n_bins, d.grid_min, d.grid_max, d.profile = n_bins, grid_min, grid_max, _dp(prof)
    L.oracle_deposition(C.byref(cfg), C.byref(res.c), C.byref(d))
    return prof, float(d.Q_sum)


def kx_profile(cfg, x, ny, nz, mode):
    """k0*nx of the cold root `mode` (1 plus, 2 minus, 3 fast, 4 slow) at (x, 0, 0), single precision as written"""
    L = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    re, im = np.zeros_like(x), np.zeros_like(x)
    L.oracle_kx_profile.argtypes = [C.POINTER(_abi.Cfg), C.c_long, _abi.c_double_p, C.c_double, C.c_double, C.c_int, _abi.c_double_p, _abi.c_double_p]
    L.oracle_kx_profile(C.byref(cfg), len(x), _dp(x), float(ny), float(nz), int(mode), _dp(re), _dp(im))
    return re, im


def ox_conv(cfg, res: ResultArrays):
    L = load()
    out = np.zeros(res.nray, dtype=np.dtype(_abi.OX_DTYPE, align=True))
    L.oracle_ox_conv.argtypes = [C.POINTER(_abi.Cfg), C.POINTER(_abi.Results), C.c_void_p]
    L.oracle_ox_conv(C.byref(cfg), C.byref(res.c), out.ctypes.data_as(C.c_void_p))
    return out, int(out["converted"].sum())


def mirror_Brz_grid(coils, n_r, r_min, r_max, n_z, z_min, z_max):
    import rays_b200 as rb
    L = load()
    arr = coils if not isinstance(coils, (list, tuple)) else rb.make_coils(coils)
    rg, zg = np.zeros(n_r), np.zeros(n_z)
    Br, Bz, Aphi = (np.zeros((n_z, n_r)) for _ in range(3))
    L.oracle_mirror_Brz_grid.argtypes = [C.POINTER(_abi.Coil), C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double] + [_abi.c_double_p] * 5
    L.oracle_mirror_Brz_grid(arr, len(arr), n_r, r_min, r_max, n_z, z_min, z_max, _dp(rg), _dp(zg), _dp(Br), _dp(Bz), _dp(Aphi))
    return rg, zg, Br, Bz, Aphi


def Brz_loop_scaled(r, z):
    L = load()
    r = np.ascontiguousarray(r, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    Br, Bz, A = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
    L.oracle_Brz_loop_scaled.argtypes = [C.c_long] + [_abi.c_double_p] * 5
    L.oracle_Brz_loop_scaled(len(r), _dp(r), _dp(z), _dp(Br), _dp(Bz), _dp(A))
    return Br, Bz, A


def elliptic(m):
    L = load()
    m = np.ascontiguousarray(m, dtype=np.float64)
    K, E = np.zeros_like(m), np.zeros_like(m)
    L.oracle_elliptic.argtypes = [C.c_long] + [_abi.c_double_p] * 3
    L.oracle_elliptic(len(m), _dp(m), _dp(K), _dp(E))
    return K, E


def binner(Q, xQ, xmin, xmax, n_bins):
    L = load()
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    xQ = np.ascontiguousarray(xQ, dtype=np.float64)
    out = np.zeros(n_bins)
    ierr = L.oracle_binner(_dp(Q), _dp(xQ), len(Q), xmin, xmax, _dp(out), n_bins)
    return out, ierr


def zfun(cfg, x, kz):
    L = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    kz = np.ascontiguousarray(kz, dtype=np.float64)
    re, im = np.zeros_like(x), np.zeros_like(x)
    L.oracle_zfun(C.byref(cfg), len(x), _dp(x), _dp(kz), _dp(re), _dp(im))
    return re, im


def num_threads() -> int:
    return load().oracle_num_threads()
