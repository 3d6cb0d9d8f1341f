"""Shared workload builders for the tests: initialise the host mirror from a namelist, override the
ODE options, build a launch fan with the ORACLE (so that both sides integrate identical rays)."""
from __future__ import annotations

import ctypes as C

import numpy as np

import rays_b200 as rb
from rays_b200 import _abi
import _oracle as orc


def init_case(name: str, **ode):
    """initialize(read_input) without device launch; returns the live rays_cfg of the host."""
    L = _abi.load()
    rc = L.rays_host_initialize(rb.config_path(name).encode(), 0)
    assert rc == 0, L.rays_host_last_error()
    if ode:
        rb.set_ode(**ode)
    return rb.host_cfg()


def init_case_text(name: str, edits, tmpdir, **ode):
    """like init_case, but with (old, new) text substitutions applied to the namelist first"""
    import os
    txt = open(rb.config_path(name)).read()
    for old, new in edits:
        assert old in txt, old
        txt = txt.replace(old, new)
    p = os.path.join(str(tmpdir), "rays.in")
    open(p, "w").write(txt)
    L = _abi.load()
    rc = L.rays_host_initialize(p.encode(), 0)
    assert rc == 0, L.rays_host_last_error()
    if ode:
        rb.set_ode(**ode)
    return rb.host_cfg()


def launch_params():
    L = _abi.load()
    sl, so, ax = _abi.SlabLaunch(), _abi.SolovevLaunch(), _abi.AxisymLaunch()
    L.rays_host_launch_params(C.byref(sl), C.byref(so), C.byref(ax))
    return sl, so, ax


def oracle_fan(cfg, cap=200000, **override):
    """Launch fan of the initialised case from the oracle's launcher; `override` patches the launcher's
    namelist values (e.g. n_rindex_phi=8)."""
    L = _abi.load()
    model = L.rays_host_ray_init_model().decode().strip()
    sl, so, ax = launch_params()
    if model == "simple_slab":
        kind, p = "slab", sl
    elif model == "solovev":
        kind, p = "solovev", so
    elif model == "axisym_toroid_ray_init_R_Z_nphi_ntheta":
        kind, p = "axisym", ax
    else:
        rin, nin = _abi.c_double_p(), _abi.c_double_p()
        k = L.rays_host_directions_in(C.byref(rin), C.byref(nin))
        return orc.launch_fan_directions(cfg, np.ctypeslib.as_array(rin, (k, 3)).copy(), np.ctypeslib.as_array(nin, (k, 3)).copy(), False) + (None, None)
    for k, v in override.items():
        setattr(p, k, v)
    return orc.launch_fan(cfg, kind, p, cap) + (kind, p)


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) per column-normalised scale"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def vec_rel_err(a, b):
    """relative error of position / wave-vector triples measured against the vector norm"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nrm = np.maximum(np.linalg.norm(b, axis=-1, keepdims=True), 1e-300)
    return float(np.max(np.abs(a - b) / nrm)) if a.size else 0.0
