"""ctypes view of include/rays_b200.h (the C ABI) and of the host-side mirror (rays_host_*).

The structures below are field-for-field copies of the C structs; `tests/test_abi.py` checks their
sizes against `rays_b200_struct_sizes` of the built library.  The library is loaded from
`rays_b200/lib/librays_b200.so` (built in-tree by `__graft_entry__.build()` / `make -C
rays_b200/csrc`); a missing library is an error — there is no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

NSPECIES = 6
NV_MAX = 19
FLAG_LEN = 60
EQ_OUT = 106

# enums (include/rays_b200.h)
EQ_SLAB, EQ_SOLOVEV, EQ_AXISYM_TOROID, EQ_MULTIPLE_MIRROR = 1, 2, 3, 4
ODE_RK4, ODE_SG = 1, 2
DERIV_COLD, DERIV_NUM = 1, 2
MAG_SOLOVEV, MAG_EQDSK_SPLINE = 1, 2
PROF_SPLINE = 7   # enum rays_prof_model: density_spline_interp / temperature_spline_interp
PARAM_ARCL, PARAM_TIME = 1, 2
DAMP_NONE, DAMP_FUND_ECH = 0, 1

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
D6 = C.c_double * NSPECIES
I6 = C.c_int32 * NSPECIES


class Spline1D(C.Structure):
    _fields_ = [("nx", C.c_int32), ("pad_", C.c_int32), ("x_grid", c_double_p), ("fspl", c_double_p)]


class Spline2D(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("x_grid", c_double_p), ("y_grid", c_double_p), ("fspl", c_double_p)]


class SlabEq(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax", "rmaj", "rmin", "x0")] + [
        (n, C.c_int32) for n in ("bx_prof_model", "by_prof_model", "bz_prof_model", "dens_prof_model")] + [
        (n, C.c_double) for n in ("bx0", "by0", "bz0", "LBy_shear_scale", "LBz_scale", "dBzdx", "Ln_scale", "dndx", "alphan1", "alphan2", "n_min")] + [
        ("t_prof_model", I6), ("LT_scale", C.c_double), ("dtdx", C.c_double), ("alphat1", D6), ("alphat2", D6), ("T_min", D6)]


class SolovevEq(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("rmaj", "kappa", "bphi0", "iota0", "outer_bound", "psiB", "inner_bound", "vert_bound", "r_Zmax",
                                           "box_rmin", "box_rmax", "box_zmin", "box_zmax")] + [
        ("dens_prof_model", C.c_int32), ("t_prof_model", I6), ("pad_", C.c_int32), ("alphan1", C.c_double), ("alphan2", C.c_double),
        ("alphat1", D6), ("alphat2", D6)]


class AxisymEq(C.Structure):
    _fields_ = [("magnetics_model", C.c_int32), ("density_prof_model", C.c_int32), ("temperature_prof_model", I6)] + [
        (n, C.c_double) for n in ("r_axis", "z_axis", "box_rmin", "box_rmax", "box_zmin", "box_zmax", "inner_bound", "outer_bound", "upper_bound",
                                  "lower_bound", "plasma_psi_limit", "alphan1", "alphan2", "d_scrape_off", "T_scrape_off")] + [
        ("alphat1", D6), ("alphat2", D6)] + [
        (n, C.c_double) for n in ("sm_rmaj", "sm_kappa", "sm_bphi0", "sm_iota0", "sm_psiB", "sm_box_rmin", "sm_box_rmax", "sm_box_zmin", "sm_box_zmax")] + [
        ("ne_spline", Spline1D), ("Te_spline", Spline1D), ("Ti_spline", Spline1D),
        ("Psi_spline", Spline2D), ("T_spline", Spline1D), ("eq_psibound", C.c_double)]


class MirrorEq(C.Structure):
    _fields_ = [("density_prof_model", C.c_int32), ("temperature_prof_model", I6), ("pad_", C.c_int32)] + [
        (n, C.c_double) for n in ("box_rmax", "box_zmin", "box_zmax", "r_LUFS", "z_LUFS", "Aphi_LUFS", "plasma_AphiN_limit", "alphan1", "alphan2",
                                  "AphiN0_d", "delta_d", "d_scrape_off", "T_scrape_off")] + [
        ("alphat1", D6), ("alphat2", D6), ("AphiN0_t", D6), ("delta_t", D6),
        ("Br_spline", Spline2D), ("Bz_spline", Spline2D), ("Aphi_spline", Spline2D)]


class Coil(C.Structure):
    """rays_coil (include/rays_b200.h): one coil of the mirror coil set"""
    _fields_ = [(n, C.c_double) for n in ("inner_radius", "outer_radius", "z_center", "z_width", "I_coil")] + [
        ("n_turns", C.c_int64), ("n_r_layers", C.c_int32), ("n_z_slices", C.c_int32)]


class OxConv(C.Structure):
    """rays_ox_conv (include/rays_b200.h): per-ray outcome of analyze_OX_conv"""
    _fields_ = [("x_max", C.c_double * 3), ("k_max", C.c_double * 3), ("alpha_max", C.c_double), ("x_cut", C.c_double * 3), ("conv_coeff", C.c_double),
                ("nvecx_c", C.c_double * 3), ("nvecy_c", C.c_double * 3), ("nvecz_c", C.c_double * 3)] + [
        (n, C.c_int32) for n in ("ray_number", "step_number", "found_max", "found_cutoff", "converted", "iteration")]


OX_DTYPE = [("x_max", "f8", 3), ("k_max", "f8", 3), ("alpha_max", "f8"), ("x_cut", "f8", 3), ("conv_coeff", "f8"), ("nvecx_c", "f8", 3),
            ("nvecy_c", "f8", 3), ("nvecz_c", "f8", 3), ("ray_number", "i4"), ("step_number", "i4"), ("found_max", "i4"), ("found_cutoff", "i4"),
            ("converted", "i4"), ("iteration", "i4")]


class Cfg(C.Structure):
    _fields_ = [
        ("clight", C.c_double), ("eps0", C.c_double),
        ("omgrf", C.c_double), ("k0", C.c_double), ("dispersion_resid_limit", C.c_double),
        ("ray_param", C.c_int32), ("wave_mode", C.c_int32), ("k0_sign", C.c_int32),
        ("nspec", C.c_int32),
        ("qs", D6), ("ms", D6), ("n0s", D6), ("t0s", D6), ("eta", D6),
        ("ode_solver", C.c_int32), ("ray_deriv", C.c_int32), ("nv", C.c_int32), ("nstep_max", C.c_int32),
        ("ds", C.c_double), ("s_max", C.c_double),
        ("rel_err0", C.c_double), ("abs_err0", C.c_double), ("SG_error_limit", C.c_double),
        ("damping_model", C.c_int32), ("multi_spec_damping", C.c_int32),
        ("total_damping_limit", C.c_double),
        ("integrate_eq_gradients", C.c_int32),
        ("equilib_model", C.c_int32),
        ("slab", SlabEq), ("solovev", SolovevEq), ("axisym", AxisymEq), ("mirror", MirrorEq),
        ("zfun_re", Spline1D),
    ]


class Fan(C.Structure):
    _fields_ = [("nray", C.c_int64), ("rvec0", c_double_p), ("rindex_vec0", c_double_p), ("ray_pwr_wt", c_double_p)]


class Results(C.Structure):
    _fields_ = [
        ("nray", C.c_int64), ("nv", C.c_int32), ("npoints_alloc", C.c_int32),
        ("ray_vec", c_double_p), ("residual", c_double_p), ("npoints", c_int32_p), ("ray_stop_code", c_int32_p),
        ("ray_stop_flag", C.c_char_p), ("initial_ray_power", c_double_p), ("ray_trace_time", c_double_p),
        ("end_residuals", c_double_p), ("max_residuals", c_double_p), ("end_ray_parameter", c_double_p),
        ("start_ray_vec", c_double_p), ("end_ray_vec", c_double_p),
        ("total_trace_time", C.c_double), ("total_ray_steps", C.c_int64),
    ]


class Deposition(C.Structure):
    _fields_ = [("n_bins", C.c_int32), ("pad_", C.c_int32), ("grid_min", C.c_double), ("grid_max", C.c_double),
                ("profile", c_double_p), ("Q_sum", C.c_double)]


class SolovevLaunch(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_r_launch", "n_theta_launch", "n_rindex_theta", "n_rindex_phi")] + [
        (n, C.c_double) for n in ("r_launch0", "dr_launch", "theta_launch0", "dtheta_launch", "rindex_theta0", "delta_rindex_theta",
                                  "rindex_phi0", "delta_rindex_phi")]


class AxisymLaunch(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_R_launch", "n_Z_launch", "n_rindex_theta", "n_rindex_phi")] + [
        (n, C.c_double) for n in ("R_launch0", "Z_launch0", "rindex_theta0", "delta_rindex_theta", "rindex_phi0", "delta_rindex_phi")]


class SlabLaunch(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_x_launch", "n_y_launch", "n_z_launch", "n_ky_launch", "n_kz_launch", "pad_")] + [
        (n, C.c_double) for n in ("x_launch0", "dx_launch", "y_launch0", "dy_launch", "z_launch0", "dz_launch", "rindex_y0", "delta_rindex_y0",
                                  "rindex_z0", "delta_rindex_z0")]


# every extern "C" symbol include/rays_b200.h declares (tests/test_abi.py checks they are exported)
ABI_SYMBOLS = [
    "rays_b200_init", "rays_b200_finalize", "rays_b200_last_error", "rays_b200_stop_string", "rays_b200_set_config",
    "rays_b200_trace", "rays_b200_fan_upload", "rays_b200_trace_device", "rays_b200_results_download",
    "rays_b200_last_trace_stats", "rays_b200_last_trace_info", "rays_b200_fan_shard",
    "rays_b200_launch_fan_solovev", "rays_b200_launch_fan_axisym", "rays_b200_launch_fan_slab", "rays_b200_launch_fan_directions",
    "rays_b200_fan_download", "rays_b200_deposition", "rays_b200_trace_device_binned",
    "rays_b200_probe_equilibrium", "rays_b200_probe_rhs", "rays_b200_probe_check_save",
    "rays_b200_fp64_peak", "rays_b200_host_alloc", "rays_b200_host_free", "rays_b200_stream", "rays_b200_version",
    "rays_b200_struct_sizes", "rays_b200_selftest_arith", "rays_b200_last_trace_breakdown", "rays_b200_mirror_brz_grid", "rays_b200_ox_conv_analysis",
    "rays_b200_deposition_fixed", "rays_b200_deposition_set_total_weight", "rays_b200_summaries_pack",
    "rays_b200_init_multi", "rays_b200_ngpu", "rays_b200_trace_multi", "rays_b200_trace_multi_binned", "rays_b200_multi_summaries", "rays_b200_finalize_multi",
]
HOST_SYMBOLS = [
    "rays_host_initialize", "rays_host_trace_rays", "rays_host_finalize_run", "rays_host_deallocate", "rays_host_last_error",
    "rays_host_cfg", "rays_host_nspec", "rays_host_run_label", "rays_host_ray_init_model", "rays_host_launch_params",
    "rays_host_directions_in", "rays_host_set_ode", "rays_host_set_fan", "rays_host_get_fan", "rays_host_results",
    "rays_host_zfun", "rays_host_cspline", "rays_host_bcspline", "rays_host_mirror_magnetics", "rays_host_write_deposition_profiles",
]

# RAYS_B200_LIB overrides the library path (A/B experiments with alternative builds)
_LIB_PATH = os.environ.get("RAYS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "librays_b200.so")
_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load librays_b200.so and declare the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C rays_b200/csrc`).  rays_b200 has no CPU or pure-Python fallback.")
    L = C.CDLL(_LIB_PATH)
    i, i64, dbl, vp, cp = C.c_int, C.c_int64, C.c_double, C.c_void_p, C.c_char_p
    P = C.POINTER
    protos = {
        "rays_b200_init": (i, [i]), "rays_b200_finalize": (i, []), "rays_b200_last_error": (cp, []),
        "rays_b200_stop_string": (i, [i, C.c_char_p, i]), "rays_b200_set_config": (i, [P(Cfg)]),
        "rays_b200_trace": (i, [P(Cfg), P(Fan), P(Results)]), "rays_b200_fan_upload": (i, [P(Fan)]),
        "rays_b200_trace_device": (i, [i]), "rays_b200_results_download": (i, [P(Results)]),
        "rays_b200_last_trace_stats": (i, [P(dbl), P(i64), P(C.c_int32)]),
        "rays_b200_last_trace_info": (i, [P(i64), C.c_char_p, i, P(C.c_int32), P(C.c_int32)]),
        "rays_b200_fan_shard": (i, [i, i]),
        "rays_b200_launch_fan_solovev": (i, [P(SolovevLaunch), P(i64)]), "rays_b200_launch_fan_axisym": (i, [P(AxisymLaunch), P(i64)]),
        "rays_b200_launch_fan_slab": (i, [P(SlabLaunch), P(i64)]),
        "rays_b200_launch_fan_directions": (i, [i64, c_double_p, c_double_p, P(i64)]),
        "rays_b200_fan_download": (i, [c_double_p, c_double_p, c_double_p]),
        "rays_b200_deposition": (i, [P(Deposition), vp]), "rays_b200_trace_device_binned": (i, [i, dbl, dbl, i]),
        "rays_b200_deposition_fixed": (i, [P(Deposition), P(i64), vp, P(dbl)]), "rays_b200_deposition_set_total_weight": (i, [dbl]),
        "rays_b200_summaries_pack": (i, [vp, i64, P(i64), P(C.c_int32)]),
        "rays_b200_init_multi": (i, [i]), "rays_b200_ngpu": (i, []), "rays_b200_finalize_multi": (i, []),
        "rays_b200_trace_multi": (i, [P(Cfg), P(Fan), P(Results)]), "rays_b200_trace_multi_binned": (i, [P(Cfg), P(Fan), P(Results), P(Deposition)]),
        "rays_b200_multi_summaries": (vp, [i, P(i64), P(C.c_int32)]),
        "rays_b200_probe_equilibrium": (i, [i64, c_double_p, c_double_p, c_int32_p]),
        "rays_b200_probe_rhs": (i, [i64, c_double_p, c_double_p, c_int32_p]),
        "rays_b200_probe_check_save": (i, [i64, c_double_p, c_double_p, c_int32_p]),
        "rays_b200_fp64_peak": (i, [P(dbl), P(dbl)]), "rays_b200_host_alloc": (i, [P(vp), C.c_size_t]), "rays_b200_host_free": (i, [vp]),
        "rays_b200_stream": (vp, []), "rays_b200_version": (i, []), "rays_b200_struct_sizes": (i, [P(C.c_int32), i]),
        "rays_b200_selftest_arith": (i, [i64, C.c_uint64, P(i64)]),
        "rays_b200_last_trace_breakdown": (i, [P(dbl), P(dbl), P(C.c_int32)]),
        "rays_host_initialize": (i, [cp, i]), "rays_host_trace_rays": (i, []), "rays_host_finalize_run": (i, [cp]),
        "rays_host_deallocate": (i, []), "rays_host_last_error": (cp, []), "rays_host_cfg": (P(Cfg), []),
        "rays_host_nspec": (i, []), "rays_host_run_label": (cp, []), "rays_host_ray_init_model": (cp, []),
        "rays_host_launch_params": (i, [P(SlabLaunch), P(SolovevLaunch), P(AxisymLaunch)]),
        "rays_host_directions_in": (i64, [P(c_double_p), P(c_double_p)]),
        "rays_host_set_ode": (i, [cp, cp, i, dbl, dbl, dbl, dbl, dbl]),
        "rays_host_set_fan": (i, [i64, c_double_p, c_double_p, c_double_p]),
        "rays_host_get_fan": (i64, [P(c_double_p), P(c_double_p), P(c_double_p)]),
        "rays_host_results": (i, [P(Results)]),
        "rays_host_zfun": (i, [dbl, dbl, P(dbl), P(dbl)]),
        "rays_b200_mirror_brz_grid": (i, [P(Coil), C.c_int32, C.c_int32, dbl, dbl, C.c_int32, dbl, dbl, c_double_p, c_double_p, c_double_p,
                                       c_double_p, c_double_p]),
        "rays_host_mirror_magnetics": (i, [cp, cp, C.c_char_p, i]),
        "rays_host_write_deposition_profiles": (i, [cp, i, cp, cp, i, c_double_p, c_double_p, c_double_p, c_double_p]),
        "rays_b200_ox_conv_analysis": (i, [vp, P(i64)]),
        "rays_host_cspline": (i, [c_double_p, i, c_double_p]), "rays_host_bcspline": (i, [c_double_p, i, c_double_p, i, c_double_p]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


STRUCTS_IN_SIZE_ORDER = [Cfg, Fan, Results, Deposition, SolovevLaunch, AxisymLaunch, SlabLaunch, Spline1D, Spline2D, SlabEq, SolovevEq, AxisymEq, MirrorEq]
