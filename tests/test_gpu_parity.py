"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): RK4 — identical step counts and stop reasons, final (x, k, power)
<= 1e-10 relative; SG — identical stop reasons, trajectories within the run's own ODE tolerance.  The
kernels evaluate the reference's expressions in the reference's order without FMA contraction, so
wherever no libm function is involved the comparison below is in fact BITWISE."""
import numpy as np
import pytest

import rays_b200 as rb
import _oracle as orc
from _cases import init_case, init_case_text, oracle_fan, vec_rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device(built):
    rb.init(0)
    yield


def _compare_traces(g, o, cfg, tol, bitwise):
    assert np.array_equal(g.npoints, o.npoints), f"npoints differ at {np.nonzero(g.npoints != o.npoints)[0][:10]}"
    assert np.array_equal(g.ray_stop_code, o.ray_stop_code)
    assert g.ray_stop_flag == o.ray_stop_flag
    assert g.total_ray_steps == int(np.sum(o.npoints - 1))
    nv = int(cfg.nv)
    if bitwise:
        # NaN end vectors are legitimate: outside the Solov'ev plasma deriv_cold divides 0*0/0
        # (deriv_cold.f90:86) and the ray stops with 'infinite_Vg' on both sides
        assert np.array_equal(g.ray_vec, o.ray_vec, equal_nan=True), f"max abs diff {np.nanmax(np.abs(g.ray_vec - o.ray_vec))}"
        assert np.array_equal(g.residual, o.residual, equal_nan=True)
        assert np.array_equal(g.end_ray_vec, o.end_ray_vec, equal_nan=True)
        assert np.array_equal(g.end_residuals, o.end_residuals, equal_nan=True)
        assert np.array_equal(g.max_residuals, o.max_residuals, equal_nan=True)
        assert np.array_equal(g.end_ray_parameter, o.end_ray_parameter, equal_nan=True)
    else:
        fin = np.isfinite(o.end_ray_vec).all(axis=1)
        assert np.array_equal(fin, np.isfinite(g.end_ray_vec).all(axis=1))
        g.end_ray_vec[~fin] = 0.0
        o.end_ray_vec[~fin] = 0.0
        assert vec_rel_err(g.end_ray_vec[:, 0:3], o.end_ray_vec[:, 0:3]) <= tol
        assert vec_rel_err(g.end_ray_vec[:, 3:6], o.end_ray_vec[:, 3:6]) <= tol
        if nv > 7 and cfg.damping_model:
            assert np.max(np.abs(g.end_ray_vec[:, 7] - o.end_ray_vec[:, 7])) <= tol
        for i in range(g.nray):
            n = int(o.npoints[i])
            assert vec_rel_err(g.ray_vec[i, :n, 0:3], o.ray_vec[i, :n, 0:3]) <= tol
            assert vec_rel_err(g.ray_vec[i, :n, 3:6], o.ray_vec[i, :n, 3:6]) <= tol
    if bitwise:
        assert np.array_equal(g.start_ray_vec, o.start_ray_vec, equal_nan=True)
    else:   # the gradient slots start from B, n_e, T_e at the launch point (libm in the profiles)
        assert np.allclose(g.start_ray_vec, o.start_ray_vec, rtol=1e-13, atol=0, equal_nan=True)
    assert np.array_equal(g.initial_ray_power, o.initial_ray_power)


def _run_both(cfg, r, n, w):
    g = rb.trace(cfg, r, n, w)
    o, st, _ = orc.trace(cfg, r, n, w)
    assert st == 0
    return g, o


# ---- config 1: slab (shipped as SG; RK4 as BASELINE.json names it) ------------------------------------
@pytest.mark.parametrize("deriv", ["cold", "numerical"])
def test_slab_rk4_bitwise(deriv):
    cfg = init_case("slab_ECH_90GHz_case_1.in", ode_solver_name="RK4_ODE", ray_deriv_name=deriv)
    r, n, w, _, _ = oracle_fan(cfg, n_kz_launch=16, delta_rindex_z0=0.02, n_ky_launch=4, delta_rindex_y0=0.05)
    assert r.shape[0] > 32
    g, o = _run_both(cfg, r, n, w)
    _compare_traces(g, o, cfg, 1e-10, bitwise=True)


def test_slab_sg_as_shipped():
    cfg = init_case("slab_ECH_90GHz_case_1.in")
    r, n, w, _, _ = oracle_fan(cfg)
    g, o = _run_both(cfg, r, n, w)
    # SG's only libm call is pow() in the step-size formula (ode_RAYS.f90:1221): last-bit differences
    # there move the trajectory by rounding error, far inside the run's tolerance (1e-4)
    _compare_traces(g, o, cfg, float(cfg.rel_err0), bitwise=False)
    assert vec_rel_err(g.end_ray_vec[:, 0:3], o.end_ray_vec[:, 0:3]) <= 1e-9


# ---- config 2: Solov'ev -------------------------------------------------------------------------------
@pytest.mark.parametrize("deriv", ["cold", "numerical"])
def test_solovev_rk4_bitwise(deriv):
    cfg = init_case("solovev_ECH_90GHz_plus_root.in", ode_solver_name="RK4_ODE", ray_deriv_name=deriv, nstep_max=300, ds=5e-11)
    r, n, w, _, _ = oracle_fan(cfg, n_r_launch=2, dr_launch=-0.02, n_theta_launch=4, theta_launch0=-0.3, dtheta_launch=0.2,
                               n_rindex_theta=4, rindex_theta0=-0.2, delta_rindex_theta=0.1, n_rindex_phi=8, rindex_phi0=0.05,
                               delta_rindex_phi=0.05)
    assert r.shape[0] > 64
    g, o = _run_both(cfg, r, n, w)
    _compare_traces(g, o, cfg, 1e-10, bitwise=True)
    assert len(set(o.ray_stop_flag)) >= 2   # the fan exercises more than one termination reason


def test_solovev_sg_as_shipped():
    cfg = init_case("solovev_ECH_90GHz_plus_root.in")
    r, n, w, _, _ = oracle_fan(cfg)
    g, o = _run_both(cfg, r, n, w)
    _compare_traces(g, o, cfg, 1e-9 * 100, bitwise=False)


def test_solovev_sg_fan_within_tolerance():
    cfg = init_case("solovev_ECH_90GHz_plus_root.in", nstep_max=200, ds=5e-11, rel_err0=1e-6, abs_err0=1e-6, SG_error_limit=0.1)
    r, n, w, _, _ = oracle_fan(cfg, n_theta_launch=4, theta_launch0=-0.3, dtheta_launch=0.2, n_rindex_theta=4, rindex_theta0=-0.2,
                               delta_rindex_theta=0.1, n_rindex_phi=8, rindex_phi0=0.05, delta_rindex_phi=0.05)
    g, o = _run_both(cfg, r, n, w)
    _compare_traces(g, o, cfg, 1e-6, bitwise=False)


# ---- config 3: MPEX mirror (bicubic tables; tanh/cosh/pow in the profiles -> libm-level agreement) ------
def test_mpex_rk4():
    cfg = init_case("mpex/rays.in")
    r, n, w, _, _ = oracle_fan(cfg)
    assert r.shape[0] == 11
    g, o = _run_both(cfg, r, n, w)
    _compare_traces(g, o, cfg, 1e-10, bitwise=False)


def test_mpex_sg():
    cfg = init_case("mpex/rays.in", ode_solver_name="SG_ODE", rel_err0=1e-7, abs_err0=1e-7, nstep_max=100, ds=5e-12)
    r, n, w, _, _ = oracle_fan(cfg)
    g, o = _run_both(cfg, r, n, w)
    _compare_traces(g, o, cfg, 1e-7, bitwise=False)


# ---- config 5 (small): axisym_toroid + damping + deposition ----------------------------------------------
def test_axisym_damping_rk4_and_deposition():
    cfg = init_case("axisym_deposition_fan.in", nstep_max=400)
    r, n, w, _, _ = oracle_fan(cfg, n_rindex_theta=8, delta_rindex_theta=0.05, n_rindex_phi=8, delta_rindex_phi=0.04)
    assert r.shape[0] > 16
    g, o = _run_both(cfg, r, n, w)
    # exp() in the Z function is CUDA's: k_i agrees to rounding, not bitwise
    _compare_traces(g, o, cfg, 1e-10, bitwise=False)
    assert np.max(o.end_ray_vec[:, 7]) > 0.01, "the fan must actually deposit power"
    # deposition on stored trajectories and fused into the trace, against the oracle's ray-ordered sum
    po, qo = orc.deposition(cfg, o, 101, 0.0, 1.0)
    rb.set_config(cfg)
    rb.fan_upload(r, n, w)
    rb.trace_device(store=True)
    pg, qg = rb.deposition(101, 0.0, 1.0)
    assert np.max(np.abs(pg - po)) <= 1e-9 * np.max(np.abs(po))
    assert abs(qg - qo) <= 1e-9 * abs(qo)
    rb.trace_device(store=False, bins=(101, 0.0, 1.0))
    pf, qf = rb.deposition(101, 0.0, 1.0)
    assert np.max(np.abs(pf - po)) <= 1e-9 * np.max(np.abs(po))
    # binner sum rule: sum(profile) == sum over rays of P_end * power (bin_to_uniform_grid test invariant)
    # (the last SAVED point counts: a step that trips 'total_absorption' is not stored)
    p_last = np.array([o.ray_vec[i, o.npoints[i] - 1, 7] for i in range(o.nray)])
    assert abs(qf - float(np.sum(p_last * o.initial_ray_power))) <= 1e-9 * abs(qo)


# ---- tabulated (splined) density / temperature profiles in axisym_toroid ---------------
@pytest.mark.parametrize("deriv", ["cold", "numerical"])
def test_axisym_spline_profiles_rk4_bitwise(deriv, tmp_path):
    # no damping: nothing on the path calls libm, the comparison is bitwise (nv = 12 with the gradient slots)
    cfg = init_case_text("axisym_spline_profiles.in", [("damping_model='damp_fund_ECH'", "damping_model='no_damp'")], tmp_path,
                         ray_deriv_name=deriv, nstep_max=300)
    assert cfg.nv == 12
    r, n, w, _, _ = oracle_fan(cfg)
    assert r.shape[0] > 128
    g, o = _run_both(cfg, r, n, w)
    _compare_traces(g, o, cfg, 1e-10, bitwise=True)


def test_axisym_spline_profiles_damping_and_sg():
    cfg = init_case("axisym_spline_profiles.in", nstep_max=400)
    r, n, w, _, _ = oracle_fan(cfg)
    g, o = _run_both(cfg, r, n, w)
    _compare_traces(g, o, cfg, 1e-10, bitwise=False)
    assert np.max(o.end_ray_vec[:, 7]) > 0.5
    cfg = init_case("axisym_spline_profiles.in", ode_solver_name="SG_ODE", nstep_max=200, rel_err0=1e-7, abs_err0=1e-7)
    g, o = _run_both(cfg, r[::4], n[::4], w[::4])
    _compare_traces(g, o, cfg, 1e-7, bitwise=False)


# ---- one-point probes -------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,box", [
    ("slab_ECH_90GHz_case_1.in", ((-0.6, 0.6), (-0.6, 0.6), (-1.1, 1.1))),
    ("solovev_ECH_90GHz_plus_root.in", ((0.5, 1.5), (-0.3, 0.3), (-0.75, 0.75))),
    ("axisym_deposition_fan.in", ((0.5, 1.5), (-0.3, 0.3), (-0.75, 0.75))),
    ("axisym_spline_profiles.in", ((0.5, 1.5), (-0.3, 0.3), (-0.75, 0.75))),
    ("mpex/rays.in", ((-0.15, 0.15), (-0.15, 0.15), (2.75, 3.65))),
])
def test_probe_equilibrium(case, box):
    cfg = init_case(case)
    rng = np.random.default_rng(20260101)
    pts = np.stack([rng.uniform(lo, hi, 4096) for lo, hi in box], axis=1)
    rb.set_config(cfg)
    g, ge = rb.probe_equilibrium(pts)
    o, oe = orc.probe_equilibrium(cfg, pts)
    assert np.array_equal(ge, oe)
    ok = oe == 0
    assert ok.sum() > 100 and (~ok).sum() > 10
    # ion temperature gradients are not evaluated on the device (nothing on the path reads them)
    cols = np.ones(o.shape[1], dtype=bool)
    cols[3 + 9 + 6 + 18 + 6 + 3:3 + 9 + 6 + 18 + 6 + 18] = False
    if case.startswith("mpex"):
        assert np.allclose(g[ok][:, cols], o[ok][:, cols], rtol=1e-12, atol=0)
    else:
        assert np.array_equal(g[ok][:, cols], o[ok][:, cols])


@pytest.mark.parametrize("deriv", ["cold", "numerical"])
def test_probe_rhs_and_check_save(deriv):
    cfg = init_case("solovev_ECH_90GHz_plus_root.in", ode_solver_name="RK4_ODE", ray_deriv_name=deriv, nstep_max=100, ds=5e-11)
    r, n, w, _, _ = oracle_fan(cfg, n_rindex_phi=8, rindex_phi0=0.05, delta_rindex_phi=0.05)
    o, _, _ = orc.trace(cfg, r, n, w)
    v = np.concatenate([o.ray_vec[i, :o.npoints[i]] for i in range(o.nray)])
    rb.set_config(cfg)
    gd, gs = rb.probe_rhs(v)
    od, os_ = orc.probe_rhs(cfg, v)
    assert np.array_equal(gs, os_) and np.array_equal(gd, od)
    gr, gc = rb.probe_check_save(v)
    orr, oc = orc.probe_check_save(cfg, v)
    assert np.array_equal(gc, oc) and np.array_equal(gr, orr)


def test_exact_arithmetic_helpers_on_device():
    """rcp_rn / sqrt_rn / qdiv (ray_physics.cuh) == IEEE 1.0/d, sqrt(x), x/d on 2^30 operand pairs."""
    import ctypes as C
    from rays_b200 import _abi
    L = _abi.load()
    mm = (C.c_int64 * 4)()
    assert L.rays_b200_selftest_arith(1 << 30, 20260101, mm) == 0
    assert list(mm)[:3] == [0, 0, 0], list(mm)
    assert mm[3] < 1000      # all-ones-significand divisors (forced by the adversarial modes): the documented exception


# ---- generic kernels: species counts other than electron + one ion, per-species damping slots -----------
THREE_SPECIES = [(" spec_name(1) = 'deuterium',", " spec_name(1) = 'deuterium',\n spec_name(2) = 'tritium',\n spec_model(2) = 'cold',\n t0s_eV(2)=1.0e2 ,\n eta(2)=0.5"),
                 (" eta(1)=1.", " eta(1)=0.5"), ("t_prof_model=2*'zero'", "t_prof_model=3*'zero'")]
ELECTRONS_ONLY = [(" eta(1)=1.", " eta(1)=0."), (" neutrality", " neutrality")]


@pytest.mark.parametrize("deriv", ["cold", "numerical"])
def test_three_species_generic_kernel_bitwise(deriv, tmp_path):
    cfg = init_case_text("solovev_ECH_90GHz_plus_root.in", THREE_SPECIES, tmp_path, ode_solver_name="RK4_ODE", ray_deriv_name=deriv,
                         nstep_max=200, ds=5e-11)
    assert cfg.nspec == 2
    r, n, w, _, _ = oracle_fan(cfg, n_theta_launch=3, theta_launch0=-0.3, dtheta_launch=0.3, n_rindex_theta=3, rindex_theta0=-0.2,
                               delta_rindex_theta=0.2, n_rindex_phi=8, rindex_phi0=0.05, delta_rindex_phi=0.05)
    assert r.shape[0] > 32
    g, o = _run_both(cfg, r, n, w)
    assert "generic" in rb.last_trace_stats()["kernel"]
    _compare_traces(g, o, cfg, 1e-10, bitwise=True)


def test_three_species_generic_sg(tmp_path):
    cfg = init_case_text("solovev_ECH_90GHz_plus_root.in", THREE_SPECIES, tmp_path, nstep_max=100, ds=5e-11, rel_err0=1e-6, abs_err0=1e-6,
                         SG_error_limit=0.1)
    r, n, w, _, _ = oracle_fan(cfg, n_rindex_phi=8, rindex_phi0=0.05, delta_rindex_phi=0.05)
    g, o = _run_both(cfg, r, n, w)
    assert "generic" in rb.last_trace_stats()["kernel"]
    _compare_traces(g, o, cfg, 1e-6, bitwise=False)


def test_multi_species_damping_generic_kernel(tmp_path):
    cfg = init_case_text("axisym_deposition_fan.in", [("multi_spec_damping = .false.", "multi_spec_damping = .true.")], tmp_path, nstep_max=300)
    assert cfg.nv == 10 and cfg.multi_spec_damping == 1
    r, n, w, _, _ = oracle_fan(cfg, n_rindex_theta=4, delta_rindex_theta=0.1, n_rindex_phi=8, delta_rindex_phi=0.04)
    g, o = _run_both(cfg, r, n, w)
    assert "generic" in rb.last_trace_stats()["kernel"]
    _compare_traces(g, o, cfg, 1e-10, bitwise=False)
    fin = np.isfinite(o.end_ray_vec).all(axis=1)
    assert np.allclose(g.end_ray_vec[fin, 8], o.end_ray_vec[fin, 8], rtol=1e-10, atol=1e-14)      # electron slot = total
    assert np.all(o.end_ray_vec[fin, 9] == 0.0) and np.all(g.end_ray_vec[fin, 9] == 0.0)           # ions absorb nothing in this model


def test_launch_fans_on_device_match_the_oracle():
    """row f1: the device launch-fan kernels + order-preserving compaction against the oracle's launchers"""
    from _cases import launch_params
    cfg = init_case("solovev_fan_1M.in")
    _, so, _ = launch_params()
    so.n_r_launch, so.n_theta_launch, so.n_rindex_theta, so.n_rindex_phi = 3, 5, 16, 16
    so.rindex_phi0, so.delta_rindex_phi = 0.05, 0.08       # the upper end of the grid is evanescent: candidates get dropped
    ro, no, wo = orc.launch_fan(cfg, "solovev", so, 4096)
    rb.set_config(cfg)
    k = rb.launch_fan("solovev", so)
    assert k == ro.shape[0] and 0 < k <= 3 * 5 * 256
    rg, ng, wg = rb.fan_download(k)
    assert np.array_equal(rg, ro) and np.array_equal(ng, no) and np.array_equal(wg, wo)
    # slab
    cfg = init_case("slab_ECH_90GHz_case_1.in")
    sl, _, _ = launch_params()
    sl.n_kz_launch, sl.delta_rindex_z0, sl.n_ky_launch, sl.delta_rindex_y0, sl.n_x_launch, sl.dx_launch = 40, 0.03, 5, 0.1, 3, 0.05
    ro, no, wo = orc.launch_fan(cfg, "slab", sl, 4096)
    rb.set_config(cfg)
    k = rb.launch_fan("slab", sl)
    assert k == ro.shape[0] and 0 < k < 600
    rg, ng, wg = rb.fan_download(k)
    assert np.array_equal(rg, ro) and np.array_equal(ng, no) and np.array_equal(wg, wo)
    # axisym
    cfg = init_case("axisym_deposition_fan.in")
    _, _, ax = launch_params()
    ax.n_rindex_theta, ax.n_rindex_phi, ax.delta_rindex_theta, ax.delta_rindex_phi = 24, 24, 0.03, 0.05
    ro, no, wo = orc.launch_fan(cfg, "axisym", ax, 4096)
    rb.set_config(cfg)
    k = rb.launch_fan("axisym", ax)
    assert k == ro.shape[0]
    rg, ng, wg = rb.fan_download(k)
    assert np.array_equal(rg, ro) and np.array_equal(ng, no) and np.array_equal(wg, wo)
    # file_input / n(theta): acos + cos are libm -> rounding-level agreement
    cfg = init_case("mpex/rays.in")
    r11, n11, w11, _, _ = oracle_fan(cfg)
    import ctypes as C
    from rays_b200 import _abi
    L = _abi.load()
    rin, nin = _abi.c_double_p(), _abi.c_double_p()
    kk = L.rays_host_directions_in(C.byref(rin), C.byref(nin))
    rb.set_config(cfg)
    k = rb.launch_fan_directions(np.ctypeslib.as_array(rin, (kk, 3)).copy(), np.ctypeslib.as_array(nin, (kk, 3)).copy())
    assert k == 11
    rg, ng, _ = rb.fan_download(k)
    assert np.array_equal(rg, r11) and np.allclose(ng, n11, rtol=1e-13, atol=0)


def test_host_mirror_program_flow(tmp_path):
    """initialize(read_input) -> trace_rays -> finalize_run through the host mirror, fan built on the device"""
    from scipy.io import netcdf_file
    rb.initialize(rb.config_path("mpex/rays.in"), ray_init=True, device=0)
    rb.trace_rays()
    res = rb.results()
    assert res["nray"] == 11 and all(f.strip() == "nstep > nstep_max" for f in res["ray_stop_flag"])
    cfg = rb.host_cfg()
    r, n, w = rb.get_fan()
    o, st, _ = orc.trace(cfg, r, n, w)
    assert np.array_equal(res["npoints"], o.npoints)
    assert vec_rel_err(res["ray_vec"][:, :, 0:3], o.ray_vec[:, :, 0:3]) <= 1e-10
    rb.finalize_run(str(tmp_path))
    f = netcdf_file(str(tmp_path / "run_results.2nd_harm_11_rays_nx.nc"), "r", mmap=False)
    assert f.variables["ray_vec"].shape == (11, 501, 12)
    assert np.array_equal(np.array(f.variables["ray_vec"].data), res["ray_vec"])
    f.close()


@pytest.mark.parametrize("slice_steps,register,staged", [("5", "1", "1"), ("5", "0", "1"), ("0", "0", "1"), ("40", "1", "1"), ("5", "0", "0"), ("0", "0", "0"),
                                                         ("0", "1", "copier"), ("5", "1", "copier")])
def test_copy_out_paths_and_time_slicing_are_bitwise_identical(slice_steps, register, staged, monkeypatch):
    """The library's execution options change the schedule and the copy-out path, never the results:
    time slicing (rays suspended after n steps and re-launched packed) x streaming copy-out by the kernel (page-locked arrays) /
    the concurrent copier kernel (page-locked arrays, the default; RAYS_B200_COPIER=0 selects the in-kernel copy) / packed rows through the page-locked ring + host threads
    (pageable arrays) / plain cudaMemcpy2D into pageable arrays."""
    monkeypatch.setenv("RAYS_B200_SLICE", slice_steps)
    monkeypatch.setenv("RAYS_B200_REGISTER_HOST", register)
    copier = staged == "copier"
    monkeypatch.setenv("RAYS_B200_COPIER", "1" if copier else "0")
    if copier:
        staged = "1"
    monkeypatch.setenv("RAYS_B200_STAGED_COPY", staged)
    monkeypatch.setenv("RAYS_B200_COPY_THREADS", "5")
    cfg = init_case("solovev_fan_1M.in", nstep_max=120)
    r, n, w, _, _ = oracle_fan(cfg, n_r_launch=2, n_theta_launch=3, n_rindex_theta=8, n_rindex_phi=8, dtheta_launch=0.2,
                               delta_rindex_theta=0.05, delta_rindex_phi=0.04)
    g, o = _run_both(cfg, r, n, w)
    st = rb.last_trace_stats()
    if slice_steps != "0":
        assert st["n_passes"] >= 2
    assert ("copy_out_kernel" in st["kernel"]) == copier
    _compare_traces(g, o, cfg, 1e-10, bitwise=True)
    # and Shampine-Gordon through the same options
    cfg = init_case("solovev_fan_1M.in", ode_solver_name="SG_ODE", ray_deriv_name="cold", nstep_max=60, rel_err0=1e-6, abs_err0=1e-6,
                    SG_error_limit=0.1)
    g, o = _run_both(cfg, r[:96], n[:96], w[:96])
    _compare_traces(g, o, cfg, 1e-6, bitwise=False)


# ---- parity at scale (VERDICT r1, next #1b): >= 64k-ray slices of the bench fans ---------------------------------------------
def test_config4_slice_65k_rays_bitwise():
    """every 16th ray of the 1 048 576-ray bench fan (RK4, deriv_num, nstep_max = 1000): 65 536 rays, bitwise against the oracle"""
    cfg = init_case("solovev_fan_1M.in")
    r, n, w, _, _ = oracle_fan(cfg, cap=1 << 21)
    assert r.shape[0] == 1048576
    idx = np.arange(0, r.shape[0], 16)
    g = rb.trace(cfg, r[idx], n[idx], w[idx])
    o, st, _ = orc.trace(cfg, r[idx], n[idx], w[idx], nthreads=0)
    assert st == 0
    assert np.array_equal(g.npoints, o.npoints) and np.array_equal(g.ray_stop_code, o.ray_stop_code)
    assert np.array_equal(g.ray_vec, o.ray_vec, equal_nan=True)
    assert np.array_equal(g.residual, o.residual, equal_nan=True)
    assert np.array_equal(g.end_ray_vec, o.end_ray_vec, equal_nan=True)


def test_sg_slot_machine_equals_the_per_lane_kernel_and_the_oracle(monkeypatch):
    """16 384 rays of the bench fan with Shampine-Gordon: the slot-machine kernel, the round-1 per-lane kernel and the oracle take the
    same steps to the same stop reasons; the two device kernels run the same arithmetic and agree bitwise"""
    cfg = init_case("solovev_fan_1M.in", ode_solver_name="SG_ODE", ray_deriv_name="cold", rel_err0=1e-6, abs_err0=1e-6, SG_error_limit=0.1)
    r, n, w, _, _ = oracle_fan(cfg, cap=1 << 21)
    idx = np.arange(0, r.shape[0], 64)
    g = rb.trace(cfg, r[idx], n[idx], w[idx])
    assert "trace" in rb.last_trace_stats()["kernel"]
    monkeypatch.setenv("RAYS_B200_SG_LANES", "1")
    g1 = rb.trace(cfg, r[idx], n[idx], w[idx])
    monkeypatch.delenv("RAYS_B200_SG_LANES")
    rb.set_config(cfg)
    assert np.array_equal(g.npoints, g1.npoints) and np.array_equal(g.ray_stop_code, g1.ray_stop_code)
    assert np.array_equal(g.ray_vec, g1.ray_vec, equal_nan=True) and np.array_equal(g.residual, g1.residual, equal_nan=True)
    # the scheduling policy of the slot machine (which kinds of macro-step share an iteration) changes the schedule, never a bit
    for policy in ("0", "1", "2"):
        monkeypatch.setenv("RAYS_B200_SG_MIXED", policy)
        gp = rb.trace(cfg, r[idx], n[idx], w[idx])
        assert np.array_equal(g.npoints, gp.npoints) and np.array_equal(g.ray_stop_code, gp.ray_stop_code), policy
        assert np.array_equal(g.ray_vec, gp.ray_vec, equal_nan=True) and np.array_equal(g.residual, gp.residual, equal_nan=True), policy
    monkeypatch.delenv("RAYS_B200_SG_MIXED")
    o, st, nrhs = orc.trace(cfg, r[idx], n[idx], w[idx], nthreads=0)
    assert np.array_equal(g.npoints, o.npoints), np.nonzero(g.npoints != o.npoints)[0][:10]
    assert np.array_equal(g.ray_stop_code, o.ray_stop_code)
    g = rb.trace(cfg, r[idx], n[idx], w[idx])
    assert abs(rb.last_trace_stats()["rhs_evals"] - nrhs) <= 1e-4 * nrhs, "the slot machine evaluates the right-hand sides the reference does"
    fin = np.isfinite(o.end_ray_vec).all(axis=1)
    assert np.max(np.abs(g.end_ray_vec[fin, :3] - o.end_ray_vec[fin, :3])) <= 1e-6


def test_config5_sharded_profile_is_bitwise_independent_of_the_sharding():
    """a 65 536-ray slice of the config-5 deposition fan: the fixed-point profile summed over 1, 2, 3 and 8 interleaved shards is the
    same int64 vector, and it equals the oracle's ray-ordered double sum to 1e-12 of the largest bin"""
    cfg = init_case("axisym_deposition_fan.in", nstep_max=400)
    r, n, w, _, _ = oracle_fan(cfg, n_rindex_theta=256, delta_rindex_theta=0.4 / 255, n_rindex_phi=256, delta_rindex_phi=0.35 / 255, cap=70000)
    assert r.shape[0] >= 60000
    rb.set_config(cfg)
    nb = 501
    accs = {}
    for world in (1, 2, 3, 8):
        tot = np.zeros(nb, dtype=np.int64)
        for rank in range(world):
            rb.fan_upload(r, n, w)
            rb.fan_shard(rank, world)
            rb.trace_device(store=False, bins=(nb, 0.0, 1.0))
            acc, unit, prof, q = rb.deposition_fixed(nb, 0.0, 1.0)
            assert np.array_equal(prof, acc.astype(np.float64) * unit)
            tot += acc
        accs[world] = tot
    for world in (2, 3, 8):
        assert np.array_equal(accs[world], accs[1]), f"profile of {world} shards differs from the unsharded one"
    # twice the same trace: the same bits (no order dependence inside one GPU either)
    rb.fan_upload(r, n, w)
    rb.trace_device(store=False, bins=(nb, 0.0, 1.0))
    acc2, unit, _, _ = rb.deposition_fixed(nb, 0.0, 1.0)
    assert np.array_equal(acc2, accs[1])
    # against the oracle (doubles summed in ray order, deposition_profiles_m.f90:244-257) on a 4096-ray sub-slice
    idx = np.arange(0, r.shape[0], 16)
    o, st, _ = orc.trace(cfg, r[idx], n[idx], w[idx], nthreads=0)
    po, qo = orc.deposition(cfg, o, nb, 0.0, 1.0)
    rb.fan_upload(r[idx], n[idx], w[idx])
    rb.deposition_set_total_weight(float(np.sum(np.abs(w))))   # the unit of the whole fan, as a host with pre-sharded fans would set it
    rb.trace_device(store=False, bins=(nb, 0.0, 1.0))
    pg, qg = rb.deposition(nb, 0.0, 1.0)
    assert np.max(np.abs(pg - po)) <= 1e-10 * np.max(np.abs(po))   # exp() of the Z function is CUDA's: rounding level, not bitwise
    assert abs(qg - qo) <= 1e-10 * abs(qo)


def test_summaries_pack_matches_the_downloaded_summaries():
    import torch
    cfg = init_case("axisym_deposition_fan.in", nstep_max=200)
    r, n, w, _, _ = oracle_fan(cfg, n_rindex_theta=16, delta_rindex_theta=0.025, n_rindex_phi=16, delta_rindex_phi=0.02)
    rb.set_config(cfg)
    rb.fan_upload(r, n, w)
    rb.trace_device(store=False)
    nv = int(cfg.nv)
    rows, rd = rb.summaries_pack()
    assert rows == r.shape[0] and rd == 6 + 2 * nv
    t = torch.zeros((rows, rd), dtype=torch.float64, device="cuda")
    rb.summaries_pack(t.data_ptr(), rows)
    res = rb.results_download(rows, nv, int(cfg.nstep_max) + 1, store=False)
    h = t.cpu().numpy()
    assert np.array_equal(h[:, 0], res.npoints) and np.array_equal(h[:, 1], res.ray_stop_code)
    assert np.array_equal(h[:, 2], res.initial_ray_power) and np.array_equal(h[:, 5], res.end_ray_parameter, equal_nan=True)
    assert np.array_equal(h[:, 6:6 + nv], res.start_ray_vec, equal_nan=True) and np.array_equal(h[:, 6 + nv:], res.end_ray_vec, equal_nan=True)


def test_trace_multi_one_process_all_gpus():
    """rays_b200_init_multi / rays_b200_trace_multi (one host process drives every visible GPU, rays iray % ngpu): the caller's arrays
    receive exactly what the single-GPU call delivers; the binned variant's NCCL-reduced profile equals the single-GPU fixed-point one"""
    ngpu = rb.init_multi(0)
    assert ngpu >= 1
    cfg = init_case("solovev_fan_1M.in", nstep_max=150)
    r, n, w, _, _ = oracle_fan(cfg, n_r_launch=2, n_theta_launch=4, n_rindex_theta=16, n_rindex_phi=16, dtheta_launch=0.2,
                               delta_rindex_theta=0.025, delta_rindex_phi=0.02)
    g1 = rb.trace(cfg, r, n, w)
    gm = rb.trace_multi(cfg, r, n, w)
    assert np.array_equal(gm.npoints, g1.npoints) and np.array_equal(gm.ray_stop_code, g1.ray_stop_code) and gm.ray_stop_flag == g1.ray_stop_flag
    assert np.array_equal(gm.ray_vec, g1.ray_vec, equal_nan=True) and np.array_equal(gm.residual, g1.residual, equal_nan=True)
    assert np.array_equal(gm.end_ray_vec, g1.end_ray_vec, equal_nan=True) and np.array_equal(gm.start_ray_vec, g1.start_ray_vec, equal_nan=True)
    assert gm.total_ray_steps == g1.total_ray_steps
    # config 5 in small: fused binning on every GPU + ncclReduce (int64) + ncclAllGather of the summaries
    cfg = init_case("axisym_deposition_fan.in", nstep_max=300)
    r, n, w, _, _ = oracle_fan(cfg, n_rindex_theta=48, delta_rindex_theta=0.4 / 47, n_rindex_phi=48, delta_rindex_phi=0.35 / 47)
    res, prof, q = rb.trace_multi(cfg, r, n, w, bins=(501, 0.0, 1.0))
    rb.set_config(cfg)
    rb.fan_upload(r, n, w)
    rb.trace_device(store=False, bins=(501, 0.0, 1.0))
    acc, unit, p1, q1 = rb.deposition_fixed(501, 0.0, 1.0)
    assert np.array_equal(prof, p1) and q == q1, "the reduced profile does not depend on the number of GPUs"
    s1 = rb.results_download(r.shape[0], int(cfg.nv), int(cfg.nstep_max) + 1, store=False)
    assert np.array_equal(res.npoints, s1.npoints) and np.array_equal(res.end_ray_vec, s1.end_ray_vec, equal_nan=True)
