"""axisym_toroid with eqdsk_magnetics_spline_interp (psi(R,Z) bicubic + R*Bphi cubic from a g-file).
The fixture rays_b200/configs/eqdsk/solovev.geqdsk is the Solov'ev equilibrium of axisym_deposition_fan.in written
as a g-file the way the reference's solovev_2_eqdsk does (which, by the two models' sign conventions, is the field
of solovev_magnetics with iota0 -> -iota0), so the eqdsk run must reproduce that solovev_magnetics run to spline accuracy (the reference's own cross-check, solovev_2_eqdsk/compare_analyt_2_interp.f90)."""
import numpy as np
import pytest

import rays_b200 as rb
from rays_b200 import _abi
import _oracle as orc
from _cases import init_case, init_case_text, oracle_fan, vec_rel_err


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    yield


def test_gfile_reader_and_geometry():
    """ReadgFile + initialize_eqdsk_magnetics_spline_interp (eqdsk_utilities_m.f90:52-108,
    eqdsk_magnetics_spline_interp_m.f90:120-185)"""
    cfg = init_case("eqdsk/rays.in")
    a = cfg.axisym
    assert a.magnetics_model == _abi.MAG_EQDSK_SPLINE
    assert (a.Psi_spline.nx, a.Psi_spline.ny, a.T_spline.nx) == (65, 65, 65)
    assert (a.r_axis, a.z_axis) == (1.0, 0.0)
    assert a.box_rmin == 0.1 and a.box_rmax == 0.1 + 1.4 and a.box_zmin == -0.7 and a.box_zmax == 0.7
    assert abs(a.inner_bound - np.sqrt(2.0 - 1.4 ** 2)) < 1e-9 and a.outer_bound == 1.4
    assert a.upper_bound == -a.lower_bound and 0.65 < a.upper_bound < 0.67
    assert a.eq_psibound == 0.0038016                       # PSIBOUND - PSIAXIS, 9 digits as the file carries them
    rg = np.ctypeslib.as_array(a.Psi_spline.x_grid, (65,))
    zg = np.ctypeslib.as_array(a.Psi_spline.y_grid, (65,))
    assert np.array_equal(rg, np.array([0.1 + (a.box_rmax - 0.1) * i / 64 for i in range(65)]))
    fs = np.ctypeslib.as_array(a.Psi_spline.fspl, (65, 65, 4, 4))      # [j][i][cy][cx]
    bp0 = 3.3 * 0.01
    psi = 0.5 * bp0 * ((rg[None, :] * zg[:, None] / 1.1) ** 2 + (rg[None, :] ** 2 - 1.0) ** 2 / 4.0)
    assert np.all(np.abs(fs[:, :, 0, 0] - psi) <= 5.1e-9 * np.abs(psi))    # e16.9: nine significant digits
    ts = np.ctypeslib.as_array(a.T_spline.fspl, (65, 4))
    assert np.all(ts[:, 0] == 3.3) and np.max(np.abs(ts[:, 1:])) < 1e-12       # R*Bphi = bphi0*rmaj, flat


def test_gfile_errors(tmp_path):
    L = _abi.load()
    txt = open(rb.config_path("eqdsk/rays.in")).read()
    p = tmp_path / "rays.in"
    p.write_text(txt)                                          # the g-file is not beside this copy
    assert L.rays_host_initialize(str(p).encode(), 0) != 0 and b"ReadgFile" in L.rays_host_last_error()
    g = open(rb.config_path("eqdsk/solovev.geqdsk")).read().split("\n")
    (tmp_path / "solovev.geqdsk").write_text("\n".join(g[:300]))   # truncated
    assert L.rays_host_initialize(str(p).encode(), 0) != 0 and b"truncated" in L.rays_host_last_error()


def test_oracle_eqdsk_reproduces_solovev_magnetics(tmp_path):
    rng = np.random.default_rng(1)
    pts = np.stack([rng.uniform(0.65, 1.42, 2000), rng.uniform(-0.3, 0.3, 2000), rng.uniform(-0.5, 0.5, 2000)], axis=1)
    cfg = init_case("eqdsk/rays.in", nstep_max=400)
    e1, err1 = orc.probe_equilibrium(cfg, pts)
    r, n, w, _, _ = oracle_fan(cfg)
    o1, st, _ = orc.trace(cfg, r, n, w)
    assert st == 0 and r.shape[0] == 256
    p1, q1 = orc.deposition(cfg, o1, 101, 0.0, 1.0)
    assert np.max(o1.npoints) > 90 and np.max(o1.end_ray_vec[:, 7]) > 0.5      # rays cross the plasma and are absorbed
    cfg = init_case_text("axisym_deposition_fan.in", [("iota0=.01", "iota0=-.01")], tmp_path, nstep_max=400)
    e2, err2 = orc.probe_equilibrium(cfg, pts)
    o2, st, _ = orc.trace(cfg, r, n, w)
    p2, q2 = orc.deposition(cfg, o2, 101, 0.0, 1.0)
    assert np.array_equal(err1, err2) and (err1 == 0).sum() > 1000 and (err1 != 0).sum() > 100
    ok = err1 == 0
    scale = np.max(np.abs(e2[ok]), axis=0)
    scale[scale == 0] = 1.0
    d = np.max(np.abs(e1[ok] - e2[ok]), axis=0) / scale
    assert np.max(d[:3]) < 2e-6            # B
    assert np.max(d) < 5e-4                # grad B needs second derivatives of a spline through 10-digit data
    assert np.array_equal(o1.npoints, o2.npoints) and o1.ray_stop_flag == o2.ray_stop_flag
    # a 65 x 65 spline through 9-digit data: the two runs agree to ~1e-6
    assert vec_rel_err(o1.end_ray_vec[:, 0:3], o2.end_ray_vec[:, 0:3]) < 2e-6
    assert vec_rel_err(o1.end_ray_vec[:, 3:6], o2.end_ray_vec[:, 3:6]) < 5e-6
    assert np.max(np.abs(o1.end_ray_vec[:, 7] - o2.end_ray_vec[:, 7])) < 1e-4
    assert abs(q1 - q2) < 1e-5 * abs(q2) and np.max(np.abs(p1 - p2)) < 2e-3 * np.max(p2)


# ---- CUDA path --------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("deriv", ["cold", "numerical"])
def test_gpu_eqdsk_rk4_bitwise(deriv, tmp_path):
    import shutil
    from test_gpu_parity import _compare_traces, _run_both
    rb.init(0)
    shutil.copy(rb.config_path("eqdsk/solovev.geqdsk"), tmp_path / "solovev.geqdsk")
    cfg = init_case_text("eqdsk/rays.in", [("damping_model='damp_fund_ECH'", "damping_model='no_damp'"),
                                           ("integrate_eq_gradients=.false.", "integrate_eq_gradients=.true.")], tmp_path,
                         ray_deriv_name=deriv, nstep_max=300)
    assert cfg.nv == 12
    r, n, w, _, _ = oracle_fan(cfg)
    g, o = _run_both(cfg, r, n, w)
    _compare_traces(g, o, cfg, 1e-10, bitwise=True)      # splines + exact division: no libm on this path


@pytest.mark.gpu
def test_gpu_eqdsk_damping_deposition_and_probe():
    from test_gpu_parity import _compare_traces, _run_both
    rb.init(0)
    cfg = init_case("eqdsk/rays.in", nstep_max=400)
    r, n, w, _, _ = oracle_fan(cfg)
    g, o = _run_both(cfg, r, n, w)
    _compare_traces(g, o, cfg, 1e-10, bitwise=False)
    assert np.max(o.end_ray_vec[:, 7]) > 0.01
    po, qo = orc.deposition(cfg, o, 101, 0.0, 1.0)
    rb.set_config(cfg)
    rb.fan_upload(r, n, w)
    rb.trace_device(store=False, bins=(101, 0.0, 1.0))      # psi_N of the binning comes from the bicubic spline
    pf, qf = rb.deposition(101, 0.0, 1.0)
    assert np.max(np.abs(pf - po)) <= 1e-9 * np.max(np.abs(po)) and abs(qf - qo) <= 1e-9 * abs(qo)
    rng = np.random.default_rng(20260101)
    pts = np.stack([rng.uniform(0.05, 1.55, 4096), rng.uniform(-0.3, 0.3, 4096), rng.uniform(-0.75, 0.75, 4096)], axis=1)
    ge, gerr = rb.probe_equilibrium(pts)
    oe, oerr = orc.probe_equilibrium(cfg, pts)
    assert np.array_equal(gerr, oerr)
    ok = oerr == 0
    cols = np.ones(oe.shape[1], dtype=bool)
    cols[3 + 9 + 6 + 18 + 6 + 3:3 + 9 + 6 + 18 + 6 + 18] = False      # ion temperature gradients: not evaluated on the device
    assert ok.sum() > 100 and np.array_equal(ge[ok][:, cols], oe[ok][:, cols])
