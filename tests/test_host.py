"""Host-side mirror of the reference's Fortran host: namelist reader, module initialisation, spline setup,
netCDF-3 reader/writer, result file contract (no GPU needed)."""
import ctypes as C
import os

import numpy as np
import pytest

import rays_b200 as rb
from rays_b200 import _abi
from _cases import init_case


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    yield


def test_constants_carry_single_precision_literals():
    """pi, clight, e, me are float32 literals widened to double."""
    cfg = init_case("slab_ECH_90GHz_case_1.in")
    pi32 = float(np.float32(3.1415926535897932385))
    clight = float(np.float32(2.997930e8))
    assert cfg.clight == clight
    assert cfg.omgrf == 2.0 * pi32 * 90.0e9
    assert cfg.k0 == cfg.omgrf / clight
    me, e = float(np.float32(9.1094e-31)), float(np.float32(1.6022e-19))
    assert cfg.ms[0] == me and cfg.ms[1] == me * 3670.0 and cfg.qs[0] == -e
    mu0 = pi32 * float(np.float32(4.e-7))
    assert cfg.eps0 == 1.0 / (mu0 * (clight * clight))
    assert cfg.total_damping_limit == 0.99            # namelist value parsed as a double


def test_namelists_of_all_configs_parse():
    nv = {"slab_ECH_90GHz_case_1.in": 7, "solovev_ECH_90GHz_plus_root.in": 12, "mpex/rays.in": 12, "solovev_fan_1M.in": 7,
          "axisym_deposition_fan.in": 8, "axisym_spline_profiles.in": 13}
    for name, want in nv.items():
        cfg = init_case(name)
        assert cfg.nv == want and cfg.nspec == 1
    cfg = init_case("mpex/rays.in")
    m = cfg.mirror
    assert (m.Br_spline.nx, m.Br_spline.ny) == (51, 201)
    assert abs(m.r_LUFS - 0.12) < 1e-12 and abs(m.z_LUFS - 3.6) < 1e-12 and m.Aphi_LUFS != 0.0
    assert cfg.ode_solver == _abi.ODE_RK4 and cfg.ds == 1e-12 and cfg.nstep_max == 500


def test_unknown_namelist_variable_is_an_error(tmp_path):
    txt = open(rb.config_path("slab_ECH_90GHz_case_1.in")).read().replace("frf=90.e9,", "frf=90.e9,\n  message_unit=11,")
    p = tmp_path / "rays.in"
    p.write_text(txt)
    L = _abi.load()
    assert L.rays_host_initialize(str(p).encode(), 0) != 0
    assert b"message_unit" in L.rays_host_last_error().lower()


def test_invalid_model_strings_fail_like_the_reference(tmp_path):
    L = _abi.load()
    for old, new in (("ode_solver_name='SG_ODE'", "ode_solver_name='EULER'"), ("equilib_model='slab'", "equilib_model='torus'"),
                     ("dens_prof_model='linear'", "dens_prof_model='cubic'")):
        txt = open(rb.config_path("slab_ECH_90GHz_case_1.in")).read()
        assert old in txt
        p = tmp_path / "bad.in"
        p.write_text(txt.replace(old, new))
        assert L.rays_host_initialize(str(p).encode(), 0) != 0


def test_cubic_and_bicubic_spline_setup_interpolate():
    """cspline / bcspline (not-a-knot): reproduce smooth functions to the accuracy test_pspline.f90 reports."""
    L = _abi.load()
    nx = 101
    x = np.linspace(0.0, 2.0, nx)
    f = np.zeros((nx, 4))
    f[:, 0] = 1.0 + np.cos(10.0 * x)
    assert L.rays_host_cspline(x.ctypes.data_as(_abi.c_double_p), nx, f.ctypes.data_as(_abi.c_double_p)) == 0
    import _oracle as orc
    s = _abi.Spline1D()
    s.nx, s.x_grid, s.fspl = nx, x.ctypes.data_as(_abi.c_double_p), f.ctypes.data_as(_abi.c_double_p)
    xs = np.linspace(0.0, 2.0, 1777)
    fo, fpo = np.zeros_like(xs), np.zeros_like(xs)
    orc.load().oracle_cspeval(C.byref(s), len(xs), xs.ctypes.data_as(_abi.c_double_p), fo.ctypes.data_as(_abi.c_double_p),
                              fpo.ctypes.data_as(_abi.c_double_p))
    assert np.max(np.abs(fo - (1.0 + np.cos(10.0 * xs)))) < 1e-4
    assert np.max(np.abs(fpo + 10.0 * np.sin(10.0 * xs))) < 5e-2
    # 2-D
    nx, ny = 41, 61
    xg, yg = np.linspace(0.1, 1.0, nx), np.linspace(-1.0, 1.0, ny)
    F = np.zeros((ny, nx, 4, 4))
    F[:, :, 0, 0] = np.sqrt(xg[None, :] ** 2 + yg[:, None] ** 2)
    assert L.rays_host_bcspline(xg.ctypes.data_as(_abi.c_double_p), nx, yg.ctypes.data_as(_abi.c_double_p), ny,
                                F.ctypes.data_as(_abi.c_double_p)) == 0
    s2 = _abi.Spline2D()
    s2.nx, s2.ny = nx, ny
    s2.x_grid, s2.y_grid, s2.fspl = xg.ctypes.data_as(_abi.c_double_p), yg.ctypes.data_as(_abi.c_double_p), F.ctypes.data_as(_abi.c_double_p)
    rng = np.random.default_rng(3)
    px, py = rng.uniform(0.1, 1.0, 500), rng.uniform(-1, 1, 500)
    f, fx, fy = np.zeros(500), np.zeros(500), np.zeros(500)
    orc.load().oracle_bcspeval(C.byref(s2), 500, px.ctypes.data_as(_abi.c_double_p), py.ctypes.data_as(_abi.c_double_p),
                               f.ctypes.data_as(_abi.c_double_p), fx.ctypes.data_as(_abi.c_double_p), fy.ctypes.data_as(_abi.c_double_p))
    rr = np.sqrt(px ** 2 + py ** 2)
    assert np.max(np.abs(f - rr)) < 2e-5 and np.max(np.abs(fx - px / rr)) < 2e-3 and np.max(np.abs(fy - py / rr)) < 2e-3


def test_run_results_netcdf_contract(tmp_path):
    """finalize_run writes run_results.<label>.nc with the reference's dims/vars/types
    (ray_results_m.f90:171-249); scipy's classic-netCDF reader stands in for post_process_RAYS."""
    import _oracle as orc
    from _cases import oracle_fan
    from scipy.io import netcdf_file
    cfg = init_case("slab_ECH_90GHz_case_1.in")
    r, n, w, _, _ = oracle_fan(cfg)
    rb.set_fan(r, n, w)
    # fill the host's ray_results_m arrays with an oracle trace (no GPU here), then write the file
    res = _abi.Results()
    L = _abi.load()
    assert L.rays_host_results(C.byref(res)) == 0
    fan, keep = rb.make_fan(r, n, w)
    st = orc.load().oracle_trace(C.byref(cfg), C.byref(fan), C.byref(res), 0, None)
    assert st == 0
    rb.finalize_run(str(tmp_path))
    f = netcdf_file(str(tmp_path / "run_results.run_1.nc"), "r", mmap=False)
    assert f.dimensions["number_of_rays"] == 3 and f.dimensions["dim_v_vector"] == 7 and f.dimensions["max_number_of_points"] == 501
    assert f.RAYS_run_label.decode().strip() == "run_1"
    rv = f.variables["ray_vec"]
    assert rv.shape == (3, 501, 7) and rv.data.dtype == np.dtype(">f8")
    npnt = f.variables["npoints"].data
    assert list(npnt) == [501, 501, 501]
    got = np.array(rv.data)
    want = np.ctypeslib.as_array(res.ray_vec, (3, 501, 7))
    assert np.array_equal(got, want)
    assert f.variables["residual"].shape == (3, 501)
    assert f.variables["initial_ray_power"].data.dtype == np.dtype(">f4")
    flag = b"".join(f.variables["ray_stop_flag"].data[0]).decode()
    assert flag == " nstep > nstep_max".ljust(60)
    assert f.variables["date_vector"].shape == (8,)
    f.close()


def test_mpex_field_file_reader_matches_scipy():
    from scipy.io import netcdf_file
    cfg = init_case("mpex/rays.in")
    f = netcdf_file(rb.config_path("mpex/Brz_fields.MPEX_9_filaments_D3-6_ECH_2nd_harm.nc"), "r", mmap=False)
    Br = np.array(f.variables["Br"].data, dtype=np.float64)        # (n_z, n_r)
    nx, ny = cfg.mirror.Br_spline.nx, cfg.mirror.Br_spline.ny
    fs = np.ctypeslib.as_array(cfg.mirror.Br_spline.fspl, (ny, nx, 4, 4))
    assert np.array_equal(fs[:, :, 0, 0], Br)                       # spline knots carry the file's data
    rg = np.ctypeslib.as_array(cfg.mirror.Br_spline.x_grid, (nx,))
    assert np.allclose(rg, np.array(f.variables["r_grid"].data))
    f.close()


def test_spline_profile_namelists_build_normalised_tables():
    """density_spline_interp_m.f90:93-104 / temperature_spline_interp_m.f90:97-108: uniform psi_N grid on [0,1],
    values normalised to the first (on-axis) one, not-a-knot cubic spline through them."""
    import re
    cfg = init_case("axisym_spline_profiles.in")
    a = cfg.axisym
    assert a.density_prof_model == _abi.PROF_SPLINE and list(a.temperature_prof_model)[:2] == [_abi.PROF_SPLINE] * 2
    txt = open(rb.config_path("axisym_spline_profiles.in")).read()
    for key, sp in (("ne_in", a.ne_spline), ("Te_in", a.Te_spline), ("Ti_in", a.Ti_spline)):
        vals = np.array([float(v) for v in re.search(key + r"\s*=\s*([^\n]*)", txt).group(1).split(",")])
        assert sp.nx == 21 == len(vals)
        grid = np.ctypeslib.as_array(sp.x_grid, (21,))
        fs = np.ctypeslib.as_array(sp.fspl, (21, 4))
        assert np.array_equal(grid, np.array([1.0 * i / 20 for i in range(21)]))
        assert np.array_equal(fs[:, 0], vals / vals[0]) and fs[0, 0] == 1.0
        # continuity of the spline and its first two derivatives at the interior knots
        h = np.diff(grid)
        end = fs[:-1, 0] + h * (fs[:-1, 1] + h * (fs[:-1, 2] + h * fs[:-1, 3]))
        assert np.max(np.abs(end - fs[1:, 0])) < 1e-14
        d1 = fs[:-1, 1] + h * (2 * fs[:-1, 2] + 3 * h * fs[:-1, 3])
        assert np.max(np.abs(d1[:-1] - fs[1:-1, 1])) < 1e-11
        # not-a-knot: the third derivative is continuous across the 2nd and the next-to-last knot
        assert abs(fs[0, 3] - fs[1, 3]) < 1e-9 * max(1.0, abs(fs[0, 3])) and abs(fs[-3, 3] - fs[-2, 3]) < 1e-9 * max(1.0, abs(fs[-3, 3]))


def test_spline_profile_errors(tmp_path):
    L = _abi.load()
    txt = open(rb.config_path("axisym_spline_profiles.in")).read()
    for old, new in (("  ngrid = 21\n  ne_in", "  ngrid = 2\n  ne_in"), ("density_prof_model = 'density_spline_interp'", "density_prof_model = 'spline'")):
        assert old in txt
        p = tmp_path / "bad.in"
        p.write_text(txt.replace(old, new))
        assert L.rays_host_initialize(str(p).encode(), 0) != 0


def test_deposition_profiles_netcdf_contract(tmp_path):
    """write_deposition_profiles_NC (deposition_profiles_m.f90:336-420): record dimension n_profiles, the per-profile
    variables and the grid of bin edges; read back with scipy's classic-netCDF reader (stand-in for plot_RAYS_*.py)."""
    import _oracle as orc
    from _cases import oracle_fan
    from scipy.io import netcdf_file
    cfg = init_case("axisym_deposition_fan.in", nstep_max=300)
    r, n, w, _, _ = oracle_fan(cfg, n_rindex_theta=4, delta_rindex_theta=0.1, n_rindex_phi=4, delta_rindex_phi=0.08)
    o, st, _ = orc.trace(cfg, r, n, w)
    prof, q = orc.deposition(cfg, o, 101, 0.0, 1.0)
    assert q > 0
    # two records exercise the record layout (the reference writes Ptotal_psi and Ptotal_rho for axisym_toroid)
    rb.write_deposition_profiles(str(tmp_path), [
        dict(profile_name="Ptotal_psi", grid_name="psi", grid_min=0.0, grid_max=1.0, profile=prof, Q_sum=q),
        dict(profile_name="Ptotal_rho", grid_name="rho", grid_min=0.0, grid_max=2.0, profile=2.0 * prof, Q_sum=2.0 * q)])
    f = netcdf_file(str(tmp_path / "deposition_profiles.axisym_deposition_fan.nc"), "r", mmap=False)
    assert f.dimensions["n_profiles"] is None and f.dimensions["n_bins"] == 101 and f.dimensions["n_bins_p1"] == 102 and f.dimensions["d20"] == 20
    assert list(f.variables) == ["Q_sum", "n_bins", "grid_min", "grid_max", "profile_name", "grid_name", "grid", "profile"]
    assert f.variables["profile"].shape == (2, 101) and f.variables["grid"].shape == (2, 102) and f.variables["profile_name"].shape == (2, 20)
    assert f.variables["profile"].data.dtype == np.dtype(">f8") and f.variables["n_bins"].data.dtype == np.dtype(">i4")
    assert np.array_equal(np.array(f.variables["profile"].data[0]), prof) and np.array_equal(np.array(f.variables["profile"].data[1]), 2.0 * prof)
    assert list(f.variables["Q_sum"].data) == [q, 2.0 * q] and list(f.variables["n_bins"].data) == [101, 101]
    assert list(f.variables["grid_max"].data) == [1.0, 2.0]
    g1 = np.array(f.variables["grid"].data[1])
    assert g1[0] == 0.0 and g1[-1] == 2.0 and np.allclose(np.diff(g1), 2.0 / 101, rtol=1e-12)
    assert b"".join(f.variables["profile_name"].data[1]).decode() == "Ptotal_rho".ljust(20)
    assert b"".join(f.variables["grid_name"].data[0]).decode() == "psi".ljust(20)
    assert f.RAYS_run_label.decode().strip() == "axisym_deposition_fan" and len(f.date_vector) == 8
    f.close()


def test_spline_setup_against_an_independent_implementation():
    """The host mirror's spline setup (splines_setup.cpp: v_spline / cspline / bcspline, not-a-knot) feeds the GPU path AND the
    oracle, so GPU-vs-oracle parity cannot see an error in it.  A not-a-knot interpolating cubic (tensor-product bicubic) spline
    is unique: scipy's CubicSpline / RectBivariateSpline (FITPACK B-splines, a different algorithm) must give the same function
    and derivatives to rounding level - on a uniform grid and on the MPEX field file's own grid and data."""
    from scipy.interpolate import CubicSpline, RectBivariateSpline
    from scipy.io import netcdf_file
    import _oracle as orc
    L = _abi.load()
    rng = np.random.default_rng(11)
    # 1-D: smooth + rough data
    nx = 57
    x = np.linspace(-0.3, 1.7, nx)
    for y in (np.exp(-x) * np.sin(7.0 * x), rng.normal(size=nx)):
        f = np.zeros((nx, 4))
        f[:, 0] = y
        assert L.rays_host_cspline(x.ctypes.data_as(_abi.c_double_p), nx, f.ctypes.data_as(_abi.c_double_p)) == 0
        s = _abi.Spline1D()
        s.nx, s.x_grid, s.fspl = nx, x.ctypes.data_as(_abi.c_double_p), f.ctypes.data_as(_abi.c_double_p)
        xs = np.sort(rng.uniform(x[0], x[-1], 2000))
        fo, fpo = np.zeros_like(xs), np.zeros_like(xs)
        orc.load().oracle_cspeval(C.byref(s), len(xs), xs.ctypes.data_as(_abi.c_double_p), fo.ctypes.data_as(_abi.c_double_p),
                                  fpo.ctypes.data_as(_abi.c_double_p))
        ref = CubicSpline(x, y, bc_type="not-a-knot")
        scale = np.max(np.abs(y))
        assert np.max(np.abs(fo - ref(xs))) <= 1e-12 * scale
        assert np.max(np.abs(fpo - ref(xs, 1))) <= 1e-10 * scale / (x[1] - x[0])
    # 2-D: the shipped MPEX field (51 x 201 grid, Br / Bz / Aphi) and its first derivatives
    nc = netcdf_file(rb.config_path("mpex/Brz_fields.MPEX_9_filaments_D3-6_ECH_2nd_harm.nc"), "r", mmap=False)
    rg, zg = np.array(nc.variables["r_grid"].data, dtype=np.float64), np.array(nc.variables["z_grid"].data, dtype=np.float64)
    for name in ("Br", "Bz", "Aphi"):
        A = np.array(nc.variables[name].data, dtype=np.float64)          # file order (z, r): Fortran A(r, z)
        assert A.shape == (len(zg), len(rg))
        F = np.zeros((len(zg), len(rg), 4, 4))
        F[:, :, 0, 0] = A
        assert L.rays_host_bcspline(rg.ctypes.data_as(_abi.c_double_p), len(rg), zg.ctypes.data_as(_abi.c_double_p), len(zg),
                                    F.ctypes.data_as(_abi.c_double_p)) == 0
        s2 = _abi.Spline2D()
        s2.nx, s2.ny = len(rg), len(zg)
        s2.x_grid, s2.y_grid, s2.fspl = rg.ctypes.data_as(_abi.c_double_p), zg.ctypes.data_as(_abi.c_double_p), F.ctypes.data_as(_abi.c_double_p)
        n = 3000
        pr, pz = rng.uniform(rg[0], rg[-1], n), rng.uniform(zg[0], zg[-1], n)
        f, fx, fy = np.zeros(n), np.zeros(n), np.zeros(n)
        orc.load().oracle_bcspeval(C.byref(s2), n, pr.ctypes.data_as(_abi.c_double_p), pz.ctypes.data_as(_abi.c_double_p),
                                   f.ctypes.data_as(_abi.c_double_p), fx.ctypes.data_as(_abi.c_double_p), fy.ctypes.data_as(_abi.c_double_p))
        ref = RectBivariateSpline(rg, zg, A.T, kx=3, ky=3, s=0)
        scale = np.max(np.abs(A))
        assert np.max(np.abs(f - ref.ev(pr, pz))) <= 1e-11 * scale
        assert np.max(np.abs(fx - ref.ev(pr, pz, dx=1))) <= 1e-9 * scale / (rg[1] - rg[0])
        assert np.max(np.abs(fy - ref.ev(pr, pz, dy=1))) <= 1e-9 * scale / (zg[1] - zg[0])
    nc.close()


def test_host_constants_carry_the_reference_single_precision_literals():
    """constants_m.f90:38-48, rf_m.f90:87-91, species_m.f90:33-34,150-158: the reference writes pi, clight, mu0's 4.e-7, me and e as
    default-REAL literals, so the doubles it computes with are float32-rounded (pi is off by 2.8e-8).  The host
    mirror feeds the same configuration to the GPU path and to the oracle, so this is checked against the arithmetic restated
    here with numpy float32 round trips, and against the values SURVEY.md A.1 lists - exactly."""
    f32 = lambda v: float(np.float32(v))
    pi, clight, me, e = f32(3.1415926535897932385), f32(2.997930e8), f32(9.1094e-31), f32(1.6022e-19)
    mu0 = pi * f32(4.e-7)
    eps0 = 1.0 / (mu0 * (clight * clight))
    assert (pi, clight, mu0, eps0) == (3.1415927410125732, 299792992.0, 1.2566371110902128e-06, 8.8541559251146982e-12)
    assert (me, e) == (9.1094003725447319e-31, 1.6021999911602857e-19)
    cfg = init_case("slab_ECH_90GHz_case_1.in")
    assert cfg.clight == clight and cfg.eps0 == eps0
    omgrf = 2.0 * pi * 90.e9
    assert cfg.omgrf == omgrf and cfg.k0 == omgrf / clight
    assert cfg.nspec == 1
    assert list(cfg.qs)[:2] == [e * -1.0, e * 1.0] and list(cfg.ms)[:2] == [me * 1.0, me * 3670.0]
    assert list(cfg.t0s)[:2] == [e * 5.0e3, e * 1.0e2]
