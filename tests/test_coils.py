"""the coil-field generator behind the mirror equilibrium's Brz field file
(mirror_magnetics_lib: mirror_magnetics_m.f90, B_loop_m.f90; math_functions_lib/complete_elliptic_int_m.f90).
The coil set of the reference's shipped field file is not in its tree, so the oracle is checked against the textbook
loop field (scipy's K, E) and against B = curl A -- the identity the reference's own shipped file satisfies -- and
the CUDA path against the oracle (bitwise off the axis)."""
import os
import shutil

import numpy as np
import pytest

import rays_b200 as rb
from rays_b200 import _abi
import _oracle as orc
from _cases import init_case_text

COILS = [dict(inner_radius=a, outer_radius=b, z_center=zc, z_width=w, I_coil=cur, n_turns=n, n_r_layers=3, n_z_slices=3)
         for a, b, w, zc, n, cur in zip([0.22, 0.22, 0.30, 0.30, 0.22, 0.22], [0.30, 0.30, 0.42, 0.42, 0.30, 0.30],
                                        [0.10, 0.10, 0.16, 0.16, 0.10, 0.10], [2.60, 2.85, 3.10, 3.35, 3.60, 3.85],
                                        [120, 120, 200, 200, 120, 120], [5200., 5200., 3100., 3100., 5200., 5200.])]   # = configs/coils/*.nml
SHIPPED = "mpex/Brz_fields.MPEX_9_filaments_D3-6_ECH_2nd_harm.nc"


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    yield


def _curl_errors(rg, zg, Br, Bz, A):
    """the file's 'Aphi' is the flux function r*A_phi: B_z = (1/r) dAphi/dr, B_r = -(1/r) dAphi/dz (central differences)"""
    dr, dz = rg[1] - rg[0], zg[1] - zg[0]
    ez = np.max(np.abs((A[:, 2:] - A[:, :-2]) / (2 * dr) / rg[1:-1] - Bz[:, 1:-1])) / np.abs(Bz).max()
    er = np.max(np.abs(-(A[2:, :] - A[:-2, :])[:, 1:] / (2 * dz) / rg[1:] - Br[1:-1, 1:])) / np.abs(Br).max()
    return ez, er


def test_complete_elliptic_integrals_match_scipy():
    from scipy.special import ellipe, ellipk
    m = np.concatenate([np.linspace(0.0, 0.999, 500), 1.0 - np.logspace(-12, -3, 40)])
    K, E = orc.elliptic(m)
    assert np.max(np.abs(K - ellipk(m)) / ellipk(m)) < 5e-14 and np.max(np.abs(E - ellipe(m)) / ellipe(m)) < 5e-14


def test_loop_field_matches_the_textbook_formulas():
    from scipy.special import ellipe, ellipk
    rng = np.random.default_rng(0)
    r, z = rng.uniform(0.002, 3.0, 4000), rng.uniform(-3.0, 3.0, 4000)
    Br, Bz, A = orc.Brz_loop_scaled(r, z)
    mu0 = 4e-7 * np.pi
    k2 = 4 * r / ((1 + r) ** 2 + z ** 2)
    K, E = ellipk(k2), ellipe(k2)
    pre = mu0 / (2 * np.pi) / np.sqrt((1 + r) ** 2 + z ** 2)
    Bz_t = pre * (K + (1 - r ** 2 - z ** 2) / ((1 - r) ** 2 + z ** 2) * E)
    Br_t = pre * z / r * (-K + (1 + r ** 2 + z ** 2) / ((1 - r) ** 2 + z ** 2) * E)
    rA_t = r * mu0 / (np.pi * np.sqrt(k2)) * np.sqrt(1 / r) * ((1 - k2 / 2) * K - E)      # the reference's Aphi is r*A_phi
    # 3e-8: pi and 4.e-7 are single-precision literals in B_loop_m.f90:28-31
    assert np.max(np.abs(Bz - Bz_t)) < 3e-8 * np.abs(Bz_t).max() and np.max(np.abs(Br - Br_t)) < 3e-8 * np.abs(Br_t).max()
    assert np.max(np.abs(A - rA_t)) < 3e-8 * np.abs(rA_t).max()
    # on the axis and across the small-r series switch at r = 1e-3 loop radii
    b0 = orc.Brz_loop_scaled(np.array([0.0]), np.array([0.7]))
    assert b0[0][0] == 0.0 and b0[2][0] == 0.0 and abs(b0[1][0] - mu0 / 2 / 1.49 ** 1.5) < 6e-8 * b0[1][0]
    lo, hi = orc.Brz_loop_scaled(np.array([0.00099]), np.array([0.7])), orc.Brz_loop_scaled(np.array([0.00101]), np.array([0.7]))
    assert abs(lo[1][0] - hi[1][0]) < 1e-6 * hi[1][0] and abs(lo[0][0] / 0.00099 - hi[0][0] / 0.00101) < 1e-5 * hi[0][0] / 0.00101
    # (R) the series for Aphi (B_loop_m.f90:226-227) carries 1/pi where the closed form has mu0: reproduced as written
    # (only points within 1e-3 loop radii of the axis but not on it take that branch; no shipped grid has one)
    assert abs(lo[2][0] * mu0 * np.pi / (hi[2][0] * (0.00099 / 0.00101) ** 2) - 1.0) < 1e-4


def test_grid_satisfies_curl_identities_like_the_shipped_file():
    from scipy.io import netcdf_file
    f = netcdf_file(rb.config_path(SHIPPED), "r", mmap=False)
    ref = [np.array(f.variables[k].data, dtype=np.float64) for k in ("r_grid", "z_grid", "Br", "Bz", "Aphi")]
    f.close()
    ez, er = _curl_errors(*ref)
    assert ez < 1e-4 and er < 2e-4          # the reference's own output: Aphi is r*A_phi, signs as in _curl_errors
    rg, zg, Br, Bz, A = orc.mirror_Brz_grid(COILS, 51, 0.0, 0.2, 201, 2.8, 3.6)
    assert np.array_equal(rg, ref[0]) and np.array_equal(zg, ref[1])          # same grid formula as the shipped file's
    ez, er = _curl_errors(rg, zg, Br, Bz, A)
    assert ez < 2e-3 and er < 5e-3          # coils closer to the window than MPEX's: larger O(h^2) difference error
    rg2, zg2, Br2, Bz2, A2 = orc.mirror_Brz_grid(COILS, 101, 0.0, 0.2, 401, 2.8, 3.6)
    ez2, er2 = _curl_errors(rg2, zg2, Br2, Bz2, A2)
    assert ez2 < 0.3 * ez and er2 < 0.3 * er      # second order in h
    assert np.all(Br[:, 0] == 0.0) and np.all(A[:, 0] == 0.0) and 2.0 < Bz.min() < Bz.max() < 4.5
    # div B = 0 on the axis: dBz/dz = -2 dBr/dr
    dz = zg[1] - zg[0]
    assert np.max(np.abs((Bz[2:, 0] - Bz[:-2, 0]) / (2 * dz) + 2.0 * Br[1:-1, 1] / rg[1])) < 2e-3 * np.abs(Bz).max() / 0.2


def test_coil_struct_layout():
    assert _abi.C.sizeof(_abi.Coil) == 56 and _abi.Coil.n_turns.offset == 40 and _abi.Coil.n_z_slices.offset == 52


def test_generator_refuses_to_run_without_a_gpu(tmp_path):
    """program mirror_magnetics goes through the CUDA kernel only: no CPU fallback"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rb.RaysError):
        rb.mirror_magnetics(rb.config_path("coils/mirror_magnetics.nml"), str(tmp_path))
    assert not list(tmp_path.iterdir())
    with pytest.raises(rb.RaysError):
        rb.mirror_Brz_grid(COILS, 5, 0.0, 0.2, 5, 2.8, 3.6)


# ---- CUDA path ----------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_grid_equals_oracle():
    rb.init(0)
    g = rb.mirror_Brz_grid(COILS, 51, 0.0, 0.2, 201, 2.8, 3.6)
    o = orc.mirror_Brz_grid(COILS, 51, 0.0, 0.2, 201, 2.8, 3.6)
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1])
    for a, b in zip(g[2:], o[2:]):
        assert np.array_equal(a[:, 1:], b[:, 1:])                       # off the axis: sqrt and division only -> bitwise
        assert np.allclose(a[:, 0], b[:, 0], rtol=1e-14, atol=0)        # on the axis: pow(1 + z^2, 1.5)
    # a window inside the small-r series (r < 1e-3 loop radii): real powers -> rounding-level agreement
    g = rb.mirror_Brz_grid(COILS, 17, 0.0, 2.0e-4, 33, 2.8, 3.6)
    o = orc.mirror_Brz_grid(COILS, 17, 0.0, 2.0e-4, 33, 2.8, 3.6)
    for a, b in zip(g[2:], o[2:]):   # B_r is a sum of 54 filament terms of both signs: tolerance against the field scale
        assert np.allclose(a, b, rtol=1e-12, atol=1e-13 * np.abs(b).max())
    # ragged shapes and a single-filament coil
    one = [dict(inner_radius=0.5, outer_radius=0.5, z_center=0.0, z_width=0.0, I_coil=1.0, n_turns=1, n_r_layers=1, n_z_slices=1)]
    g = rb.mirror_Brz_grid(one, 7, 0.0, 1.5, 1, 0.3, 0.3)
    o = orc.mirror_Brz_grid(one, 7, 0.0, 1.5, 1, 0.3, 0.3)
    assert g[2].shape == (1, 7) and np.array_equal(g[3][:, 1:], o[3][:, 1:]) and np.array_equal(g[4][:, 1:], o[4][:, 1:])


@pytest.mark.gpu
def test_gpu_mirror_magnetics_program_and_round_trip(tmp_path):
    """program mirror_magnetics: namelists -> Brz_fields.<name>.nc with the shipped file's structure; the file then
    drives the multiple_mirror equilibrium of the MPEX example (mirror_magnetics_spline_interp reads it)."""
    from scipy.io import netcdf_file
    from test_gpu_parity import _compare_traces, _run_both
    from _cases import oracle_fan
    rb.init(0)
    path = rb.mirror_magnetics(rb.config_path("coils/mirror_magnetics.nml"), str(tmp_path))
    assert os.path.basename(path) == "Brz_fields.synthetic_6_coils_case_A_ECH_window.nc"
    f, ref = netcdf_file(path, "r", mmap=False), netcdf_file(rb.config_path(SHIPPED), "r", mmap=False)
    assert dict(f.dimensions) == dict(ref.dimensions) and list(f.variables) == list(ref.variables)
    for k in ref.variables:
        assert f.variables[k].dimensions == ref.variables[k].dimensions and f.variables[k].data.dtype == ref.variables[k].data.dtype
    assert f.NC_file_name.decode() == os.path.basename(path)
    o = orc.mirror_Brz_grid(COILS, 51, 0.0, 0.2, 201, 2.8, 3.6)
    for k, b in zip(("r_grid", "z_grid", "Br", "Bz", "Aphi"), o):
        a = np.array(f.variables[k].data, dtype=np.float64)
        assert np.array_equal(a[..., 1:], b[..., 1:]) and np.allclose(a, b, rtol=1e-14, atol=0)
    assert float(f.variables["r_LUFS"].data) == 0.12 and float(f.variables["z_max"].data) == 3.6
    f.close(); ref.close()
    # round trip: the generated field replaces the shipped one in the MPEX example
    for name in ("ray_init_2nd_harm_11_rays_nx.in",):
        shutil.copy(rb.config_path("mpex/" + name), tmp_path / name)
    cfg = init_case_text("mpex/rays.in", [("Brz_fields.MPEX_9_filaments_D3-6_ECH_2nd_harm.nc", os.path.basename(path))], tmp_path, nstep_max=200)
    r, n, w, _, _ = oracle_fan(cfg)
    assert r.shape[0] >= 5
    g, oo = _run_both(cfg, r, n, w)
    _compare_traces(g, oo, cfg, 1e-10, bitwise=False)
