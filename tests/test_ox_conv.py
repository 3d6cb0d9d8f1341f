"""O-X mode conversion analysis of stored trajectories (post_process_lib/OX_conv_analysis_m.f90):
per ray the saved point of maximum electron density, the nearest point on the O-mode cutoff surface (alpha_e = 1)
and the conversion coefficient there.  The MPEX examples run it (`do_OX_conv_analysis = .true.`) but ship no numbers:
the oracle is checked on the physics it must show, the CUDA kernel against the oracle."""
import numpy as np
import pytest

import rays_b200 as rb
import _oracle as orc
from _cases import init_case, oracle_fan


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    yield


def _oracle_case(name, **ode):
    cfg = init_case(name, **ode)
    r, n, w, _, _ = oracle_fan(cfg)
    o, st, _ = orc.trace(cfg, r, n, w)
    assert st == 0
    return cfg, r, n, w, o


def test_oracle_ox_analysis_on_the_mpex_scans():
    # the nx scan is mirror symmetric in x about the central ray (ray 6): so must the analysis be
    cfg, r, n, w, o = _oracle_case("mpex/rays.in")
    ox, nconv = orc.ox_conv(cfg, o)
    assert nconv == 11 and np.all(ox["found_max"] == 1) and np.all(ox["found_cutoff"] == 1) and list(ox["ray_number"]) == list(range(1, 12))
    assert np.all((ox["iteration"] >= 1) & (ox["iteration"] <= 10))
    c = ox["conv_coeff"]
    assert np.all((c > 1e-4) & (c <= 1.0)) and np.argmax(c) == 5 and c[5] > 0.99
    assert np.allclose(c, c[::-1], rtol=1e-6) and np.all(np.diff(c[:6]) > 0)
    assert np.allclose(ox["x_cut"][:, 0], -ox["x_cut"][::-1, 0], atol=1e-9) and np.allclose(ox["x_cut"][:, 1:], ox["x_cut"][::-1, 1:], atol=1e-9)
    # x_cut lies on the cutoff surface, x_max is the saved point just below it
    e, err = orc.probe_equilibrium(cfg, np.ascontiguousarray(ox["x_cut"]))
    alpha_e = e[:, 3 + 9 + 6 + 18 + 6 + 18 + 1 + 3 + 3 + 9 + 6 + 6]          # RAYS_EQ_OUT layout: ... omgc6, omgp2 6, alpha6
    assert np.all(err == 0) and np.all(np.abs(alpha_e - 1.0) <= 1e-4)
    assert np.all(ox["alpha_max"] < 1.0) and np.all(ox["alpha_max"] > 0.9)
    for k in range(11):
        i = ox["step_number"][k] - 1
        assert np.array_equal(ox["x_max"][k], o.ray_vec[k, i, 0:3]) and np.array_equal(ox["k_max"][k], o.ray_vec[k, i, 3:6])
    # n components in the cutoff frame: |n_x|^2 + |n_y|^2 + |n_z|^2 = |k_max|^2 / k0^2
    n2 = (ox["nvecx_c"] ** 2).sum(1) + (ox["nvecy_c"] ** 2).sum(1) + (ox["nvecz_c"] ** 2).sum(1)
    assert np.allclose(n2, (ox["k_max"] ** 2).sum(1) / cfg.k0 ** 2, rtol=1e-12)
    # the nz scan walks through the optimum angle: the coefficient peaks inside the scan
    cfg, r, n, w, o = _oracle_case("mpex_nz/rays.in")
    ox, nconv = orc.ox_conv(cfg, o)
    c = ox["conv_coeff"]
    assert nconv == 11 and 0 < np.argmax(c) < 10 and c.max() > 0.995 and c.min() > 0.8


def test_oracle_ox_analysis_when_there_is_no_cutoff():
    # Solov'ev O-mode rays far above the plasma frequency: density maximum found, but no cutoff surface nearby
    cfg, r, n, w, o = _oracle_case("examples/solovev_ECH_90GHz_plus_root.in", nstep_max=300)
    ox, nconv = orc.ox_conv(cfg, o)
    assert nconv == 0 and np.all(ox["conv_coeff"] == 0.0) and np.all(ox["found_cutoff"] == 0)
    assert np.all(ox["x_cut"] == 0.0) and np.all(ox["iteration"][ox["found_max"] == 1] == 11)
    # a ray that never passes a density maximum: records stay zero (slab, density rising along x)
    cfg, r, n, w, o = _oracle_case("examples/slab_ECH_90GHz_case_2.in", nstep_max=20)
    ox, nconv = orc.ox_conv(cfg, o)
    assert nconv == 0 and np.all(ox["found_max"] == 0) and np.all(ox["step_number"] == 0) and np.all(ox["x_max"] == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name,ode", [("mpex/rays.in", {}), ("mpex_nz/rays.in", {}), ("examples/solovev_ECH_90GHz_plus_root.in", {"nstep_max": 300}),
                                      ("examples/slab_ECH_90GHz_case_2.in", {"nstep_max": 20}), ("axisym_deposition_fan.in", {"nstep_max": 200})])
def test_gpu_ox_analysis_equals_oracle(name, ode):
    rb.init(0)
    cfg = init_case(name, **ode)
    kw = dict(n_rindex_theta=8, delta_rindex_theta=0.05, n_rindex_phi=8, delta_rindex_phi=0.04) if name.startswith("axisym") else {}
    r, n, w, _, _ = oracle_fan(cfg, **kw)
    rb.set_config(cfg)
    rb.fan_upload(r, n, w)
    rb.trace_device(store=True)
    g, gn = rb.ox_conv_analysis(r.shape[0])
    res = rb.results_download(r.shape[0], int(cfg.nv), int(cfg.nstep_max) + 1, store=True)
    o, on = orc.ox_conv(cfg, res)          # the oracle analyses the very trajectories the device stored
    assert gn == on
    for f in ("ray_number", "step_number", "found_max", "found_cutoff", "converted", "iteration"):
        assert np.array_equal(g[f], o[f]), f
    assert np.array_equal(g["x_max"], o["x_max"]) and np.array_equal(g["k_max"], o["k_max"])
    # alpha at the maximum is the density profile itself: bitwise unless the profile calls libm (the mirror's hyperbolic profile
    # takes tanh and cosh from one exponential on the device: 3e-16 absolute, tests/test_exact_division.py has the formula's check)
    if name.startswith("mpex"):
        assert np.allclose(g["alpha_max"], o["alpha_max"], rtol=1e-14, atol=0)
    else:
        assert np.array_equal(g["alpha_max"], o["alpha_max"])
    # the mirror / Gaussian / hyperbolic profiles call libm (tanh, cosh, pow, exp): rounding-level agreement from there on
    assert np.allclose(g["x_cut"], o["x_cut"], rtol=1e-12, atol=1e-15)
    assert np.allclose(g["conv_coeff"], o["conv_coeff"], rtol=1e-9, atol=0)
    for f in ("nvecx_c", "nvecy_c", "nvecz_c"):
        assert np.allclose(g[f], o[f], rtol=1e-9, atol=1e-12)
