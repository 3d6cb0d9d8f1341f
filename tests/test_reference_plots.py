"""Parity against the REFERENCE'S OWN OUTPUT.  The example directories of the reference ship vector PDFs of the
rays the Fortran code traced (examples_RAYS/*/ray_plots.*.pdf).  tests/golden/make_ref_plot_vectors.py turned
their polylines into coordinates (tests/golden/ref_plot_vectors.json): 29 rays of seven example inputs, slab and
Solov'ev (two of the slab inputs are not shipped and were reconstructed from the figures, see that script), all
integrated with SG_ODE, 2 164 plotted trajectory points with a resolution of 1e-6 pt = 1.3e-9 ...
8.7e-9 m (3e-14 m for z on the x10^-5 axis of the equatorial-plane runs).

Every plotted point must coincide with a saved point of our trajectory of the same ray to that resolution, in
order, and the plotted ray must end where ours ends.  The oracle is pinned here on CPU; the CUDA path is held to
the same figures in the gpu-marked test."""
import json
import os

import numpy as np
import pytest

import rays_b200 as rb
import _oracle as orc
from _cases import init_case, oracle_fan

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_plot_vectors.json")))
NAMELISTS = sorted({f["namelist"] for f in GOLD["figures"]})

# The figures of the Solov'ev example were drawn by the OLDER generation of the code (RAYS_code/), which ends a ray at the plasma
# edge differently from RAYS_project: RAYS_code/solovev_eq_m.f90:140 raises 'psi >1 out_of_plasma' (the current solovev_eq has no
# such test) and RAYS_code/ray_tracing.f90:131-153 ends the ray on ANY flag of check_save without counting that point.  With that
# ending rule (oracle_set_old_generation, a test switch of the oracle) every one of the 29 figure rays ends on exactly the last
# plotted point.  The CURRENT generation (what the oracle restates and the CUDA path implements) carries plus_root rays 3 and 5 ONE
# saved point further, into psi_N > 1, before Shampine-Gordon stops them with 'equations stiff'; all other rays have the same length
# in both generations.  ENDS_EARLY lists that known generation difference for the current-generation runs.
ENDS_EARLY = {("examples/solovev_ECH_90GHz_plus_root.in", 2): 1, ("examples/solovev_ECH_90GHz_plus_root.in", 4): 1}
TOL_QUANTA = 1.5      # 0.5 from the PDF's rounding per coordinate + the tick-calibration fit


def _coord(tr, name):
    return {"x": tr[:, 0], "y": tr[:, 1], "z": tr[:, 2], "r": np.sqrt(tr[:, 0] ** 2 + tr[:, 1] ** 2)}[name]


def _check_against_figures(namelist, res, ends_early=ENDS_EARLY):
    figs = [f for f in GOLD["figures"] if f["namelist"] == namelist]
    assert figs
    n_pts = 0
    for F in figs:
        assert len(F["rays"]) == res.nray, "the figure shows every ray of the run"
        for i, R in enumerate(F["rays"]):
            tr = res.ray_vec[i, :res.npoints[i]]
            H, V = _coord(tr, F["h"]), _coord(tr, F["v"])
            gh, gv = np.array(R["h"]), np.array(R["v"])
            last = -1
            for a, b in zip(gh, gv):
                d = np.maximum(np.abs(H - a) / F["quantum_h"], np.abs(V - b) / F["quantum_v"])
                k = int(np.argmin(d))
                assert d[k] <= TOL_QUANTA, (F["pdf"], i, k, float(d[k]))
                assert k >= last, "plotted points follow the ray"
                last = k
            n_pts += len(gh)
            assert res.npoints[i] - 1 - last == ends_early.get((namelist, i), 0), (F["pdf"], i, int(res.npoints[i]), last)
    return n_pts


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    yield


@pytest.mark.parametrize("namelist", NAMELISTS)
def test_oracle_reproduces_the_reference_figures(namelist):
    cfg = init_case(namelist)
    assert cfg.ode_solver == 2      # every shipped example figure is a Shampine-Gordon run
    r, n, w, _, _ = oracle_fan(cfg)
    o, st, _ = orc.trace(cfg, r, n, w)
    assert st == 0
    assert _check_against_figures(namelist, o) >= 80


@pytest.mark.parametrize("namelist", [n for n in NAMELISTS if "solovev" in n])
def test_old_generation_ending_rule_reproduces_every_figure_ray_exactly(namelist):
    """with the older generation's ending rule (psi_N > 1 is 'out_of_plasma', a flagged point ends the ray uncounted:
    RAYS_code/solovev_eq_m.f90:140, RAYS_code/ray_tracing.f90:131-153) EVERY figure ray ends on its last plotted point: no exceptions"""
    L = orc.load()
    cfg = init_case(namelist)
    r, n, w, _, _ = oracle_fan(cfg)
    was = L.oracle_set_old_generation(1)
    try:
        o, st, _ = orc.trace(cfg, r, n, w)
    finally:
        L.oracle_set_old_generation(was)
    assert st == 0
    assert _check_against_figures(namelist, o, ends_early={}) >= 80
    # the rays that end at the plasma edge do so with the older generation's flag
    assert all(f.strip() in ("out_of_plasma", "nstep > nstep_max") for f in o.ray_stop_flag)


def test_figure_inventory():
    assert len(GOLD["figures"]) == 10 and sum(len(f["rays"]) for f in GOLD["figures"]) == 47
    assert sum(len(r["h"]) for f in GOLD["figures"] for r in f["rays"]) == 2164


@pytest.mark.gpu
@pytest.mark.parametrize("namelist", NAMELISTS)
def test_cuda_path_reproduces_the_reference_figures(namelist):
    rb.init(0)
    cfg = init_case(namelist)
    r, n, w, _, _ = oracle_fan(cfg)
    g = rb.trace(cfg, r, n, w)
    _check_against_figures(namelist, g)
    # and through the device launcher (ray_init on the GPU) as `program rays` would run the example
    o, st, _ = orc.trace(cfg, r, n, w)
    assert np.array_equal(g.npoints, o.npoints) and g.ray_stop_flag == o.ray_stop_flag


# ---- RK4 + mirror equilibrium: the MPEX example's raster figure ------------------------------------------------
RASTER = json.load(open(os.path.join(HERE, "golden", "ref_raster_vectors.json")))


def _check_against_raster(res, G):
    """One of the reference's MPEX figures (11 rays each, RK4_ODE, spline-interpolated mirror field; launch scans drawn in
    the z-y plane ('nz', 'nz_30deg') or the x-y plane ('nx', 'x'), 0.87 mm per pixel; tests/golden/make_ref_raster_vectors.py).
    Rays overdraw each other, so the comparison is between the band of ray-coloured pixels and the band of our 11
    trajectories inside the figure's window (clear of the guide lines), both ways:
      * every trajectory point has a ray-coloured pixel within 2.5 px (1.2 px for the rays whose colour survives the
        anti-aliasing: all but brown and grey), i.e. our rays run where the reference drew rays;
      * 98 % of the ray-coloured pixels lie within 1.5 px of one of our trajectories, i.e. the reference drew no ray
        where we have none -- in particular both edges of the fan and its end on the last flux surface agree."""
    P = np.stack([np.array(G["pixels_h"]) / G["pixel_h"], np.array(G["pixels_v"]) / G["pixel_v"]], axis=1)
    col = {"x": 0, "y": 1, "z": 2}
    ih, iv = col[G["h"]], col[G["v"]]
    w = G["window"]
    assert res.nray == 11
    bands = []
    for i in range(res.nray):
        tr = res.ray_vec[i, :res.npoints[i]]
        tr = tr[(tr[:, ih] > w[0]) & (tr[:, ih] < w[1]) & (tr[:, iv] > w[2]) & (tr[:, iv] < w[3] - 0.0015)]
        Q = np.stack([tr[:, ih] / G["pixel_h"], tr[:, iv] / G["pixel_v"]], axis=1)
        bands.append(Q)
        d = np.sqrt(((Q[::3, None, :] - P[None, :, :]) ** 2).sum(2)).min(1)
        assert len(Q) > 100 and d.max() <= (2.5 if i in (5, 7) else 1.2), (i, float(d.max()))
    Q = np.concatenate(bands)
    keep = P[:, 1] * G["pixel_v"] < -0.0275         # the rows above hold the plasma-boundary guide line
    d = np.sqrt(((P[keep][:, None, :] - Q[None, :, :]) ** 2).sum(2)).min(1)
    # (the dashed resonance guide line crosses the z-y windows: a few dozen of its blended pixels pass the colour filter)
    assert np.percentile(d, 98) <= 1.5 and (d > 3.0).mean() < 0.015, (float(np.percentile(d, 98)), float((d > 3.0).mean()))
    if G["h"] == "z":   # the fan's extent along z at the level where it leaves the window: first and last ray
        z_end = [float(b[-1, 0] * G["pixel_h"]) for b in bands]
        low = P[P[:, 1] * G["pixel_v"] < -0.114][:, 0] * G["pixel_h"]
        assert abs(min(z_end) - low.min()) < 2.5 * G["pixel_h"] and abs(max(z_end) - low.max()) < 2.5 * G["pixel_h"]
    else:               # x-y view: the returning rays' outermost reach in x on both sides
        x_our = np.concatenate([b[:, 0] for b in bands]) * G["pixel_h"]
        x_pix = P[:, 0] * G["pixel_h"]
        assert abs(x_our.min() - x_pix.min()) < 2.5 * G["pixel_h"] and abs(x_our.max() - x_pix.max()) < 2.5 * G["pixel_h"]


@pytest.mark.parametrize("fig", range(len(RASTER["figures"])))
def test_oracle_reproduces_the_reference_mpex_raster(fig):
    G = RASTER["figures"][fig]
    cfg = init_case(G["namelist"])
    assert cfg.ode_solver == 1 and cfg.equilib_model == 4      # RK4_ODE, multiple_mirror
    r, n, w, _, _ = oracle_fan(cfg)
    o, st, _ = orc.trace(cfg, r, n, w)
    assert st == 0 and set(s.strip() for s in o.ray_stop_flag) <= {"out_of_plasma", "nstep > nstep_max"}   # nx runs 500 steps only
    _check_against_raster(o, G)


@pytest.mark.gpu
@pytest.mark.parametrize("fig", range(len(RASTER["figures"])))
def test_cuda_path_reproduces_the_reference_mpex_raster(fig):
    rb.init(0)
    G = RASTER["figures"][fig]
    cfg = init_case(G["namelist"])
    r, n, w, _, _ = oracle_fan(cfg)
    g = rb.trace(cfg, r, n, w)
    _check_against_raster(g, G)
    o, st, _ = orc.trace(cfg, r, n, w)
    assert np.array_equal(g.npoints, o.npoints) and g.ray_stop_flag == o.ray_stop_flag


# ---- cold dispersion roots of the launchers: the slab example's kx-profile figures --------------------------
KX = json.load(open(os.path.join(HERE, "golden", "ref_kx_profiles.json")))


@pytest.mark.parametrize("page", range(len(KX["pages"])))
def test_oracle_reproduces_the_reference_kx_profiles(page):
    """examples_RAYS/ECH_90GHz_slab/pdf_plots/kx_plots.run_{1,2}.pdf: k0*nx (re, im) of two cold roots at 101 x positions
    across the slab (write_kx_profiles, slab_processor_m.f90:729-827); run_3's input was reconstructed from its pages, so for
    those six pages this test is a consistency check of that reconstruction and the ray figure is the independent one.  The tick labels are outlined, so the page's y
    calibration (scale, offset) is fitted: 2 numbers against 404 plotted values, which then agree to the PDF's 1e-4 pt.
    Where the two roots are a complex pair the fast/slow pages label them the other way round than the current source
    (an older build drew them): those pages are compared as unordered pairs (re) and by magnitude (im)."""
    G = KX["pages"][page]
    cfg = init_case(G["namelist"])
    r, n, w, _, _ = oracle_fan(cfg)
    ny, nz = n[G["ray"] - 1, 1], n[G["ray"] - 1, 2]
    x = np.array([cfg.slab.xmin + i * ((cfg.slab.xmax - cfg.slab.xmin) / 100) for i in range(101)])
    xp = np.array(G["x_pt"])
    assert np.max(np.abs(xp - np.linspace(xp[0], xp[-1], 101))) < 2e-4          # the plotted abscissae are that uniform grid
    modes = (1, 2) if G["roots"] == "plus_minus" else (3, 4)
    a_re, a_im = orc.kx_profile(cfg, x, ny, nz, modes[0])
    b_re, b_im = orc.kx_profile(cfg, x, ny, nz, modes[1])
    Y = [np.array(c) for c in G["curves_pt"]]
    assert np.ptp(a_re) > 100.0 and max(np.ptp(a_im), np.ptp(b_im)) > 100.0        # propagating and evanescent stretches both present
    # calibration from the sum of the two real parts (independent of which curve is which root)
    A = np.stack([a_re + b_re, 2.0 * np.ones(101)], axis=1)
    scale, off = np.linalg.lstsq(A, Y[0] + Y[2], rcond=None)[0]
    assert 0.001 < scale < 0.1
    y = [(c - off) / scale for c in Y]
    tol = 3e-4 / scale + 2e-7 * np.maximum(np.abs(a_re), np.abs(b_re))             # PDF quantisation + single-precision output
    tol_i = 3e-4 / scale + 2e-7 * np.abs(a_im)
    if G["ordered"]:      # strict: red = first root (re, im), blue = second root (re, im), signs included
        for got, want, t in ((y[0], a_re, tol), (y[1], a_im, tol_i), (y[2], b_re, tol), (y[3], b_im, tol_i)):
            assert np.all(np.abs(got - want) <= t), float(np.max(np.abs(got - want)))
    else:
        same = (np.abs(y[0] - a_re) <= tol) & (np.abs(y[2] - b_re) <= tol)
        swapped = (np.abs(y[0] - b_re) <= tol) & (np.abs(y[2] - a_re) <= tol)
        assert np.all(same | swapped)
        assert np.all(np.abs(np.abs(y[1]) - np.abs(a_im)) <= tol_i) and np.all(np.abs(np.abs(y[3]) - np.abs(b_im)) <= tol_i)


# ---- the reference author's Mathematica evaluation of the cold roots in the Solov'ev equilibrium ---------------------------------
MMA = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_solovev_nx_roots.json")))


@pytest.mark.parametrize("page", range(len(MMA["pages"])))
def test_oracle_reproduces_the_reference_solovev_root_plots(page):
    """kx_plots_Solovev_90GHz_ECH.pdf (a printed Mathematica notebook of the RAYS author, independent of the Fortran): Re and Im of n_x
    of the plus (O) and minus (X) roots along the equatorial plane of the two Solov'ev example cases.  Every one of the 1 454 plotted
    samples must lie on one of the oracle's four root branches (|Re n_x|, |Im n_x| of plus / minus from solve_n1_vs_n2_n3 on the
    Solov'ev equilibrium) to the PDF's resolution: pins solovev_eq's B(R), n(psi), RLSDP_cold and the n1^2(n3) quadratic with its
    cutoffs, the right-hand resonance and the evanescent stretch."""
    import ctypes as C
    pg = MMA["pages"][page]
    cfg = init_case(pg["namelist"])
    L = orc.load()

    def nx(x, mode):
        cfg.wave_mode = mode
        re, im = C.c_double(0), C.c_double(0)
        rv = np.array([min(max(x, 0.2000001), 1.3999999), 0.0, 0.0])
        assert L.oracle_solve_n1(C.byref(cfg), rv.ctypes.data_as(C.POINTER(C.c_double)), 0.0, pg["n_z"], C.byref(re), C.byref(im)) == 0
        return complex(re.value, im.value)
    devs = []
    for cv in pg["curves"]:
        x, y = np.array(cv["x"]), np.array(cv["y"])
        cand = []
        for mode in (1, 2):
            v = np.array([nx(xi, mode) for xi in x])
            cand.append(np.abs(np.abs(v.real if cv["part"] == "re" else v.imag) - y))
        d = cand[0] if np.median(cand[0]) < np.median(cand[1]) else cand[1]     # the curve is ONE branch
        assert np.median(d) <= 1.2e-4, (cv["part"], len(x), np.median(d))      # PDF resolution: 1e-3 pt = 1.7e-5 in n_x
        devs.append(d)
    d = np.concatenate(devs)
    # steep stretches (next to the resonance / cutoffs) turn the x resolution of the PDF into a larger n_x deviation
    assert np.percentile(d, 99) <= 1.0e-3 and d.max() <= 2.5e-3, (np.percentile(d, 99), d.max())
