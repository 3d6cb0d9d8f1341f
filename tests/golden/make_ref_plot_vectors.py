"""Golden vectors from the reference's OWN ray output: the example directories ship matplotlib-written vector
PDFs of the rays the Fortran code traced (examples_RAYS/*/ray_plots.*.pdf, plotted by graphics_RAYS/plot_RAYS_*.py
from run_results.<label>.nc).  A PDF polyline keeps the plotted trajectory points with 1e-6 pt resolution
(~2.5e-9 m on these axes), so the figures pin positions along every example ray -- including the end point
after 500-1000 Shampine-Gordon segments -- far below the north-star tolerances.

This script (run in the build container, where /root/reference exists) extracts, per figure page: the axis
calibration from the tick marks + tick labels, and every ray polyline in data coordinates; it also writes the
example namelists translated to the current namelist names (the shipped inputs predate three renames:
`message_unit` and `b0` were dropped, `t0s` became `t0s_eV`, `ray_deriv_name` was added) into rays_b200/configs/examples/.

    python tests/golden/make_ref_plot_vectors.py   ->  tests/golden/ref_plot_vectors.json
"""
import json
import os
import re
import zlib

import numpy as np

REF = "/root/reference/examples_RAYS"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
NUM = r"-?\d+(?:\.\d+)?"

FIGURES = [   # (pdf, example input, what the two plotted coordinates are per page)
    ("ECH_90GHz_slab/pdf_plots/ray_plots.run_1.pdf", "ECH_90GHz_slab/slab_ECH_90GHz_case_1.in", [("z", "x")]),
    ("ECH_90GHz_slab/pdf_plots/ray_plots.run_2.pdf", "ECH_90GHz_slab/slab_ECH_90GHz_case_2.in", [("z", "x")]),
    # cases 3 and 4 share one figure (rays 1-3: case 3, rays 4-6: case 4); their inputs are not shipped: RECONSTRUCTED below
    ("ECH_90GHz_slab/pdf_plots/ray_plots.runs_3_and_4.pdf", ("slab_ECH_90GHz_case_3.in", "slab_ECH_90GHz_case_4.in"), [("z", "x")]),
    ("ECH_90GHz_solovev_SG_eq_plane/ray_plots.plus_root.pdf", "ECH_90GHz_solovev_SG_eq_plane/solovev_ECH_90GHz_plus_root.in", [("r", "z"), ("x", "y")]),
    ("ECH_90GHz_solovev_SG_eq_plane/ray_plots.minus_root.pdf", "ECH_90GHz_solovev_SG_eq_plane/solovev_ECH_90GHz_minus_root.in", [("r", "z"), ("x", "y")]),
    ("ECH_90GHz_solovev_SG_eq_plane/ray_plots.minus_root_2.pdf", "ECH_90GHz_solovev_SG_eq_plane/solovev_ECH_90GHz_minus_root_case_2.in", [("r", "z"), ("x", "y")]),
]


def pages(pdf):
    d = open(pdf, "rb").read()
    out = []
    for s in re.findall(rb"stream\r?\n(.*?)\r?\nendstream", d, re.S):
        try:
            t = zlib.decompress(s).decode("latin1")
        except Exception:
            continue
        if " re W n" in t and "Tj" in t:
            out.append(t)
    return out


def parse_page(t):
    """-> list of events in drawing order: ('path', op, [(x, y)...]) and ('text', string, has_minus_sign)"""
    lines = [l.strip() for l in t.split("\n")]
    ev, cur, placed = [], [], False
    for k, L in enumerate(lines):
        m = re.fullmatch(rf"({NUM}) ({NUM}) ([ml])", L)
        if m:
            p = (float(m.group(1)), float(m.group(2)))
            cur = [p] if m.group(3) == "m" else cur + [p]
            continue
        if L in ("S", "B", "f"):
            if cur:
                ev.append(("path", L, cur))
                cur = []
            continue
        if re.fullmatch(rf"1 0 -0 1 {NUM} {NUM} cm", L):
            placed = True          # labels with a minus sign are placed through a cm matrix ...
        m = re.fullmatch(r"\((.*)\) Tj(?: ET)?", L)
        if m:
            s = m.group(1).replace("\\(", "(").replace("\\)", ")")
            minus = placed and any("minus Do" in x for x in lines[k + 1:k + 6])   # ... and the sign is an XObject
            ev.append(("text", s, minus))
            placed = False
    return ev


def calibrate(ev):
    """tick mark (2-point path leaving the axes box) followed by its label -> linear map pt -> data, per axis"""
    xt, yt = [], []
    for a, b in zip(ev, ev[1:]):
        if a[0] == "path" and len(a[2]) == 2 and b[0] == "text" and re.fullmatch(r"\d+(\.\d+)?", b[1]):
            (x0, y0), (x1, y1) = a[2]
            v = float(b[1]) * (-1.0 if b[2] else 1.0)
            if x0 == x1 and abs(y1 - y0) == 3.5:
                xt.append((x0, v))
            elif y0 == y1 and abs(x1 - x0) == 3.5:
                yt.append((y0, v))
    fits = []
    for tk in (xt, yt):
        p, v = np.array(tk).T
        a, b = np.polyfit(p, v, 1)
        assert len(tk) >= 3 and np.max(np.abs(a * p + b - v)) < 2e-8 / 1.0 + 1e-6 * abs(a), (tk, a, b)
        fits.append((float(a), float(b)))
    return fits


def translate_namelist(src):
    t = open(src).read()
    t = t[:t.rindex("/") + 1] + "\n"                       # text after the last group ('NSTX') is not namelist input
    t = re.sub(r"^\s*message_unit\s*=.*\n", "", t, flags=re.M)
    t = re.sub(r"^\s*b0\s*=.*\n", "", t, flags=re.M)
    t = t.replace("t0s(", "t0s_eV(")
    t = re.sub(r"verbosity\s*=\s*\d+", "verbosity=0", t)
    # ray_deriv_name (ode_m.f90:95,107) is newer than the examples and has no default: they ran the analytic derivatives
    t = re.sub(r"^(\s*ode_solver_name\s*=.*\n)", r"\1  ray_deriv_name='cold'\n", t, flags=re.M)
    return t


# Cases 3 and 4 of the slab example (README: constant density, linearly increasing Bz with the cyclotron resonance at x = 0,
# nz = 0.1, 0.4, 0.7; case 3 = X mode launched below the right-hand cutoff, case 4 = X mode launched above the cyclotron
# resonance) ship figures but no input files.  The inputs were reconstructed from the figures themselves: the kx-profile
# pages of kx_plots.run_3.pdf fix the equilibrium -- n0 = 4.0e19, bz0 = 3.2151 (the resonant field to five digits) and
# LBz_scale = 0.53585 are the only values for which all 404 plotted root values of a page, including the 45 185 rad/m spike
# next to the upper-hybrid resonance, agree to 1e-6 pt (a 1e-5 change of LBz_scale moves that spike by 1 %) -- and the ray
# figure fixes the launch (x = -0.45 / +0.45, z = -0.6), the root (minus / plus), k0_sign (+1 / -1) and the stepper settings
# (SG_ODE, ds = 5e-11, nstep_max = 500, tolerances 1e-4 as in case 1: with 1e-8 the rays miss by 10^3 resolutions).
def reconstructed_case(case1_text, which):
    t = case1_text
    edits = [("run_label='run_1'", "run_label='run_%d'" % which), ("n0=1.0e20", "n0=4.0e19"),
             ("bz_prof_model='constant'", "bz_prof_model='linear'"), ("bz0=1.286", "bz0=3.2151"), ("LBz_scale = 1.125", "LBz_scale = 0.53585"),
             ("dens_prof_model='linear'", "dens_prof_model='constant'"), ("rindex_z0=0.4", "rindex_z0=0.1"), ("delta_rindex_z0=0.1", "delta_rindex_z0=0.3")]
    edits += [("x_launch0= -0.08", "x_launch0= -0.45")] if which == 3 else \
             [("x_launch0= -0.08", "x_launch0= 0.45"), ("wave_mode='minus'", "wave_mode='plus'"), ("k0_sign = 1", "k0_sign = -1")]
    for old, new in edits:
        assert old in t, old
        t = t.replace(old, new)
    return "! RECONSTRUCTED from the reference's figures (tests/golden/make_ref_plot_vectors.py): the example ships no input for this case\n" + t


def main():
    out = {"_doc": "made by tests/golden/make_ref_plot_vectors.py from the reference's example PDFs; coordinates in metres",
           "figures": []}
    for pdf, nml, axes in FIGURES:
        if isinstance(nml, tuple):       # the shared figure of the reconstructed cases 3 and 4
            case1 = translate_namelist(os.path.join(REF, "ECH_90GHz_slab/slab_ECH_90GHz_case_1.in"))
            for k, nm in enumerate(nml):
                open(os.path.join(ROOT, "rays_b200", "configs", "examples", nm), "w").write(reconstructed_case(case1, 3 + k))
            name = None
        else:
            name = os.path.basename(nml)
            open(os.path.join(ROOT, "rays_b200", "configs", "examples", name), "w").write(translate_namelist(os.path.join(REF, nml)))
        pg = pages(os.path.join(REF, pdf))
        assert len(pg) == len(axes), (pdf, len(pg))
        for t, (hname, vname) in zip(pg, axes):
            ev = parse_page(t)
            (ax, bx), (ay, by) = calibrate(ev)
            # '(x10^-5)' multiplier on the z axis of the r-z page of the equatorial-plane runs (z stays within 1e-5 m of 0)
            texts = [e[1] for e in ev if e[0] == "text"]
            offset_text = "×" in texts
            if offset_text:
                k = texts.index("×")
                assert texts[k + 1:k + 4] == ["1", "0", "5"] and vname == "z", texts[k:k + 5]
                ay, by = ay * 1e-5, by * 1e-5
            rays = []
            for e in ev:
                if e[0] == "path" and e[1] == "S" and len(e[2]) > 2 and e[2][0] != e[2][-1]:
                    P = np.array(e[2])
                    rays.append({"h": (ax * P[:, 0] + bx).tolist(), "v": (ay * P[:, 1] + by).tolist()})
            title = [e[1] for e in ev if e[0] == "text" and "geometry" in e[1]]
            groups = [(name, rays)] if name else [(nml[0], rays[:3]), (nml[1], rays[3:])]
            for nm, rr in groups:
                out["figures"].append({"pdf": "examples_RAYS/" + pdf, "namelist": "examples/" + nm, "title": title[0] if title else "",
                                       "h": hname, "v": vname, "v_has_multiplier": bool(offset_text),
                                       "quantum_h": abs(ax) * 1e-6, "quantum_v": abs(ay) * 1e-6, "rays": rr})
            print(pdf, hname, vname, len(rays), "rays", [len(r["h"]) for r in rays], "multiplier" if offset_text else "")
    json.dump(out, open(os.path.join(HERE, "ref_plot_vectors.json"), "w"))


if __name__ == "__main__":
    main()
