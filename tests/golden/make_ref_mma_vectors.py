"""Golden vectors for the cold dispersion roots in the Solov'ev equilibrium from the reference author's own, independent
evaluation: examples_RAYS/ECH_90GHz_solovev_SG_eq_plane/kx_plots_Solovev_90GHz_ECH.pdf is a printed Mathematica notebook
(Solovev_examples_temp.nb) that plots Re and Im of n_x of the plus and minus roots of the 4th-order cold-plasma dispersion relation
along the equatorial plane (0.2 m <= x <= 1.4 m) for the two example cases (n0 = 5e19, n_z = 0.3 and n0 = 7.5e19, n_z = 0.4), with the
Solov'ev profiles B = b0 rMaj / x, n = n0 (1 - psi_N) of the RAYS inputs next to it.  The curves are vector polylines (adaptive
sampling, 1e-3 pt); the tick labels are text, so both axes are calibrated from the page itself: x from the first / last sample at the
plot range [0.2, 1.4], y from the tick-label spacing (57.198 pt per unit) and the Im = 0 line.

    python tests/golden/make_ref_mma_vectors.py   ->  tests/golden/ref_solovev_nx_roots.json
"""
import json
import os
import re
import zlib

import numpy as np

PDF = "/root/reference/examples_RAYS/ECH_90GHz_solovev_SG_eq_plane/kx_plots_Solovev_90GHz_ECH.pdf"
HERE = os.path.dirname(os.path.abspath(__file__))
# page -> (translated example namelist, n_z of the notebook's data set, y (pt) of the n_x = 0 line)
PAGES = [(0, "examples/solovev_ECH_90GHz_minus_root.in", 0.3), (1, "examples/solovev_ECH_90GHz_minus_root_case_2.in", 0.4)]
X0_PT, X_PT_PER_M, Y_PT_PER_UNIT = 153.418, 272.956, 57.198


def page_streams():
    d = open(PDF, "rb").read()
    out = []
    for s in re.findall(rb"stream\r?\n(.*?)\r?\nendstream", d, re.S):
        try:
            out.append(zlib.decompress(s).decode("latin1"))
        except Exception:
            pass
    return out


def curves(t):
    paths, cur, nums, col = [], [], [], None
    for tk in t.split():
        try:
            nums.append(float(tk))
            continue
        except ValueError:
            pass
        if tk == "m":
            cur = [tuple(nums[-2:])]
        elif tk == "l":
            cur.append(tuple(nums[-2:]))
        elif tk == "S":
            if len(cur) > 2:
                paths.append((col, np.array(cur)))
            cur = []
        elif tk in ("RG", "rg"):
            col = tuple(nums[-3:])
        nums = []
    return paths


def main():
    st = page_streams()
    out = {"_doc": "made by tests/golden/make_ref_mma_vectors.py from kx_plots_Solovev_90GHz_ECH.pdf (Mathematica); x in m, y = |Re n_x| or |Im n_x|", "pages": []}
    for pi, nml, nz in PAGES:
        cs = curves(st[pi])
        y0 = max(p[:, 1].max() for c, p in cs if c[2] > 0.99)        # the pure-blue Im curve rests on n_x = 0 (y grows downwards)
        page = {"pdf": "examples_RAYS/ECH_90GHz_solovev_SG_eq_plane/kx_plots_Solovev_90GHz_ECH.pdf", "page": pi, "namelist": nml, "n_z": nz, "curves": []}
        for c, p in cs:
            x = 0.2 + (p[:, 0] - X0_PT) / X_PT_PER_M
            y = (y0 - p[:, 1]) / Y_PT_PER_UNIT
            keep = y < 1.2        # the plot range ends at n_x = 1.27: samples next to the resonance are clipped there
            page["curves"].append({"part": "im" if c[2] > 0.99 else "re", "x": [float(v) for v in x[keep]], "y": [float(v) for v in y[keep]]})
        out["pages"].append(page)
    json.dump(out, open(os.path.join(HERE, "ref_solovev_nx_roots.json"), "w"))
    print(sum(len(c["x"]) for pg in out["pages"] for c in pg["curves"]), "points")


if __name__ == "__main__":
    main()
