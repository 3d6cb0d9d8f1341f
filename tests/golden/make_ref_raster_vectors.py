"""Golden vectors for the RK4 + mirror-equilibrium path from the reference's own output: the MPEX example directory
ships a raster plot (matplotlib PNG, z-y plane) of the 11 rays the Fortran code traced with RK4_ODE through the
spline-interpolated mirror field.  Resolution 0.87 mm per pixel on ~0.2 m long rays -- coarse next to the vector
PDFs of the slab/Solov'ev examples (make_ref_plot_vectors.py), but it is the reference's result for this path.

Extracts the ray-coloured pixels (saturated colours that are not the plot's pure blue / green / red guide lines)
inside the window the rays occupy, in data coordinates; the axes frame (1-px spines) carries the axis limits
z in [2.8, 3.6], y in [-0.2, 0.2] (set_XY_lim).  Also copies the example's input next to our configs.

    python tests/golden/make_ref_raster_vectors.py   ->  tests/golden/ref_raster_vectors.json
"""
import json
import os
import shutil

import numpy as np
from PIL import Image

REFROOT = "/root/reference/examples_RAYS/MPEX_examples"
CASES = [   # (example directory, our config directory, ray_init file, plotted plane)
    ("MPX_2nd_harm_11_rays_nz_delta_d_0.05_psiP_0.05", "mpex_nz", "ray_init_2nd_harm_11_rays_nz.in", "zy"),
    ("MPX_2nd_harm_11_rays_nz_30deg_delta_d_0.05_psiP_0.05", "mpex_nz_30deg", "ray_init_2nd_harm_11_rays_nz_30deg.in", "zy"),
    ("MPX_2nd_harm_11_rays_nx_delta_d_0.05_psiP_0.05", "mpex", "ray_init_2nd_harm_11_rays_nx.in", "xy"),
    ("MPX_2nd_harm_11_rays_x_delta_d_0.05_psiP_0.05", "mpex_x", "ray_init_2nd_harm_11_rays_x.in", "xy"),
]
# axis limits (set_XY_lim) and the window the rays occupy, clear of the guide lines (last flux surface / cutoff circle)
# `sat`: minimum colour saturation of a ray pixel.  The z-y figures also carry black dashed flux lines (grey when blended), the
# x-y figures only the cutoff circle -- and their rays overlap for most of their length, which washes the colours out.
PLANES = {"zy": dict(h="z", v="y", lim=(2.8, 3.6, -0.2, 0.2), win=(3.15, 3.31, -0.1165, -0.02), sat=60),
          "xy": dict(h="x", v="y", lim=(-0.2, 0.2, -0.2, 0.2), win=(-0.05, 0.05, -0.1005, -0.0285), sat=20)}
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def one_case(exdir, cfgname, ray_init, plane):
    REF = os.path.join(REFROOT, exdir)
    cfgdir = os.path.join(ROOT, "rays_b200", "configs", cfgname)
    os.makedirs(cfgdir, exist_ok=True)
    shutil.copyfile(os.path.join(REF, ray_init), os.path.join(cfgdir, ray_init))
    os.chmod(os.path.join(cfgdir, ray_init), 0o644)
    if cfgname != "mpex":   # (configs/mpex/ holds the nx example as shipped, field file included; the others point to that file)
        txt = open(os.path.join(REF, "rays.in")).read()
        open(os.path.join(cfgdir, "rays.in"), "w").write(txt.replace("mirror_field_NC_file = 'Brz_fields", "mirror_field_NC_file = '../mpex/Brz_fields"))

    im = np.array(Image.open(os.path.join(REF, "Ray_trajectories.png")).convert("RGB")).astype(int)
    dark = im.max(2) < 140
    cs, rs = dark.sum(0), dark.sum(1)
    cols = [i for i in range(im.shape[1]) if cs[i] > 0.8 * cs.max()]
    rows = [i for i in range(im.shape[0]) if rs[i] > 0.8 * rs.max()]
    left, right, top, bottom = cols[0], cols[-1], rows[0], rows[-1]          # the four 1-px spines
    P = PLANES[plane]
    zmin, zmax, ymin, ymax = P["lim"]
    px_z, px_y = (zmax - zmin) / (right - left), (ymax - ymin) / (bottom - top)
    r, g, b = im[:, :, 0], im[:, :, 1], im[:, :, 2]
    sat = (im.max(2) - im.min(2) > P["sat"]) | ((P["sat"] < 60) & (im.max(2) < 200))   # x-y: dark blends count too
    # guide lines are drawn in pure blue / red / green and anti-aliased against white: (t, t, 255), (255, t, t), (t, g, t)
    guide = ((b >= 235) & (np.abs(r - g) < 20) & (b > r + 10)) | ((r >= 245) & (np.abs(g - b) < 15)) | ((np.abs(r - b) < 6) & (g > r + 40) & (g < 140 + r // 2))
    ys, xs = np.nonzero(sat & ~guide)
    z = zmin + (xs - left) * px_z
    y = ymax - (ys - top) * px_y
    w = P["win"]
    win = (z > w[0]) & (z < w[1]) & (y > w[2]) & (y < w[3])
    out = {"png": "examples_RAYS/MPEX_examples/" + exdir + "/Ray_trajectories.png",
           "namelist": cfgname + "/rays.in", "h": P["h"], "v": P["v"], "window": list(w), "pixel_h": px_z, "pixel_v": px_y,
           "frame_px": [int(left), int(right), int(top), int(bottom)],
           "pixels_h": [round(float(v), 6) for v in z[win]], "pixels_v": [round(float(v), 6) for v in y[win]]}
    print(exdir, "frame", left, right, top, bottom, "pixel", px_z, px_y, "ray pixels", int(win.sum()))
    return out


def main():
    out = {"_doc": "made by tests/golden/make_ref_raster_vectors.py; ray-coloured pixel centres of the reference's MPEX figures, metres",
           "figures": [one_case(*c) for c in CASES]}
    json.dump(out, open(os.path.join(HERE, "ref_raster_vectors.json"), "w"))


if __name__ == "__main__":
    main()
