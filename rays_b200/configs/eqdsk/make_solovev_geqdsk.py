"""Writes solovev.geqdsk: the Solov'ev equilibrium of ../axisym_deposition_fan.in as a g-file, the way the
reference's own converter does it (RAYS_project/solovev_2_eqdsk/solovev_2_eqdsk.f90:60-160: psi on a uniform
(R,Z) grid from solovev_magnetics_psi, T = R*Bphi = bphi0*rmaj, P = TTp = Pp = Q = 0, analytic boundary
contour, one dummy limiter point) in the format of WritegFile (RAYS_lib/eqdsk_utilities_m.f90:114-143).

Sign convention: eqdsk_magnetics_spline_interp takes B_R = +psi_Z/R, B_Z = -psi_R/R
(eqdsk_magnetics_spline_interp_m.f90:236-238), solovev_magnetics the opposite (solovev_magnetics_m.f90:160-166), and the
launcher needs psi to grow outwards (axisym_toroid_ray_init_R_Z_nphi_ntheta_m.f90:173,208).  The file therefore
describes the field of solovev_magnetics with iota0 -> -iota0; tests/test_eqdsk.py compares against that run."""
import os

import numpy as np

rmaj, outer_bound, kappa, bphi0, iota0 = 1.0, 1.4, 1.1, 3.3, 0.01
box_rmin, box_rmax, box_zmin, box_zmax = 0.1, 1.5, -0.7, 0.7
NRBOX, NZBOX, NBOUND = 65, 65, 101

bp0 = bphi0 * iota0
psiB = 0.5 * bp0 * (outer_bound ** 2 - rmaj ** 2) ** 2 / rmaj ** 2 / 4.0
inner_bound = np.sqrt(2.0 * rmaj ** 2 - outer_bound ** 2)
R = box_rmin + (box_rmax - box_rmin) * np.arange(NRBOX) / (NRBOX - 1)
Z = box_zmin + (box_zmax - box_zmin) * np.arange(NZBOX) / (NZBOX - 1)
RR, ZZ = np.meshgrid(R, Z, indexing="xy")          # [j, i]: R fastest, as ((Psi(i,j), i=1,NRBOX), j=1,NZBOX)
psi = 0.5 * bp0 * ((RR * ZZ / (rmaj * kappa)) ** 2 + ((RR ** 2 - rmaj ** 2) ** 2) / rmaj ** 2 / 4.0)

rb, zb = np.zeros(NBOUND), np.zeros(NBOUND)
rb[0] = rb[-1] = inner_bound
rb[(NBOUND - 1) // 2] = outer_bound
dR = 2.0 * (outer_bound - inner_bound) / (NBOUND - 1)
for i in range(2, (NBOUND - 1) // 2 + 1):          # Fortran i = 2 .. (NBOUND-1)/2
    r = inner_bound + i * dR
    zsq = kappa ** 2 / (4.0 * r ** 2) * (outer_bound ** 4 + 2.0 * (r ** 2 - outer_bound ** 2) * rmaj ** 2 - r ** 4)
    rb[i - 1], zb[i - 1] = r, np.sqrt(zsq)
    rb[NBOUND - (i - 1) - 1], zb[NBOUND - (i - 1) - 1] = r, -np.sqrt(zsq)


def e16(v):
    """Fortran e16.9: 0.dddddddddE+xx"""
    if v == 0.0:
        return " 0.000000000E+00"
    e = int(np.floor(np.log10(abs(v)))) + 1
    m = v / 10.0 ** e
    if abs(round(m, 9)) >= 1.0:
        m, e = m / 10.0, e + 1
    return f"{m:12.9f}E{e:+03d}".rjust(16)


def block(vals):
    vals = list(vals)
    return "".join("".join(e16(v) for v in vals[k:k + 5]) + "\n" for k in range(0, len(vals), 5))


out = f"{'Solovev equilibrium (solovev_2_eqdsk layout)':<48s}{0:4d}{NRBOX:4d}{NZBOX:4d}\n"
out += block([box_rmax - box_rmin, box_zmax - box_zmin, rmaj, box_rmin, 0.0])
out += block([rmaj, 0.0, 0.0, psiB, bphi0])
out += block([0.0, 0.0, 0.0, rmaj, 0.0])
out += block([0.0, 0.0, psiB, 0.0, 0.0])
out += block([bphi0 * rmaj] * NRBOX)
for _ in range(3):
    out += block([0.0] * NRBOX)
out += block(psi.ravel())
out += block([0.0] * NRBOX)
out += f"{NBOUND:5d}{1:5d}\n"
out += block(np.stack([rb, zb], axis=1).ravel())
out += block([0.0, 0.0])
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "solovev.geqdsk"), "w").write(out)
print("wrote solovev.geqdsk", NRBOX, NZBOX, "psiB", psiB)
