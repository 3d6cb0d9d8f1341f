"""Extracts the reference's own known-answer table for the plasma Z function into zfun_kat.json.

Source: /root/reference/RAYS_project/math_functions_lib/"Splined Z function results.txt", lines 48-85
(output of test_zfun.f90 for zfun_real_arg_D at integer x in [-15, 15], 13 digits) and the Mathematica
values the author pasted below it (lines 86-127).  Run in the build container (the reference tree is
not present on the GPU box); the JSON is committed."""
import json
import re

SRC = "/root/reference/RAYS_project/math_functions_lib/Splined Z function results.txt"
lines = open(SRC).read().splitlines()
fortran = []
for ln in lines:
    m = re.match(r"\s*z real =\s*(-?[\d.]+)\s+zfun_D\s+=\s+(-?[\d.]+E[+-]\d+)\s+(-?[\d.]+E[+-]\d+)", ln)
    if m:
        fortran.append({"x": float(m.group(1)), "re": float(m.group(2)), "im": float(m.group(3))})
# only the first table ("Results for real argument, zfun_real_arg_D") -- 31 entries
fortran = fortran[:31]
assert len(fortran) == 31 and fortran[0]["x"] == -15.0 and fortran[-1]["x"] == 15.0
# Mathematica block: full-precision entries for x = 6..15
math = []
for ln in lines:
    m = re.match(r"\s*(-0\.\d{16})\s+(\d\.\d+)\*10\^(-\d+) I", ln)
    if m:
        math.append({"re": float(m.group(1)), "im": float(m.group(2)) * 10.0 ** int(m.group(3))})
assert len(math) == 10
for i, e in enumerate(math):
    e["x"] = float(6 + i)
json.dump({"source": "RAYS_project/math_functions_lib/Splined Z function results.txt:48-85 (Fortran zfun_real_arg_D) and :86-127 (Mathematica)",
           "zfun_real_arg_D": fortran, "mathematica_x6_15": math}, open(__file__.replace("make_zfun_kat.py", "zfun_kat.json"), "w"), indent=1)
print(len(fortran), len(math))
