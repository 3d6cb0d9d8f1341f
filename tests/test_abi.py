"""The C-ABI library loads on a machine without a GPU, exports every symbol include/rays_b200.h declares,
the ctypes mirrors of the structs have the C layout, and the GPU entry points fail loudly (no fallback)."""
import ctypes as C
import os
import re

from rays_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    L = _abi.load()
    hdr = open(os.path.join(ROOT, "include", "rays_b200.h")).read()
    declared = set(re.findall(r"\b(rays_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_abi.ABI_SYMBOLS), declared ^ set(_abi.ABI_SYMBOLS)
    for name in sorted(declared) + _abi.HOST_SYMBOLS:
        assert getattr(L, name) is not None


def test_struct_layouts_match(built):
    L = _abi.load()
    out = (C.c_int32 * 13)()
    assert L.rays_b200_struct_sizes(out, 13) == 13
    for i, st in enumerate(_abi.STRUCTS_IN_SIZE_ORDER):
        assert C.sizeof(st) == out[i], (st.__name__, C.sizeof(st), out[i])


def test_no_cpu_fallback(built):
    """Without a GPU the hot path must refuse to run (RAYS_ERR_CUDA), never fall back."""
    import torch
    if torch.cuda.is_available():
        return
    L = _abi.load()
    assert L.rays_b200_init(0) == 2
    assert b"no CUDA device" in L.rays_b200_last_error()
    assert L.rays_b200_trace_device(1) == 3          # not initialised
    assert L.rays_b200_set_config(None) == 3


def test_stop_strings_are_the_references(built):
    """strings verbatim, blank padded to 60."""
    import rays_b200 as rb
    want = {1: "sout > s_max", 2: " nstep > nstep_max", 3: "infinite Vg", 4: "ray stalled", 5: "dispersion_residual", 6: "infinite_Vg",
            7: "total_absorption", 8: "ODE total error", 9: "step number .ge. maxnum", 10: "equations stiff", 20: "x out_of_bounds",
            23: "R out_of_box", 24: "z out_of_box", 25: "R_out_of_box", 26: "Z_out_of_box", 27: "out_of_plasma", 28: "R out_of_bounds",
            30: "negative_dens", 31: "negative_temp", 0: ""}
    for code, s in want.items():
        got = rb.stop_string(code)
        assert len(got) == 60 and got == s.ljust(60)


def test_product_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under rays_b200/ or include/ may reference it."""
    bad = []
    for base in ("rays_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle/|_oracle|rays_oracle|librays_oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
