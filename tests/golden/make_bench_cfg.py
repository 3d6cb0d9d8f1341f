#!/usr/bin/env python
"""Writes tests/golden/bench_cfg_solovev_fan_1M.json: the marshalled module state (rays_cfg, byte for byte) and the launcher
parameters of bench.py's workload, as the product's host mirror forms them from rays_b200/configs/solovev_fan_1M.in.

`bench.py --impl reference` reads this file instead of parsing the namelist, so that the reference arm (the CPU oracle) runs
without mapping rays_b200/lib/librays_b200.so at all.  tests/test_bench_contract.py checks the file against the live host
mirror, so it cannot go stale unnoticed.  The workload's rays_cfg holds no table pointers (Solov'ev equilibrium, no damping)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "bench_cfg_solovev_fan_1M.json")


def null_pointers(st):
    """table pointers are process-local addresses (and unused by this workload: Solov'ev equilibrium, no damping)"""
    for name, typ in st._fields_:
        v = getattr(st, name)
        if isinstance(v, C.Structure):
            null_pointers(v)
        elif isinstance(v, C._Pointer) or typ in (C.c_void_p, C.c_char_p):
            setattr(st, name, typ())


def snapshot():
    import rays_b200 as rb
    from rays_b200 import _abi
    from _cases import launch_params
    L = _abi.load()
    assert L.rays_host_initialize(rb.config_path("solovev_fan_1M.in").encode(), 0) == 0, L.rays_host_last_error()
    cfg = _abi.Cfg.from_buffer_copy(bytes(rb.host_cfg()))
    assert cfg.damping_model == 0 and cfg.equilib_model == _abi.EQ_SOLOVEV, "this snapshot drops the table pointers"
    null_pointers(cfg)
    _, so, _ = launch_params()
    return {"namelist": "rays_b200/configs/solovev_fan_1M.in", "rays_cfg_hex": bytes(cfg).hex(), "sizeof_rays_cfg": C.sizeof(_abi.Cfg),
            "solovev_launch_hex": bytes(so).hex(), "ray_init_model": L.rays_host_ray_init_model().decode().strip()}


if __name__ == "__main__":
    json.dump(snapshot(), open(OUT, "w"), indent=1)
    print("wrote", OUT)
