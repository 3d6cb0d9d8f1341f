"""bench.py's reference arm (the CPU implementation of the path on the host cores) runs without a GPU and prints the
contract's JSON line; the product arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(built):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--rays", "2048", "--cpu-rays", "256"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ray_steps_per_sec" and line["unit"] == "ray-steps/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["config"]["workload"] == "solovev_fan_1M"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "rays" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the sampled value is validated by one trace of the whole fan
    assert line["full_fan"]["rays"] == line["config"]["rays_per_gpu"] and line["full_fan"]["value"] > 0
    # both arms print the same config keys (the driver compares them)
    assert set(line["config"]) == {"workload", "rays_per_gpu", "ray_steps_per_fan", "ode", "ray_deriv", "ds", "nstep_max", "nv", "sharding", "l2"}


def test_reference_arm_does_not_map_the_product_library(built):
    """the reference arm is the CPU oracle alone: it must run with rays_b200/lib/librays_b200.so out of reach"""
    code = ("import sys, os; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--rays', '2048', '--cpu-rays', '256', '--no-cpu'];"
            "import runpy; runpy.run_path(os.path.join(%r, 'bench.py'), run_name='__main__');"
            "maps = open('/proc/self/maps').read(); assert 'librays_b200' not in maps, 'product library mapped'; assert 'librays_oracle' in maps" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]


def test_bench_cfg_snapshot_is_current(built):
    """tests/golden/bench_cfg_solovev_fan_1M.json (what the reference arm traces) == what the host mirror makes of the namelist today"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_bench_cfg
    live = make_bench_cfg.snapshot()
    saved = json.load(open(make_bench_cfg.OUT))
    assert live == saved, "run tests/golden/make_bench_cfg.py"


def test_product_arm_needs_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--rays", "2048", "--no-cpu"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "ray_steps_per_sec" not in out.stdout
