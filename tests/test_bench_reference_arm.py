"""The reference arm of bench.py (`--impl reference`) on the CPU: one JSON line with the keys the driver reads, the same `config`
dictionary as the GPU arm, and a process that never maps the product library (the arm times the oracle restatement -- the real Spheral
is unbuildable here -- so the product must not be on its path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line_and_never_loads_the_product():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-sample', '16'];\n"
            "try:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "maps = open('/proc/self/maps').read()\n"
            "sys.stderr.write('PRODUCT_LOADED=%%d\\n' %% ('libsphb200' in maps))\n" % os.path.join(ROOT, "bench.py"))
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["metric"].startswith("particle-updates/sec") and d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"]) > 0
    assert "Noh-spherical-3d ASPH" in d["config"]["workload"] and d["config"]["particles"] == 8000000       # the GPU arm's configuration
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "16^3" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "PRODUCT_LOADED=0" in p.stderr, "the reference arm mapped libsphb200.so"
