"""bench.py's CPU arm runs without a GPU and prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2",
                          "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "DOF-stage-updates/s" and line["unit"] == "DOF-stage-updates/s"
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["dtype"] == "f64"
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == line["value"]
    assert "workload" in line["config"] and "model" not in line["config"]


def test_algorithmic_bytes_match_the_survey_formulas():
    sys.path.insert(0, ROOT)
    import bench
    total, elem, edge = bench.algorithmic_bytes(4)
    assert total == 3960.0 and abs(elem + edge - total) < 1e-9          # B_inv(4) = 3,960 B
    assert round(bench.algorithmic_bytes(2)[0], 1) == 2193.6             # B_inv(2) = 2,194 B
    assert bench.algorithmic_bytes_visc(4) == 10248.0                    # B_visc(4) = 10,248 B


def test_cpu_arm_covers_the_dissipation_workload():
    """The CPU arm of the shock-capturing workload (c3): the C port's PerssonC0 path on a bounded 4:1 Sod tube sample."""
    sys.path.insert(0, ROOT)
    import bench
    r = bench.cpu_run(2, steps=1, warmup=1, nx=80, dissipation=True)
    assert r["kind"] == "port" and r["value"] > 0 and r["cores"] >= 1
    assert "PerssonC0" in r["sample"] and "80x20x2=3200 triangles" in r["sample"]
