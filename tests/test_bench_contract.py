"""The bench.py JSON contract, checked on the last line recorded on a B200 (profiles/) and on the
reference arm run here on the CPU."""
import glob
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"]


def test_recorded_b200_line_has_every_contract_key():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_upwind512.json")))
    assert files, "no recorded bench line under profiles/"
    line = [l for l in open(files[-1]) if l.startswith("{")][-1]
    j = json.loads(line)
    for k in REQUIRED:
        assert k in j, k
    assert j["unit"] == "GCUPS" and j["dtype"] == "f64" and j["higher_is_better"] is True
    assert j["vs_baseline"] is None  # BASELINE.md publishes no GCUPS figure
    assert "workload" in j["config"] and "model" not in j["config"]
    r = j["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = j["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] == 8 * 512 ** 3 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < j["value"]  # host copies inside the timed region
    assert j["gpu_launches"] > 0
    assert j["clocks"] is None or "sm_mhz" in j["clocks"]


def test_reference_arm_single_process():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--workload", "upwind128", "--tsteps", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert p.returncode == 0, p.stderr
    j = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert j["impl"] == "reference" and j["gpu_launches"] == 0 and j["value"] > 0
    assert j["cpu_baseline"]["value"] == j["value"] == j["e2e"]["value"]


def test_recorded_b200_laplacian_line_and_its_reference_arm():
    """BASELINE config 4 (`bench.py --workload lap1024`): the line recorded on a B200 and the reference arm here."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_lap1024.json")))
    assert files, "no recorded laplacian bench line under profiles/"
    j = json.loads([l for l in open(files[-1]) if l.startswith("{")][-1])
    for k in REQUIRED:
        assert k in j, k
    assert j["unit"] == "GCUPS" and j["dtype"] == "f64" and "laplacian" in j["metric"]
    assert j["config"]["kernel"] == "lap7_fused2_kernel" and j["config"]["applies_per_sweep"] == 2
    r = j["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    assert j["e2e"]["h2d_bytes_per_step"] == 8 * 1024 ** 3 and j["e2e"]["value"] < j["value"]
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["gpu_launches"] > 0
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--workload", "lap1024"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr
    ref = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert ref["impl"] == "reference" and ref["metric"] == j["metric"] and ref["unit"] == j["unit"]
    assert ref["value"] > 0 and ref["cpu_baseline"]["value"] == ref["value"] == ref["e2e"]["value"]
