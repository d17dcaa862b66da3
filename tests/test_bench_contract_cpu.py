"""bench.py's output contract: ONE JSON line on stdout with the keys the driver reads.  The reference arm (CPU oracle)
runs here; our arm needs a GPU, so its keys are checked on the line committed under profiles/."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["metric"].split()[0] in base["metric"]
    assert d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("C2")
    assert d["value"] > 0 and abs(d["value"] - 128 * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_committed_gpu_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_n1.json")))
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["impl"] == "ours" and d["n_gpus"] == 1 and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert d["gpu_launches"] == d["steps"] * d["config"]["launches_per_step"] > 0
    assert abs(d["value"] - 128 * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] > 3 * 128 * 3 * 32 * 32 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    ro = d["roofline"]
    assert ro["bound"] in ("hbm", "tensor") and ro["unit"] in ("GB/s", "TFLOP/s") and abs(ro["frac"] - ro["achieved"] / ro["peak"]) < 1e-9
    assert ro["traffic"] is None or ro["traffic"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and not d["clocks"]["reasons"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "images/s" and cb["value"] > 0


def test_launch_summary_reads_the_committed_ncu_list():
    """profiles/r01_launches_n1.csv (ncu launch list of three steady-state steps) -> per-family table"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(ROOT, "profiles", "r01_launches_n1.csv"),
                        "357"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    assert "3.00 steps of 357 launches" in r.stdout
    assert "conv fprop/dgrad (sv_igemm_fprop)" in r.stdout and "BatchNorm backward" in r.stdout
