"""bench.py's output contract, checked on the CPU arm (the GPU arm needs a device): exactly one JSON line on stdout with the
keys the driver reads, whatever libraries print along the way."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "queries/s" and d["n_gpus"] == 1
    for key in ("metric", "value", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["config"]["workload"].startswith("10k x 128") and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "[bench r0]" in out.stderr          # the log goes to stderr


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1", "--steps", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    assert out.stdout.strip() == ""            # nothing that could be mistaken for a measurement


def test_roofline_record_carries_the_random_gather_ceiling():
    """The copy peak is the roofline of long rows only: every roofline record also states what random gathers of rows of that
    length reach on the part (committed microbenchmarks under profiles/) and the fraction of THAT."""
    sys.path.insert(0, ROOT)
    import bench
    for row_bytes, lo, hi in ((3072, 6000, 8000), (512, 3000, 4500), (128, 1000, 2500)):
        g = bench.gather_ceiling(row_bytes)
        assert g and g["measured_row_bytes"] == row_bytes and lo < g["gbs"] < hi, g
        assert g["source"].startswith("profiles/")
    r = bench.roofline_of(12e9, 11.9e9, 5.2, {"hbm_gbs": 6546.9}, row_bytes=512)
    assert abs(r["achieved"] - 12e9 / 5.2e-3 / 1e9) < 0.1 and r["peak"] == 6546.9 and abs(r["frac"] - r["achieved"] / 6546.9) < 1e-3
    assert abs(r["random_gather_ceiling"]["frac"] - r["achieved"] / r["random_gather_ceiling"]["gbs"]) < 1e-3
    assert "random_gather_ceiling" not in bench.roofline_of(1e9, 1e9, 1.0, {})     # no row length, no ceiling


def test_deadline_guard_prints_the_line_once_and_marks_the_unfinished_extra():
    """The extras run after the headline was measured but before the one JSON line is printed: when they outlast the run's
    deadline the guard prints the line as it stands — once — and the process leaves with status 0."""
    code = r'''
import json, os, sys, time
sys.path.insert(0, %r)
import bench
os.environ["HB_BENCH_DEADLINE"] = "1.5"
line = {"metric": "m", "value": 1.0, "other_workloads": {"c1": {"value": 2.0}}}
bench.arm_deadline_guard(time.time(), 0, lambda: line, "sharded")
time.sleep(30)            # an extra that never finishes
bench.emit({"late": True})
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["value"] == 1.0 and d["other_workloads"] == {"c1": {"value": 2.0}} and "skipped" in d["sharded"]
