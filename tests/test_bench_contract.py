"""bench.py's output contract, checked on the CPU box: the reference arm (the CPU port on the host cores) prints one JSON line with
the keys the driver reads, and the product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--steps", "1", "--warmup", "0", "--width", "320", "--height", "180", "--radius", "4"]


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    r = run_bench("--impl", "reference", *SMALL)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["metric"].startswith("Mrays/s") and d["unit"] == "Mrays/s" and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "generated-terrain" in d["config"]["workload"] and "320x180" in d["config"]["workload"]
    # `config` names the workload and is IDENTICAL in both arms (the driver compares them): it is bench.shared_config(args), nothing else
    import argparse
    import bench
    sys_argv, sys.argv = sys.argv, ["bench.py", "--impl", "reference", *SMALL]
    try:
        assert d["config"] == bench.shared_config(bench.parse())
    finally:
        sys.argv = sys_argv
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present: the product arm runs (covered by the -m gpu tests and the bench itself)")
    r = run_bench(*SMALL)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr) and "no CPU fallback" in (r.stdout + r.stderr)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]   # no metric line from a run that measured nothing
