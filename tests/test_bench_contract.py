"""bench.py's contract that can be checked without a GPU: the reference arm (the unmodified reference on the host cores)
prints one JSON line with the keys the driver reads, and the product arm refuses loudly when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

import parity as P

BENCH = os.path.join(P.REPO, "bench.py")


def test_reference_arm_prints_one_json_line():
    if not P.have_ref():
        pytest.skip("oracle/_ref/ref_driver not built")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mpixels_per_s" and d["unit"] == "Mpixels/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(d["config"]) == {"workload", "width", "height", "gpus"} and "sample" in d["cpu_baseline"]
    assert (d["config"]["width"], d["config"]["height"]) == (1280, 800)


def test_reference_arm_of_the_scaling_workload_uses_the_same_config():
    """--gpus N > 1 defaults to C5; the reference arm reports C5's config and says which bounded sample it timed."""
    if not P.have_ref():
        pytest.skip("oracle/_ref/ref_driver not built")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    assert d["impl"] == "reference" and d["n_gpus"] == 2
    assert (d["config"]["width"], d["config"]["height"], d["config"]["gpus"]) == (7680, 4320, 2) and "10 008 338" in d["config"]["workload"]
    assert "c5_golden" in d["cpu_baseline"]["sample"] and d["value"] > 0


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, BENCH, "--workload", "c1", "--steps", "1", "--no-cpu-baseline"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode != 0 and not r.stdout.strip(), "the product arm must not produce a number without the CUDA path"
