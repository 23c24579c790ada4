"""CPU tests (-m "not gpu") of bench.py's contract: the reference arm runs the unmodified reference (oracle/_ref) on the
host cores and prints the one JSON line the driver parses; the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

from cpupathtrace_b200 import REPO_ROOT, pth

BENCH = os.path.join(REPO_ROOT, "bench.py")
SMALL = ["--steps", "1", "--warmup", "0", "--width", "96", "--height", "64", "--spp", "4", "--mesh", "40x30", "--cpu-seconds", "0.5"]


def test_reference_arm_prints_the_contract_line():
    if not os.path.exists(pth.REF_FAST):
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", *SMALL], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1  # stdout carries the JSON line and nothing else
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == line["unit"] == "Msamples/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0 and line["ms_per_step"] > 0
    assert line["vs_baseline"] is None and line["dtype"] == "f32" and line["data"] == "synthetic" and "workload" in line["config"]
    base = line["cpu_baseline"]
    assert base["kind"] == "reference" and base["cores"] >= 1 and base["value"] == line["value"] and base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, BENCH, *SMALL], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith("{")]  # and no number is printed
