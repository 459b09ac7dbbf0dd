"""CPU suite: the reference arm of bench.py (`--impl reference`: the reference's CPU path through oracle/_ref + the C
port, no GPU) prints the contract's JSON line, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--nodes", "2000", "--edges", "60000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "ms" and line["higher_is_better"] is False
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0
    assert line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert "workload" in line["config"] and "model" not in line["config"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_has_no_cpu_fallback():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--nodes", "500",
                        "--edges", "5000", "--no-generated", "--no-kernels", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0          # fails loudly: nothing on this path runs on the host
