"""GPU test: the libtorch shim (host/gala_b200_torch.h), driven by a C++ program that spells
its calls exactly like a generated gala.cu (host/shim_selftest.cpp), against dense torch math."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "gala-gnn-acceleration-language_b200", "host", "shim_selftest")


def test_libtorch_shim_selftest():
    if not os.path.exists(BIN):
        pytest.skip("host/shim_selftest not built (make -C gala-gnn-acceleration-language_b200/host)")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SHIM SELFTEST OK" in r.stdout


def test_plain_c_program_over_the_c_abi():
    """host/c_abi_example.c: C99, no torch -- COO -> CSR -> column tiling -> plan -> SpMM / fused GAT."""
    exe = os.path.join(ROOT, "gala-gnn-acceleration-language_b200", "host", "c_abi_example")
    if not os.path.exists(exe):
        pytest.skip("host/c_abi_example not built (make -C gala-gnn-acceleration-language_b200/host)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C ABI EXAMPLE OK" in r.stdout
