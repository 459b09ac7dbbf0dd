"""CPU suite: the C-ABI library loads and exports every symbol include/gala_b200.h
declares; argument validation that needs no GPU; host-side logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from gala_b200 import lib as L
from gala_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gala_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gala_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gala_b200.h but not exported"
    assert set(names) == set(L.EXPORTS)


def test_abi_version_and_error_strings():
    lib = L.load()
    assert lib.gala_b200_abi_version() == 4
    assert lib.gala_b200_error_string(0) == b"success"
    for code in (-1, -2, -3, -4, -5):
        assert b"gala_b200" in lib.gala_b200_error_string(code)


def test_argument_errors_return_codes_not_exits():
    """The reference printf+exit()s on error (cuda.h:980-998); the ABI returns codes."""
    lib = L.load()
    assert lib.gala_spmm_f32(None, None, None, 4, None, None, None, None) == -1
    g = L.GalaGraph(offsets=None, cols=None, bounds=None, nrows=-3, ncols=0, segments=1, nvals=0)
    assert lib.gala_spmm_f32(C.byref(g), None, None, 4, None, None, None, None) == -2
    # any number of column segments is accepted (the reference has no limit); beyond 64 the device copy of `bounds`
    # is required as well
    import numpy as np
    b = np.zeros(2000, np.int32)
    g = L.GalaGraph(offsets=None, cols=None, bounds=None, nrows=0, ncols=0, segments=1000, nvals=0)
    assert lib.gala_spmm_f32(C.byref(g), None, None, 4, None, None, None, None) == -1
    g = L.GalaGraph(offsets=None, cols=None, bounds=b.ctypes.data, nrows=0, ncols=0, segments=1000, nvals=0)
    assert lib.gala_spmm_f32(C.byref(g), None, None, 4, None, None, None, None) == -1
    g = L.GalaGraph(offsets=None, cols=None, bounds=None, nrows=8, ncols=8, segments=1, nvals=0)
    assert lib.gala_spmm_f32(C.byref(g), None, None, 4, None, None, None, None) == -1
    with pytest.raises(L.GalaError):
        L.check(-3)


def test_tiled_graph_host_view():
    off = torch.tensor([0, 1, 2, 0, 1, 1], dtype=torch.int32)
    cols = torch.tensor([0, 1, 1], dtype=torch.int32)
    g = ops.TiledGraph(off, cols, nrows=2, ncols=2, bounds=[0, 2, 2, 3], segments=2)
    assert g.c.nrows == 2 and g.c.segments == 2 and g.c.nvals == 3
    assert np.array_equal(g.bounds, [0, 2, 2, 3])
    assert L.load().gala_plan_workspace_bytes(C.byref(g.c)) >= 4 * 2


def test_no_cpu_fallback_when_library_missing(monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libgala_b200.so")
    with pytest.raises(ImportError):
        L.load()


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++-isms, no torch types)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "gala_b200.h"\nint main(void) { return gala_b200_abi_version() > 0 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_new_entry_points_validate_arguments():
    lib = L.load()
    assert lib.gala_csr_reorder(-1, 0, None, None, None, None, None, None, None, None, 0, None) == -2
    assert lib.gala_csr_reorder(4, 3, None, None, None, None, None, None, None, None, 0, None) == -1
    assert lib.gala_permute_rows_f32(None, None, None, 4, 4, 0, None) == -1
    assert lib.gala_permute_rows_f32(None, None, None, 0, 4, 0, None) == 0          # empty: nothing to do
    assert lib.gala_degree_order(0, None, None, None, None, 0, None) == 0
    assert lib.gala_degree_order(5, None, None, None, None, 0, None) == -1
    assert lib.gala_gat_backward_att_f32(None, None, None, None, None, 0.2, None, None, None) == -1
    assert lib.gala_b200_probe_read(None, 1024, 1, None, None) == -1
    assert lib.gala_degree_order_workspace_bytes(1000) > 0


def test_reflection_helper_from_plain_c(tmp_path):
    """gala_reflection_f32 is host arithmetic behind the C ABI: a C99 program linked against libgala_b200.so (no GPU,
    no torch) builds the Householder vector and checks H e_last = sR^-1 w and |v| = 1."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "refl.c"
    src.write_text(r'''
#include <math.h>
#include <stdio.h>
#include "gala_b200.h"
int main(void) {
    float w[8] = {0.3f, -1.2f, 0.7f, 0.05f, -0.4f, 2.0f, -0.9f, -0.6f}, v[8], sR, zero[8] = {0};
    double n2 = 0.0, err = 0.0;
    int i;
    if (gala_reflection_f32(w, 8, v, &sR) != GALA_OK) return 1;
    for (i = 0; i < 8; ++i) n2 += (double)v[i] * v[i];
    if (fabs(n2 - 1.0) > 1e-6) return 2;
    /* column K-1 of H = I - 2 v v^T is e - 2 v[K-1] v; scaled by sR it must be w */
    for (i = 0; i < 8; ++i) {
        double h = (i == 7 ? 1.0 : 0.0) - 2.0 * (double)v[7] * (double)v[i];
        err = fmax(err, fabs(sR * h - (double)w[i]));
    }
    if (err > 1e-5) return 3;
    if (!(sR > 0.0f)) return 4;                    /* w[7] < 0: the stable branch picks sR = +|w| */
    if (gala_reflection_f32(zero, 8, v, &sR) != GALA_ERR_BAD_SHAPE) return 5;
    if (gala_reflection_f32(w, 8, v, (float *)0) != GALA_ERR_NULL_POINTER) return 6;
    /* the column-mode layer refuses widths it cannot hold in one warp pass, before touching the device */
    {
        gala_graph_t g = {0};
        if (gala_gat_forward_col_f32(&g, 0, 1.0f, 0.0f, 0, 12, 0.2f, 0, 0, 0, 0, 0, 0, 0, 0) == GALA_OK) return 7;
    }
    printf("ok\n");
    return 0;
}
''')
    exe = tmp_path / "refl"
    libdir = os.path.dirname(L.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-o", str(exe), "-L", libdir, "-lgala_b200", "-lm", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout, r.stderr)
