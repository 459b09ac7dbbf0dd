"""Host-side algebra of the reflected-basis GAT layer (gala_gat_forward_col_f32, GAT2.fold_reflected): the Householder
helper of the C-ABI and the folding of the reflections into the weights, checked in numpy against the literal op
sequence of the generated program (reference src/codegen/common.h:622-675, 735-810, 835-927, 1185-1281).  No GPU:
the kernel itself is checked by tests/test_ops_gpu.py::test_gat_col_variant_matches_oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from gala_b200 import lib as _l
from gala_b200 import ops


def _reflection(w):
    K = len(w)
    w_arr = (C.c_float * K)(*[float(x) for x in w])
    v_arr = (C.c_float * K)()
    sR = C.c_float()
    rc = _l.load().gala_reflection_f32(w_arr, K, v_arr, C.byref(sR))
    return rc, np.array(list(v_arr), dtype=np.float64), float(sR.value)


@pytest.mark.parametrize("K", [1, 4, 8, 16, 32, 41])
@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_reflection_maps_last_unit_vector_onto_w(K, sign):
    rng = np.random.default_rng(K)
    w = rng.normal(size=K).astype(np.float32)
    w[-1] = sign * abs(w[-1])
    rc, v, sR = _reflection(w)
    assert rc == 0
    assert abs(np.linalg.norm(v) - 1.0) < 1e-6
    H = np.eye(K) - 2.0 * np.outer(v, v)
    assert np.abs(H @ H - np.eye(K)).max() < 1e-6
    assert np.abs(sR * H[:, -1] - w).max() <= 2e-6 * np.abs(w).max()
    assert np.sign(sR) == -sign                  # the stable branch: no cancellation in v's last component
    X = rng.uniform(-1, 1, (50, K))
    assert np.abs(sR * (X @ H)[:, -1] - X @ w).max() < 1e-5


def test_reflection_degenerate_and_pivot_only():
    rc, _, _ = _reflection(np.zeros(8, dtype=np.float32))
    assert rc != 0
    # w already along the last axis: H = I - 2 e e^T, sR = -w_last
    w = np.zeros(8, dtype=np.float32)
    w[-1] = 3.0
    rc, v, sR = _reflection(w)
    assert rc == 0 and abs(sR + 3.0) < 1e-6 and abs(abs(v[-1]) - 1.0) < 1e-6


def test_ops_reflect_matches_dense_householder():
    rng = np.random.default_rng(0)
    w = torch.tensor(rng.normal(size=32).astype(np.float32))
    v, sR = ops.reflection(w)
    H = torch.eye(32, dtype=torch.float64) - 2.0 * torch.outer(v.double(), v.double())
    T = torch.tensor(rng.normal(size=(5, 32)).astype(np.float32))
    assert torch.allclose(ops.reflect(T, v).double(), T.double() @ H, atol=1e-6)
    W = torch.tensor(rng.normal(size=(32, 7)).astype(np.float32))
    assert torch.allclose(ops.reflect(W, v, dim=0).double(), H @ W.double(), atol=1e-6)
    b = torch.tensor(rng.normal(size=32).astype(np.float32))
    assert torch.allclose(ops.reflect(b, v).double(), H @ b.double(), atol=1e-6)


def _gat_layer(A, aL, aR, X, slope, relu):
    """Dense restatement of the emitted layer: exp -> clamp(1e12) -> row-normalise with the 1e-12 seed."""
    z = aL[:, None] + aR[None, :]
    z = np.where(z > 0, z, slope * z)
    e = np.minimum(np.exp(z), 1e12) * A
    alpha = e / (e.sum(1, keepdims=True) + 1e-12)
    Y = alpha @ X
    return np.maximum(Y, 0) if relu else Y


def _gat_layer_col(A, aL, sR, bR, Xr, slope, relu, v_in, v_out):
    """What gala_gat_forward_col_f32 computes, restated."""
    Y = _gat_layer(A, aL, sR * Xr[:, -1] + bR, Xr, slope, False)
    if v_in is not None:
        Y = Y - 2.0 * np.outer(Y @ v_in, v_in)
    if relu:
        Y = np.maximum(Y, 0)
    if v_out is not None:
        Y = Y - 2.0 * np.outer(Y @ v_out, v_out)
    return Y


def test_reflected_model_equals_literal_model():
    """GAT2.fold_reflected + the restated column-mode layer reproduce the literal 2-layer forward."""
    from gala_b200.gat_model import GAT2
    n, F_in, hidden, classes = 120, 20, 32, 7
    rng = np.random.default_rng(1)
    A = (rng.uniform(size=(n, n)) < 0.1).astype(np.float64)
    A[np.arange(n), np.arange(n)] = 1.0
    m = GAT2(F_in, hidden, classes, "cpu", seed=5)
    X = rng.uniform(-0.5, 0.5, (n, F_in))
    d = lambda t: t.double().numpy()
    lin = lambda x, wb: x @ d(wb[0]).T + d(wb[1])
    # literal (common.h order; layer 2 under the FFN-recompute rewrite)
    res = lin(X, m.fc0)
    res = _gat_layer(A, lin(res, m.efc0)[:, 0], lin(res, m.efc1)[:, 0], res, 0.2, True)
    t = lin(res, m.fc1)
    agg = _gat_layer(A, lin(t, m.efc2)[:, 0], lin(t, m.efc3)[:, 0], res, 0.2, False)
    want = lin(agg, m.fc1)
    # reflected
    r = m.fold_reflected()
    res = X @ d(r["W0"]).T + d(r["b0"])
    aL = res @ d(r["W_att1"])[0] + m.b_att1_host[0]
    res = _gat_layer_col(A, aL, r["s1"], m.bR1, res, 0.2, True, d(r["v1"]), d(r["v2"]))
    aL = res @ d(r["W_att2"])[0] + m.bL2
    agg = _gat_layer_col(A, aL, r["s2"], m.bR2, res, 0.2, False, d(r["v2"]), None)
    got = lin(agg, m.fc1)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 2e-6


def test_reflected_gatn_equals_literal_model():
    """GATN.fold_reflected (3 layers: hidden rows reflected back after each aggregation, the last hidden layer's rows
    leaving in the final layer's basis) + the restated column-mode layer reproduce GATN.forward_literal's math."""
    from gala_b200.gat_model import GATN
    n, dims = 90, [12, 32, 16, 5]
    rng = np.random.default_rng(2)
    A = (rng.uniform(size=(n, n)) < 0.12).astype(np.float64)
    A[np.arange(n), np.arange(n)] = 1.0
    m = GATN(dims, "cpu", seed=4)
    m.host_biases()
    assert m.reflected_ok()
    X = rng.uniform(-0.5, 0.5, (n, dims[0]))
    d = lambda t: t.double().numpy()
    lin = lambda x, wb: x @ d(wb[0]).T + d(wb[1])
    L = m.L
    res = X
    for i in range(L - 1):
        t = lin(res, m.fc[i])
        res = _gat_layer(A, lin(t, m.efcL[i])[:, 0], lin(t, m.efcR[i])[:, 0], t, 0.2, True)
    t = lin(res, m.fc[-1])
    agg = _gat_layer(A, lin(t, m.efcL[-1])[:, 0], lin(t, m.efcR[-1])[:, 0], res, 0.2, False)
    want = lin(agg, m.fc[-1])
    r = m.fold_reflected()
    res, aL_last = X, None
    for i in range(L - 1):
        t = res @ d(r["W"][i]).T + d(r["b"][i])
        aL = t @ d(r["W_att"][i])[0] + m._bh[i][0]
        last = i == L - 2
        res = _gat_layer_col(A, aL, r["s"][i], r["bR"][i], t, 0.2, True, d(r["v"][i]), d(r["v"][L - 1]) if last else None)
        if last:
            aL_last = res @ d(r["W_att"][L - 1])[0] + m._bh[L - 1][0]
    agg = _gat_layer_col(A, aL_last, r["s"][L - 1], r["bR"][L - 1], res, 0.2, False, d(r["v"][L - 1]), None)
    got = lin(agg, m.fc[-1])
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 2e-6
