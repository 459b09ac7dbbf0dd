"""Shared helpers for the parity tests."""
import os

import numpy as np

from gala_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["small_a", "small_dup", "small_hub"]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def make_csr(n, e, seed, empty_rows=0):
    """Seeded power-law CSR (self loops, symmetric); optionally blank out some rows."""
    src, dst = synth.powerlaw_coo_np(n, e, seed=seed)
    if empty_rows:
        rng = np.random.default_rng(seed + 1000)
        kill = rng.choice(n, empty_rows, replace=False)
        keep = ~np.isin(src, kill)
        src, dst = src[keep], dst[keep]
    return synth.coo_to_csr_np(n, src, dst)


def rel_err(a, b):
    """Norm-wise relative error ||a-b||_F / ||b||_F (the fp32 bound
    is defined norm-wise per layer)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))


def max_rel_to_rowscale(a, b):
    """max_i ||a_i - b_i||_inf / max(||b_i||_inf, tiny) over rows."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.ndim == 1:
        a, b = a[:, None], b[:, None]
    num = np.abs(a - b).max(axis=1)
    den = np.maximum(np.abs(b).max(axis=1), 1e-30)
    return float((num / den).max()) if a.size else 0.0


FP32_TOL = 1e-5  # BASELINE.json north_star: fp32 outputs within 1e-5 relative error per layer
