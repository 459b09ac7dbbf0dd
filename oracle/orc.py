"""ctypes bindings for the checker libraries (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.

  lib()  -> oracle/_build/libgala_oracle.so  (C restatement, oracle/gala_oracle.c)
  ref()  -> oracle/_ref/libgala_ref.so       (the reference's own headers compiled from
                                              /root/reference by oracle/Makefile; may be
                                              absent on a machine without that tree and
                                              without the prebuilt file)

All arrays are numpy, int32 indices / float32 values, C-contiguous.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def _opt(a):
    """Nullable float array -> ctypes pointer or None."""
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def build(quiet=True):
    """Compile the checker libraries (building the checker is not using it)."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "_build", "libgala_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_mask_subgraph_offsets.restype = C.c_int64
    return _LIB


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libgala_ref.so"))


def ref():
    global _REF
    if _REF is None:
        import torch  # noqa: F401  (libgala_ref.so links libtorch; load it first)

        _REF = C.CDLL(os.path.join(HERE, "_ref", "libgala_ref.so"))
    return _REF


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Tiled:
    """Column-segmented graph in the layout the generated code carries around
    (src/ops/tiling.h:222-283): offsets[S*(N+1)], cols[E], vals[E], bounds[2S]."""

    def __init__(self, nrows, ncols, S, offsets, cols, vals, bounds):
        self.nrows, self.ncols, self.S = int(nrows), int(ncols), int(S)
        self.offsets = _c(offsets, np.int32)
        self.cols = _c(cols, np.int32)
        self.vals = _c(vals, np.float32)
        self.bounds = _c(bounds, np.int32)
        self.nvals = int(self.cols.shape[0])

    @staticmethod
    def from_csr(nrows, ncols, offset, ids, vals=None):
        ids = _c(ids, np.int32)
        if vals is None:
            vals = np.ones(ids.shape[0], np.float32)
        return Tiled(nrows, ncols, 1, offset, ids, vals, np.array([0, ids.shape[0]], np.int32))


# ----------------------------------------------------------------------------- oracle (C port)
def gspmm_wsum(nrows, offset, ids, vals, B, out=None):
    K = B.shape[1]
    if out is None:
        out = np.zeros((nrows, K), np.float32)
    lib().orc_gspmm_wsum(C.c_int(nrows), _c(offset, np.int32).ctypes, _c(ids, np.int32).ctypes,
                         _c(vals, np.float32).ctypes, _c(B, np.float32).ctypes, C.c_int(K),
                         out.ctypes)
    return out


def spmm(g, X, weighted=True, Y=None, vals=None):
    X = _c(X, np.float32)
    K = X.shape[1]
    if Y is None:
        Y = np.zeros((g.nrows, K), np.float32)
    v = (g.vals if vals is None else _c(vals, np.float32)) if weighted else None
    lib().orc_spmm_tiled(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes, g.cols.ctypes, _opt(v),
                         g.bounds.ctypes, X.ctypes, C.c_int(K), Y.ctypes)
    return Y


def spmm_f64(g, X, weighted=True, vals=None):
    X = _c(X, np.float32)
    K = X.shape[1]
    Y = np.zeros((g.nrows, K), np.float64)
    v = (g.vals if vals is None else _c(vals, np.float32)) if weighted else None
    lib().orc_spmm_tiled_f64acc(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes, g.cols.ctypes,
                                _opt(v), g.bounds.ctypes, X.ctypes, C.c_int(K), Y.ctypes)
    return Y


def spmm_sampled(g, X, nsamples, ra, rb, weighted=False):
    X = _c(X, np.float32)
    K = X.shape[1]
    Y = np.zeros((g.nrows, K), np.float32)
    lib().orc_spmm_sampled_tiled(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes, g.cols.ctypes,
                                 _opt(g.vals if weighted else None), g.bounds.ctypes, X.ctypes,
                                 C.c_int(K), C.c_int(nsamples), C.c_int(ra), C.c_int(rb), Y.ctypes)
    return Y


def edge_rowsum(g, vals):
    out = np.zeros(g.nrows, np.float32)
    lib().orc_edge_rowsum_tiled(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes,
                                _c(vals, np.float32).ctypes, g.bounds.ctypes, out.ctypes)
    return out


def edge_scale_rows(g, vals, rowval):
    v = np.array(vals, dtype=np.float32, copy=True)
    lib().orc_edge_scale_rows_tiled(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes,
                                    g.bounds.ctypes, v.ctypes, _c(rowval, np.float32).ctypes)
    return v


def sddvv(g, A, B, op="add"):
    out = np.zeros(g.nvals, np.float32)
    fn = lib().orc_sddvv_add_tiled if op == "add" else lib().orc_sddvv_mul_tiled
    fn(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes, g.cols.ctypes, g.bounds.ctypes,
       _c(A, np.float32).ctypes, _c(B, np.float32).ctypes, out.ctypes)
    return out


def sddmm(g, A, B):
    A = _c(A, np.float32)
    B = _c(B, np.float32)
    out = np.zeros(g.nvals, np.float32)
    lib().orc_sddmm_dot_tiled(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes, g.cols.ctypes,
                              g.bounds.ctypes, A.ctypes, B.ctypes, C.c_int(A.shape[1]), out.ctypes)
    return out


def leaky_relu(x, slope=0.2):
    x = _c(x, np.float32)
    y = np.empty_like(x)
    lib().orc_leaky_relu(C.c_int64(x.size), x.ctypes, C.c_float(slope), y.ctypes)
    return y


def edge_softmax_fwd(g, x):
    x = _c(x, np.float32)
    alpha = np.empty(g.nvals, np.float32)
    recip = np.empty(g.nrows, np.float32)
    lib().orc_edge_softmax_fwd_tiled(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes,
                                     g.bounds.ctypes, C.c_int64(g.nvals), x.ctypes, alpha.ctypes,
                                     recip.ctypes)
    return alpha, recip


def edge_softmax_bwd(g, alpha, dalpha):
    out = np.empty(g.nvals, np.float32)
    lib().orc_edge_softmax_bwd_tiled(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes,
                                     g.bounds.ctypes, C.c_int64(g.nvals),
                                     _c(alpha, np.float32).ctypes, _c(dalpha, np.float32).ctypes,
                                     out.ctypes)
    return out


def gat_backward_att(g, alpha, dalpha, aL, aR, slope=0.2):
    out = np.zeros(g.nrows, np.float32)
    lib().orc_gat_backward_att_tiled(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes, g.cols.ctypes,
                                     g.bounds.ctypes, C.c_int64(g.nvals), _c(alpha, np.float32).ctypes,
                                     _c(dalpha, np.float32).ctypes, _c(aL, np.float32).ctypes,
                                     _c(aR, np.float32).ctypes, C.c_float(slope), out.ctypes)
    return out


def gat_forward(g, aL, aR, X, slope=0.2):
    X = _c(X, np.float32)
    K = X.shape[1]
    Y = np.zeros((g.nrows, K), np.float32)
    alpha = np.empty(g.nvals, np.float32)
    lib().orc_gat_forward_tiled(C.c_int(g.nrows), C.c_int(g.S), g.offsets.ctypes, g.cols.ctypes,
                                g.bounds.ctypes, C.c_int64(g.nvals), _c(aL, np.float32).ctypes,
                                _c(aR, np.float32).ctypes, X.ctypes, C.c_int(K), C.c_float(slope),
                                Y.ctypes, alpha.ctypes)
    return Y, alpha


def csr_build(nrows, row_ids, col_ids, vals=None, _l=None):
    row_ids = _c(row_ids, np.int32)
    col_ids = _c(col_ids, np.int32)
    E = row_ids.shape[0]
    vals = np.ones(E, np.float32) if vals is None else _c(vals, np.float32)
    offset = np.zeros(nrows + 1, np.int32)
    ids = np.zeros(E, np.int32)
    ov = np.zeros(E, np.float32)
    rc = lib().orc_csr_build(C.c_int(nrows), C.c_int64(E), row_ids.ctypes, col_ids.ctypes,
                             vals.ctypes, offset.ctypes, ids.ctypes, ov.ctypes)
    assert rc == 0
    return offset, ids, ov


def csr_transpose(nrows, ncols, offset, ids, vals):
    offset = _c(offset, np.int32)
    ids = _c(ids, np.int32)
    vals = _c(vals, np.float32)
    E = ids.shape[0]
    to = np.zeros(ncols + 1, np.int32)
    ti = np.zeros(E, np.int32)
    tv = np.zeros(E, np.float32)
    rc = lib().orc_csr_transpose(C.c_int(nrows), C.c_int(ncols), offset.ctypes, ids.ctypes,
                                 vals.ctypes, to.ctypes, ti.ctypes, tv.ctypes)
    assert rc == 0
    return to, ti, tv


def col_breakpoints(ncols, T):
    out = np.zeros(ncols // max(T, 1) + 3, np.int32)
    n = lib().orc_col_breakpoints(C.c_int(ncols), C.c_int(T), out.ctypes)
    return out[:n].copy()


def col_tile(nrows, ncols, offset, ids, vals, T):
    offset = _c(offset, np.int32)
    ids = _c(ids, np.int32)
    vals = _c(vals, np.float32)
    bp = col_breakpoints(ncols, T)
    S = bp.shape[0] - 1
    E = ids.shape[0]
    o = np.zeros((nrows + 1) * S, np.int32)
    c = np.zeros(E, np.int32)
    v = np.zeros(E, np.float32)
    b = np.zeros(2 * S, np.int32)
    lib().orc_col_tile(C.c_int(nrows), offset.ctypes, ids.ctypes, vals.ctypes, C.c_int(S + 1),
                       bp.ctypes, o.ctypes, c.ctypes, v.ctypes, b.ctypes)
    return Tiled(nrows, ncols, S, o, c, v, b)


def sample_ab(nrows, offset, ids, vals, s, ra, rb):
    offset = _c(offset, np.int32)
    ids = _c(ids, np.int32)
    vals = _c(vals, np.float32)
    no = np.zeros(nrows + 1, np.int32)
    ni = np.zeros(nrows * s, np.int32)
    nv = np.zeros(nrows * s, np.float32)
    rc = lib().orc_sample_ab(C.c_int(nrows), offset.ctypes, ids.ctypes, vals.ctypes, C.c_int(s),
                             C.c_int(ra), C.c_int(rb), no.ctypes, ni.ctypes, nv.ctypes)
    return rc, no, ni, nv


def mask_subgraphs(nrows, ncols, offset, ids, vals, mask, layers):
    """Restatement of getMaskSubgraphs (tests/common.h:20-105) with a zero-initialised
    propagated mask.  Returns [(fwd_offset, fwd_ids, fwd_vals, bwd_offset, bwd_ids, bwd_vals)]."""
    offset = _c(offset, np.int32)
    ids = _c(ids, np.int32)
    vals = _c(vals, np.float32)
    cur = _c(mask, np.uint8).copy()
    res = []
    for _ in range(layers):
        no = np.zeros(nrows + 1, np.int32)
        nv = lib().orc_mask_subgraph_offsets(C.c_int(nrows), offset.ctypes, cur.ctypes, no.ctypes)
        ni = np.zeros(max(nv, 1), np.int32)
        nvl = np.zeros(max(nv, 1), np.float32)
        lib().orc_mask_subgraph_fill(C.c_int(nrows), offset.ctypes, ids.ctypes, vals.ctypes,
                                     no.ctypes, ni.ctypes, nvl.ctypes)
        ni, nvl = ni[:nv].copy(), nvl[:nv].copy()
        to, ti, tv = csr_transpose(nrows, ncols, no, ni, nvl)
        res.append((no, ni, nvl, to, ti, tv))
        nxt = np.zeros(nrows, np.uint8)
        lib().orc_mask_next(C.c_int(nrows), offset.ctypes, ids.ctypes, cur.ctypes, nxt.ctypes)
        cur = nxt
    return res


# ----------------------------------------------------------------------------- reference (_ref)
def ref_csr_build(nrows, ncols, row_ids, col_ids, vals=None):
    row_ids = _c(row_ids, np.int32)
    col_ids = _c(col_ids, np.int32)
    E = row_ids.shape[0]
    vals = np.ones(E, np.float32) if vals is None else _c(vals, np.float32)
    offset = np.zeros(nrows + 1, np.int32)
    ids = np.zeros(E, np.int32)
    ov = np.zeros(E, np.float32)
    rc = ref().ref_csr_build(C.c_int(nrows), C.c_int(ncols), C.c_int64(E), row_ids.ctypes,
                             col_ids.ctypes, vals.ctypes, offset.ctypes, ids.ctypes, ov.ctypes)
    assert rc == 0
    return offset, ids, ov


def ref_csr_transpose(nrows, ncols, offset, ids, vals):
    offset = _c(offset, np.int32)
    ids = _c(ids, np.int32)
    vals = _c(vals, np.float32)
    E = ids.shape[0]
    to = np.zeros(ncols + 1, np.int32)
    ti = np.zeros(E, np.int32)
    tv = np.zeros(E, np.float32)
    ref().ref_csr_transpose(C.c_int(nrows), C.c_int(ncols), offset.ctypes, ids.ctypes, vals.ctypes,
                            to.ctypes, ti.ctypes, tv.ctypes)
    return to, ti, tv


def ref_gspmm_wsum(nrows, ncols, offset, ids, vals, B, out=None):
    B = _c(B, np.float32)
    K = B.shape[1]
    # DenseMatrix pads 4*ncols elements (dense_matrix.h:52,79); gSpMM never reads them.
    if out is None:
        out = np.zeros((nrows, K), np.float32)
    ref().ref_gspmm_wsum(C.c_int(nrows), C.c_int(ncols), _c(offset, np.int32).ctypes,
                         _c(ids, np.int32).ctypes, _c(vals, np.float32).ctypes, B.ctypes,
                         C.c_int(K), out.ctypes)
    return out


def ref_col_tile(nrows, ncols, offset, ids, vals, T):
    offset = _c(offset, np.int32)
    ids = _c(ids, np.int32)
    vals = _c(vals, np.float32)
    bp = np.zeros(ncols // max(T, 1) + 3, np.int32)
    n = ref().ref_col_breakpoints(C.c_int(nrows), C.c_int(ncols), offset.ctypes, ids.ctypes,
                                  vals.ctypes, C.c_int(T), bp.ctypes, C.c_int(bp.shape[0]))
    assert n > 0
    bp = bp[:n].copy()
    S = n - 1
    E = ids.shape[0]
    o = np.zeros((nrows + 1) * S, np.int32)
    c = np.zeros(E, np.int32)
    v = np.zeros(E, np.float32)
    b = np.zeros(2 * S, np.int32)
    ref().ref_col_tile(C.c_int(nrows), C.c_int(ncols), offset.ctypes, ids.ctypes, vals.ctypes,
                       C.c_int(n), bp.ctypes, o.ctypes, c.ctypes, v.ctypes, b.ctypes)
    return bp, Tiled(nrows, ncols, S, o, c, v, b)


def ref_sample_ab(nrows, ncols, offset, ids, vals, s, ra, rb):
    offset = _c(offset, np.int32).copy()
    ids = _c(ids, np.int32).copy()
    vals = _c(vals, np.float32).copy()
    no = np.zeros(nrows + 1, np.int32)
    ni = np.zeros(nrows * s, np.int32)
    nv = np.zeros(nrows * s, np.float32)
    ref().ref_sample_ab(C.c_int(nrows), C.c_int(ncols), offset.ctypes, ids.ctypes, vals.ctypes,
                        C.c_int(s), C.c_int(ra), C.c_int(rb), no.ctypes, ni.ctypes, nv.ctypes)
    return no, ni, nv


def ref_mask_subgraphs(nrows, ncols, offset, ids, vals, mask, layers):
    offset = _c(offset, np.int32).copy()
    ids = _c(ids, np.int32).copy()
    vals = _c(vals, np.float32).copy()
    E = ids.shape[0]
    fo = np.zeros(layers * (nrows + 1), np.int32)
    fi = np.zeros(layers * max(E, 1), np.int32)
    fv = np.zeros(layers * max(E, 1), np.float32)
    fn = np.zeros(layers, np.int32)
    bo = np.zeros(layers * (ncols + 1), np.int32)
    bi = np.zeros(layers * max(E, 1), np.int32)
    bv = np.zeros(layers * max(E, 1), np.float32)
    ref().ref_mask_subgraphs(C.c_int(nrows), C.c_int(ncols), offset.ctypes, ids.ctypes,
                             vals.ctypes, _c(mask, np.uint8).ctypes, C.c_int(layers), fo.ctypes,
                             fi.ctypes, fv.ctypes, fn.ctypes, bo.ctypes, bi.ctypes, bv.ctypes)
    res, pos = [], 0
    for l in range(layers):
        n = int(fn[l])
        res.append((fo[l * (nrows + 1):(l + 1) * (nrows + 1)].copy(), fi[pos:pos + n].copy(),
                    fv[pos:pos + n].copy(), bo[l * (ncols + 1):(l + 1) * (ncols + 1)].copy(),
                    bi[pos:pos + n].copy(), bv[pos:pos + n].copy()))
        pos += n
    return res


def csr_reorder(nrows, offset, ids, vals, perm):
    offset, ids, vals, perm = _c(offset, np.int32), _c(ids, np.int32), _c(vals, np.float32), _c(perm, np.int32)
    E = ids.shape[0]
    no, ni, nv = np.zeros(nrows + 1, np.int32), np.zeros(E, np.int32), np.zeros(E, np.float32)
    lib().orc_csr_reorder(C.c_int(nrows), offset.ctypes, ids.ctypes, vals.ctypes, perm.ctypes,
                          no.ctypes, ni.ctypes, nv.ctypes)
    return no, ni, nv


def permute_rows(X, perm, from_=False):
    X, perm = _c(X, np.float32), _c(perm, np.int32)
    Y = np.zeros_like(X)
    lib().orc_permute_rows(C.c_int(X.shape[0]), C.c_int(X.shape[1]), X.ctypes, perm.ctypes,
                           C.c_int(int(from_)), Y.ctypes)
    return Y


def degree_order(nrows, offset):
    offset = _c(offset, np.int32)
    perm, order = np.zeros(nrows, np.int32), np.zeros(nrows, np.int32)
    lib().orc_degree_order(C.c_int(nrows), offset.ctypes, perm.ctypes, order.ctypes)
    return perm, order


def ref_row_reorder_to_adj(nrows, offset, ids, vals, perm):
    offset, ids, vals, perm = _c(offset, np.int32), _c(ids, np.int32), _c(vals, np.float32), _c(perm, np.int32)
    E = ids.shape[0]
    no, ni, nv = np.zeros(nrows + 1, np.int32), np.zeros(E, np.int32), np.zeros(E, np.float32)
    ref().ref_row_reorder_to_adj(C.c_int(nrows), offset.ctypes, ids.ctypes, vals.ctypes, perm.ctypes,
                                 no.ctypes, ni.ctypes, nv.ctypes)
    return no, ni, nv


def ref_row_permute_dense(X, perm, from_=False):
    X = np.array(X, dtype=np.float32, order="C", copy=True)
    perm = _c(perm, np.int32)
    ref().ref_row_permute_dense(C.c_int(X.shape[0]), C.c_int(X.shape[1]), X.ctypes, perm.ctypes,
                                C.c_int(int(from_)))
    return X
