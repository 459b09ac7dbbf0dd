"""Graph format construction on the GPU (C-ABI gala_csr_from_coo & co.): the device-side
replacement of the host data prep the generated program runs (reference
src/formats/csrc_matrix.h:148-376, src/ops/tiling.h:222-283,454-508,1594-1608,
tests/common.h:20-123).  Same function names as the reference where one exists."""
import ctypes as C

import numpy as np
import torch

from . import lib as _l
from .ops import TiledGraph


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def csr_build(nrows, ncols, row_ids, col_ids, vals=None):
    """CSRCMatrix::build (CSR): COO -> (offsets, ids, vals) sorted by (row, col)."""
    lib = _l.load()
    E = int(row_ids.numel())
    dev = row_ids.device
    offsets = torch.empty(nrows + 1, dtype=torch.int32, device=dev)
    ids = torch.empty(E, dtype=torch.int32, device=dev)
    out_vals = torch.empty(E, dtype=torch.float32, device=dev) if vals is not None else None
    nb = lib.gala_csr_from_coo_workspace_bytes(nrows, ncols, E)
    ws = _ws(nb, dev)
    _l.check(lib.gala_csr_from_coo(nrows, ncols, E, _l.ptr(row_ids), _l.ptr(col_ids), _l.ptr(vals),
                                   _l.ptr(offsets), _l.ptr(ids), _l.ptr(out_vals), _l.ptr(ws), nb,
                                   _l.stream_ptr()))
    return offsets, ids, out_vals


def buildTranspose(nrows, ncols, offsets, ids, vals=None):
    lib = _l.load()
    E = int(ids.numel())
    dev = ids.device
    t_off = torch.empty(ncols + 1, dtype=torch.int32, device=dev)
    t_ids = torch.empty(E, dtype=torch.int32, device=dev)
    t_vals = torch.empty(E, dtype=torch.float32, device=dev) if vals is not None else None
    nb = lib.gala_csr_from_coo_workspace_bytes(ncols, nrows, E)
    ws = _ws(nb, dev)
    _l.check(lib.gala_csr_transpose(nrows, ncols, E, _l.ptr(offsets), _l.ptr(ids), _l.ptr(vals),
                                    _l.ptr(t_off), _l.ptr(t_ids), _l.ptr(t_vals), _l.ptr(ws), nb,
                                    _l.stream_ptr()))
    return t_off, t_ids, t_vals


def ord_col_tiling(nrows, ncols, offsets, ids, vals, cols_per_partition):
    """static_ord_col_breakpoints + ord_col_tiling_torch -> TiledGraph (vals attached)."""
    lib = _l.load()
    E = int(ids.numel())
    dev = ids.device
    S = lib.gala_col_tile_segments(ncols, cols_per_partition)
    out_off = torch.empty(S * (nrows + 1), dtype=torch.int32, device=dev)
    out_cols = torch.empty(E, dtype=torch.int32, device=dev)
    out_vals = torch.empty(E, dtype=torch.float32, device=dev)
    bounds = np.zeros(2 * S, np.int32)
    nb = lib.gala_col_tile_workspace_bytes(nrows, ncols, cols_per_partition)
    ws = _ws(nb, dev)
    _l.check(lib.gala_col_tile(nrows, ncols, E, _l.ptr(offsets), _l.ptr(ids), _l.ptr(vals),
                               cols_per_partition, _l.ptr(out_off), _l.ptr(out_cols), _l.ptr(out_vals),
                               C.c_void_p(bounds.ctypes.data), _l.ptr(ws), nb, _l.stream_ptr()))
    return TiledGraph(out_off, out_cols, nrows, ncols, bounds, S, vals=out_vals)


def inplace_sample_graph_ab(nrows, offsets, ids, vals, sample_size, ra=5, rb=7):
    lib = _l.load()
    dev = ids.device
    no = torch.empty(nrows + 1, dtype=torch.int32, device=dev)
    ni = torch.empty(nrows * sample_size, dtype=torch.int32, device=dev)
    nv = torch.empty(nrows * sample_size, dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _l.check(lib.gala_sample_ab(nrows, _l.ptr(offsets), _l.ptr(ids), _l.ptr(vals), sample_size, ra, rb,
                                _l.ptr(no), _l.ptr(ni), _l.ptr(nv), _l.ptr(status), _l.stream_ptr()))
    return no, ni, nv, status


def getMaskSubgraphs(nrows, ncols, offsets, ids, vals, mask, layers):
    """Per layer: (fwd_offsets, fwd_ids, fwd_vals, bwd_offsets, bwd_ids, bwd_vals); mask is a
    uint8 device tensor.  Largest sub-graph last, as the reference returns them."""
    lib = _l.load()
    dev = ids.device
    E = int(ids.numel())
    cur = mask.to(torch.uint8).contiguous()
    nb = lib.gala_mask_subgraph_workspace_bytes(nrows)
    ws = _ws(nb, dev)
    res = []
    for _ in range(layers):
        no = torch.empty(nrows + 1, dtype=torch.int32, device=dev)
        ni = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        nv = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
        nxt = torch.empty(nrows, dtype=torch.uint8, device=dev)
        total = C.c_int64(0)
        _l.check(lib.gala_mask_subgraph(nrows, _l.ptr(offsets), _l.ptr(ids), _l.ptr(vals), _l.ptr(cur),
                                        _l.ptr(no), _l.ptr(ni), _l.ptr(nv), C.byref(total), _l.ptr(nxt),
                                        _l.ptr(ws), nb, _l.stream_ptr()))
        n = int(total.value)
        ni, nv = ni[:n].contiguous(), nv[:n].contiguous()
        to, ti, tv = buildTranspose(nrows, ncols, no, ni, nv)
        res.append((no, ni, nv, to, ti, tv))
        cur = nxt
    return res


def rowReorderToAdj(nrows, offsets, ids, vals, perm):
    """reordering.h:940-1013: relabel node i as perm[i] (rows and columns), rows column-sorted."""
    lib = _l.load()
    E = int(ids.numel())
    dev = ids.device
    no = torch.empty(nrows + 1, dtype=torch.int32, device=dev)
    ni = torch.empty(E, dtype=torch.int32, device=dev)
    nv = torch.empty(E, dtype=torch.float32, device=dev) if vals is not None else None
    nb = lib.gala_csr_from_coo_workspace_bytes(nrows, nrows, E)
    ws = _ws(nb, dev)
    _l.check(lib.gala_csr_reorder(nrows, E, _l.ptr(offsets), _l.ptr(ids), _l.ptr(vals), _l.ptr(perm),
                                  _l.ptr(no), _l.ptr(ni), _l.ptr(nv), _l.ptr(ws), nb, _l.stream_ptr()))
    return no, ni, nv


def _permute_rows(X, perm, from_):
    lib = _l.load()
    n, K = X.shape
    Y = torch.empty_like(X)
    _l.check(lib.gala_permute_rows_f32(_l.ptr(X), _l.ptr(perm), _l.ptr(Y), n, K, from_, _l.stream_ptr()))
    return Y


def rowPermuteDenseTo(X, perm):
    """reordering.h:244-283: Y[perm[i]] = X[i] (returned; the reference overwrites in place)."""
    return _permute_rows(X, perm, 0)


def rowPermuteDenseFrom(X, perm):
    """reordering.h:207-236: Y[i] = X[perm[i]]."""
    return _permute_rows(X, perm, 1)


def getAcendingOrder(nrows, device):
    """reordering.h:1085-1093 (identity; the spelling is the reference's)."""
    return torch.arange(nrows, dtype=torch.int32, device=device)


def getDecendingOrder(nrows, device):
    """reordering.h:1095-1103: ret[nrows-1-i] = i."""
    return torch.arange(nrows - 1, -1, -1, dtype=torch.int32, device=device)


def degree_order(nrows, offsets):
    """Nodes by descending degree (ties by id): (perm, order), perm in the "to" form the two
    functions above take.  Not in the reference (its rabbit order is commented out)."""
    lib = _l.load()
    dev = offsets.device
    perm = torch.empty(nrows, dtype=torch.int32, device=dev)
    order = torch.empty(nrows, dtype=torch.int32, device=dev)
    nb = lib.gala_degree_order_workspace_bytes(nrows)
    ws = _ws(nb, dev)
    _l.check(lib.gala_degree_order(nrows, _l.ptr(offsets), _l.ptr(perm), _l.ptr(order), _l.ptr(ws), nb,
                                   _l.stream_ptr()))
    return perm, order
