"""GPU test: out-of-bounds writes.  compute-sanitizer is not available on the GPU pool, so every
output buffer is carved out of a sentinel-filled arena and the sentinels on both sides must
survive each kernel (odd sizes, tails, hub rows, several segments)."""
import numpy as np
import pytest
import torch

from gala_b200 import formats, ops
from util import make_csr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PAD = 257          # elements of sentinel on each side (odd on purpose: misaligns 16-byte vector stores)
SENT = -12345.0


class Arena:
    def __init__(self):
        self.views = []

    def out(self, shape, dtype=torch.float32, align16=True):
        n = int(np.prod(shape))
        pad = PAD + (3 if align16 else 0)
        buf = torch.full((n + 2 * pad + 8,), SENT, dtype=torch.float32, device=DEV).to(dtype)
        start = pad
        if align16:
            while (buf.data_ptr() + start * buf.element_size()) % 16:
                start += 1
        view = buf[start:start + n].view(shape)
        self.views.append((buf, start, n))
        return view

    def check(self):
        for buf, start, n in self.views:
            assert bool((buf[:start] == SENT).all()), "sentinel before the output was overwritten"
            assert bool((buf[start + n:] == SENT).all()), "sentinel after the output was overwritten"


@pytest.mark.parametrize("n,e,T,thr,K", [(1000, 30000, None, None, 32), (1003, 30011, 333, 64, 41), (517, 9000, 100, 32, 100),
                                          (2049, 100000, None, 128, 7), (301, 4000, 17, 16, 1)])
def test_outputs_stay_in_bounds(orc, n, e, T, thr, K):
    offset, ids = make_csr(n, e, n + e, empty_rows=3)
    t = orc.Tiled.from_csr(n, n, offset, ids) if T is None else orc.col_tile(n, n, offset, ids, np.ones(ids.shape[0], np.float32), T)
    g = ops.TiledGraph(torch.from_numpy(t.offsets).to(DEV), torch.from_numpy(t.cols).to(DEV), n, n, t.bounds, t.S)
    if thr:
        g.build_plan(thr)
    E = g.nvals
    X = torch.rand(n, K, device=DEV) - 0.5
    a = torch.randn(n, device=DEV)
    w = torch.rand(E, device=DEV)
    ar = Arena()
    ops.spmm(g, X, out=ar.out((n, K)))
    ops.spmm(g, X, vals=w, out=ar.out((n, K), align16=False), relu=True)
    ops.spmm_sampled(g, X, 20, 5, 7, out=ar.out((n, K)))
    ops.sddvv(g, a, a, "add", out=ar.out((E,)))
    ops.sddvv(g, a, a, "mul", leaky_slope=0.2, out=ar.out((E,), align16=False))
    ops.sddmm(g, X, X, out=ar.out((E,)))
    ops.edge_rowsum(g, w, out=ar.out((n, 1)))
    ops.edge_softmax_fwd(g, w, out=ar.out((E,)), recip=ar.out((n,)))
    ops.edge_softmax_bwd(g, w, w, out=ar.out((E,), align16=False))
    ops.gat_forward(g, a, a, X, out=ar.out((n, K)), alpha_out=ar.out((E,)))
    ops.gat_backward_att(g, w, w, a, a, 0.2, out=ar.out((n, 1)))
    if K % 4 == 0 and K <= 32:
        ops.gat_forward_dot(g, a, X[0].contiguous(), 0.1, X, out=ar.out((n, K)), alpha_out=ar.out((E,)))
    torch.cuda.synchronize()
    ar.check()


@pytest.mark.parametrize("M,K,N", [(1, 5, 8), (127, 33, 32), (129, 602, 32), (1000, 64, 41), (5001, 100, 47), (700, 31, 64)])
def test_linear_outputs_stay_in_bounds(M, K, N):
    X = torch.rand(M, K, device=DEV) - 0.5
    W = torch.rand(N, K, device=DEV) - 0.5
    b = torch.rand(N, device=DEV)
    ar = Arena()
    ops.linear(X, W, b, out=ar.out((M, N)))
    ops.linear(X, W, b, relu=True, out=ar.out((M, N), align16=False))
    torch.cuda.synchronize()
    ar.check()


def test_format_outputs_are_fully_written_and_bounded(orc):
    n = 777
    offset, ids = make_csr(n, 20000, 5)
    rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(offset))
    p = np.random.default_rng(0).permutation(ids.shape[0])
    r, c = torch.from_numpy(rows[p]).to(DEV), torch.from_numpy(ids[p]).to(DEV)
    off, idd, _ = formats.csr_build(n, n, r, c)
    assert int(off[-1]) == ids.shape[0] and int(idd.min()) >= 0 and int(idd.max()) < n
    tg = formats.ord_col_tiling(n, n, off, idd, torch.ones(idd.numel(), device=DEV), 100)
    assert int(tg.cols.min()) >= 0 and int(tg.cols.max()) < n and int(tg.bounds[-1]) == idd.numel()


def test_reorder_outputs_stay_in_bounds_and_errors_are_reported():
    from gala_b200 import lib as _l

    n = 777
    offset, ids = make_csr(n, 9001, 5, empty_rows=4)
    E = ids.shape[0]
    off, col = torch.from_numpy(offset).to(DEV), torch.from_numpy(ids).to(DEV)
    w = torch.rand(E, device=DEV)
    X = torch.rand(n, 7, device=DEV)
    lib = _l.load()
    ar = Arena()
    perm = ar.out((n,), torch.int32)
    order = ar.out((n,), torch.int32)
    ws = torch.empty(lib.gala_degree_order_workspace_bytes(n), dtype=torch.uint8, device=DEV)
    _l.check(lib.gala_degree_order(n, _l.ptr(off), _l.ptr(perm), _l.ptr(order), _l.ptr(ws), ws.numel(), _l.stream_ptr()))
    no, ni, nv = ar.out((n + 1,), torch.int32), ar.out((E,), torch.int32), ar.out((E,))
    nb = lib.gala_csr_from_coo_workspace_bytes(n, n, E)
    ws2 = torch.empty(nb, dtype=torch.uint8, device=DEV)
    _l.check(lib.gala_csr_reorder(n, E, _l.ptr(off), _l.ptr(col), _l.ptr(w), _l.ptr(perm.contiguous()), _l.ptr(no),
                                  _l.ptr(ni), _l.ptr(nv), _l.ptr(ws2), nb, _l.stream_ptr()))
    Y = ar.out((n, 7))
    _l.check(lib.gala_permute_rows_f32(_l.ptr(X), _l.ptr(perm.contiguous()), _l.ptr(Y), n, 7, 0, _l.stream_ptr()))
    torch.cuda.synchronize()
    ar.check()
    assert sorted(perm.tolist()) == list(range(n)) and int(no[-1]) == E
    # error convention: negative GALA_ERR_* codes, nothing launched
    assert lib.gala_csr_reorder(n, E, _l.ptr(off), _l.ptr(col), None, None, _l.ptr(no), _l.ptr(ni), None, _l.ptr(ws2), nb,
                                _l.stream_ptr()) == -1                                      # NULL perm
    assert lib.gala_csr_reorder(n, E, _l.ptr(off), _l.ptr(col), None, _l.ptr(perm.contiguous()), _l.ptr(no), _l.ptr(ni),
                                None, _l.ptr(ws2), 16, _l.stream_ptr()) < 0                 # workspace too small
    assert lib.gala_permute_rows_f32(_l.ptr(X), _l.ptr(perm.contiguous()), _l.ptr(X), n, 7, 0, _l.stream_ptr()) < 0   # in place
    assert lib.gala_degree_order(-1, _l.ptr(off), _l.ptr(perm), None, _l.ptr(ws), ws.numel(), _l.stream_ptr()) < 0


def test_read_probe_reports_l2_above_hbm():
    l2 = ops.probe_read_gbs(24 << 20, 100, DEV)
    hbm = ops.probe_read_gbs(1 << 30, 2, DEV)
    assert 1000 < hbm < 9000, hbm           # B200 HBM3e: ~6.5 TB/s measured peak
    assert l2 > hbm, (l2, hbm)
