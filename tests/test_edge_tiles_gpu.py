"""GPU parity tests (-m gpu) of the edge-parallel (nnz-split / merge-path) streaming edge kernels
(csrc/edge_tiles.cuh): row sum (K3, reference src/codegen/cuda.h:505-524), row scaling (K4, :525-562),
edge-softmax forward / backward (src/codegen/common.h:760-799), each against the
oracle AND against the row-structured kernels of the same library (a plan without the tile table).

Graphs are built from prescribed degree sequences so that every path of the tile kernel runs: rows inside a
tile's shared-memory window, rows longer than the window (streamed by the whole CTA) with ordinary rows behind them
in the same tile, tiles in which no row starts, empty rows at both ends, an edge count that is an exact multiple
of the tile size, and mean degrees that select 4 / 8 / 16 / 32 lanes per row."""
import numpy as np
import pytest
import torch

from gala_b200 import ops
from util import FP32_TOL, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TILE = 4096     # csrc/edge_tiles.cuh: kTileEdges (window: 8192 edges, 1024 staged row pointers)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def csr_from_degrees(deg, ncols, seed):
    """CSR with the given row degrees, sorted random columns (duplicates allowed, as the reference keeps them)."""
    deg = np.asarray(deg, np.int64)
    rng = np.random.default_rng(seed)
    offset = np.zeros(deg.shape[0] + 1, np.int64)
    np.cumsum(deg, out=offset[1:])
    cols = rng.integers(0, ncols, int(offset[-1]), dtype=np.int64)
    rows = np.repeat(np.arange(deg.shape[0]), deg)
    order = np.lexsort((cols, rows))
    return offset.astype(np.int32), cols[order].astype(np.int32)


def degree_cases():
    rng = np.random.default_rng(7)
    cases = {}
    # mean degree ~130 -> 32 lanes per row; power-law-ish with a few thousand-edge rows
    d = np.minimum(rng.pareto(1.2, 3000) * 40 + 1, 6000).astype(np.int64)
    cases["powerlaw_g32"] = d
    # one row far longer than the staged window, ordinary rows right behind it in the same tile
    d = rng.integers(1, 300, 2000)
    d[5] = 40000
    d[700] = 17000
    d[701] = 9000
    cases["rows_longer_than_window"] = d
    # short rows: 4, 8 and 16 lanes per row; degree 1 puts more rows into a tile than row pointers are staged
    cases["degree_0_1_2_g4"] = rng.integers(0, 3, 30000)
    cases["mean_degree_5_g4"] = rng.integers(0, 11, 60000)
    cases["mean_degree_20_g8"] = rng.integers(10, 31, 20000)
    cases["mean_degree_50_g16"] = rng.integers(20, 81, 8000)
    # empty rows at both ends, and an edge count that is an exact multiple of the tile size
    d = np.concatenate([np.zeros(50, np.int64), rng.integers(1, 200, 1500), np.zeros(70, np.int64)])
    d[60] += (-int(d.sum())) % TILE
    cases["empty_ends_exact_multiple"] = d
    # a single tile, and a graph whose only row spans several tiles
    cases["single_small_tile"] = rng.integers(0, 9, 300)
    cases["one_row_three_tiles"] = np.array([3 * TILE + 5])
    return cases


CASES = degree_cases()


def graphs(orc, name):
    deg = CASES[name]
    n = deg.shape[0]
    offset, ids = csr_from_degrees(deg, n, seed=len(name))
    w = np.random.default_rng(3).uniform(-1, 1, ids.shape[0]).astype(np.float32)
    t = orc.Tiled.from_csr(n, n, offset, ids, w)
    tiled = ops.TiledGraph(dev(t.offsets), dev(t.cols), n, n, t.bounds, 1).build_plan(2048)
    rowwise = ops.TiledGraph(dev(t.offsets), dev(t.cols), n, n, t.bounds, 1).build_plan(2048)
    assert tiled.plan.tile_rows and tiled.plan.n_tiles == (t.nvals + TILE - 1) // TILE
    tiled.plan.tile_policy = 1           # GALA_TILES_ALWAYS: also where the default policy prefers the row-structured kernels
    rowwise.plan.tile_rows = None        # same plan without the tile table: the row-structured kernels run
    return t, tiled, rowwise


def test_tile_table_is_the_nnz_split(orc):
    """plan.tile_rows[t] = (first row starting at or after edge t * TILE, its first edge), closed by (nrows, E):
    bit-exact against numpy's searchsorted."""
    for name in CASES:
        t, g, _ = graphs(orc, name)
        nt = g.plan.n_tiles
        ws = g._plan_ws.view(torch.int32)
        base = (g.plan.tile_rows - g._plan_ws.data_ptr()) // 4
        got = ws[base:base + 2 * (nt + 1)].cpu().numpy().reshape(nt + 1, 2)
        off = t.offsets[:t.nrows + 1]
        want = np.searchsorted(off, np.arange(nt) * TILE, side="left")
        assert np.array_equal(got[:nt, 0], want) and np.array_equal(got[:nt, 1], off[want])
        assert got[nt, 0] == t.nrows and got[nt, 1] == t.nvals


@pytest.mark.parametrize("name", list(CASES))
def test_edge_tile_kernels_match_oracle_and_rowwise_kernels(orc, name):
    t, g, gr = graphs(orc, name)
    n = t.nrows
    rng = np.random.default_rng(11)
    deg = np.diff(t.offsets[:n + 1])
    # K3 row sum
    pos = np.abs(t.vals) + 0.01
    got = ops.edge_rowsum(g, dev(pos)).cpu().numpy().ravel()
    assert rel_err(got, orc.edge_rowsum(t, pos)) < FP32_TOL
    assert rel_err(got, ops.edge_rowsum(gr, dev(pos)).cpu().numpy().ravel()) < FP32_TOL
    assert np.allclose(got[deg == 0], np.float32(1e-12), rtol=1e-6)
    # K4 row scaling, in place: one rounding per edge -> bit-exact
    A = rng.normal(size=n).astype(np.float32)
    v = dev(t.vals.copy())
    ops.edge_scale_rows_(g, v, dev(A))
    assert np.array_equal(v.cpu().numpy(), orc.edge_scale_rows(t, t.vals, A))
    # edge-softmax forward (+ reciprocal row sums), out of place and in place
    x = rng.normal(scale=2.0, size=t.nvals).astype(np.float32)
    want, recip = orc.edge_softmax_fwd(t, x)
    r = torch.empty(n, device=DEV)
    got = ops.edge_softmax_fwd(g, dev(x), recip=r).cpu().numpy()
    assert rel_err(got, want) < FP32_TOL
    assert np.allclose(got, want, rtol=1e-5, atol=1e-12)
    assert rel_err(r.cpu().numpy()[deg > 0], recip[deg > 0]) < FP32_TOL
    assert np.allclose(got, ops.edge_softmax_fwd(gr, dev(x)).cpu().numpy(), rtol=1e-5, atol=1e-12)
    xi = dev(x.copy())
    ops.edge_softmax_fwd(g, xi, out=xi)
    assert np.array_equal(xi.cpu().numpy(), got)
    sums = np.add.reduceat(got.astype(np.float64), t.offsets[:n][deg > 0]) if t.nvals else np.zeros(0)
    assert np.allclose(sums, 1.0, atol=1e-5)            # attention rows sum to one
    # edge-softmax backward
    da = rng.normal(size=t.nvals).astype(np.float32)
    got_b = ops.edge_softmax_bwd(g, dev(want), dev(da)).cpu().numpy()
    assert rel_err(got_b, orc.edge_softmax_bwd(t, want, da)) < FP32_TOL
    assert rel_err(got_b, ops.edge_softmax_bwd(gr, dev(want), dev(da)).cpu().numpy()) < FP32_TOL


def test_misaligned_edge_arrays_take_the_rowwise_kernels(orc):
    """Bulk copies need 16-byte aligned edge arrays; a view that starts 4 bytes in must still be served."""
    t, g, _ = graphs(orc, "powerlaw_g32")
    x = np.random.default_rng(5).normal(size=t.nvals).astype(np.float32)
    buf = torch.empty(t.nvals + 1, device=DEV)
    view = buf[1:]
    view.copy_(dev(x))
    want, _ = orc.edge_softmax_fwd(t, x)
    got = ops.edge_softmax_fwd(g, view).cpu().numpy()
    assert rel_err(got, want) < FP32_TOL
    assert rel_err(ops.edge_rowsum(g, view).cpu().numpy().ravel(), orc.edge_rowsum(t, x)) < FP32_TOL


def test_tile_kernels_are_bit_reproducible(orc):
    t, g, _ = graphs(orc, "rows_longer_than_window")
    x = dev(np.random.default_rng(2).normal(size=t.nvals).astype(np.float32))
    a = ops.edge_softmax_fwd(g, x)
    for _ in range(3):
        assert torch.equal(a, ops.edge_softmax_fwd(g, x))
