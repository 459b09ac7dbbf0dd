"""CPU suite (-m "not gpu"): the C restatement in oracle/ against the golden vectors the
reference itself produced (tests/golden/*.npz, see make_golden.py) and, when
oracle/_ref is present, against the reference library on fresh random inputs."""
import numpy as np
import pytest

from util import GOLDEN_CASES, golden, make_csr, rel_err


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_csr_build_matches_reference_golden(orc, name):
    g = golden(name)
    n = int(g["n"])
    offset, ids, vals = orc.csr_build(n, g["coo_src"], g["coo_dst"])
    assert np.array_equal(offset, g["offset"])      # integer work: bit-exact
    assert np.array_equal(ids, g["ids"])
    assert np.all(vals == 1.0)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_transpose_matches_reference_golden(orc, name):
    g = golden(name)
    n = int(g["n"])
    ones = np.ones(g["ids"].shape[0], np.float32)
    to, ti, tv = orc.csr_transpose(n, n, g["offset"], g["ids"], ones)
    assert np.array_equal(to, g["t_offset"])
    assert np.array_equal(ti, g["t_ids"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_gspmm_matches_reference_golden(orc, name):
    g = golden(name)
    n = int(g["n"])
    Yw = orc.gspmm_wsum(n, g["offset"], g["ids"], g["w"], g["X"])
    assert rel_err(Yw, g["Y_w"]) < 1e-6
    tiled = orc.Tiled.from_csr(n, n, g["offset"], g["ids"], g["w"])
    assert rel_err(orc.spmm(tiled, g["X"], weighted=True), g["Y_w"]) < 1e-6
    assert rel_err(orc.spmm(tiled, g["X"], weighted=False), g["Y_1"]) < 1e-6
    # the double-accumulate arbiter brackets both
    assert rel_err(orc.spmm_f64(tiled, g["X"], weighted=True), g["Y_w"]) < 1e-5


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_col_tile_matches_reference_golden(orc, name):
    g = golden(name)
    n, T = int(g["n"]), int(g["T"])
    assert np.array_equal(orc.col_breakpoints(n, T), g["breakpoints"])
    t = orc.col_tile(n, n, g["offset"], g["ids"], g["w"], T)
    assert np.array_equal(t.offsets, g["tile_offsets"])
    assert np.array_equal(t.cols, g["tile_cols"])
    assert np.array_equal(t.bounds, g["tile_bounds"])
    assert np.array_equal(t.vals, g["tile_vals"])
    # tiled SpMM == untiled SpMM (same per-row order)
    assert rel_err(orc.spmm(t, g["X"], weighted=True), g["Y_w"]) < 1e-6


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_sample_ab_matches_reference_golden(orc, name):
    g = golden(name)
    n = int(g["n"])
    rc, so, si, sv = orc.sample_ab(n, g["offset"], g["ids"], g["w"], 20, 5, 7)
    assert rc == 0
    assert np.array_equal(so, g["s_offset"])
    assert np.array_equal(si, g["s_ids"])
    assert np.array_equal(sv, g["s_vals"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mask_subgraph_matches_reference_golden(orc, name):
    g = golden(name)
    n = int(g["n"])
    ones = np.ones(g["ids"].shape[0], np.float32)
    (fo, fi, fv, bo, bi, bv), = orc.mask_subgraphs(n, n, g["offset"], g["ids"], ones, g["mask"], 1)
    assert np.array_equal(fo, g["m_fwd_offset"]) and np.array_equal(fi, g["m_fwd_ids"])
    assert np.array_equal(bo, g["m_bwd_offset"]) and np.array_equal(bi, g["m_bwd_ids"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_reorder_matches_reference_golden(orc, name):
    """rowReorderToAdj / rowPermuteDenseTo / rowPermuteDenseFrom (src/ops/reordering.h)."""
    g = golden(name)
    n = int(g["n"])
    ro, ri, rv = orc.csr_reorder(n, g["offset"], g["ids"], g["w"], g["perm"])
    assert np.array_equal(ro, g["r_offset"]) and np.array_equal(ri, g["r_ids"])
    assert np.array_equal(rv, g["r_vals"])          # small_dup: duplicate edges ordered by value
    assert np.array_equal(orc.permute_rows(g["X"], g["perm"], False), g["X_to"])
    assert np.array_equal(orc.permute_rows(g["X"], g["perm"], True), g["X_from"])


def test_degree_order_is_stable_descending(orc):
    offset, ids = make_csr(500, 6000, 3, empty_rows=20)
    perm, order = orc.degree_order(500, offset)
    deg = np.diff(offset)
    assert np.array_equal(order, np.argsort(-deg, kind="stable"))
    assert np.array_equal(perm[order], np.arange(500))


def test_sampled_spmm_equals_spmm_on_sampled_graph(orc):
    """K1s (sampling inside the kernel, cuda.h:313-320) and inplace_sample_graph_ab
    (tiling.h:454-508) pick the same multiset of edges per row."""
    n = 300
    offset, ids = make_csr(n, 6000, 5)
    ones = np.ones(ids.shape[0], np.float32)
    X = np.random.default_rng(0).uniform(-0.5, 0.5, (n, 16)).astype(np.float32)
    g = orc.Tiled.from_csr(n, n, offset, ids)
    Y1 = orc.spmm_sampled(g, X, 20, 5, 7)
    rc, so, si, sv = orc.sample_ab(n, offset, ids, ones, 20, 5, 7)
    Y2 = orc.spmm(orc.Tiled.from_csr(n, n, so, si, sv), X, weighted=False)
    assert rel_err(Y1, Y2) < 1e-6


def test_edge_kernels_hand_checked(orc):
    """K3..K7 + softmax exist in the reference only as CUDA text: pin the restatement on a
    4-node graph worked out by hand (2 segments to exercise the per-segment 1e-12 seed)."""
    # rows: 0:{0,2} 1:{1} 2:{0,2,3} 3:{}   columns split at 2 -> seg0 cols {0,1}, seg1 {2,3}
    offsets = np.array([0, 1, 2, 3, 3, 0, 1, 1, 3, 3], np.int32)
    cols = np.array([0, 1, 0, 2, 2, 3], np.int32)
    bounds = np.array([0, 3, 3, 6], np.int32)
    g = orc.Tiled(4, 4, 2, offsets, cols, np.ones(6, np.float32), bounds)
    A = np.array([1., 2., 3., 4.], np.float32)
    B = np.array([10., 20., 30., 40.], np.float32)
    add = orc.sddvv(g, A, B, "add")
    assert np.array_equal(add, np.array([11, 22, 13, 31, 33, 43], np.float32))
    mul = orc.sddvv(g, A, B, "mul")
    assert np.array_equal(mul, np.array([10, 40, 30, 30, 90, 120], np.float32))
    rs = orc.edge_rowsum(g, add)
    assert np.allclose(rs, [11 + 31, 22, 13 + 33 + 43, 0], rtol=1e-6)
    assert rs[3] == np.float32(2e-12)   # empty row: one 1e-12 seed per segment
    sc = orc.edge_scale_rows(g, add, A)
    assert np.array_equal(sc, np.array([11, 44, 39, 31, 99, 129], np.float32))
    x = np.array([0., 1., -1., 2., 0.5, 30.], np.float32)
    alpha, recip = orc.edge_softmax_fwd(g, x)
    ex = np.minimum(np.exp(x.astype(np.float64)), 1e12)
    want = np.array([ex[0] / (ex[0] + ex[3]), 1.0, ex[2] / (ex[2] + ex[4] + ex[5]),
                     ex[3] / (ex[0] + ex[3]), ex[4] / (ex[2] + ex[4] + ex[5]),
                     ex[5] / (ex[2] + ex[4] + ex[5])])
    assert np.allclose(alpha, want, rtol=1e-6)
    da = np.array([1., -2., 0.5, 3., 1., 2.], np.float32)
    out = orc.edge_softmax_bwd(g, alpha, da)
    s = alpha.astype(np.float64) * da
    acc = np.array([s[0] + s[3], s[1], s[2] + s[4] + s[5]])
    row = np.array([0, 1, 2, 0, 2, 2])
    assert np.allclose(out, s - alpha * acc[row], rtol=1e-5, atol=1e-7)
    Am = np.arange(12, dtype=np.float32).reshape(4, 3)
    Bm = (np.arange(12, dtype=np.float32).reshape(4, 3) - 5) * 0.5
    dd = orc.sddmm(g, Am, Bm)
    assert np.allclose(dd, [(Am[r] * Bm[c]).sum() for r, c in zip(row, cols)])
    assert np.allclose(orc.leaky_relu(np.array([-1., 2.], np.float32)), [-0.2, 2.0])


def test_softmax_clamp_no_max_subtraction(orc):
    """exp overflow is clamped to 1e12, not rescued by max-subtraction (common.h:760-761)."""
    offset = np.array([0, 2], np.int32)
    g = orc.Tiled.from_csr(1, 1, offset, np.array([0, 0], np.int32))
    alpha, _ = orc.edge_softmax_fwd(g, np.array([100.0, 0.0], np.float32))
    assert np.isclose(alpha[0], 1.0, rtol=1e-6) and np.isclose(alpha[1], 1e-12, rtol=1e-3)
    alpha, _ = orc.edge_softmax_fwd(g, np.array([100.0, 90.0], np.float32))
    assert np.allclose(alpha, [0.5, 0.5])


def test_gat_forward_composition(orc):
    n = 200
    offset, ids = make_csr(n, 3000, 9)
    rng = np.random.default_rng(1)
    aL, aR = rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)
    X = rng.uniform(-0.5, 0.5, (n, 8)).astype(np.float32)
    g = orc.Tiled.from_csr(n, n, offset, ids)
    Y, alpha = orc.gat_forward(g, aL, aR, X)
    rows = np.repeat(np.arange(n), np.diff(offset))
    sums = np.bincount(rows, weights=alpha, minlength=n)
    assert np.allclose(sums, 1.0, rtol=1e-5)
    assert rel_err(Y, orc.spmm(g, X, vals=alpha)) < 1e-7


def test_oracle_matches_reference_on_random_inputs(orc):
    if not orc.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    rng = np.random.default_rng(3)
    for seed, (n, e, K, T) in enumerate([(150, 1500, 7, 40), (700, 20000, 33, 300), (64, 64, 1, 1000)]):
        offset0, ids0 = make_csr(n, e, 20 + seed)
        rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(offset0))
        p = rng.permutation(ids0.shape[0])
        ro, ri, rv = orc.ref_csr_build(n, n, rows[p], ids0[p])
        oo, oi, ov = orc.csr_build(n, rows[p], ids0[p])
        assert np.array_equal(ro, oo) and np.array_equal(ri, oi)
        assert np.array_equal(ro, offset0) and np.array_equal(ri, ids0)
        w = rng.uniform(-1, 1, ids0.shape[0]).astype(np.float32)
        X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
        assert rel_err(orc.gspmm_wsum(n, ro, ri, w, X), orc.ref_gspmm_wsum(n, n, ro, ri, w, X)) < 1e-6
        bp, rt = orc.ref_col_tile(n, n, ro, ri, w, T)
        ot = orc.col_tile(n, n, ro, ri, w, T)
        for a, b in ((rt.offsets, ot.offsets), (rt.cols, ot.cols), (rt.vals, ot.vals), (rt.bounds, ot.bounds)):
            assert np.array_equal(a, b)
        rs = orc.ref_sample_ab(n, n, ro, ri, w, 20, 5, 7)
        rc, *os_ = orc.sample_ab(n, ro, ri, w, 20, 5, 7)
        assert rc == 0 and all(np.array_equal(a, b) for a, b in zip(rs, os_))
        tt = orc.ref_csr_transpose(n, n, ro, ri, np.ones_like(w))
        ot2 = orc.csr_transpose(n, n, ro, ri, np.ones_like(w))
        assert all(np.array_equal(a, b) for a, b in zip(tt, ot2))


def test_cora_gcn_forward_matches_reference_cpu_path(orc):
    """BASELINE.json configs[0]: 2-layer GCN on the Cora shape.  The golden logits come from the reference's own
    CPU path (CSRCMatrix::build + gSpMM<wsumAgg> + torch-CPU Linear, tests/golden/make_golden.py)."""
    import torch
    import torch.nn.functional as F

    g = golden("cora_gcn")
    n, feats = int(g["n"]), int(g["feats"])
    X = np.random.default_rng(int(g["x_seed"])).uniform(-0.5, 0.5, (n, feats)).astype(np.float32)
    ones = np.ones(g["ids"].shape[0], np.float32)

    def agg(x):
        return orc.gspmm_wsum(n, g["offset"], g["ids"], ones, np.ascontiguousarray(x, np.float32))
    deg = agg(np.ones((n, 1), np.float32))
    assert np.array_equal(deg.ravel(), np.diff(g["offset"]).astype(np.float32))      # degrees are exact
    norm = torch.pow(torch.from_numpy(deg), -0.5).numpy()
    assert np.array_equal(norm.ravel(), g["norm"])
    res = F.linear(torch.from_numpy(X), torch.from_numpy(g["W0"]), torch.from_numpy(g["b0"])).numpy()
    res = np.maximum(norm * agg(norm * res), 0)
    res = norm * agg(norm * res)
    logits = F.linear(torch.from_numpy(res), torch.from_numpy(g["W1"]), torch.from_numpy(g["b1"])).numpy()
    assert rel_err(logits, g["logits"]) < 1e-6
