"""GPU parity tests (-m gpu): every kernel, called through the C-ABI, against the oracle
on the same seeded inputs.  Integer / index work bit-exact; fp32 within 1e-5 relative
error (BASELINE.json north_star), measured norm-wise per output with the
double-accumulate arbiter alongside (SURVEY.md section 7 'summation order')."""
import numpy as np
import pytest
import torch

from gala_b200 import emitted, ops
from util import FP32_TOL, GOLDEN_CASES, golden, make_csr, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def to_gpu_graph(t, plan_threshold=None):
    g = ops.TiledGraph(dev(t.offsets), dev(t.cols), t.nrows, t.ncols, t.bounds, t.S)
    if plan_threshold is not None:
        g.build_plan(plan_threshold)
    return g


def graph_case(orc, n, e, seed, T=None, empty_rows=0):
    offset, ids = make_csr(n, e, seed, empty_rows)
    w = np.random.default_rng(seed).uniform(-1, 1, ids.shape[0]).astype(np.float32)
    if T is None:
        return orc.Tiled.from_csr(n, n, offset, ids, w)
    return orc.col_tile(n, n, offset, ids, w, T)


CASES = [
    # n, e, seed, T (None = untiled), empty rows, hub threshold (None = no plan)
    (2708, 13264, 1, None, 0, None),        # Cora shape
    (2708, 13264, 1, 100000, 0, 64),        # Cora shape, shipped col_tile(100000) -> S=1, plan
    (3000, 400000, 2, None, 0, 256),        # dense-ish power law, hub rows on CTAs
    (3000, 400000, 2, 700, 0, 256),         # 5 column segments + hubs
    (1500, 30000, 3, 400, 40, None),        # empty rows, 4 segments, no plan
    (257, 3000, 4, 16, 5, 32),              # 17 segments
]
KS = [1, 7, 32, 41, 64, 100, 128, 602]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("K", KS)
@pytest.mark.parametrize("weighted", [False, True])
def test_spmm_matches_oracle(orc, case, K, weighted):
    n, e, seed, T, empty, thr = case
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    X = np.random.default_rng(seed + K).uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    want = orc.spmm(t, X, weighted=weighted)
    got = ops.spmm(g, dev(X), vals=dev(t.vals) if weighted else None).cpu().numpy()
    exact = orc.spmm_f64(t, X, weighted=weighted)
    assert rel_err(got, want) < FP32_TOL
    # no worse than the reference's own serial fp32 order, judged by the fp64 arbiter
    assert rel_err(got, exact) <= max(2.0 * rel_err(want, exact), 2e-7)
    # element-wise: within 1e-5 of the magnitude that was summed (backward-error bound)
    mag = orc.spmm_f64(orc.Tiled(t.nrows, t.ncols, t.S, t.offsets, t.cols, np.abs(t.vals), t.bounds),
                       np.abs(X), weighted=weighted)
    assert np.all(np.abs(got - exact) <= 1e-5 * mag + 1e-30)


@pytest.mark.parametrize("case", [CASES[3], CASES[4], CASES[5]])
@pytest.mark.parametrize("K", [1, 32, 100])
def test_spmm_segment_major_schedule(orc, case, K):
    """One launch per column segment (the schedule used when X exceeds the L2) == single launch == oracle,
    including the fused row/col scale + ReLU epilogue and user-level accumulation."""
    n, e, seed, T, empty, thr = case
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(seed + K)
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    norm = rng.uniform(0.1, 1.0, n).astype(np.float32)
    want = orc.spmm(t, X, weighted=True)
    a = ops.spmm(g, dev(X), vals=dev(t.vals), schedule="segment_major").cpu().numpy()
    b = ops.spmm(g, dev(X), vals=dev(t.vals), schedule="row_major").cpu().numpy()
    assert rel_err(a, want) < FP32_TOL and rel_err(b, want) < FP32_TOL
    want2 = np.maximum(norm[:, None] * orc.spmm(t, norm[:, None] * X, weighted=False), 0)
    a2 = ops.spmm(g, dev(X), row_scale=dev(norm), col_scale=dev(norm), relu=True, schedule="segment_major").cpu().numpy()
    assert rel_err(a2, want2) < FP32_TOL
    Y0 = rng.uniform(-1, 1, (n, K)).astype(np.float32)
    out = dev(Y0.copy())
    ops.spmm(g, dev(X), vals=dev(t.vals), out=out, accumulate=True, schedule="segment_major")
    assert rel_err(out.cpu().numpy(), Y0 + want) < FP32_TOL


def test_spmm_degrees_are_exact(orc):
    """SpMM(A, ones) = degrees: small integers, exact in fp32 whatever the order
    (the generated GCN computes its normalisation this way, codegen/gala.cu:437)."""
    t = graph_case(orc, 3000, 400000, 2, 700)
    g = to_gpu_graph(t, 128)
    got = ops.spmm(g, torch.ones(3000, 1, device=DEV)).cpu().numpy().ravel()
    deg = np.zeros(3000, np.int64)
    for s in range(t.S):
        deg += np.diff(t.offsets[s * 3001:(s + 1) * 3001])
    assert np.array_equal(got, deg.astype(np.float32))


def test_spmm_epilogue_and_accumulate(orc):
    n, K = 2000, 32
    t = graph_case(orc, n, 60000, 5, 500)
    g = to_gpu_graph(t, 128)
    rng = np.random.default_rng(0)
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    norm = rng.uniform(0.1, 1.0, n).astype(np.float32)
    Y0 = rng.uniform(-1, 1, (n, K)).astype(np.float32)
    # GCN layer body: norm * (A @ (norm * X)) then relu  (codegen/gala.cu:441-450)
    want = norm[:, None] * orc.spmm(t, norm[:, None] * X, weighted=False)
    got = ops.spmm(g, dev(X), row_scale=dev(norm), col_scale=dev(norm), relu=True).cpu().numpy()
    assert rel_err(got, np.maximum(want, 0)) < FP32_TOL
    # accumulate: Y += A @ X, the reference's per-segment `C = C + ...` (cuda.h:309-351)
    out = dev(Y0.copy())
    ops.spmm(g, dev(X), vals=dev(t.vals), out=out, accumulate=True)
    assert rel_err(out.cpu().numpy(), orc.spmm(t, X, weighted=True, Y=Y0.copy())) < FP32_TOL


@pytest.mark.parametrize("K", [1, 32, 47, 100])
@pytest.mark.parametrize("T", [None, 600])
def test_spmm_sampled_matches_oracle(orc, K, T):
    n = 2500
    t = graph_case(orc, n, 80000, 6, T)
    g = to_gpu_graph(t)
    X = np.random.default_rng(K).uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    for (s, ra, rb) in ((20, 5, 7), (3, 97, 13), (33, 0, 100)):
        want = orc.spmm_sampled(t, X, s, ra, rb)
        got = ops.spmm_sampled(g, dev(X), s, ra, rb).cpu().numpy()
        assert rel_err(got, want) < FP32_TOL
    want = orc.spmm_sampled(t, X, 20, 5, 7, weighted=True)
    got = ops.spmm_sampled(g, dev(X), 20, 5, 7, vals=dev(t.vals)).cpu().numpy()
    assert rel_err(got, want) < FP32_TOL


@pytest.mark.parametrize("case", CASES)
def test_edge_kernels_match_oracle(orc, case):
    n, e, seed, T, empty, thr = case
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(seed)
    A = rng.normal(size=n).astype(np.float32)
    B = rng.normal(size=n).astype(np.float32)
    # K5 / K7: one rounding per edge -> bit-exact
    for op in ("add", "mul"):
        got = ops.sddvv(g, dev(A), dev(B), op).cpu().numpy()
        assert np.array_equal(got, orc.sddvv(t, A, B, op))
    got = ops.sddvv(g, dev(A), dev(B), "add", leaky_slope=0.2).cpu().numpy()
    assert np.array_equal(got, orc.leaky_relu(orc.sddvv(t, A, B, "add"), 0.2))
    # K4: bit-exact
    v = dev(t.vals.copy())
    ops.edge_scale_rows_(g, v, dev(A))
    assert np.array_equal(v.cpu().numpy(), orc.edge_scale_rows(t, t.vals, A))
    # K3
    pos = np.abs(t.vals) + 0.01
    got = ops.edge_rowsum(g, dev(pos)).cpu().numpy().ravel()
    assert rel_err(got, orc.edge_rowsum(t, pos)) < FP32_TOL
    deg = got * 0
    for s in range(t.S):
        deg += np.diff(t.offsets[s * (n + 1):(s + 1) * (n + 1)])
    assert np.allclose(got[deg == 0], t.S * np.float32(1e-12), rtol=1e-6)   # per-segment seed
    # softmax forward / backward
    x = rng.normal(scale=2.0, size=t.nvals).astype(np.float32)
    want, recip = orc.edge_softmax_fwd(t, x)
    r = torch.empty(n, device=DEV)
    got = ops.edge_softmax_fwd(g, dev(x), recip=r).cpu().numpy()
    assert rel_err(got, want) < FP32_TOL
    assert np.allclose(got, want, rtol=1e-5, atol=1e-12)
    assert rel_err(r.cpu().numpy()[deg > 0], recip[deg > 0]) < FP32_TOL
    da = rng.normal(size=t.nvals).astype(np.float32)
    got_b = ops.edge_softmax_bwd(g, dev(want), dev(da)).cpu().numpy()
    want_b = orc.edge_softmax_bwd(t, want, da)
    assert rel_err(got_b, want_b) < FP32_TOL
    # edge side of the GAT layer backward in one kernel (softmax bwd + LeakyReLU bwd + row sum)
    aL = rng.normal(size=n).astype(np.float32)
    aR = rng.normal(size=n).astype(np.float32)
    got_g = ops.gat_backward_att(g, dev(want), dev(da), dev(aL), dev(aR), 0.2).cpu().numpy().ravel()
    want_g = orc.gat_backward_att(t, want, da, aL, aR, 0.2)
    # ds = alpha*dalpha - alpha*tot cancels, and so does its row sum: backward-error bound against the
    # magnitudes that enter the sums (sum |alpha*dalpha| + |tot| * sum alpha)
    mag = orc.edge_rowsum(t, np.abs(want * da)) + np.abs(orc.edge_rowsum(t, want * da)) * orc.edge_rowsum(t, want)
    assert np.all(np.abs(got_g - want_g) <= 4e-6 * mag + 1e-10)
    # in place (x aliases alpha), as the generated code does with val_exp
    xi = dev(x.copy())
    ops.edge_softmax_fwd(g, xi, out=xi)
    assert np.array_equal(xi.cpu().numpy(), got)


@pytest.mark.parametrize("case", CASES[:5])
@pytest.mark.parametrize("K", [1, 7, 32, 64, 100, 602, 1433])
def test_sddmm_matches_oracle(orc, case, K):
    n, e, seed, T, empty, thr = case
    if K > 602 and n > 2708:
        pytest.skip("large K only on the Cora shape")
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(seed + K)
    A = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    B = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    got = ops.sddmm(g, dev(A), dev(B)).cpu().numpy()
    want = orc.sddmm(t, A, B)
    assert rel_err(got, want) < FP32_TOL


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("K", [8, 32, 41, 128])
def test_gat_fused_forward_matches_composition(orc, case, K):
    """The fused kernel against the oracle's op-by-op composition of the emitted GAT layer
    (edge_sddvv -> LeakyReLU -> softmax -> weighted SpMM)."""
    n, e, seed, T, empty, thr = case
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(seed + K)
    aL = rng.normal(size=n).astype(np.float32)
    aR = rng.normal(size=n).astype(np.float32)
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    want_Y, want_alpha = orc.gat_forward(t, aL, aR, X)
    alpha = torch.empty(t.nvals, device=DEV)
    got_Y = ops.gat_forward(g, dev(aL), dev(aR), dev(X), alpha_out=alpha).cpu().numpy()
    assert rel_err(got_Y, want_Y) < FP32_TOL
    assert rel_err(alpha.cpu().numpy(), want_alpha) < FP32_TOL
    got_relu = ops.gat_forward(g, dev(aL), dev(aR), dev(X), relu=True).cpu().numpy()
    assert rel_err(got_relu, np.maximum(want_Y, 0)) < FP32_TOL
    # and against the unfused sequence of our own kernels
    att = ops.sddvv(g, dev(aL), dev(aR), "add", leaky_slope=0.2)
    ops.edge_softmax_fwd(g, att, out=att)
    unfused = ops.spmm(g, dev(X), vals=att).cpu().numpy()
    assert rel_err(got_Y, unfused) < FP32_TOL


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("K", [4, 8, 12, 16, 32])
def test_gat_dot_variant_matches_composition(orc, case, K):
    """gala_gat_forward_dot_f32: aR recomputed from the gathered rows, aR[j] = X[j,:].wR + bR."""
    n, e, seed, T, empty, thr = case
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(seed + K)
    aL = rng.normal(size=n).astype(np.float32)
    wR = rng.normal(size=K).astype(np.float32)
    bR = 0.3
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    aR = (X.astype(np.float64) @ wR.astype(np.float64) + bR).astype(np.float32)
    want_Y, want_alpha = orc.gat_forward(t, aL, aR, X)
    alpha = torch.empty(t.nvals, device=DEV)
    got = ops.gat_forward_dot(g, dev(aL), dev(wR), bR, dev(X), alpha_out=alpha, relu=True).cpu().numpy()
    assert rel_err(got, np.maximum(want_Y, 0)) < FP32_TOL
    assert rel_err(alpha.cpu().numpy(), want_alpha) < FP32_TOL
    # a shape outside the kernel's range takes the documented fallback (materialised aR)
    if K == 12:
        X2 = rng.uniform(-0.5, 0.5, (n, 41)).astype(np.float32)
        w2 = rng.normal(size=41).astype(np.float32)
        aR2 = (X2.astype(np.float64) @ w2.astype(np.float64) + bR).astype(np.float32)
        want2, _ = orc.gat_forward(t, aL, aR2, X2)
        got2 = ops.gat_forward_dot(g, dev(aL), dev(w2), bR, dev(X2)).cpu().numpy()
        assert rel_err(got2, want2) < FP32_TOL


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("K", [4, 8, 16, 32])
def test_gat_col_variant_matches_oracle(orc, case, K):
    """gala_gat_forward_col_f32: features handed over in the reflected basis X' = X H (H = I - 2 v v^T maps the last
    unit vector onto the direction of the right-hand projection), aR read from the last column of the gathered row.
    The result must be the oracle's layer on the ORIGINAL X with the materialised aR = X.wR + bR."""
    n, e, seed, T, empty, thr = case
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(seed + 17 * K)
    aL = rng.normal(size=n).astype(np.float32)
    wR = rng.normal(size=K).astype(np.float32)
    if seed % 2:
        wR[-1] = -abs(wR[-1])            # both signs of the pivot component
    bR = 0.3
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    aR = (X.astype(np.float64) @ wR.astype(np.float64) + bR).astype(np.float32)
    want_Y, want_alpha = orc.gat_forward(t, aL, aR, X)
    v, sR = ops.reflection(dev(wR))
    vd = v.double().cpu().numpy()
    H = np.eye(K) - 2.0 * np.outer(vd, vd)
    assert np.abs(H @ H - np.eye(K)).max() < 1e-6          # float32 v: orthogonal to rounding
    assert np.abs(sR * H[:, -1] - wR).max() < 1e-5 * np.abs(wR).max()
    Xr = (X.astype(np.float64) @ H).astype(np.float32)
    alpha = torch.empty(t.nvals, device=DEV)
    # back to the original basis, then ReLU
    got = ops.gat_forward_col(g, dev(aL), sR, bR, dev(Xr), relu=True, reflect_in=v, alpha_out=alpha).cpu().numpy()
    assert rel_err(got, np.maximum(want_Y, 0)) < FP32_TOL
    assert rel_err(alpha.cpu().numpy(), want_alpha) < FP32_TOL
    # ... and on into the next layer's basis (another reflection, after the ReLU)
    w2 = rng.normal(size=K).astype(np.float32)
    v2, _ = ops.reflection(dev(w2))
    v2d = v2.double().cpu().numpy()
    H2 = np.eye(K) - 2.0 * np.outer(v2d, v2d)
    got2 = ops.gat_forward_col(g, dev(aL), sR, bR, dev(Xr), relu=True, reflect_in=v, reflect_out=v2).cpu().numpy()
    assert rel_err(got2, np.maximum(want_Y, 0) @ H2) < FP32_TOL
    # no reflections: the sum stays in the gathered basis
    got3 = ops.gat_forward_col(g, dev(aL), sR, bR, dev(Xr)).cpu().numpy()
    assert rel_err(got3, want_Y @ H) < FP32_TOL


def test_forward_host_back_to_back_calls_do_not_clobber_the_staging_buffer(orc):
    """GAT2.forward_host lets the next call's upload start while the previous call's aggregation still runs; three
    un-synchronised calls with different inputs through ONE staging buffer must each give their own forward."""
    from gala_b200.gat_model import GAT2
    n, F_in = 30000, 602
    t = graph_case(orc, n, 3000000, 21)
    g = to_gpu_graph(t, 2048)
    model = GAT2(F_in, 32, 41, DEV, seed=9)
    gen = torch.Generator().manual_seed(4)
    Xs = [(torch.rand(n, F_in, generator=gen) - 0.5).pin_memory() for _ in range(2)]
    outs = [torch.empty(n, 41).pin_memory() for _ in range(3)]
    stage = torch.empty(n, F_in, device=DEV)
    for mode in ("reflected", "folded"):
        for i, X in enumerate((Xs[0], Xs[1], Xs[0])):
            model.forward_host(g, X, outs[i], chunks=4, mode=mode, stage=stage)
        torch.cuda.synchronize()
        want = [model.forward(g, X.to(DEV), mode="literal", dense="torch").cpu() for X in Xs]
        for i, w in enumerate((want[0], want[1], want[0])):
            assert float((outs[i] - w).double().norm() / w.double().norm()) < FP32_TOL, (mode, i)


def test_gatn_reflected_and_folded_forwards_match_literal(orc):
    """3-layer GATN (the Papers-shape program): own-kernel forward in the original and in the reflected basis against
    the op-by-op forward with cuBLAS dense parts."""
    from gala_b200.gat_model import GATN
    n = 4000
    t = graph_case(orc, n, 150000, 7)
    g = to_gpu_graph(t, 256)
    for dims in ([64, 32, 32, 41], [20, 16, 32, 172], [33, 32, 7]):
        model = GATN(dims, DEV, seed=5).host_biases()
        X = torch.rand(n, dims[0], device=DEV) - 0.5
        want = model.forward_literal(g, X)
        for mode in ("folded", "reflected"):
            got = model.forward(g, X, mode=mode)
            assert float((got - want).double().norm() / want.double().norm()) < FP32_TOL, (dims, mode)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("K,C", [(32, 41), (8, 3), (16, 47)])
def test_gat_col_dense_epilogue(orc, case, K, C):
    """gala_gat_forward_col_f32 with the dense epilogue: projections and classifier of the FINAL row (after both
    reflections), against the oracle's layer + numpy."""
    n, e, seed, T, empty, thr = case
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(seed + K + C)
    aL = rng.normal(size=n).astype(np.float32)
    wR = rng.normal(size=K).astype(np.float32)
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    aR = (X.astype(np.float64) @ wR.astype(np.float64) + 0.1).astype(np.float32)
    want_Y, _ = orc.gat_forward(t, aL, aR, X)
    want_Y = np.maximum(want_Y, 0).astype(np.float64)
    v, sR = ops.reflection(dev(wR))
    v2, _ = ops.reflection(dev(rng.normal(size=K).astype(np.float32)))
    H = np.eye(K) - 2.0 * np.outer(v.double().cpu().numpy(), v.double().cpu().numpy())
    H2 = np.eye(K) - 2.0 * np.outer(v2.double().cpu().numpy(), v2.double().cpu().numpy())
    Xr = (X.astype(np.float64) @ H).astype(np.float32)
    att_w = rng.normal(size=(2, K)).astype(np.float32)
    att_b = [0.25, -0.5]
    cls_w = rng.normal(size=(C, K)).astype(np.float32)
    cls_b = rng.normal(size=C).astype(np.float32)
    Y, att, cls = ops.gat_forward_col_ex(g, dev(aL), sR, 0.1, dev(Xr), relu=True, reflect_in=v, reflect_out=v2,
                                         att_w=dev(att_w), att_b=att_b, cls_wT=dev(np.ascontiguousarray(cls_w.T)),
                                         cls_b=dev(cls_b))
    final = want_Y @ H2
    assert rel_err(Y.cpu().numpy(), final) < FP32_TOL
    assert rel_err(att.cpu().numpy(), (final @ att_w.T.astype(np.float64) + np.array(att_b)).T) < FP32_TOL
    assert rel_err(cls.cpu().numpy(), final @ cls_w.T.astype(np.float64) + cls_b) < FP32_TOL
    # projections alone: the register path writes the left-hand one only
    Y3, att3, _ = ops.gat_forward_col_ex(g, dev(aL), sR, 0.1, dev(Xr), relu=True, reflect_in=v, reflect_out=v2,
                                         att_w=dev(att_w), att_b=att_b)
    assert rel_err(Y3.cpu().numpy(), final) < FP32_TOL
    assert rel_err(att3[0].cpu().numpy(), final @ att_w[0].astype(np.float64) + att_b[0]) < FP32_TOL
    # classifier only (Y not written)
    Y2, _, cls2 = ops.gat_forward_col_ex(g, dev(aL), sR, 0.1, dev(Xr), relu=True, reflect_in=v,
                                         cls_wT=dev(np.ascontiguousarray(cls_w.T)), cls_b=dev(cls_b), want_y=False)
    assert Y2 is None
    assert rel_err(cls2.cpu().numpy(), want_Y @ cls_w.T.astype(np.float64) + cls_b) < FP32_TOL


def test_gat_col_rejects_unsupported_widths(orc):
    t = graph_case(orc, 300, 3000, 5)
    g = to_gpu_graph(t, 64)
    X = torch.rand(300, 12, device=DEV)
    with pytest.raises(Exception):
        ops.gat_forward_col(g, X[:, 0].contiguous(), 1.0, 0.0, X)


def test_gat_model_dot_and_materialised_paths_agree(orc):
    from gala_b200.gat_model import GAT2
    n = 3000
    t = graph_case(orc, n, 200000, 11)
    g = to_gpu_graph(t, 256)
    model = GAT2(64, 32, 41, DEV, seed=3)
    X = torch.rand(n, 64, device=DEV) - 0.5
    b = model.forward(g, X, mode="literal", dense="torch")
    for mode in ("folded", "dot", "literal", "fused", "reflected", "reflected_fused"):
        for dense in ("torch", "tcgen05"):
            a = model.forward(g, X, mode=mode, dense=dense)
            assert float((a - b).double().norm() / b.double().norm()) < FP32_TOL


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("K,C", [(32, 41), (8, 3), (64, 64), (100, 47), (128, 7)])
def test_gat_dense_epilogue(orc, case, K, C):
    """gala_gat_forward_ex_f32: next-layer attention projections + classifier fused on the output rows."""
    n, e, seed, T, empty, thr = case
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(seed + K + C)
    aL, aR = rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    att_w = rng.normal(size=(2, K)).astype(np.float32)
    att_b = [0.1, -0.2]
    Wc = rng.normal(size=(C, K)).astype(np.float32)
    bc = rng.normal(size=C).astype(np.float32)
    want_Y, _ = orc.gat_forward(t, aL, aR, X)
    want_Y = np.maximum(want_Y, 0)
    y, att, cls = ops.gat_forward_ex(g, dev(aL), dev(aR), dev(X), relu=True, att_w=dev(att_w), att_b=att_b,
                                     cls_wT=dev(np.ascontiguousarray(Wc.T)), cls_b=dev(bc))
    assert rel_err(y.cpu().numpy(), want_Y) < FP32_TOL
    assert rel_err(att.cpu().numpy().T, want_Y.astype(np.float64) @ att_w.T.astype(np.float64) + np.array(att_b)) < FP32_TOL
    assert rel_err(cls.cpu().numpy(), want_Y.astype(np.float64) @ Wc.T.astype(np.float64) + bc) < FP32_TOL
    _, _, cls2 = ops.gat_forward_ex(g, dev(aL), dev(aR), dev(X), relu=True, cls_wT=dev(np.ascontiguousarray(Wc.T)),
                                    cls_b=dev(bc), want_y=False)
    assert torch.equal(cls2, cls)


def test_gcn_model_fused_matches_literal_and_oracle(orc):
    """2-layer GCN as generated (codegen/gala.cu:422-459): fused epilogues == op-by-op == oracle."""
    from gala_b200.gcn_model import GCN2
    n = 3000
    t = graph_case(orc, n, 200000, 12, 700)
    g = to_gpu_graph(t, 256)
    model = GCN2(100, 32, 47, DEV, seed=4).prepare(g)
    X = torch.rand(n, 100, device=DEV) - 0.5
    lit = model.forward_literal(g, X)
    for dense in ("torch", "tcgen05"):
        fused = model.forward(g, X, dense=dense)
        assert float((fused - lit).double().norm() / lit.double().norm()) < FP32_TOL
    # oracle composition on the host
    import torch.nn.functional as F
    Xc = X.cpu()
    deg = orc.spmm(t, np.ones((n, 1), np.float32), weighted=False).ravel()
    norm = (deg ** -0.5).astype(np.float32)
    res = F.linear(Xc, *[w.cpu() for w in model.fc0]).numpy() * norm[:, None]
    res = np.maximum(orc.spmm(t, res, weighted=False) * norm[:, None], 0) * norm[:, None]
    res = orc.spmm(t, res.astype(np.float32), weighted=False) * norm[:, None]
    want = F.linear(torch.from_numpy(res.astype(np.float32)), *[w.cpu() for w in model.fc1]).numpy()
    assert rel_err(lit.cpu().numpy(), want) < FP32_TOL


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_fixtures_on_gpu(orc, name):
    """The reference's own outputs (tests/golden) reproduced by the CUDA path."""
    gd = golden(name)
    n = int(gd["n"])
    S = gd["tile_bounds"].shape[0] // 2
    g = ops.TiledGraph(dev(gd["tile_offsets"]), dev(gd["tile_cols"]), n, n, gd["tile_bounds"], S)
    g.build_plan(64)
    got = ops.spmm(g, dev(gd["X"]), vals=dev(gd["tile_vals"])).cpu().numpy()
    assert rel_err(got, gd["Y_w"]) < FP32_TOL
    got = ops.spmm(g, dev(gd["X"])).cpu().numpy()
    assert rel_err(got, gd["Y_1"]) < FP32_TOL
    g1 = ops.TiledGraph(dev(gd["offset"]), dev(gd["ids"]), n)
    got = ops.spmm(g1, dev(gd["X"]), vals=dev(gd["w"])).cpu().numpy()
    assert rel_err(got, gd["Y_w"]) < FP32_TOL
    # sampled kernel == SpMM over the reference's sampled graph
    gs = ops.TiledGraph(dev(gd["s_offset"]), dev(gd["s_ids"]), n)
    a = ops.spmm(gs, dev(gd["X"]), vals=dev(gd["s_vals"])).cpu().numpy()
    b = ops.spmm_sampled(g1, dev(gd["X"]), 20, 5, 7, vals=dev(gd["w"])).cpu().numpy()
    assert rel_err(a, b) < FP32_TOL


def test_emitted_interface(orc):
    """Calls spelled exactly as the generated gala.cu spells them (cuda.h:441-952)."""
    n = 1200
    t = graph_case(orc, n, 50000, 8, 300)
    emitted.clear_cache()
    emitted.global_nrows = n
    off, col, val = dev(t.offsets), dev(t.cols), dev(t.vals)
    bounds = torch.from_numpy(t.bounds)   # CPU tensor, as in the reference
    rng = np.random.default_rng(0)
    X = rng.uniform(-0.5, 0.5, (n, 32)).astype(np.float32)
    y = emitted.aggregate_node_mul_sum_call(dev(X), off, col, val, bounds, t.S)
    assert rel_err(y.cpu().numpy(), orc.spmm(t, X, weighted=True)) < FP32_TOL
    y = emitted.aggregate_node_mul_sum_direct_call(dev(X), off, col, val, bounds, t.S)
    assert rel_err(y.cpu().numpy(), orc.spmm(t, X, weighted=False)) < FP32_TOL
    ones = torch.ones(n, 1, device=DEV)
    deg = emitted.aggregate_node_mul_sum_direct_call(ones, off, col, val, bounds, t.S)
    a = rng.normal(size=(n, 1)).astype(np.float32)
    b = rng.normal(size=(n, 1)).astype(np.float32)
    att = emitted.edge_sddvv(dev(a), dev(b), off, col, val, bounds, n, t.S)
    assert np.array_equal(att.cpu().numpy(), orc.sddvv(t, a.ravel(), b.ravel(), "add"))
    att = torch.nn.functional.leaky_relu(att, 0.2)
    # body of non_lnr_op_softmax_AutoGrad::forward, common.h:760-773, op by op
    val_exp = torch.clamp(torch.exp(att), 0.0, 1e12)
    row_sum = emitted.node_spmv_backward_of_sddmm_nln(off, col, val_exp, bounds, n, t.S)
    row_sum = torch.reciprocal(row_sum)
    val_exp = emitted.inplace_softmax_sddvv(row_sum, off, col, val_exp, bounds, n, t.S)
    want, _ = orc.edge_softmax_fwd(t, orc.leaky_relu(orc.sddvv(t, a.ravel(), b.ravel(), "add")))
    assert rel_err(val_exp.cpu().numpy(), want) < FP32_TOL
    fused = emitted.non_lnr_op_softmax_forward(att, off, col, bounds, t.S)
    assert rel_err(fused.cpu().numpy(), want) < FP32_TOL
    nrm = torch.pow(deg, -0.5)
    ev = emitted.aggregate_edge_mul(nrm, nrm, off, col, val, bounds, t.S)
    dnp = deg.cpu().numpy().ravel() ** -0.5
    assert rel_err(ev.cpu().numpy(), orc.sddvv(t, dnp.astype(np.float32), dnp.astype(np.float32), "mul")) < 1e-6
    dz = rng.uniform(-0.5, 0.5, (n, 32)).astype(np.float32)
    dd = emitted.edge_sddmm(dev(dz), dev(X), off, col, val, bounds, n, t.S)
    assert rel_err(dd.cpu().numpy(), orc.sddmm(t, dz, X)) < FP32_TOL


def test_empty_and_degenerate_graphs(orc):
    # no rows
    g = ops.TiledGraph(torch.zeros(1, dtype=torch.int32, device=DEV),
                       torch.zeros(0, dtype=torch.int32, device=DEV), 0, 0)
    assert ops.spmm(g, torch.zeros(0, 8, device=DEV)).shape == (0, 8)
    # rows but no edges
    g = ops.TiledGraph(torch.zeros(6, dtype=torch.int32, device=DEV),
                       torch.zeros(0, dtype=torch.int32, device=DEV), 5, 5)
    y = ops.spmm(g, torch.ones(5, 8, device=DEV))
    assert torch.all(y == 0)
    r = ops.edge_rowsum(g, torch.zeros(0, device=DEV))
    assert torch.allclose(r, torch.full_like(r, 1e-12))
    y = ops.gat_forward(g, torch.zeros(5, device=DEV), torch.zeros(5, device=DEV),
                        torch.ones(5, 8, device=DEV))
    assert torch.all(y == 0)
    # one row holding every edge (max-size row on a hub CTA), K not a multiple of 4
    n = 20000
    offset = np.zeros(n + 1, np.int32)
    offset[1:] = n
    ids = np.arange(n, dtype=np.int32)
    t = orc.Tiled.from_csr(n, n, offset, ids)
    gg = to_gpu_graph(t, 1024)
    assert gg.plan.n_hub == 1
    X = np.random.default_rng(0).uniform(-0.5, 0.5, (n, 7)).astype(np.float32)
    got = ops.spmm(gg, dev(X)).cpu().numpy()
    assert rel_err(got[0], X.astype(np.float64).sum(0)) < FP32_TOL
    assert np.all(got[1:] == 0)


def test_misaligned_views_fall_back_to_narrow_loads(orc):
    n, K = 500, 32
    t = graph_case(orc, n, 8000, 10)
    g = to_gpu_graph(t)
    X = np.random.default_rng(1).uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    buf = torch.zeros(n * K + 1, device=DEV)
    buf[1:] = dev(X).ravel()
    Xv = buf[1:].view(n, K)                      # 4-byte aligned only
    assert Xv.data_ptr() % 16 != 0
    got = ops.spmm(g, Xv, pad=None).cpu().numpy()           # gathered where it lies: 4-byte loads
    assert rel_err(got, orc.spmm(t, X, weighted=False)) < FP32_TOL
    got = ops.spmm(g, Xv).cpu().numpy()                     # default: re-pitched once, 128-bit loads
    assert rel_err(got, orc.spmm(t, X, weighted=False)) < FP32_TOL


@pytest.mark.parametrize("case", [CASES[2], CASES[3], CASES[4]])
@pytest.mark.parametrize("K", [5, 7, 41, 47, 100, 602])
def test_row_pitched_operands(orc, case, K):
    """ABI v2 row pitches (ldx / ldy): odd-width rows stored with a 16-byte pitch are gathered with 128-bit loads
    (the reference runs K % 32 remainder kernels for them, cuda.h:58-168).  Every combination of packed / pitched
    X and Y gives the same result as the oracle on the packed data; the padding of Y is never written."""
    n, e, seed, T, empty, thr = case
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(seed + K)
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    want = orc.spmm(t, X, weighted=True)
    Xd = dev(X)
    ld = ops.pad_pitch(K)
    Xp = ops.pad_rows(Xd)
    assert Xp.stride(0) == ld and Xp.data_ptr() % 16 == 0 and torch.equal(Xp, Xd)
    if ld > K:
        assert float(Xp.as_strided((n, ld - K), (ld, 1), K).abs().max()) == 0.0   # padding zeroed
    packed_scalar = ops.spmm(g, Xd, vals=dev(t.vals), pad=None)           # packed rows, narrow loads
    pitched = ops.spmm(g, Xp, vals=dev(t.vals))                            # caller-provided pitched rows
    auto = ops.spmm(g, Xd, vals=dev(t.vals))                               # default: re-pitched internally (K <= 256)
    tight = ops.spmm(g, ops.pad_rows(Xd, (K + 3) // 4 * 4), vals=dev(t.vals))   # any 16-byte pitch works
    for got in (packed_scalar, pitched, auto, tight):
        assert got.is_contiguous() and rel_err(got.cpu().numpy(), want) < FP32_TOL
    assert torch.equal(pitched, tight)
    if K <= 256:
        assert torch.equal(pitched, auto)
    # pitched OUTPUT: the columns past K keep their sentinel
    Yp = torch.full((n, ld + 4), 7.0, device=DEV)
    ops.spmm(g, Xp, vals=dev(t.vals), out=Yp[:, :K])
    assert torch.equal(Yp[:, :K], pitched) and bool((Yp[:, K:] == 7.0).all())
    # accumulate + epilogue on pitched rows
    norm = dev(rng.uniform(0.1, 1.0, n).astype(np.float32))
    Y0 = dev(rng.uniform(-1, 1, (n, K)).astype(np.float32))
    Ya = Y0.clone()
    ops.spmm(g, Xp, vals=dev(t.vals), out=Ya, accumulate=True, relu=True, schedule="row_major")
    assert rel_err(Ya.cpu().numpy(), np.maximum(want + Y0.cpu().numpy(), 0)) < FP32_TOL
    y = ops.spmm(g, Xp, vals=dev(t.vals), row_scale=norm, col_scale=norm)
    ws = orc.Tiled(t.nrows, t.ncols, t.S, t.offsets, t.cols, t.vals * norm.cpu().numpy()[t.cols], t.bounds)
    assert rel_err(y.cpu().numpy(), norm.cpu().numpy()[:, None] * orc.spmm(ws, X, weighted=True)) < FP32_TOL
    # sampled flavour and the fused GAT layer take pitched rows too
    if K in (41, 47):
        ys = ops.spmm_sampled(g, Xd, 20, 5, 7)
        assert rel_err(ys.cpu().numpy(), orc.spmm_sampled(t, X, 20, 5, 7)) < FP32_TOL
        aL, aR = rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)
        yg, _ = orc.gat_forward(t, aL, aR, X)
        a_out = torch.empty(g.nvals, device=DEV)
        got = ops.gat_forward(g, dev(aL), dev(aR), Xd, alpha_out=a_out)
        assert rel_err(got.cpu().numpy(), yg) < FP32_TOL


@pytest.mark.parametrize("n,e,T,thr", [(2000, 40000, 20, 64), (1500, 30000, 7, None)])
def test_more_than_64_column_segments(orc, n, e, T, thr):
    """The reference accepts any number of column segments (tiling.h:222-283); beyond 64 the kernels derive the
    segment starts from the per-segment row pointers instead of taking them as parameters."""
    t = graph_case(orc, n, e, 77, T)
    assert t.S > 64
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(5)
    X = rng.uniform(-0.5, 0.5, (n, 32)).astype(np.float32)
    aL, aR = rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)
    assert rel_err(ops.spmm(g, dev(X), vals=dev(t.vals)).cpu().numpy(), orc.spmm(t, X, weighted=True)) < FP32_TOL
    assert rel_err(ops.spmm(g, dev(X), vals=dev(t.vals), schedule="segment_major").cpu().numpy(),
                   orc.spmm(t, X, weighted=True)) < FP32_TOL
    assert np.array_equal(ops.sddvv(g, dev(aL), dev(aR), "add").cpu().numpy(), orc.sddvv(t, aL, aR, "add"))
    assert rel_err(ops.edge_rowsum(g, dev(t.vals)).cpu().numpy().ravel(), orc.edge_rowsum(t, t.vals)) < FP32_TOL
    a_want, _ = orc.edge_softmax_fwd(t, t.vals)
    assert rel_err(ops.edge_softmax_fwd(g, dev(t.vals)).cpu().numpy(), a_want) < FP32_TOL
    y_want, _ = orc.gat_forward(t, aL, aR, X)
    assert rel_err(ops.gat_forward(g, dev(aL), dev(aR), dev(X)).cpu().numpy(), y_want) < FP32_TOL
    assert rel_err(ops.spmm_sampled(g, dev(X), 20, 5, 7).cpu().numpy(), orc.spmm_sampled(t, X, 20, 5, 7)) < FP32_TOL
    assert rel_err(ops.sddmm(g, dev(X), dev(X)).cpu().numpy(), orc.sddmm(t, X, X)) < FP32_TOL


def test_results_are_bit_reproducible(orc):
    t = graph_case(orc, 3000, 400000, 2, 700)
    g = to_gpu_graph(t, 256)
    X = dev(np.random.default_rng(2).uniform(-0.5, 0.5, (3000, 32)).astype(np.float32))
    a = ops.spmm(g, X, vals=dev(t.vals))
    for _ in range(3):
        assert torch.equal(a, ops.spmm(g, X, vals=dev(t.vals)))


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("K", [8, 32, 64, 256])
def test_bf16_feature_storage_matches_fp32_path_on_rounded_features(orc, case, K):
    """Optional bf16 feature rows: the kernels must equal the fp32 oracle evaluated on the bf16-ROUNDED features
    (1e-5, same bar as the fp32 path) and stay within the north-star bound for bf16 features (1e-2) of the
    un-rounded fp32 result."""
    n, e, seed, T, empty, thr = case
    t = graph_case(orc, n, e, seed, T, empty)
    g = to_gpu_graph(t, thr)
    rng = np.random.default_rng(seed + K)
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    Xb = dev(X).to(torch.bfloat16)
    Xr = Xb.float().cpu().numpy()
    for weighted in (False, True):
        got = ops.spmm_bf16(g, Xb, vals=dev(t.vals) if weighted else None).cpu().numpy()
        assert rel_err(got, orc.spmm(t, Xr, weighted=weighted)) < FP32_TOL
        assert rel_err(got, orc.spmm(t, X, weighted=weighted)) < 1e-2
    aL = rng.normal(size=n).astype(np.float32)
    aR = rng.normal(size=n).astype(np.float32)
    alpha = torch.empty(t.nvals, device=DEV)
    got = ops.gat_forward_bf16(g, dev(aL), dev(aR), Xb, 0.2, relu=True, alpha_out=alpha).cpu().numpy()
    want, want_alpha = orc.gat_forward(t, aL, aR, Xr, 0.2)
    assert rel_err(got, np.maximum(want, 0)) < FP32_TOL
    assert rel_err(alpha.cpu().numpy(), want_alpha) < FP32_TOL
    assert rel_err(got, np.maximum(orc.gat_forward(t, aL, aR, X, 0.2)[0], 0)) < 1e-2
    # epilogue options and accumulate on the bf16 path
    rs = rng.uniform(0.5, 1.5, n).astype(np.float32)
    Y0 = rng.uniform(-1, 1, (n, K)).astype(np.float32)
    out = dev(Y0.copy())
    ops.spmm_bf16(g, Xb, out=out, accumulate=True)
    assert rel_err(out.cpu().numpy(), Y0 + orc.spmm(t, Xr, weighted=False)) < FP32_TOL
    got = ops.spmm_bf16(g, Xb, row_scale=dev(rs), relu=True).cpu().numpy()
    assert rel_err(got, np.maximum(orc.spmm(t, Xr, weighted=False) * rs[:, None], 0)) < FP32_TOL


def test_bf16_feature_storage_rejects_unsupported_widths():
    from gala_b200 import lib as _l
    offset, ids = make_csr(64, 400, 1)
    g = ops.TiledGraph(dev(offset), dev(ids), 64)
    for K in (4, 24, 41, 512):
        with pytest.raises(_l.GalaError):
            ops.spmm_bf16(g, torch.zeros(64, K, device=DEV, dtype=torch.bfloat16))


@pytest.mark.parametrize("dense", ["tcgen05", "torch"])
def test_cora_gcn_forward_matches_reference_cpu_golden(dense):
    """BASELINE.json configs[0] on the GPU: the generated 2-layer GCN forward (fused epilogues and op by op)
    against the logits the reference's CPU path produced for the same graph, features and weights."""
    from gala_b200 import formats
    from gala_b200.gcn_model import GCN2

    gd = golden("cora_gcn")
    n, feats = int(gd["n"]), int(gd["feats"])
    X = dev(np.random.default_rng(int(gd["x_seed"])).uniform(-0.5, 0.5, (n, feats)).astype(np.float32))
    ones = torch.ones(gd["ids"].shape[0], device=DEV)
    g = formats.ord_col_tiling(n, n, dev(gd["offset"]), dev(gd["ids"]), ones, 100000).build_plan(64)   # shipped schedule
    assert g.segments == 1
    model = GCN2(feats, 32, 7, DEV)
    model.fc0, model.fc1 = (dev(gd["W0"]), dev(gd["b0"])), (dev(gd["W1"]), dev(gd["b1"]))
    model.prepare(g)
    assert rel_err(model.norm.cpu().numpy(), gd["norm"]) < 1e-6
    assert rel_err(model.forward(g, X, dense=dense).cpu().numpy(), gd["logits"]) < FP32_TOL
    assert rel_err(model.forward_literal(g, X).cpu().numpy(), gd["logits"]) < FP32_TOL
