"""GPU parity (-m gpu): device-side format construction vs the oracle / the reference's own
golden outputs.  Integer outputs must be bit-exact (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from gala_b200 import formats, ops, synth
from util import GOLDEN_CASES, golden, make_csr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def eq(t, a):
    return np.array_equal(t.cpu().numpy(), a)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_formats_on_gpu(name):
    """The reference's own CSR / transpose / tiling / sampling / sub-graph outputs."""
    g = golden(name)
    n = int(g["n"])
    off, ids, _ = formats.csr_build(n, n, dev(g["coo_src"]), dev(g["coo_dst"]))
    assert eq(off, g["offset"]) and eq(ids, g["ids"])
    to, ti, _ = formats.buildTranspose(n, n, off, ids)
    assert eq(to, g["t_offset"]) and eq(ti, g["t_ids"])
    tg = formats.ord_col_tiling(n, n, off, ids, dev(g["w"]), int(g["T"]))
    assert eq(tg.offsets, g["tile_offsets"]) and eq(tg.cols, g["tile_cols"])
    assert np.array_equal(tg.bounds, g["tile_bounds"]) and eq(tg.vals, g["tile_vals"])
    so, si, sv, status = formats.inplace_sample_graph_ab(n, off, ids, dev(g["w"]), 20, 5, 7)
    assert eq(so, g["s_offset"]) and eq(si, g["s_ids"]) and eq(sv, g["s_vals"]) and int(status) == 0
    ones = torch.ones(ids.numel(), device=DEV)
    (fo, fi, fv, bo, bi, bv), = formats.getMaskSubgraphs(n, n, off, ids, ones, dev(g["mask"]), 1)
    assert eq(fo, g["m_fwd_offset"]) and eq(fi, g["m_fwd_ids"])
    assert eq(bo, g["m_bwd_offset"]) and eq(bi, g["m_bwd_ids"])


@pytest.mark.parametrize("n,e,seed", [(1, 1, 0), (97, 400, 1), (5000, 300000, 2), (70000, 700000, 3)])
def test_csr_build_and_transpose_match_oracle(orc, n, e, seed):
    offset0, ids0 = make_csr(n, e, seed, empty_rows=min(n // 10, 50))
    rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(offset0))
    rng = np.random.default_rng(seed)
    dup = rng.integers(0, ids0.shape[0], ids0.shape[0] // 20)      # duplicates are kept
    r = np.concatenate([rows, rows[dup]])
    c = np.concatenate([ids0, ids0[dup]])
    p = rng.permutation(r.shape[0])
    r, c = r[p], c[p]
    v = rng.integers(1, 100, r.shape[0]).astype(np.float32)
    oo, oi, ov = orc.csr_build(n, r, c, v)
    off, ids, vals = formats.csr_build(n, n, dev(r), dev(c), dev(v))
    assert eq(off, oo) and eq(ids, oi)
    # values travel with their edge; among duplicate (row,col) pairs only the multiset is defined
    key = oi.astype(np.int64) + np.repeat(np.arange(n, dtype=np.int64), np.diff(oo)) * n
    got = vals.cpu().numpy()
    order_a = np.lexsort((ov, key))
    order_b = np.lexsort((got, key))
    assert np.array_equal(ov[order_a], got[order_b])
    to, ti, tv = orc.csr_transpose(n, n, oo, oi, np.ones_like(ov))
    gto, gti, _ = formats.buildTranspose(n, n, off, ids)
    assert eq(gto, to) and eq(gti, ti)
    # transpose of a symmetric duplicate-free graph is itself
    offset_s, ids_s = make_csr(n, e, seed)
    o2, i2, _ = formats.buildTranspose(n, n, dev(offset_s), dev(ids_s))
    assert eq(o2, offset_s) and eq(i2, ids_s)


@pytest.mark.parametrize("T", [1, 7, 64, 1000, 100000])
def test_col_tiling_matches_oracle(orc, T):
    n = 1000 if T > 1 else 40
    offset, ids = make_csr(n, 20 * n, 7, empty_rows=5)
    w = np.random.default_rng(0).uniform(-1, 1, ids.shape[0]).astype(np.float32)
    t = orc.col_tile(n, n, offset, ids, w, T)
    g = formats.ord_col_tiling(n, n, dev(offset), dev(ids), dev(w), T)
    assert g.segments == t.S
    assert eq(g.offsets, t.offsets) and eq(g.cols, t.cols) and eq(g.vals, t.vals)
    assert np.array_equal(g.bounds, t.bounds)
    # and the tiled graph feeds the kernels: tiled SpMM == untiled SpMM
    if t.S <= 64:
        X = torch.rand(n, 16, device=DEV) - 0.5
        g1 = ops.TiledGraph(dev(offset), dev(ids), n)
        a, b = ops.spmm(g, X, vals=g.vals), ops.spmm(g1, X, vals=dev(w))
        assert float((a - b).abs().max()) < 1e-5


def test_sample_ab_matches_oracle_and_flags_empty_rows(orc):
    n = 3000
    offset, ids = make_csr(n, 60000, 4)
    w = np.random.default_rng(1).uniform(0, 1, ids.shape[0]).astype(np.float32)
    for (s, ra, rb) in ((20, 5, 7), (1, 0, 0), (128, 97, 3)):
        rc, so, si, sv = orc.sample_ab(n, offset, ids, w, s, ra, rb)
        go, gi, gv, status = formats.inplace_sample_graph_ab(n, dev(offset), dev(ids), dev(w), s, ra, rb)
        assert rc == 0 and int(status) == 0
        assert eq(go, so) and eq(gi, si) and eq(gv, sv)
    offset2, ids2 = make_csr(n, 60000, 4, empty_rows=3)
    *_, status = formats.inplace_sample_graph_ab(n, dev(offset2), dev(ids2), torch.ones(ids2.shape[0], device=DEV), 20)
    assert int(status) == 1      # `% 0` in the reference (tiling.h:482)


def test_mask_subgraphs_two_layers_match_oracle(orc):
    n = 4000
    offset, ids = make_csr(n, 40000, 5)
    ones = np.ones(ids.shape[0], np.float32)
    mask = (np.random.default_rng(2).random(n) < 0.05).astype(np.uint8)
    want = orc.mask_subgraphs(n, n, offset, ids, ones, mask, 2)
    got = formats.getMaskSubgraphs(n, n, dev(offset), dev(ids), dev(ones), dev(mask), 2)
    for w_, g_ in zip(want, got):
        for a, b in zip(w_, g_):
            assert eq(b, a)
    assert got[1][1].numel() > got[0][1].numel()     # the mask grows by one hop


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_reorder_on_gpu(name):
    """The reference's own rowReorderToAdj / rowPermuteDense{To,From} outputs."""
    g = golden(name)
    n = int(g["n"])
    perm = dev(g["perm"])
    ro, ri, rv = formats.rowReorderToAdj(n, dev(g["offset"]), dev(g["ids"]), dev(g["w"]), perm)
    assert eq(ro, g["r_offset"]) and eq(ri, g["r_ids"]) and eq(rv, g["r_vals"])
    assert eq(formats.rowPermuteDenseTo(dev(g["X"]), perm), g["X_to"])
    assert eq(formats.rowPermuteDenseFrom(dev(g["X"]), perm), g["X_from"])


@pytest.mark.parametrize("n,e,K", [(1, 1, 3), (300, 5000, 7), (20000, 600000, 32)])
def test_reorder_matches_oracle_and_commutes_with_spmm(orc, n, e, K):
    offset, ids = make_csr(n, e, 9, empty_rows=min(n // 10, 30))
    rng = np.random.default_rng(n)
    w = rng.uniform(-1, 1, ids.shape[0]).astype(np.float32)
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    perm, order = formats.degree_order(n, dev(offset))
    operm, oorder = orc.degree_order(n, offset)
    assert eq(perm, operm) and eq(order, oorder)
    want = orc.csr_reorder(n, offset, ids, w, operm)
    ro, ri, rv = formats.rowReorderToAdj(n, dev(offset), dev(ids), dev(w), perm)
    assert eq(ro, want[0]) and eq(ri, want[1]) and eq(rv, want[2])
    assert np.all(np.diff(np.diff(want[0])) <= 0)                       # degrees now descend
    Xp = formats.rowPermuteDenseTo(dev(X), perm)
    assert eq(Xp, orc.permute_rows(X, operm, False))
    assert eq(formats.rowPermuteDenseFrom(Xp, perm), X)                 # from undoes to
    # P A P^T (P X) = P (A X): aggregation on the relabelled graph is the relabelled aggregation
    y = ops.spmm(ops.TiledGraph(dev(offset), dev(ids), n).build_plan(), dev(X), vals=dev(w))
    yp = ops.spmm(ops.TiledGraph(ro, ri, n).build_plan(), Xp, vals=rv)
    back = formats.rowPermuteDenseFrom(yp, perm)
    assert float((back - y).abs().max()) <= 1e-5 * max(1.0, float(y.abs().max()))


def test_csr_build_at_reddit_scale_properties():
    """Full BASELINE size: sortedness, row-pointer consistency, idempotence, involution."""
    n, e, *_ = synth.SHAPES["reddit"]
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=DEV)
    E = int(ids.numel())
    rows = torch.repeat_interleave(torch.arange(n, device=DEV, dtype=torch.int32), (offset[1:] - offset[:-1]).long())
    perm = torch.randperm(E, device=DEV)
    off2, ids2, _ = formats.csr_build(n, n, rows[perm].contiguous(), ids[perm].contiguous())
    assert torch.equal(off2, offset) and torch.equal(ids2, ids)          # shuffled COO -> the same CSR
    t_off, t_ids, _ = formats.buildTranspose(n, n, off2, ids2)
    assert torch.equal(t_off, offset) and torch.equal(t_ids, ids)        # symmetric graph
    tg = formats.ord_col_tiling(n, n, offset, ids, torch.ones(E, device=DEV), 37000)   # shipped Reddit schedule
    assert tg.segments == 7 and int(tg.bounds[-1]) == E
    assert int(tg.offsets.view(7, n + 1)[:, -1].sum()) == E
    X = torch.rand(n, 32, device=DEV) - 0.5
    a = ops.spmm(tg.build_plan(), X)
    b = ops.spmm(ops.TiledGraph(offset, ids, n).build_plan(), X)
    assert float((a - b).double().norm() / b.double().norm()) < 1e-6


def test_npy_ingest_builds_the_reference_csr_on_the_device(orc, tmp_path):
    """readSM_npy32 / readDM_npy (tests/common.h:331-389) replaced by mmap -> pinned -> device -> GPU build."""
    from gala_b200 import ingest
    n, f, c = 3000, 19, 5
    offset, ids = make_csr(n, 50000, 4)
    rows = np.repeat(np.arange(n, dtype=np.uint32), np.diff(offset))
    rng = np.random.default_rng(0)
    p = rng.permutation(rows.shape[0])                       # the file need not be sorted
    np.save(tmp_path / "Adj_src.npy", np.concatenate([np.array([n, n], np.uint32), rows[p]]))
    np.save(tmp_path / "Adj_dst.npy", ids[p].astype(np.uint32))
    feat = rng.uniform(-0.5, 0.5, (n, f)).astype(np.float32)
    lab = rng.integers(0, c, (n, 1)).astype(np.int64)
    lab[:c, 0] = np.arange(c)
    np.save(tmp_path / "Feat.npy", feat)
    np.save(tmp_path / "Lab.npy", lab)
    for name in ("TnMsk", "VlMsk", "TsMsk"):
        np.save(tmp_path / (name + ".npy"), (rng.random((n, 1)) < 0.3).astype(np.int32))
    old = ingest.CHUNK_BYTES
    ingest.CHUNK_BYTES = 4096 * 4                            # force several staging rounds
    try:
        d = ingest.load_dataset(str(tmp_path) + "/", DEV)
    finally:
        ingest.CHUNK_BYTES = old
    oo, oi, _ = orc.csr_build(n, rows[p].astype(np.int32), ids[p].astype(np.int32))
    assert d["nrows"] == n and eq(d["offsets"], oo) and eq(d["ids"], oi) and bool((d["vals"] == 1).all())
    assert eq(d["input_emb"], feat) and eq(d["labels"], lab) and d["classes"] == c
    assert d["train_mask"].dtype == torch.bool and d["train_mask"].shape == (n, 1)
