"""Generates tests/golden/*.npz by running the REFERENCE's own code (oracle/_ref, built
from /root/reference by oracle/Makefile) on small seeded inputs.  Run in the authoring
container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Reference entry points exercised (file:line relative to /root/reference):
  CSRCMatrix::build src/formats/csrc_matrix.h:148-376, buildTranspose tests/common.h:107-123,
  gSpMM<wsumAgg> src/ops/aggregators.h:55-127, static_ord_col_breakpoints +
  ord_col_tiling_torch src/ops/tiling.h:1594-1608,222-283, inplace_sample_graph_ab
  src/ops/tiling.h:454-508, getMaskSubgraphs tests/common.h:20-105 (1 layer), rowReorderToAdj / rowPermuteDenseTo /
  rowPermuteDenseFrom src/ops/reordering.h:940-1013,244-283,207-236.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
from oracle import orc  # noqa: E402
from gala_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(name, n, e, seed, K, T, dup=False):
    rng = np.random.default_rng(seed)
    src, dst = synth.powerlaw_coo_np(n, e, seed=seed)
    if dup:  # duplicates are kept by build() (has_dup ignored, csrc_matrix.h:149-150)
        extra = rng.integers(0, src.shape[0], size=src.shape[0] // 10)
        src = np.concatenate([src, src[extra]])
        dst = np.concatenate([dst, dst[extra]])
    p = rng.permutation(src.shape[0])  # COO arrives unsorted
    src, dst = src[p].astype(np.int32), dst[p].astype(np.int32)
    offset, ids, vals = orc.ref_csr_build(n, n, src, dst)           # vals all ones
    t_off, t_ids, t_vals = orc.ref_csr_transpose(n, n, offset, ids, vals)
    w = rng.uniform(-1, 1, ids.shape[0]).astype(np.float32)
    X = rng.uniform(-0.5, 0.5, (n, K)).astype(np.float32)
    Y_w = orc.ref_gspmm_wsum(n, n, offset, ids, w, X)
    Y_1 = orc.ref_gspmm_wsum(n, n, offset, ids, vals, X)
    bp, tiled = orc.ref_col_tile(n, n, offset, ids, w, T)
    s_off, s_ids, s_vals = orc.ref_sample_ab(n, n, offset, ids, w, 20, 5, 7)
    mask = (rng.random(n) < 0.1).astype(np.uint8)
    (m_fo, m_fi, m_fv, m_bo, m_bi, m_bv), = orc.ref_mask_subgraphs(n, n, offset, ids, vals, mask, 1)
    perm = rng.permutation(n).astype(np.int32)                      # drawn last: earlier vectors unchanged
    r_off, r_ids, r_vals = orc.ref_row_reorder_to_adj(n, offset, ids, w, perm)
    X_to = orc.ref_row_permute_dense(X, perm, False)
    X_from = orc.ref_row_permute_dense(X, perm, True)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), n=n, K=K, T=T, coo_src=src, coo_dst=dst, offset=offset,
        ids=ids, t_offset=t_off, t_ids=t_ids, w=w, X=X, Y_w=Y_w, Y_1=Y_1, breakpoints=bp,
        tile_offsets=tiled.offsets, tile_cols=tiled.cols, tile_vals=tiled.vals,
        tile_bounds=tiled.bounds, s_offset=s_off, s_ids=s_ids, s_vals=s_vals, mask=mask,
        m_fwd_offset=m_fo, m_fwd_ids=m_fi, m_bwd_offset=m_bo, m_bwd_ids=m_bi,
        perm=perm, r_offset=r_off, r_ids=r_ids, r_vals=r_vals, X_to=X_to, X_from=X_from)
    print(name, "n", n, "E", ids.shape[0], "segments", tiled.S)


def cora_gcn():
    """BASELINE.json configs[0]: the 2-layer GCN forward of the generated program (codegen/gala.cu:422-459) on a
    Cora-shape graph through the reference's CPU path -- CSRCMatrix::build, gSpMM<wsumAgg> (oracle/_ref) and
    torch-CPU Linear (what nn::Linear is).  The features are regenerated from the seed by the tests; graph,
    weights and logits are stored."""
    import torch
    import torch.nn.functional as F

    n, e, feats, hidden, classes = synth.SHAPES["cora"]
    src, dst = synth.powerlaw_coo_np(n, e, seed=21)
    offset, ids, ones = orc.ref_csr_build(n, n, src.astype(np.int32), dst.astype(np.int32))
    rng = np.random.default_rng(22)
    X = rng.uniform(-0.5, 0.5, (n, feats)).astype(np.float32)

    def lin(o, i):
        b = 1.0 / np.sqrt(i)
        return rng.uniform(-b, b, (o, i)).astype(np.float32), rng.uniform(-b, b, o).astype(np.float32)
    W0, b0 = lin(hidden, feats)
    W1, b1 = lin(classes, hidden)

    def agg(x):
        return orc.ref_gspmm_wsum(n, n, offset, ids, ones, np.ascontiguousarray(x, np.float32))
    deg = agg(np.ones((n, 1), np.float32))                       # degrees via A @ 1 (codegen/gala.cu:437)
    norm = torch.pow(torch.from_numpy(deg), -0.5).numpy()
    res = F.linear(torch.from_numpy(X), torch.from_numpy(W0), torch.from_numpy(b0)).numpy()
    res = norm * res
    res = agg(res)
    res = np.maximum(norm * res, 0)
    res = norm * res
    res = agg(res)
    res = norm * res
    logits = F.linear(torch.from_numpy(res), torch.from_numpy(W1), torch.from_numpy(b1)).numpy()
    np.savez_compressed(os.path.join(OUT, "cora_gcn.npz"), n=n, feats=feats, x_seed=22, offset=offset, ids=ids,
                        W0=W0, b0=b0, W1=W1, b1=b1, norm=norm.ravel(), logits=logits)
    print("cora_gcn n", n, "E", ids.shape[0], "|logits|", float(np.abs(logits).sum()))


if __name__ == "__main__":
    assert orc.have_ref(), "build oracle/_ref first: make -C oracle ref"
    case("small_a", 257, 3000, 11, 8, 64)
    case("small_dup", 193, 2200, 12, 5, 50, dup=True)
    case("small_hub", 600, 30000, 13, 32, 100000)
    cora_gcn()
