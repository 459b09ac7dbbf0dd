"""CPU suite: the synthetic-graph generators behind bench.py and the partitioned runs.  Every rank of a
multi-GPU run synthesises its own copy, so they must be bit-reproducible functions of (shape, seed)."""
import numpy as np
import torch

from gala_b200 import synth


def test_numpy_and_torch_power_law_graphs_have_the_reference_data_pipeline_properties():
    n, e = 3000, 90000
    src, dst = synth.powerlaw_coo_np(n, e, seed=5)
    offset, ids = synth.coo_to_csr_np(n, src, dst)
    assert offset[-1] == ids.shape[0] and abs(int(offset[-1]) - e) <= 1
    rows = np.repeat(np.arange(n), np.diff(offset))
    key = rows.astype(np.int64) * n + ids
    assert np.all(np.diff(key) > 0)                                   # sorted by (row, col), duplicate-free
    assert np.array_equal(np.sort(key), np.sort(ids.astype(np.int64) * n + rows))   # symmetric (set_undirected)
    has_loop = np.zeros(n, bool)
    has_loop[rows[rows == ids]] = True
    assert has_loop.all()                                             # add_self_loop (gala_export_npy.py:73-74)
    t_off, t_ids = synth.powerlaw_csr_torch(n, e, seed=5, device="cpu")
    t_off2, t_ids2 = synth.powerlaw_csr_torch(n, e, seed=5, device="cpu")
    assert torch.equal(t_off, t_off2) and torch.equal(t_ids, t_ids2)  # deterministic in the seed
    assert int(t_off[-1]) == t_ids.numel() and (t_off[1:] - t_off[:-1]).max() > 20 * (e // n)   # power-law hubs


def test_papers_scale_generator_is_reproducible_and_every_node_has_a_self_loop():
    n, e = 20000, 600000
    r1, c1 = synth.powerlaw_multigraph_coo_torch(n, e, seed=3, device="cpu", chunk=100000)
    r2, c2 = synth.powerlaw_multigraph_coo_torch(n, e, seed=3, device="cpu", chunk=100000)
    assert torch.equal(r1, r2) and torch.equal(c1, c2)
    assert r1.numel() == e and int(r1.min()) >= 0 and int(max(r1.max(), c1.max())) < n
    deg = torch.bincount(r1.long(), minlength=n)
    assert int(deg.min()) >= 1                                         # self loop per node: GCN norms stay finite
    assert torch.equal(r1[-n:], torch.arange(n, dtype=torch.int32)) and torch.equal(r1[-n:], c1[-n:])
    assert int(deg.max()) > 10 * e // n                                # heavy tail
    r3, _ = synth.powerlaw_multigraph_coo_torch(n, e, seed=4, device="cpu", chunk=100000)
    assert not torch.equal(r1, r3)
