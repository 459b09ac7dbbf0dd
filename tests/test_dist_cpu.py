"""CPU suite: the multi-GPU layer's host logic (nnz-balanced 1-D row partition, column remap
into the padded all-gather layout, per-layer feature exchange) with the gloo backend,
world_size 2 -- each rank runs the oracle on its slab, the assembled result must equal the
single-process oracle forward."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gala_b200 import dist_gat, synth
from gala_b200.gat_model import GAT2
from util import rel_err


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_aggregate(orc, part):
    t = orc.Tiled.from_csr(part.rows, part.padded_n, part.offset.numpy(), part.cols.numpy())

    def agg(aL, aR, feats, relu):
        y, _ = orc.gat_forward(t, aL.numpy(), aR.numpy(), feats.numpy())
        y = torch.from_numpy(y)
        return torch.relu(y) if relu else y
    return agg


def _oracle_aggregate_col(orc, part):
    """What gala_gat_forward_col_f32 computes on the slab, through the oracle's layer: the right-hand term from the last
    column of the gathered rows, the finished rows reflected back / ReLU'd / reflected on."""
    t = orc.Tiled.from_csr(part.rows, part.padded_n, part.offset.numpy(), part.cols.numpy())

    def refl(y, v):
        return y - 2.0 * (y @ v)[:, None] * v[None, :]

    def agg(aL, sR, bR, feats, relu, v_in, v_out):
        f = feats.numpy()
        aR = (sR * f[:, -1].astype(np.float64) + bR).astype(np.float32)
        y, _ = orc.gat_forward(t, aL.numpy(), aR, f)
        y = torch.from_numpy(y).double()
        if v_in is not None:
            y = refl(y, v_in.double())
        if relu:
            y = torch.relu(y)
        if v_out is not None:
            y = refl(y, v_out.double())
        return y.float()
    return agg


def rank_mode(rank):
    return "folded"     # both ranks must take the same collective sequence


def _worker(rank, world, port, n, e, feats, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc

    torch.set_num_threads(2)
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=3, device="cpu")
    model = GAT2(feats, 8, 5, "cpu", seed=1)
    X = torch.rand(n, feats, generator=torch.Generator().manual_seed(2)) - 0.5
    part = dist_gat.RowPartition(offset, ids, n, rank, world)
    fwd = dist_gat.gat2_forward_partitioned if rank_mode(rank) == 'literal' else dist_gat.gat2_forward_partitioned_folded
    out_loc = fwd(model, part, X[part.row_lo:part.row_hi], _oracle_aggregate(orc, part))
    gathered = part.unpad(part.all_gather(out_loc))
    # rows exchanged in the reflected basis (hidden width 8: one of the widths the column-mode kernel takes)
    out_refl = dist_gat.gat2_forward_partitioned_reflected(model, part, X[part.row_lo:part.row_hi],
                                                           _oracle_aggregate_col(orc, part))
    gathered_refl = part.unpad(part.all_gather(out_refl))
    if rank == 0:
        np.save(out_path, gathered.numpy())
        np.save(out_path + ".reflected.npy", gathered_refl.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_partition_boundaries_balance_nnz():
    offset, ids = synth.powerlaw_csr_torch(3000, 90000, seed=0, device="cpu")
    for world in (1, 2, 4, 8):
        b = dist_gat.partition_rows_by_nnz(offset, world)
        assert b[0] == 0 and b[-1] == 3000 and len(b) == world + 1
        assert all(b[i] <= b[i + 1] for i in range(world))
        nnz = [int(offset[b[i + 1]] - offset[b[i]]) for i in range(world)]
        assert max(nnz) - min(nnz) <= 2 * int((offset[1:] - offset[:-1]).max())


def test_column_remap_roundtrip():
    n = 500
    offset, ids = synth.powerlaw_csr_torch(n, 6000, seed=1, device="cpu")
    X = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3)
    for world in (2, 3):
        parts = [dist_gat.RowPartition(offset, ids, n, r, world) for r in range(world)]
        padded = torch.cat([p.pad(X[p.row_lo:p.row_hi]) for p in parts], 0)   # what all_gather produces
        for p in parts:
            e_lo = int(offset[p.row_lo])
            want = ids[e_lo:e_lo + p.local_nvals].to(torch.float32)
            assert torch.equal(padded[p.cols.long(), 0], want)                 # remapped col -> same node
            assert torch.equal(p.unpad(padded), X)


def test_two_rank_gloo_forward_matches_single_process(orc, tmp_path):
    n, e, feats = 600, 9000, 12
    out_path = str(tmp_path / "dist_out.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, e, feats, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=3, device="cpu")
    model = GAT2(feats, 8, 5, "cpu", seed=1)
    X = torch.rand(n, feats, generator=torch.Generator().manual_seed(2)) - 0.5
    t = orc.Tiled.from_csr(n, n, offset.numpy(), ids.numpy())

    def agg(aL, aR, f, relu):
        y = torch.from_numpy(orc.gat_forward(t, aL.numpy(), aR.numpy(), f.numpy())[0])
        return torch.relu(y) if relu else y

    import torch.nn.functional as F
    res = F.linear(X, *model.fc0)
    y = agg(F.linear(res, *model.efc0).reshape(-1), F.linear(res, *model.efc1).reshape(-1), res, True)
    tt = F.linear(y, *model.fc1)
    a2 = agg(F.linear(tt, *model.efc2).reshape(-1), F.linear(tt, *model.efc3).reshape(-1), y, False)
    want = F.linear(a2, *model.fc1).numpy()
    assert rel_err(got, want) < 1e-5
    assert rel_err(np.load(out_path + ".reflected.npy"), want) < 1e-5


def _worker_deep(rank, world, port, n, e, dims, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc
    from gala_b200.gat_model import GATN

    torch.set_num_threads(2)
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=5, device="cpu")
    model = GATN(dims, "cpu", seed=4)
    X = torch.rand(n, dims[0], generator=torch.Generator().manual_seed(6)) - 0.5
    part = dist_gat.RowPartition(offset, ids, n, rank, world)
    out_loc = dist_gat.gatn_forward_partitioned(model, part, X[part.row_lo:part.row_hi], _oracle_aggregate(orc, part))
    gathered = part.unpad(part.all_gather(out_loc))
    if rank == 0:
        np.save(out_path, gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_three_layer_gat_two_ranks_matches_dense_reference(orc, tmp_path):
    """BASELINE.json configs[4] in miniature: 3-layer GAT, 1-D row partition, one exchange per layer --
    against a dense fp64 evaluation of the same model."""
    import torch.nn.functional as F
    from gala_b200.gat_model import GATN

    n, e, dims = 400, 5000, [10, 8, 8, 6]
    out_path = str(tmp_path / "deep_out.npy")
    mp.spawn(_worker_deep, args=(2, _free_port(), n, e, dims, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=5, device="cpu")
    model = GATN(dims, "cpu", seed=4)
    X = (torch.rand(n, dims[0], generator=torch.Generator().manual_seed(6)) - 0.5).double()
    A = torch.zeros(n, n, dtype=torch.bool)
    rows = torch.repeat_interleave(torch.arange(n), (offset[1:] - offset[:-1]).long())
    A[rows, ids.long()] = True

    def lin(x, wb):
        return F.linear(x, wb[0].double(), wb[1].double())

    def attend(aL, aR, feats):
        s = F.leaky_relu(aL[:, None] + aR[None, :], 0.2).exp().masked_fill(~A, 0.0)
        return (s / s.sum(1, keepdim=True)) @ feats

    res = X
    for i in range(model.L - 1):
        t = lin(res, model.fc[i])
        res = torch.relu(attend(lin(t, model.efcL[i]).reshape(-1), lin(t, model.efcR[i]).reshape(-1), t))
    t = lin(res, model.fc[-1])
    want = lin(attend(lin(t, model.efcL[-1]).reshape(-1), lin(t, model.efcR[-1]).reshape(-1), res), model.fc[-1])
    assert rel_err(got, want.float().numpy()) < 1e-5


def _worker_gcn(rank, world, port, n, e, dims, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc
    from gala_b200.gcn_model import GCNN

    torch.set_num_threads(2)
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=7, device="cpu")
    model = GCNN(dims, "cpu", seed=8)
    X = torch.rand(n, dims[0], generator=torch.Generator().manual_seed(9)) - 0.5
    part = dist_gat.RowPartition(offset, ids, n, rank, world)
    t = orc.Tiled.from_csr(part.rows, part.padded_n, part.offset.numpy(), part.cols.numpy())

    def aggregate(feats_all, row_scale, relu):
        y = torch.from_numpy(orc.spmm(t, feats_all.numpy(), weighted=False)) * row_scale[:, None]
        return torch.relu(y) if relu else y
    deg = torch.from_numpy(orc.spmm(t, np.ones((part.padded_n, 1), np.float32), weighted=False)).reshape(-1)
    out_loc = dist_gat.gcnn_forward_partitioned(model, part, X[part.row_lo:part.row_hi], torch.pow(deg, -0.5), aggregate)
    gathered = part.unpad(part.all_gather(out_loc))
    if rank == 0:
        np.save(out_path, gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_three_layer_gcn_two_ranks_matches_dense_reference(tmp_path):
    """3-layer GCN, 1-D row partition (BASELINE.json configs[4] in miniature) against dense fp64 math."""
    import torch.nn.functional as F
    from gala_b200.gcn_model import GCNN

    n, e, dims = 400, 5000, [10, 8, 8, 6]
    out_path = str(tmp_path / "gcn_out.npy")
    mp.spawn(_worker_gcn, args=(2, _free_port(), n, e, dims, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=7, device="cpu")
    model = GCNN(dims, "cpu", seed=8)
    X = (torch.rand(n, dims[0], generator=torch.Generator().manual_seed(9)) - 0.5).double()
    A = torch.zeros(n, n, dtype=torch.float64)
    rows = torch.repeat_interleave(torch.arange(n), (offset[1:] - offset[:-1]).long())
    A[rows, ids.long()] = 1.0
    nrm = A.sum(1).pow(-0.5)[:, None]
    res = X
    for i in range(model.L - 1):
        res = torch.relu(nrm * (A @ (nrm * F.linear(res, model.fc[i][0].double(), model.fc[i][1].double()))))
    want = F.linear(nrm * (A @ (nrm * res)), model.fc[-1][0].double(), model.fc[-1][1].double())
    assert rel_err(got, want.float().numpy()) < 1e-5


def _worker_needed(rank, world, port, n, e, dims, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc
    from gala_b200.gat_model import GATN
    from gala_b200.gcn_model import GCNN

    torch.set_num_threads(2)
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=5, device="cpu")
    X = torch.rand(n, dims[0], generator=torch.Generator().manual_seed(6)) - 0.5
    part = dist_gat.NeededRowsPartition(offset, ids, n, rank, world)
    full = dist_gat.RowPartition(offset, ids, n, rank, world)
    assert part.rows == full.rows and part.padded_n <= n and part.n_remote <= n - part.rows
    # the remapped columns address the same nodes
    probe = torch.arange(n, dtype=torch.float32)[:, None]
    got = part.all_gather(probe[part.row_lo:part.row_hi])[part.cols.long(), 0]
    e_lo = int(offset[part.row_lo])
    assert torch.equal(got, ids[e_lo:e_lo + part.local_nvals].float())
    t = orc.Tiled.from_csr(part.rows, part.padded_n, part.offset.numpy(), part.cols.numpy())

    def agg_gat(aL, aR, feats, relu):
        y = torch.from_numpy(orc.gat_forward(t, aL.numpy(), aR.numpy(), feats.numpy())[0])
        return torch.relu(y) if relu else y

    def agg_gcn(feats_all, row_scale, relu):
        y = torch.from_numpy(orc.spmm(t, feats_all.numpy(), weighted=False)) * row_scale[:, None]
        return torch.relu(y) if relu else y
    Xl = X[part.row_lo:part.row_hi]
    out_gat = dist_gat.gatn_forward_partitioned(GATN(dims, "cpu", seed=4), part, Xl, agg_gat)
    deg = torch.from_numpy(orc.spmm(t, np.ones((part.padded_n, 1), np.float32), weighted=False)).reshape(-1)
    out_gcn = dist_gat.gcnn_forward_partitioned(GCNN(dims, "cpu", seed=8), part, Xl, torch.pow(deg, -0.5), agg_gcn)
    g1, g2 = part.gather_full(out_gat), part.gather_full(out_gcn)
    if rank == 0:
        np.save(out_path, torch.stack([g1, g2]).numpy())
        np.save(out_path + ".frac.npy", np.array([part.exchange_fraction()]))
    dist.barrier()
    dist.destroy_process_group()


def test_needed_rows_exchange_matches_all_gather_results(orc, tmp_path):
    """The all-to-all-v exchange of only the referenced remote rows gives the same 3-layer GAT / GCN outputs
    as the full all-gather (compared with the dense fp64 evaluations used above)."""
    import torch.nn.functional as F
    from gala_b200.gat_model import GATN
    from gala_b200.gcn_model import GCNN

    n, e, dims = 400, 2400, [10, 8, 8, 6]           # sparse enough that a slab does not reference every node
    out_path = str(tmp_path / "needed.npy")
    mp.spawn(_worker_needed, args=(2, _free_port(), n, e, dims, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    assert float(np.load(out_path + ".frac.npy")[0]) < 1.0
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=5, device="cpu")
    X = (torch.rand(n, dims[0], generator=torch.Generator().manual_seed(6)) - 0.5).double()
    rows = torch.repeat_interleave(torch.arange(n), (offset[1:] - offset[:-1]).long())
    A = torch.zeros(n, n, dtype=torch.float64)
    A[rows, ids.long()] = 1.0
    mask = A > 0

    def lin(x, wb):
        return F.linear(x, wb[0].double(), wb[1].double())

    def attend(aL, aR, feats):
        s = F.leaky_relu(aL[:, None] + aR[None, :], 0.2).exp().masked_fill(~mask, 0.0)
        return (s / s.sum(1, keepdim=True)) @ feats
    m = GATN(dims, "cpu", seed=4)
    res = X
    for i in range(m.L - 1):
        t = lin(res, m.fc[i])
        res = torch.relu(attend(lin(t, m.efcL[i]).reshape(-1), lin(t, m.efcR[i]).reshape(-1), t))
    t = lin(res, m.fc[-1])
    want_gat = lin(attend(lin(t, m.efcL[-1]).reshape(-1), lin(t, m.efcR[-1]).reshape(-1), res), m.fc[-1])
    assert rel_err(got[0], want_gat.float().numpy()) < 1e-5
    c = GCNN(dims, "cpu", seed=8)
    nrm = A.sum(1).pow(-0.5)[:, None]
    res = X
    for i in range(c.L - 1):
        res = torch.relu(nrm * (A @ (nrm * lin(res, c.fc[i]))))
    want_gcn = lin(nrm * (A @ (nrm * res)), c.fc[-1])
    assert rel_err(got[1], want_gcn.float().numpy()) < 1e-5


def _worker_masks(rank, world, port, n, e, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=9, device="cpu")
    part = dist_gat.RowPartition(offset, ids, n, rank, world)
    mask = part.need_masks(ids, offset)
    np.save(f"{out_path}.{rank}.npy", mask.numpy())
    np.save(f"{out_path}.{rank}.bounds.npy", np.array(part.bounds))
    dist.barrier()
    dist.destroy_process_group()


def test_need_masks_mark_exactly_the_rows_each_peer_references(tmp_path):
    """RowPartition.need_masks (the predicate of the fused needed-rows exchange): bit q of row r's mask is set iff
    rank q's slab has an edge whose column is r, and every row carries its owner's bit."""
    n, e, world = 500, 3000, 3
    out_path = str(tmp_path / "mask")
    mp.spawn(_worker_masks, args=(world, _free_port(), n, e, out_path), nprocs=world, join=True)
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=9, device="cpu")
    offset, ids = offset.numpy(), ids.numpy()
    bounds = np.load(f"{out_path}.0.bounds.npy")
    some_unneeded = False
    for owner in range(world):
        mask = np.load(f"{out_path}.{owner}.npy")
        lo, hi = bounds[owner], bounds[owner + 1]
        assert mask.shape[0] == max(hi - lo, 1)
        for q in range(world):
            cols_q = ids[offset[bounds[q]]:offset[bounds[q + 1]]]            # columns referenced by rank q's slab
            ref = np.zeros(n, bool)
            ref[cols_q] = True
            want = ref[lo:hi] if q != owner else np.ones(hi - lo, bool)
            got = ((mask[:hi - lo] >> q) & 1).astype(bool)
            assert np.array_equal(got, want), (owner, q)
            some_unneeded |= not want.all()
    assert some_unneeded          # the graph is sparse enough for the masks to matter
