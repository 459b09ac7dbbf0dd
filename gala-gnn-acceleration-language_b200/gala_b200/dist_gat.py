"""1-D row-partitioned execution over the GPUs of one node.

The reference is single-GPU (`device(torch::kCUDA, 0)` everywhere, src/codegen/cuda.h:226,462);
this layer is new.  Rows (output nodes) are split into P contiguous blocks balanced by nnz;
rank p holds the CSR slab of its rows with column ids remapped into the padded all-gather
layout, so that one `all_gather_into_tensor` of the hidden-width features per layer is the
only exchange.  Attention needs no extra exchange: aR is recomputed from the gathered
features (an N x K GEMV), aL only for the rank's own rows.

Host-side logic (partition, remap, padded layout) is plain torch and runs on CPU tensors
too -- tests/test_dist_cpu.py drives it with the gloo backend, world_size 2.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F


def partition_rows_by_nnz(offset, world):
    """Row boundaries [r_0=0, ..., r_P=N] such that every block holds ~E/P edges
    (prefix-sum split on the row pointers; cf. nnz_ord_row_tile_info,
    reference src/ops/tiling.h:1656-1708)."""
    n = offset.numel() - 1
    e = int(offset[-1])
    targets = torch.arange(1, world, dtype=torch.int64, device=offset.device) * e // world
    cuts = torch.searchsorted(offset.to(torch.int64), targets, right=False).clamp_(0, n)
    b = torch.cat([torch.zeros(1, dtype=torch.int64, device=offset.device), cuts,
                   torch.full((1,), n, dtype=torch.int64, device=offset.device)])
    return torch.cummax(b, 0).values.tolist()


class RowPartition:
    """Rank `rank`'s slab of a CSR graph in the padded all-gather layout.

    Gathered buffers have shape [world * max_rows, K]; row j of rank q sits at
    q * max_rows + (j - bounds[q]).  `cols` are remapped into that layout once here."""

    def __init__(self, offset, ids, n, rank, world):
        self.n, self.rank, self.world = n, rank, world
        self.bounds = partition_rows_by_nnz(offset, world)
        self.row_lo, self.row_hi = self.bounds[rank], self.bounds[rank + 1]
        self.rows = self.row_hi - self.row_lo
        self.max_rows = max(self.bounds[q + 1] - self.bounds[q] for q in range(world))
        e_lo, e_hi = int(offset[self.row_lo]), int(offset[self.row_hi])
        self.local_nvals = e_hi - e_lo
        self.offset = (offset[self.row_lo:self.row_hi + 1] - e_lo).to(torch.int32).contiguous()
        bt = torch.tensor(self.bounds, dtype=torch.int64, device=ids.device)
        cols = ids[e_lo:e_hi].to(torch.int64)
        owner = torch.searchsorted(bt, cols, right=True) - 1
        self.cols = (owner * self.max_rows + (cols - bt[owner])).to(torch.int32).contiguous()
        self.padded_n = world * self.max_rows

    def need_masks(self, ids, offset):
        """uint8 [rows]: bit q set iff rank q's slab references the row (bit `rank` always set).  With it the fused
        exchange stores a row only into the buffers of the GPUs that will gather it ("exchange only the unique
        remote rows each peer needs") -- on the random Papers-shape graph a slab references
        93.5 / 79 / 59 % of the remote rows at 2 / 4 / 8 GPUs.  One all_to_all of counts and one of indices, at
        partition time.  `ids`, `offset`: the global CSR the partition was cut from (this rank reads only its slab)."""
        dev = ids.device
        e_lo, e_hi = int(offset[self.row_lo]), int(offset[self.row_hi])
        bt = torch.tensor(self.bounds, dtype=torch.int64, device=dev)
        cols = ids[e_lo:e_hi].to(torch.int64)
        owner = torch.searchsorted(bt, cols, right=True) - 1
        want = []
        for q in range(self.world):
            want.append(torch.unique(cols[owner == q]) - self.bounds[q] if q != self.rank
                        else torch.empty(0, dtype=torch.int64, device=dev))
        counts = torch.tensor([int(w.numel()) for w in want], dtype=torch.int64, device=dev)
        incoming = torch.empty(self.world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(incoming, counts)
        in_splits = [int(c) for c in incoming.tolist()]
        req = torch.cat(want)
        got = torch.empty(sum(in_splits), dtype=torch.int64, device=dev)
        dist.all_to_all_single(got, req, in_splits, [int(c) for c in counts.tolist()])
        mask = torch.full((max(self.rows, 1),), 1 << self.rank, dtype=torch.int32, device=dev)
        pos = 0
        for q, c in enumerate(in_splits):
            if c:
                mask[got[pos:pos + c]] |= (1 << q)       # unique indices per requester: no write conflicts
            pos += c
        self.need_fraction = float(sum(in_splits)) / max(self.rows * max(self.world - 1, 1), 1)
        return mask.to(torch.uint8)

    def pad(self, x_local):
        """[rows, K] -> [max_rows, K] (zero rows at the end)."""
        if x_local.shape[0] == self.max_rows:
            return x_local.contiguous()
        out = x_local.new_zeros((self.max_rows,) + tuple(x_local.shape[1:]))
        out[:self.rows] = x_local
        return out

    def all_gather(self, x_local):
        """Every rank's rows, padded layout: [world * max_rows, K]."""
        send = self.pad(x_local)
        out = send.new_empty((self.padded_n,) + tuple(send.shape[1:]))
        dist.all_gather_into_tensor(out, send)
        return out

    def local_slice(self, gathered):
        lo = self.rank * self.max_rows
        return gathered[lo:lo + self.rows]

    def unpad(self, gathered):
        """Padded layout -> natural node order [n, K] (used by tests / final gathers)."""
        parts = [gathered[q * self.max_rows:q * self.max_rows + self.bounds[q + 1] - self.bounds[q]]
                 for q in range(self.world)]
        return torch.cat(parts, 0)


class NeededRowsPartition:
    """Same interface as RowPartition, but a rank only ever receives the feature rows its slab references
    ("all-to-all-v of only the unique remote rows each peer needs, index lists
    precomputed at partition time").  Local column numbering: [own rows | rows needed from rank 0 | ... |
    rows needed from rank P-1] (the own block first, the requester's own rank contributes nothing).

    `all_gather(x_local)` keeps its name for the runners above; here it is one all_to_all_single with the
    precomputed splits and returns [rows + n_remote, K].  Only valid with exchange="nccl": the fused peer
    exchange (PeerExchange) addresses the padded all-gather layout and rejects this partition."""

    def __init__(self, offset, ids, n, rank, world):
        self.n, self.rank, self.world = n, rank, world
        self.bounds = partition_rows_by_nnz(offset, world)
        self.row_lo, self.row_hi = self.bounds[rank], self.bounds[rank + 1]
        self.rows = self.row_hi - self.row_lo
        self.max_rows = max(self.bounds[q + 1] - self.bounds[q] for q in range(world))
        dev = ids.device
        e_lo, e_hi = int(offset[self.row_lo]), int(offset[self.row_hi])
        self.local_nvals = e_hi - e_lo
        self.offset = (offset[self.row_lo:self.row_hi + 1] - e_lo).to(torch.int32).contiguous()
        bt = torch.tensor(self.bounds, dtype=torch.int64, device=dev)
        cols = ids[e_lo:e_hi].to(torch.int64)
        owner = torch.searchsorted(bt, cols, right=True) - 1
        new_cols = torch.empty_like(cols)
        mine = owner == rank
        new_cols[mine] = cols[mine] - self.row_lo
        base = self.rows
        want = []                      # per owner: the sorted unique global rows this rank needs from it
        for q in range(world):
            if q == rank:
                want.append(torch.empty(0, dtype=torch.int64, device=dev))
                continue
            sel = owner == q
            uq, inv = torch.unique(cols[sel], return_inverse=True)
            new_cols[sel] = base + inv
            base += uq.numel()
            want.append(uq - self.bounds[q])          # owner-local row indices
        self.cols = new_cols.to(torch.int32).contiguous()
        self.padded_n = base                          # columns of the local graph
        self.recv_splits = [int(w.numel()) for w in want]
        # tell every owner which of its rows to send here (one exchange of counts, one of indices)
        counts = torch.tensor(self.recv_splits, dtype=torch.int64, device=dev)
        send_counts = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(send_counts, counts)
        self.send_splits = [int(c) for c in send_counts.tolist()]
        req = torch.cat(want) if base > self.rows else torch.empty(0, dtype=torch.int64, device=dev)
        self.send_idx = torch.empty(sum(self.send_splits), dtype=torch.int64, device=dev)
        dist.all_to_all_single(self.send_idx, req, self.send_splits, self.recv_splits)
        self.n_remote = base - self.rows

    def all_gather(self, x_local):
        """[rows, K] -> [rows + n_remote, K]: own rows followed by the needed remote rows."""
        x_local = x_local.contiguous()
        tail = tuple(x_local.shape[1:])
        send = x_local[self.send_idx].contiguous()
        recv = x_local.new_empty((self.n_remote,) + tail)
        dist.all_to_all_single(recv, send, self.recv_splits, self.send_splits)
        return torch.cat([x_local, recv], 0)

    def local_slice(self, gathered):
        return gathered[:self.rows]

    def gather_full(self, x_local):
        """Every rank's rows in natural node order [n, K] (tests / final gathers only)."""
        pad = x_local.new_zeros((self.max_rows,) + tuple(x_local.shape[1:]))
        pad[:self.rows] = x_local
        out = pad.new_empty((self.world * self.max_rows,) + tuple(x_local.shape[1:]))
        dist.all_gather_into_tensor(out, pad)
        return torch.cat([out[q * self.max_rows:q * self.max_rows + self.bounds[q + 1] - self.bounds[q]]
                          for q in range(self.world)], 0)

    def exchange_fraction(self):
        """Rows received per exchange relative to the full all-gather (n - rows)."""
        return self.n_remote / max(self.n - self.rows, 1)


def gat2_forward_partitioned(model, part, X_local, aggregate, hook=None):
    """The 2-layer GAT forward of gat_model.GAT2 on a row partition.
    aggregate(aL_local, aR_all, feats_all, relu) -> [rows, K] runs the fused GAT kernel on
    the rank's slab (or, in the CPU tests, the oracle)."""
    run = hook if hook is not None else (lambda name, fn: fn())
    res_loc = F.linear(X_local, *model.fc0)
    res_all = part.all_gather(res_loc)
    aL = F.linear(res_loc, *model.efc0).reshape(-1)
    aR = F.linear(res_all, *model.efc1).reshape(-1)
    y_loc = run("gat_layer1", lambda: aggregate(aL, aR, res_all, True))
    y_all = part.all_gather(y_loc)
    t_all = F.linear(y_all, *model.fc1)
    aL = F.linear(part.local_slice(t_all), *model.efc2).reshape(-1)
    aR = F.linear(t_all, *model.efc3).reshape(-1)
    agg = run("gat_layer2", lambda: aggregate(aL, aR, y_all, False))
    return F.linear(agg, *model.fc1)


def gat2_forward_partitioned_folded(model, part, X_local, aggregate, hook=None):
    """Folded attention projections (gat_model mode="folded"): one [K,2] matmul per layer on the
    gathered features gives aR for every column and aL for the rank's own rows."""
    run = hook if hook is not None else (lambda name, fn: fn())
    res_loc = F.linear(X_local, *model.fc0)
    res_all = part.all_gather(res_loc)
    a = F.linear(res_all, model.W_att1, model.b_att1).t().contiguous()
    y_loc = run("gat_layer1", lambda: aggregate(part.local_slice(a[0]), a[1], res_all, True))
    y_all = part.all_gather(y_loc)
    a = F.linear(y_all, model.W_att2, model.b_att2).t().contiguous()
    agg = run("gat_layer2", lambda: aggregate(part.local_slice(a[0]), a[1], y_all, False))
    return F.linear(agg, *model.fc1)


def gat2_forward_partitioned_reflected(model, part, X_local, aggregate_col, hook=None):
    """The same forward with the hidden rows exchanged in the reflected basis of the layer that gathers them
    (gat_model.GAT2.fold_reflected): the right-hand attention term is the last column of every exchanged row, so
    nothing but the rows is exchanged and nothing is derived from gathered rows.  All-gather exchange (NCCL / gloo).
      aggregate_col(aL_local, sR, bR, feats_all, relu, v_in, v_out) -> [rows, K]   (ops.gat_forward_col on the slab)"""
    run = hook if hook is not None else (lambda name, fn: fn())
    r = model.fold_reflected()
    res_loc = F.linear(X_local, r["W0"], r["b0"])
    aL = F.linear(res_loc, r["W_att1"][:1], model.b_att1[:1]).reshape(-1)
    res_all = part.all_gather(res_loc)
    y_loc = run("gat_layer1", lambda: aggregate_col(aL, r["s1"], model.bR1, res_all, True, r["v1"], r["v2"]))
    aL = F.linear(y_loc, r["W_att2"][:1], model.b_att2[:1]).reshape(-1)
    y_all = part.all_gather(y_loc)
    agg = run("gat_layer2", lambda: aggregate_col(aL, r["s2"], model.bR2, y_all, False, r["v2"], None))
    return F.linear(agg, *model.fc1)


def gat2_forward_partitioned_dot(model, part, X_local, aggregate_dot, hook=None):
    """Same forward with the right-hand attention term recomputed inside the kernel
    (ops.gat_forward_dot): nothing but the hidden features is exchanged or re-derived."""
    run = hook if hook is not None else (lambda name, fn: fn())
    res_loc = F.linear(X_local, *model.fc0)
    aL = F.linear(res_loc, *model.efc0).reshape(-1)
    res_all = part.all_gather(res_loc)
    y_loc = run("gat_layer1", lambda: aggregate_dot(aL, model.wR1, model.bR1, res_all, True))
    aL = torch.addmv(torch.full((y_loc.shape[0],), model.bL2, device=y_loc.device), y_loc, model.wL2)
    y_all = part.all_gather(y_loc)
    agg = run("gat_layer2", lambda: aggregate_dot(aL, model.wR2, model.bR2, y_all, False))
    return F.linear(agg, *model.fc1)


def gatn_forward_partitioned(model, part, X_local, aggregate, hook=None, linear_att=None, aggregate_att=None):
    """L-layer GAT (gat_model.GATN) on a row partition, all-gather exchange.  Per hidden layer the rank
    transforms its OWN rows and projects them onto the two attention vectors; the hidden-width rows and the
    one right-hand attention scalar per row are what is exchanged (nothing is recomputed on gathered rows).
    The last hidden output is exchanged once more for the final aggregation.
      aggregate(aL_local, aR_all, feats_all, relu) -> [rows, K]
      linear_att(res_local, i) -> (t_local [rows, K], att [2, rows])          default: F.linear
      aggregate_att(aL, aR_all, feats_all) -> (y_local, att_next [2, rows])    last hidden layer; default composes"""
    run = hook if hook is not None else (lambda name, fn: fn())

    def gather_vec(v):
        return part.all_gather(v.reshape(-1, 1)).reshape(-1)

    def default_linear_att(res_loc, i):
        t = F.linear(res_loc, *model.fc[i])
        return t, F.linear(t, model.W_att[i], model.b_att[i]).t().contiguous()

    def default_aggregate_att(aL, aR_all, feats_all):
        y = aggregate(aL, aR_all, feats_all, True)
        return y, F.linear(y, model.W_att[-1], model.b_att[-1]).t().contiguous()

    linear_att = linear_att or default_linear_att
    aggregate_att = aggregate_att or default_aggregate_att
    res_loc, att = X_local, None
    for i in range(model.L - 1):
        t_loc, a = run(f"linear{i + 1}", lambda: linear_att(res_loc, i))
        t_all, aR_all = part.all_gather(t_loc), gather_vec(a[1])
        if i == model.L - 2:
            res_loc, att = run(f"gat_layer{i + 1}", lambda: aggregate_att(a[0], aR_all, t_all))
        else:
            res_loc = run(f"gat_layer{i + 1}", lambda: aggregate(a[0], aR_all, t_all, True))
    y_all, aR_all = part.all_gather(res_loc), gather_vec(att[1])
    agg = run(f"gat_layer{model.L}", lambda: aggregate(att[0], aR_all, y_all, False))
    return F.linear(agg, *model.fc[-1])

class PeerExchange:
    """Symmetric (peer-mapped) gathered buffers: every rank's kernels push their output rows
    straight into all GPUs' copies while they compute (NVLS multicast store when the fabric
    supports it, per-peer stores otherwise); a device-side barrier replaces the all-gather."""

    def __init__(self, part, widths, device, need_mask=None):
        """need_mask (RowPartition.need_masks): rows are stored only into the GPUs that reference them -- predicated
        peer stores instead of the NVLS multicast store, which by construction delivers every row to every GPU."""
        import torch.distributed._symmetric_memory as symm

        # every rank's slab sits at rank * max_rows of a [world * max_rows, K] buffer: only the padded
        # all-gather layout of RowPartition has that shape (NeededRowsPartition numbers its columns per rank)
        if part.padded_n != part.world * part.max_rows or isinstance(part, NeededRowsPartition):
            raise ValueError("PeerExchange needs the padded all-gather layout (RowPartition); "
                             "use exchange=\"nccl\" with NeededRowsPartition")
        self.part = part
        self.bufs, self.hdls, self.mos = [], [], []
        self.att_bufs, self.att_mos = [], []
        self._peer = []
        self.copy_streams = None
        self._addr = []      # per exchange: (peer bases of the rows, multicast base, peer bases of the scalars, multicast, K)
        from . import ops
        for K in widths:
            # one symmetric allocation per exchange: [padded_n, K] feature rows followed by [padded_n] floats, the
            # right-hand attention scalar of every row (the only other quantity a GAT layer gathers by column)
            flat = symm.empty((part.padded_n * (K + 1),), dtype=torch.float32, device=device)
            flat.zero_()   # padding rows are never written
            hdl = symm.rendezvous(flat, dist.group.WORLD)
            off = part.rank * part.max_rows * K * 4
            att_off = part.padded_n * K * 4 + part.rank * part.max_rows * 4
            mc = mc_att = None
            try:
                if need_mask is None and hdl.has_multicast_support and hdl.multicast_ptr:
                    mc, mc_att = hdl.multicast_ptr + off, hdl.multicast_ptr + att_off
            except Exception:
                mc = mc_att = None
            self.bufs.append(flat[:part.padded_n * K].view(part.padded_n, K))
            self.att_bufs.append(flat[part.padded_n * K:])
            self.hdls.append(hdl)
            self.mos.append(ops.make_multi_out([p + off for p in hdl.buffer_ptrs], mc, need_mask))
            self.att_mos.append(ops.make_multi_out([p + att_off for p in hdl.buffer_ptrs], mc_att, need_mask))
            self._addr.append(([p + off for p in hdl.buffer_ptrs], mc, [p + att_off for p in hdl.buffer_ptrs], mc_att, K))
            # every GPU's copy of the buffer as a local tensor (peer-mapped): destinations of copy-engine transfers
            self._peer.append([hdl.get_buffer(q, (part.padded_n * (K + 1),), torch.float32) for q in range(part.world)])
        self.need_mask = need_mask
        torch.cuda.synchronize()
        dist.barrier()

    def barrier(self, i):
        self.hdls[i].barrier(channel=i)

    def dma_push(self, i, lo, hi, rows, scalars, after):
        """Rows [lo, hi) of this rank's slab (and their right-hand attention scalars) into every GPU's gathered buffer
        of exchange i with device-to-device copies -- the copy engines move them, no SM is involved, so an aggregation
        kernel running at the same time keeps the whole GPU.  One stream per destination; every copy waits for the
        event `after` (the producer of the block).  Sends every row to every GPU (a copy engine cannot skip the rows a
        destination does not reference)."""
        part = self.part
        if self.copy_streams is None:
            self.copy_streams = [torch.cuda.Stream(device=rows.device) for _ in range(part.world)]
        K = rows.shape[1]
        r0 = part.rank * part.max_rows
        for d in range(part.world):
            q = (part.rank + d) % part.world          # staggered: the ranks do not all start on the same destination
            st = self.copy_streams[q]
            st.wait_event(after)
            with torch.cuda.stream(st):
                flat = self._peer[i][q]
                flat[(r0 + lo) * K:(r0 + hi) * K].view(hi - lo, K).copy_(rows, non_blocking=True)
                if scalars is not None:
                    a0 = part.padded_n * K + r0
                    flat[a0 + lo:a0 + hi].copy_(scalars, non_blocking=True)

    def dma_join(self, main):
        for st in self.copy_streams or ():
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)

    def block_out(self, i, row0):
        """(rows, scalars) multi-outs of exchange i for a producer that starts at local row `row0` of this rank's slab
        (the row-block pipeline launches one transform per block)."""
        from . import ops
        bases, mc, att_bases, mc_att, K = self._addr[i]
        need = self.need_mask[row0:] if self.need_mask is not None else None
        mo = ops.make_multi_out([b + row0 * K * 4 for b in bases], mc + row0 * K * 4 if mc else None, need)
        att_mo = ops.make_multi_out([b + row0 * 4 for b in att_bases], mc_att + row0 * 4 if mc_att else None, need)
        return mo, att_mo


class RowBlocks:
    """A rank's slab cut into `nblocks` row blocks of equal nnz, each with its own graph view and plan: the unit of
    the row-block pipeline -- while block j of layer l is aggregated, the transformed rows of block j-1 for layer
    l+1 are already on their way to the other GPUs (side stream), so that exchange hides behind this aggregation
    instead of following it (SURVEY.md section 8e "the local part computes while remote rows arrive")."""

    def __init__(self, part, nblocks, device):
        from . import ops
        self.cuts = partition_rows_by_nnz(part.offset.to(torch.int64), nblocks)
        self.graphs = []
        for j in range(nblocks):
            lo, hi = self.cuts[j], self.cuts[j + 1]
            g = ops.TiledGraph(part.offset[lo:hi + 1], part.cols, hi - lo, ncols=part.padded_n).build_plan()
            g.plan.tile_rows = None      # (the edge-tile table assumes row pointers that start at 0; not used here)
            self.graphs.append(g)
        self.side = torch.cuda.Stream(device=device, priority=-1)
        self.n = nblocks

    def span(self, j):
        return self.cuts[j], self.cuts[j + 1]


class PartitionedGAT:
    """GPU runner: slab graph + plan + fused kernel.  exchange="p2p": kernels push their rows to every
    GPU (PeerExchange); exchange="nccl": all_gather_into_tensor between layers."""

    launches_per_step = 4     # reflected mode: transform, two aggregation layers, classifier ("folded": 3)

    def __init__(self, model, offset, ids, n, rank, world, device, exchange="p2p"):
        from . import ops

        self.model, self.ops = model, ops
        self.part = RowPartition(offset, ids, n, rank, world)
        self.row_lo, self.row_hi = self.part.row_lo, self.part.row_hi
        self.local_nvals = self.part.local_nvals
        self.graph = ops.TiledGraph(self.part.offset, self.part.cols, self.part.rows,
                                    ncols=self.part.padded_n).build_plan()
        self.px = None
        self.exchange = "nccl"
        if exchange in ("p2p", "p2p-needed"):
            try:
                hidden = model.fc0[0].shape[0]
                need = self.part.need_masks(ids, offset) if exchange == "p2p-needed" else None
                self.px = PeerExchange(self.part, [hidden, hidden], device, need_mask=need)
                self.exchange = ("p2p-needed" if need is not None else
                                 "p2p-multicast" if self.px.mos[0].multicast_base else "p2p")
            except Exception as ex:   # no symmetric memory on this system: NCCL all-gather instead
                import sys
                sys.stderr.write(f"gala_b200.dist_gat: peer exchange unavailable ({type(ex).__name__}: {ex}); using NCCL\n")
                self.px = None

    def forward_p2p(self, X_local, hook=None):
        """Folded-projection forward with both exchanges fused into the producing kernels: every kernel pushes its
        output rows AND the right-hand attention scalar of those rows (computed in its epilogue) into all GPUs'
        gathered buffers, the left-hand scalar stays local.  Three launches and two device barriers per step; nothing
        is recomputed on gathered rows."""
        run = hook if hook is not None else (lambda name, fn: fn())
        m, px, ops = self.model, self.px, self.ops
        _, a = run("linear1", lambda: ops.linear(X_local, m.fc0[0], m.fc0[1], att_w=m.W_att1, att_b=m.b_att1_host,
                                                 multi_out=px.mos[0], att_multi_out=px.att_mos[0]))
        px.barrier(0)
        _, a2, _ = run("gat_layer1", lambda: ops.gat_forward_ex(self.graph, a[0], px.att_bufs[0], px.bufs[0], m.slope,
                                                                relu=True, att_w=m.W_att2, att_b=m.b_att2_host,
                                                                multi_out=px.mos[1], att_multi_out=px.att_mos[1]))
        px.barrier(1)
        _, _, out = run("gat_layer2", lambda: ops.gat_forward_ex(self.graph, a2[0], px.att_bufs[1], px.bufs[1], m.slope,
                                                                 relu=False, cls_wT=m.fc1_wT, cls_b=m.fc1[1],
                                                                 want_y=False))
        return out

    def forward_p2p_reflected(self, X_local, hook=None):
        """forward_p2p with the hidden rows exchanged in the reflected basis of the layer that gathers them
        (gat_model.GAT2.fold_reflected, gala_gat_forward_col_f32): the right-hand attention term is the last column
        of every pushed row, so no scalar is exchanged or gathered, and each edge costs one gather instead of two."""
        run = hook if hook is not None else (lambda name, fn: fn())
        m, px, ops = self.model, self.px, self.ops
        r = m.fold_reflected()
        _, a = run("linear1", lambda: ops.linear(X_local, r["W0"], r["b0"], att_w=r["W_att1"], att_b=m.b_att1_host,
                                                 multi_out=px.mos[0]))
        px.barrier(0)
        _, a2, _ = run("gat_layer1", lambda: ops.gat_forward_col_ex(
            self.graph, a[0], r["s1"], m.bR1, px.bufs[0], m.slope, relu=True, reflect_in=r["v1"], reflect_out=r["v2"],
            att_w=r["W_att2"], att_b=m.b_att2_host, multi_out=px.mos[1]))
        px.barrier(1)
        agg = run("gat_layer2", lambda: ops.gat_forward_col(self.graph, a2[0], r["s2"], m.bR2, px.bufs[1], m.slope,
                                                            relu=False, reflect_in=r["v2"]))
        # (the classifier as its own streaming launch: in the aggregation kernel's epilogue it costs more than the
        #  row-per-thread transform takes -- 0.048 vs 0.025 ms per rank at 2 GPUs)
        return run("classifier", lambda: ops.linear_small(agg, m.fc1[0], m.fc1[1]))

    def _aggregate(self, aL, aR, feats, relu):
        return self.ops.gat_forward(self.graph, aL.contiguous(), aR.contiguous(), feats, self.model.slope, relu=relu)

    def _aggregate_col(self, aL, sR, bR, feats, relu, v_in, v_out):
        return self.ops.gat_forward_col(self.graph, aL.contiguous(), sR, bR, feats, self.model.slope, relu=relu,
                                        reflect_in=v_in, reflect_out=v_out)

    def _aggregate_dot(self, aL, wR, bR, feats, relu):
        return self.ops.gat_forward_dot(self.graph, aL.contiguous(), wR, bR, feats, self.model.slope, relu=relu)

    def forward(self, X_local, hook=None, mode="reflected"):
        if mode == "reflected":
            if self.model.fc0[0].shape[0] not in (4, 8, 16, 32):
                mode = "folded"
            elif self.px is not None:
                return self.forward_p2p_reflected(X_local, hook)
            else:
                return gat2_forward_partitioned_reflected(self.model, self.part, X_local, self._aggregate_col, hook)
        if self.px is not None and mode in ("folded", "fused"):
            return self.forward_p2p(X_local, hook)
        if mode == "dot":
            return gat2_forward_partitioned_dot(self.model, self.part, X_local, self._aggregate_dot, hook)
        if mode in ("folded", "fused"):   # the row-partitioned runner has no separate fused variant
            return gat2_forward_partitioned_folded(self.model, self.part, X_local, self._aggregate, hook)
        return gat2_forward_partitioned(self.model, self.part, X_local, self._aggregate, hook)


class PartitionedGATN:
    """L-layer runner (gat_model.GATN).  `part` is any object with RowPartition's interface, so that a
    rank can build its slab without ever holding the whole graph."""

    def __init__(self, model, part, device, exchange="p2p", need_mask=None, pipeline=0, push_ctas=0, phases="m",
                 reflected=False):
        """reflected (fused exchanges, hidden widths in {4, 8, 16, 32}): every exchanged row travels in the reflected
        basis of the layer that gathers it (gat_model.GATN.fold_reflected, gala_gat_forward_col_f32) -- its last column
        is the right-hand attention term, so no scalar is exchanged and an edge costs one gather; kernel pushes only
        (no copy-engine phases).
        pipeline = B > 1 (fused exchanges only): producers of exchanged rows run in B row blocks and a finished block
        is sent from a side stream under the computation of the following blocks.  phases: which exchanges do so --
        f(irst transform), m(iddle: aggregation || next transform + push), l(ast hidden aggregation); the others keep
        the push in the producing kernel's epilogue.  Upper case (F, M, L): the finished block travels by copy engine
        (PeerExchange.dma_push: device-to-device copies on per-destination streams, no SM involved) instead of by a
        kernel on the side stream."""
        from . import ops

        self.model, self.ops, self.part = model, ops, part
        self.graph = ops.TiledGraph(part.offset, part.cols, part.rows, ncols=part.padded_n).build_plan()
        self.px = None
        self.exchange = "nccl"
        self.blocks = None
        self.push_ctas, self.phases = push_ctas, phases
        self.reflected = bool(reflected) and exchange in ("p2p", "p2p-needed") and model.reflected_ok()
        if self.reflected:
            assert not any(c in phases for c in "FML"), "reflected rows are pushed by kernels, not by copy engines"
            self.r = model.fold_reflected()
        if exchange in ("p2p", "p2p-needed"):
            assert (exchange == "p2p-needed") == (need_mask is not None), "p2p-needed takes RowPartition.need_masks()"
            self.px = PeerExchange(part, [model.dims[i + 1] for i in range(model.L - 1)] + [model.dims[-2]], device,
                                   need_mask=need_mask)
            self.exchange = ("p2p-needed" if need_mask is not None else
                             "p2p-multicast" if self.px.mos[0].multicast_base else "p2p")
            if self.reflected:
                self.exchange += "+reflected"
            if pipeline > 1 and max(model.dims[1:-1]) <= 64:
                self.blocks = RowBlocks(part, pipeline, device)
                self.exchange += f"+pipeline{pipeline}" + ("" if phases == "m" else ":" + phases)
                # per-forward buffers are allocated once: they are written on one stream and read on the other
                L = model.L
                self._t0 = torch.empty(part.rows, model.dims[1], device=device)
                self._t = [torch.empty(part.rows, model.dims[i + 2], device=device) for i in range(max(L - 2, 0))]
                self._res = [torch.empty(part.rows, model.dims[i + 1], device=device) for i in range(L - 1)]
                self._att = [[torch.empty(2, hi - lo, device=device) for lo, hi in map(self.blocks.span, range(pipeline))]
                             for _ in range(max(L - 2, 0))]
                self._aL = [torch.empty(part.rows, device=device) for _ in range(L)]
                self._outs = [[self.px.block_out(i, self.blocks.cuts[j]) for j in range(pipeline)] for i in range(L)]

    def _aggregate(self, aL, aR, feats, relu):
        return self.ops.gat_forward(self.graph, aL.contiguous(), aR.contiguous(), feats, self.model.slope, relu=relu)

    # ---- the three kinds of aggregation of the fused-exchange forward, in either basis -----------------------------
    def _fc(self, i):
        """(W, b, W_att) of transform i as the kernels take them."""
        if self.reflected:
            return self.r["W"][i], self.r["b"][i], self.r["W_att"][i]
        return self.model.fc[i][0], self.model.fc[i][1], self.model.W_att[i]

    def _amo(self, att_mo):
        """The right-hand attention scalar's exchange: none in the reflected basis."""
        return None if self.reflected else att_mo

    def _agg_hidden(self, graph, i, aL, out=None):
        m, px, ops = self.model, self.px, self.ops
        if self.reflected:
            r = self.r
            return ops.gat_forward_col(graph, aL, r["s"][i], r["bR"][i], px.bufs[i], m.slope, relu=True,
                                       reflect_in=r["v"][i], out=out)
        return ops.gat_forward(graph, aL, px.att_bufs[i], px.bufs[i], m.slope, relu=True, out=out)

    def _agg_last_hidden(self, graph, aL, out=None, multi_out=None, att_multi_out=None):
        """-> (y or None, att [2, rows]: att[0] = the final layer's left-hand term; att[1] only in the original basis)"""
        m, px, ops = self.model, self.px, self.ops
        i, L = m.L - 2, m.L
        if self.reflected:
            r = self.r
            y, att, _ = ops.gat_forward_col_ex(graph, aL, r["s"][i], r["bR"][i], px.bufs[i], m.slope, relu=True,
                                               reflect_in=r["v"][i], reflect_out=r["v"][L - 1], att_w=r["W_att"][L - 1],
                                               att_b=m._bh[-1], out=out, multi_out=multi_out)
            return y, att
        y, att, _ = ops.gat_forward_ex(graph, aL, px.att_bufs[i], px.bufs[i], m.slope, relu=True, att_w=m.W_att[-1],
                                       att_b=m._bh[-1], out=out, multi_out=multi_out, att_multi_out=att_multi_out)
        return y, att

    def _agg_final(self, aL):
        m, px, ops = self.model, self.px, self.ops
        L = m.L
        if self.reflected:
            r = self.r
            return ops.gat_forward_col(self.graph, aL, r["s"][L - 1], r["bR"][L - 1], px.bufs[L - 1], m.slope,
                                       relu=False, reflect_in=r["v"][L - 1])
        return ops.gat_forward(self.graph, aL, px.att_bufs[L - 1], px.bufs[L - 1], m.slope, relu=False)

    def forward(self, X_local, hook=None, mark=None):
        """mark(name), if given, is called at every phase boundary (bench scripts record an event there)."""
        m, px, part, ops = self.model, self.px, self.part, self.ops
        mark = mark if mark is not None else (lambda name: None)
        if not hasattr(m, "_bh"):
            m.host_biases()
        bh = m._bh
        L = m.L
        if px is None:
            def linear_att(res_loc, i):
                return ops.linear(res_loc, m.fc[i][0], m.fc[i][1], att_w=m.W_att[i], att_b=bh[i])

            def aggregate_att(aL, aR_all, feats_all):
                y, att, _ = ops.gat_forward_ex(self.graph, aL, aR_all, feats_all, m.slope, relu=True,
                                               att_w=m.W_att[-1], att_b=bh[-1])
                return y, att
            return gatn_forward_partitioned(m, part, X_local, self._aggregate, hook, linear_att, aggregate_att)
        run = hook if hook is not None else (lambda name, fn: fn())
        res_loc, att = X_local, None
        if self.blocks is not None:
            return self._forward_pipelined(X_local, mark)
        for i in range(L - 1):
            # the transform pushes its rows to every GPU while it computes, projects them onto the attention
            # vectors in its epilogue and pushes the right-hand scalar of every row the same way
            W, b, Wa = self._fc(i)
            _, a = run(f"linear{i + 1}", lambda: ops.linear(res_loc, W, b, att_w=Wa, att_b=bh[i], multi_out=px.mos[i],
                                                            att_multi_out=self._amo(px.att_mos[i])))
            mark(f"linear{i + 1}+push")
            px.barrier(i)
            mark(f"exchange{i + 1}")
            if i == L - 2:
                _, att = run(f"gat_layer{i + 1}", lambda: self._agg_last_hidden(
                    self.graph, a[0], multi_out=px.mos[L - 1], att_multi_out=px.att_mos[L - 1]))
                mark(f"gat_layer{i + 1}+push")
            else:
                res_loc = run(f"gat_layer{i + 1}", lambda: self._agg_hidden(self.graph, i, a[0]))
                mark(f"gat_layer{i + 1}")
        px.barrier(L - 1)
        mark(f"exchange{L}")
        agg = run(f"gat_layer{L}", lambda: self._agg_final(att[0]))
        mark(f"gat_layer{L}")
        out = run("classifier", lambda: ops.dense(agg, *m.fc[-1]))
        mark("classifier")
        return out

    def _forward_pipelined(self, X_local, mark):
        """Row-block pipeline: every producer of exchanged rows runs block by block on the main stream, and as soon as
        a block is done the side stream sends it to the GPUs that gather it -- the first transform and the last hidden
        aggregation through the copy kernel (gala_push_rows_f32: whole lines per store), the transforms in between
        fused with their push in the light linear_small_ex kernel (so the aggregation that runs next to it keeps its
        registers).  Each exchange is then (nearly) over when its producer ends, instead of starting there."""
        m, px, ops, bl = self.model, self.px, self.ops, self.blocks
        bh, L = m._bh, m.L
        main = torch.cuda.current_stream()
        keep = []        # tensors handed from the main to the side stream stay referenced until the streams have joined

        def after_block(fn):
            ev = torch.cuda.Event()
            ev.record(main)
            bl.side.wait_event(ev)
            with torch.cuda.stream(bl.side):
                fn()

        def join():
            done = torch.cuda.Event()
            done.record(bl.side)
            main.wait_event(done)

        # first transform (tensor cores, local rows) || push
        t0, aL = self._t0, self._aL[0]
        if "F" in self.phases:
            for j in range(bl.n):
                lo, hi = bl.span(j)
                _, a = ops.linear(X_local[lo:hi], m.fc[0][0], m.fc[0][1], att_w=m.W_att[0], att_b=bh[0], out=t0[lo:hi])
                aL[lo:hi].copy_(a[0])
                keep.append(a)
                ev = torch.cuda.Event()
                ev.record(main)
                px.dma_push(0, lo, hi, t0[lo:hi], a[1], ev)
            px.dma_join(main)
            mark("linear1 || copy")
        elif "f" in self.phases:
            W, b, Wa = self._fc(0)
            for j in range(bl.n):
                lo, hi = bl.span(j)
                _, a = ops.linear(X_local[lo:hi], W, b, att_w=Wa, att_b=bh[0], out=t0[lo:hi])
                aL[lo:hi].copy_(a[0])
                keep.append(a)
                mo, att_mo = self._outs[0][j]
                after_block(lambda: ops.push_rows(t0[lo:hi], mo, scalars=None if self.reflected else a[1],
                                                  scalar_multi_out=self._amo(att_mo), max_ctas=self.push_ctas))
            join()
            mark("linear1 || push")
        else:
            W, b, Wa = self._fc(0)
            _, a = ops.linear(X_local, W, b, att_w=Wa, att_b=bh[0], multi_out=px.mos[0],
                              att_multi_out=self._amo(px.att_mos[0]))
            aL = a[0]
            mark("linear1+push")
        px.barrier(0)
        mark("exchange1")
        # hidden aggregations || the next layer's transform + push
        for i in range(L - 2):
            res, aL_next = self._res[i], self._aL[i + 1]
            if "M" in self.phases:
                tn = self._t[i]
                for j in range(bl.n):
                    lo, hi = bl.span(j)
                    ops.gat_forward(bl.graphs[j], aL[lo:hi], px.att_bufs[i], px.bufs[i], m.slope, relu=True, out=res[lo:hi])
                    _, a = ops.linear(res[lo:hi], m.fc[i + 1][0], m.fc[i + 1][1], att_w=m.W_att[i + 1], att_b=bh[i + 1],
                                      out=tn[lo:hi])
                    aL_next[lo:hi].copy_(a[0])
                    keep.append(a)
                    ev = torch.cuda.Event()
                    ev.record(main)
                    px.dma_push(i + 1, lo, hi, tn[lo:hi], a[1], ev)
                px.dma_join(main)
                mark(f"gat_layer{i + 1}, linear{i + 2} || copy")
                px.barrier(i + 1)
                mark(f"exchange{i + 2}")
                aL = aL_next
                continue
            W, b, Wa = self._fc(i + 1)
            if "m" not in self.phases:
                self._agg_hidden(self.graph, i, aL, out=res)
                mark(f"gat_layer{i + 1}")
                _, a = ops.linear(res, W, b, att_w=Wa, att_b=bh[i + 1], multi_out=px.mos[i + 1],
                                  att_multi_out=self._amo(px.att_mos[i + 1]))
                mark(f"linear{i + 2}+push")
                px.barrier(i + 1)
                mark(f"exchange{i + 2}")
                aL = a[0]
                continue
            for j in range(bl.n):
                lo, hi = bl.span(j)
                self._agg_hidden(bl.graphs[j], i, aL[lo:hi], out=res[lo:hi])
                mo, att_mo = self._outs[i + 1][j]

                def transform_push(lo=lo, hi=hi, j=j, mo=mo, att_mo=att_mo, W=W, b=b, Wa=Wa):
                    ops.linear_small_ex(res[lo:hi], W, b, att_w=Wa, att_b=bh[i + 1],
                                        att_out=self._att[i][j], multi_out=mo, att_multi_out=self._amo(att_mo),
                                        max_ctas=self.push_ctas)
                    aL_next[lo:hi].copy_(self._att[i][j][0])
                after_block(transform_push)
            join()
            mark(f"gat_layer{i + 1} || linear{i + 2}+push")
            px.barrier(i + 1)
            mark(f"exchange{i + 2}")
            aL = aL_next
        # last hidden aggregation (next layer's attention projections in its epilogue) || push
        res, aL_last = self._res[L - 2], self._aL[L - 1]
        if "L" in self.phases:
            for j in range(bl.n):
                lo, hi = bl.span(j)
                _, att, _ = ops.gat_forward_ex(bl.graphs[j], aL[lo:hi], px.att_bufs[L - 2], px.bufs[L - 2], m.slope,
                                               relu=True, att_w=m.W_att[-1], att_b=bh[-1], out=res[lo:hi])
                aL_last[lo:hi].copy_(att[0])
                keep.append(att)
                ev = torch.cuda.Event()
                ev.record(main)
                px.dma_push(L - 1, lo, hi, res[lo:hi], att[1], ev)
            px.dma_join(main)
            mark(f"gat_layer{L - 1} || copy")
        elif "l" in self.phases:
            for j in range(bl.n):
                lo, hi = bl.span(j)
                _, att = self._agg_last_hidden(bl.graphs[j], aL[lo:hi], out=res[lo:hi])
                aL_last[lo:hi].copy_(att[0])
                keep.append(att)
                mo, att_mo = self._outs[L - 1][j]
                after_block(lambda: ops.push_rows(res[lo:hi], mo, scalars=None if self.reflected else att[1],
                                                  scalar_multi_out=self._amo(att_mo), max_ctas=self.push_ctas))
            join()
            mark(f"gat_layer{L - 1} || push")
        else:
            _, att = self._agg_last_hidden(self.graph, aL, multi_out=px.mos[L - 1], att_multi_out=px.att_mos[L - 1])
            aL_last = att[0]
            mark(f"gat_layer{L - 1}+push")
        px.barrier(L - 1)
        mark(f"exchange{L}")
        agg = self._agg_final(aL_last)
        mark(f"gat_layer{L}")
        out = ops.dense(agg, *m.fc[-1])
        mark("classifier")
        keep.clear()
        return out


def gcnn_forward_partitioned(model, part, X_local, norm_local, aggregate, hook=None, linear=None):
    """L-layer GCN (gcn_model.GCNN) on a row partition with all-gather exchanges: hidden layers exchange the
    transformed, norm-scaled rows; the last hidden output (already pre-scaled) is exchanged once more.
      aggregate(feats_all, row_scale_local, relu) -> [rows, K];  linear(res_local, i) -> norm * fc_i(res_local)"""
    run = hook if hook is not None else (lambda name, fn: fn())
    n2 = norm_local * norm_local
    if linear is None:
        def linear(res_loc, i):
            return norm_local[:, None] * F.linear(res_loc, *model.fc[i])
    res_loc = X_local
    for i in range(model.L - 1):
        t_all = part.all_gather(run(f"linear{i + 1}", lambda: linear(res_loc, i)))
        res_loc = run(f"gcn_aggregate{i + 1}", lambda: aggregate(t_all, n2 if i == model.L - 2 else norm_local, True))
    agg = run(f"gcn_aggregate{model.L}", lambda: aggregate(part.all_gather(res_loc), norm_local, False))
    return F.linear(agg, *model.fc[-1])


class PartitionedGCNN:
    """L-layer GCN runner: the transform of every hidden layer pushes its (norm-scaled) rows to all GPUs from
    its epilogue (exchange="p2p"); the single exchange of an aggregation OUTPUT (before the last layer) goes
    through NCCL all-gather.  exchange="nccl": all-gather everywhere."""

    def __init__(self, model, part, device, exchange="p2p", need_mask=None, pipeline=0, push_ctas=0, phases="m"):
        from . import ops

        self.model, self.ops, self.part = model, ops, part
        self.blocks = None
        self.push_ctas, self.phases = push_ctas, phases
        self.graph = ops.TiledGraph(part.offset, part.cols, part.rows, ncols=part.padded_n).build_plan()
        ones = torch.ones(part.padded_n, 1, device=device)
        self.norm = torch.pow(ops.spmm(self.graph, ones).reshape(-1), -0.5).contiguous()    # own rows' degrees
        self.norm2 = (self.norm * self.norm).contiguous()
        self.px = None
        self.exchange = "nccl"
        if exchange in ("p2p", "p2p-needed"):
            # one buffer per hidden transform, plus one for the aggregated rows the last layer gathers
            self.px = PeerExchange(part, [model.dims[i + 1] for i in range(model.L - 1)] + [model.dims[-2]], device,
                                   need_mask=need_mask)
            self.exchange = ("p2p-needed" if need_mask is not None else
                             "p2p-multicast" if self.px.mos[0].multicast_base else "p2p")
            if pipeline > 1 and max(model.dims[1:-1]) <= 64:
                self.blocks = RowBlocks(part, pipeline, device)
                self.exchange += f"+pipeline{pipeline}" + ("" if phases == "m" else ":" + phases)
                self._t0 = torch.empty(part.rows, model.dims[1], device=device)
                self._t = [torch.empty(part.rows, model.dims[i + 2], device=device) for i in range(max(model.L - 2, 0))]
                self._res = [torch.empty(part.rows, model.dims[i + 1], device=device) for i in range(model.L - 1)]
                self._outs = [[self.px.block_out(i, self.blocks.cuts[j]) for j in range(pipeline)] for i in range(model.L)]

    def _aggregate(self, feats_all, row_scale, relu):
        return self.ops.spmm(self.graph, feats_all, row_scale=row_scale, relu=relu)

    def forward(self, X_local, hook=None, mark=None):
        m, px, part, ops = self.model, self.px, self.part, self.ops
        mark = mark if mark is not None else (lambda name: None)
        if px is None:
            def linear(res_loc, i):
                return ops.linear(res_loc, m.fc[i][0], m.fc[i][1], row_scale=self.norm)
            return gcnn_forward_partitioned(m, part, X_local, self.norm, self._aggregate, hook, linear)
        run = hook if hook is not None else (lambda name, fn: fn())
        res_loc = X_local
        if self.blocks is not None:
            return self._forward_pipelined(X_local, mark)
        for i in range(m.L - 1):
            run(f"linear{i + 1}", lambda: ops.linear(res_loc, m.fc[i][0], m.fc[i][1], row_scale=self.norm, multi_out=px.mos[i]))
            mark(f"linear{i + 1}+push")
            px.barrier(i)
            mark(f"exchange{i + 1}")
            rs = self.norm2 if i == m.L - 2 else self.norm
            if i == m.L - 2:
                # the last hidden aggregation pushes its finished rows to the GPUs that gather them (same fused
                # exchange as the transforms) instead of an NCCL all-gather of its output
                run(f"gcn_aggregate{i + 1}", lambda: ops.spmm(self.graph, px.bufs[i], row_scale=rs, relu=True,
                                                             multi_out=px.mos[m.L - 1]))
                mark(f"gcn_aggregate{i + 1}+push")
            else:
                res_loc = run(f"gcn_aggregate{i + 1}", lambda: ops.spmm(self.graph, px.bufs[i], row_scale=rs, relu=True))
                mark(f"gcn_aggregate{i + 1}")
        px.barrier(m.L - 1)
        y_all = px.bufs[m.L - 1]
        mark(f"exchange{m.L}")
        agg = run(f"gcn_aggregate{m.L}", lambda: ops.spmm(self.graph, y_all, row_scale=self.norm))
        mark(f"gcn_aggregate{m.L}")
        out = run("classifier", lambda: ops.dense(agg, *m.fc[-1]))
        mark("classifier")
        return out

    def _forward_pipelined(self, X_local, mark):
        """Row-block pipeline (see PartitionedGATN._forward_pipelined): every producer of exchanged rows runs block by
        block, the side stream sends a finished block while the following blocks are computed."""
        m, px, ops, bl = self.model, self.px, self.ops, self.blocks
        L = m.L
        main = torch.cuda.current_stream()

        def after_block(fn):
            ev = torch.cuda.Event()
            ev.record(main)
            bl.side.wait_event(ev)
            with torch.cuda.stream(bl.side):
                fn()

        def join():
            done = torch.cuda.Event()
            done.record(bl.side)
            main.wait_event(done)

        t0 = self._t0
        if "F" in self.phases:
            for j in range(bl.n):
                lo, hi = bl.span(j)
                ops.linear(X_local[lo:hi], m.fc[0][0], m.fc[0][1], row_scale=self.norm[lo:hi], out=t0[lo:hi])
                ev = torch.cuda.Event()
                ev.record(main)
                px.dma_push(0, lo, hi, t0[lo:hi], None, ev)
            px.dma_join(main)
            mark("linear1 || copy")
        elif "f" in self.phases:
            for j in range(bl.n):
                lo, hi = bl.span(j)
                ops.linear(X_local[lo:hi], m.fc[0][0], m.fc[0][1], row_scale=self.norm[lo:hi], out=t0[lo:hi])
                mo = self._outs[0][j][0]
                after_block(lambda: ops.push_rows(t0[lo:hi], mo, max_ctas=self.push_ctas))
            join()
            mark("linear1 || push")
        else:
            ops.linear(X_local, m.fc[0][0], m.fc[0][1], row_scale=self.norm, multi_out=px.mos[0])
            mark("linear1+push")
        px.barrier(0)
        mark("exchange1")
        for i in range(L - 2):
            res = self._res[i]
            if "M" in self.phases:
                tn = self._t[i]
                for j in range(bl.n):
                    lo, hi = bl.span(j)
                    ops.spmm(bl.graphs[j], px.bufs[i], row_scale=self.norm[lo:hi], relu=True, out=res[lo:hi])
                    ops.linear(res[lo:hi], m.fc[i + 1][0], m.fc[i + 1][1], row_scale=self.norm[lo:hi], out=tn[lo:hi])
                    ev = torch.cuda.Event()
                    ev.record(main)
                    px.dma_push(i + 1, lo, hi, tn[lo:hi], None, ev)
                px.dma_join(main)
                mark(f"gcn_aggregate{i + 1}, linear{i + 2} || copy")
                px.barrier(i + 1)
                mark(f"exchange{i + 2}")
                continue
            if "m" not in self.phases:
                ops.spmm(self.graph, px.bufs[i], row_scale=self.norm, relu=True, out=res)
                mark(f"gcn_aggregate{i + 1}")
                ops.linear(res, m.fc[i + 1][0], m.fc[i + 1][1], row_scale=self.norm, multi_out=px.mos[i + 1])
                mark(f"linear{i + 2}+push")
                px.barrier(i + 1)
                mark(f"exchange{i + 2}")
                continue
            for j in range(bl.n):
                lo, hi = bl.span(j)
                ops.spmm(bl.graphs[j], px.bufs[i], row_scale=self.norm[lo:hi], relu=True, out=res[lo:hi])
                mo = self._outs[i + 1][j][0]
                after_block(lambda: ops.linear_small_ex(res[lo:hi], m.fc[i + 1][0], m.fc[i + 1][1],
                                                        row_scale=self.norm[lo:hi], multi_out=mo, max_ctas=self.push_ctas))
            join()
            mark(f"gcn_aggregate{i + 1} || linear{i + 2}+push")
            px.barrier(i + 1)
            mark(f"exchange{i + 2}")
        res = self._res[L - 2]
        if "L" in self.phases:
            for j in range(bl.n):
                lo, hi = bl.span(j)
                ops.spmm(bl.graphs[j], px.bufs[L - 2], row_scale=self.norm2[lo:hi], relu=True, out=res[lo:hi])
                ev = torch.cuda.Event()
                ev.record(main)
                px.dma_push(L - 1, lo, hi, res[lo:hi], None, ev)
            px.dma_join(main)
            mark(f"gcn_aggregate{L - 1} || copy")
        elif "l" in self.phases:
            for j in range(bl.n):
                lo, hi = bl.span(j)
                ops.spmm(bl.graphs[j], px.bufs[L - 2], row_scale=self.norm2[lo:hi], relu=True, out=res[lo:hi])
                mo = self._outs[L - 1][j][0]
                after_block(lambda: ops.push_rows(res[lo:hi], mo, max_ctas=self.push_ctas))
            join()
            mark(f"gcn_aggregate{L - 1} || push")
        else:
            ops.spmm(self.graph, px.bufs[L - 2], row_scale=self.norm2, relu=True, multi_out=px.mos[L - 1])
            mark(f"gcn_aggregate{L - 1}+push")
        px.barrier(L - 1)
        mark(f"exchange{L}")
        agg = ops.spmm(self.graph, px.bufs[L - 1], row_scale=self.norm)
        mark(f"gcn_aggregate{L}")
        out = ops.dense(agg, *m.fc[-1])
        mark("classifier")
        return out
