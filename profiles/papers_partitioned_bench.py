"""BASELINE.json configs[4]: 3-layer GAT, 1-D row-partitioned, on the ogbn-papers100M shape
(111 059 956 nodes, 1 615 685 872 edges, 128 feats, hidden 32, 172 classes) at N = 1/2/4/8 B200.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
           --master-port 29544 profiles/papers_partitioned_bench.py [scale]

Every rank synthesises the same seeded COO, builds the CSR on its GPU (gala_csr_from_coo) and keeps the
slab of its rows (nnz-balanced).  Phase 1 checks the partitioned forward against the single-GPU op-by-op
forward on a small graph; phase 2 times the Papers-shape forward (CUDA events, max over ranks) with the
exchange fused into the kernels (multimem / peer stores) and with NCCL all-gather."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gala_b200 import dist_gat, formats, ops, synth  # noqa: E402
from gala_b200.gat_model import GATN  # noqa: E402
from gala_b200.gcn_model import GCNN  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
scale = float(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else 1.0
# --exchanges p2p,p2p-needed,nccl : which exchanges to time (default all three);  p2p-needed = rows stored only into the
# GPUs whose slabs reference them (predicated peer stores from the producing kernels' epilogues)
# a mode may carry "+pipeB": the row-block pipeline with B blocks (dist_gat.RowBlocks), e.g. p2p-needed+pipe4
EXCHANGES = ("p2p", "p2p-needed", "nccl")
PUSH_CTAS = 0
for i, a in enumerate(sys.argv):
    if a == "--exchanges":
        EXCHANGES = tuple(sys.argv[i + 1].split(","))
    if a == "--push-ctas":
        PUSH_CTAS = int(sys.argv[i + 1])


def split_mode(mode):
    """"p2p-needed+pipe4:fm@74" -> ("p2p-needed", 4 row blocks, at most 74 CTAs for the side-stream push kernels,
    phases f(irst) and m(iddle) pipelined -- default m; upper case = by copy engine)"""
    mode, _, ctas = mode.partition("@")
    mode, _, phases = mode.partition(":")
    base, _, pipe = mode.partition("+pipe")
    return base, int(pipe) if pipe else 0, int(ctas) if ctas else PUSH_CTAS, phases or "m"


def split_refl(mode):
    """"p2p-needed+refl+pipe8" -> ("p2p-needed+pipe8", True): rows exchanged in the reflected basis (GAT only)"""
    return mode.replace("+refl", ""), "+refl" in mode


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def max_over_ranks(ms):
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def timed(fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    dist.barrier()
    return max_over_ranks(a.elapsed_time(b) / steps)


def build(n, e, seed):
    rows, cols = synth.powerlaw_multigraph_coo_torch(n, e, seed=seed, device=dev)
    offset, ids, _ = formats.csr_build(n, n, rows, cols)
    del rows, cols
    torch.cuda.empty_cache()
    # every rank must hold the same graph (bit-identical synthesis): compare a checksum across ranks
    chk = torch.stack([offset.long().sum(), (ids.long() * 31 % 1000003).sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks synthesised different graphs"
    return offset, ids


# ---- phase 1: parity on a small graph --------------------------------------------------------------
n, e, dims = 60000, 4_000_000, [64, 32, 32, 41]
offset, ids = build(n, e, 3)
model = GATN(dims, dev, seed=0).host_biases()
X = torch.rand(n, dims[0], generator=torch.Generator(device=dev).manual_seed(1), device=dev) - 0.5
want = model.forward_literal(ops.TiledGraph(offset, ids, n).build_plan(), X)
worst = 0.0
g_small = ops.TiledGraph(offset, ids, n).build_plan()
err = float((model.forward(g_small, X, mode="reflected") - want).double().norm() / want.double().norm())
worst = max(worst, err)
say(f"parity [single GPU, reflected basis] 3-layer GAT: rel err {err:.2e}")
del g_small
for mode in (("nccl", "p2p", "p2p-needed", "p2p+pipe3:fml", "p2p-needed+pipe4", "p2p-needed+pipe3:FML", "p2p+refl",
              "p2p-needed+refl", "p2p+refl+pipe3:fml", "p2p-needed+refl+pipe4") if world > 1 else ("nccl",)):
    mode, refl = split_refl(mode)
    exchange, pipe, _, phases = split_mode(mode)
    part = dist_gat.RowPartition(offset, ids, n, rank, world)
    need = part.need_masks(ids, offset) if exchange == "p2p-needed" else None
    runner = dist_gat.PartitionedGATN(model, part, dev, exchange=exchange, need_mask=need, pipeline=pipe, phases=phases,
                                      reflected=refl)
    for _ in range(3):
        out_loc = runner.forward(X[part.row_lo:part.row_hi].contiguous())
    full = part.unpad(part.all_gather(out_loc))
    err = float((full - want).double().norm() / want.double().norm())
    worst = max(worst, err)
    say(f"parity [{runner.exchange}] 3-layer GAT on {world} rank(s): rel err {err:.2e}")
    del runner, part
gfull = ops.TiledGraph(offset, ids, n).build_plan()
gcn = GCNN(dims, dev, seed=2).prepare(gfull)
want = gcn.forward_literal(gfull, X)
for mode in (("nccl", "p2p", "p2p-needed", "p2p+pipe3:fml", "p2p-needed+pipe4", "p2p-needed+pipe3:FML") if world > 1 else ("nccl",)):
    exchange, pipe, _, phases = split_mode(mode)
    part = dist_gat.RowPartition(offset, ids, n, rank, world)
    need = part.need_masks(ids, offset) if exchange == "p2p-needed" else None
    runner = dist_gat.PartitionedGCNN(gcn, part, dev, exchange=exchange, need_mask=need, pipeline=pipe, phases=phases)
    for _ in range(3):
        out_loc = runner.forward(X[part.row_lo:part.row_hi].contiguous())
    full = part.unpad(part.all_gather(out_loc))
    err = float((full - want).double().norm() / want.double().norm())
    worst = max(worst, err)
    say(f"parity [{runner.exchange}] 3-layer GCN on {world} rank(s): rel err {err:.2e}")
    del runner, part
assert worst < 1e-5
del offset, ids, X, want, full, out_loc, gfull, gcn
torch.cuda.empty_cache()

# ---- phase 2: Papers shape --------------------------------------------------------------------------
n, e, feats, hidden, classes = synth.SHAPES["papers"]
n, e = int(n * scale), int(e * scale)
dims = [feats, hidden, hidden, classes]
offset, ids = build(n, e, 0)
part = dist_gat.RowPartition(offset, ids, n, rank, world)
need_mask = None
if world > 1 and any(split_mode(split_refl(m)[0])[0] == "p2p-needed" for m in EXCHANGES):
    need_mask = part.need_masks(ids, offset)
    say(f"needed-rows masks: rank 0's rows are referenced by {part.need_fraction:.1%} of the (row, peer) pairs")
del ids
torch.cuda.empty_cache()
model = GATN(dims, dev, seed=0).host_biases()
X_loc = torch.rand(part.rows, feats, device=dev) - 0.5
say(f"papers shape x{scale}: n={n} E={e}; rank 0 holds rows [{part.row_lo},{part.row_hi}) nnz {part.local_nvals}")
res = {"workload": f"3-layer GAT forward, papers100M shape x{scale}", "n_gpus": world, "nodes": n, "edges": e}
for mode in (EXCHANGES if world > 1 else ()):
    mode, refl = split_refl(mode)
    exchange, pipe, ctas, phases = split_mode(mode)
    runner = dist_gat.PartitionedGATN(model, part, dev, exchange=exchange, pipeline=pipe, push_ctas=ctas, phases=phases,
                                      need_mask=need_mask if exchange == "p2p-needed" else None, reflected=refl)
    if ctas:
        runner.exchange += f"@{ctas}"
    ms = timed(lambda: runner.forward(X_loc))
    res[f"ms_{runner.exchange}"] = round(ms, 3)
    say(f"  [{runner.exchange}] forward {ms:.2f} ms (max over {world} ranks)")
    if runner.px is not None:      # one more step with an event at every phase boundary (rank 0's view)
        marks = []

        def mark(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))
        dist.barrier()
        mark("start")
        runner.forward(X_loc, mark=mark)
        torch.cuda.synchronize()
        res[f"phases_ms_rank0_{runner.exchange}"] = {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 2) for i in range(1, len(marks))}
        say("  phases (rank 0):", res[f"phases_ms_rank0_{runner.exchange}"])
    del runner
    torch.cuda.empty_cache()
if world > 1 and "--needed" in sys.argv:
    # all-to-all-v of only the rows a slab references (dist_gat.NeededRowsPartition) instead of the full all-gather
    offset2, ids2 = build(n, e, 0)
    npart = dist_gat.NeededRowsPartition(offset2, ids2, n, rank, world)
    del offset2, ids2
    torch.cuda.empty_cache()
    runner = dist_gat.PartitionedGATN(model, npart, dev, exchange="nccl")
    ms = timed(lambda: runner.forward(X_loc))
    res["ms_needed_rows_a2a"] = round(ms, 3)
    res["needed_rows_fraction_rank0"] = round(npart.exchange_fraction(), 4)
    say(f"  [needed-rows all-to-all] forward {ms:.2f} ms; rank 0 receives {npart.exchange_fraction():.1%} of the all-gather rows")
    del runner, npart
    torch.cuda.empty_cache()
gcn = GCNN(dims, dev, seed=2)
for mode in (EXCHANGES if world > 1 and "--no-gcn" not in sys.argv else ()):
    if "+refl" in mode:          # GCN has no attention term to fold
        continue
    exchange, pipe, ctas, phases = split_mode(mode)
    runner = dist_gat.PartitionedGCNN(gcn, part, dev, exchange=exchange, pipeline=pipe, push_ctas=ctas, phases=phases,
                                      need_mask=need_mask if exchange == "p2p-needed" else None)
    if ctas:
        runner.exchange += f"@{ctas}"
    ms = timed(lambda: runner.forward(X_loc))
    res[f"gcn_ms_{runner.exchange}"] = round(ms, 3)
    say(f"  GCN [{runner.exchange}] forward {ms:.2f} ms (max over {world} ranks)")
    if runner.px is not None:
        marks = []

        def mark(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))
        dist.barrier()
        mark("start")
        runner.forward(X_loc, mark=mark)
        torch.cuda.synchronize()
        res[f"gcn_phases_ms_rank0_{runner.exchange}"] = {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 2) for i in range(1, len(marks))}
        say("  GCN phases (rank 0):", res[f"gcn_phases_ms_rank0_{runner.exchange}"])
    del runner
    torch.cuda.empty_cache()
if world == 1:
    g1 = ops.TiledGraph(part.offset, part.cols, part.rows, ncols=part.padded_n).build_plan()
    # [N, 172] logits = 71 GiB do not fit next to the 53 GiB of features on one GPU: the classifier writes
    # 8 M-row chunks into one re-used buffer (same bytes computed and written, not retained)
    sink = (8_000_000, torch.empty(8_000_000, classes, device=dev))
    ms = timed(lambda: model.forward(g1, X_loc, logits_chunk=sink))
    res["ms_single_gpu_model"] = round(ms, 3)
    phases = {}

    def hook(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        phases[name] = round(a.elapsed_time(b), 2)
        return out
    model.forward(g1, X_loc, hook=hook, logits_chunk=sink)
    res["phases_ms"] = phases
    say(f"  single-GPU GATN.forward {ms:.2f} ms; kernels: {phases}")
    ms = timed(lambda: model.forward(g1, X_loc, logits_chunk=sink, mode="reflected"))
    res["ms_single_gpu_model_reflected"] = round(ms, 3)
    phases = {}
    model.forward(g1, X_loc, hook=hook, logits_chunk=sink, mode="reflected")
    res["phases_ms_reflected"] = phases
    say(f"  single-GPU GATN.forward, reflected basis {ms:.2f} ms; kernels: {phases}")
    if "--no-gcn" in sys.argv:
        say(json.dumps(res))
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0)
    gcn.prepare(g1)

    def gcn_step():
        r = X_loc
        for i in range(gcn.L - 1):
            t = ops.linear(r, gcn.fc[i][0], gcn.fc[i][1], row_scale=gcn.norm)
            r = ops.spmm(g1, t, row_scale=gcn.norm2 if i == gcn.L - 2 else gcn.norm, relu=True)
        agg = ops.spmm(g1, r, row_scale=gcn.norm)
        for lo in range(0, agg.shape[0], sink[0]):       # logits in re-used chunks, as above
            hi = min(agg.shape[0], lo + sink[0])
            ops.dense(agg[lo:hi], gcn.fc[-1][0], gcn.fc[-1][1], out=sink[1][:hi - lo])
    ms = timed(gcn_step)
    res["gcn_ms_single_gpu_model"] = round(ms, 3)
    say(f"  single-GPU 3-layer GCN forward {ms:.2f} ms")
say(json.dumps(res))
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"papers_partitioned_n{world}.json"), "w") as f:
        json.dump(res, f)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
