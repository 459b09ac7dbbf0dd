#!/usr/bin/env python
"""bench.py -- GAT 2-layer forward on a synthetic Reddit-shape graph (BASELINE.json
configs[1]) through the B200-native sparse path, plus per-kernel roofline numbers.

  python bench.py --gpus N --steps K --warmup W              # our arm
  python bench.py --impl reference --gpus N --steps K ...    # reference CPU arm

One "step" = one 2-layer GAT inference forward over the whole graph (dense transforms
via libtorch/cuBLAS as in the generated program + 2 launches of the fused GAT kernel).
Metric: ms per step (lower is better).  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))

# The reference arm is the reference's CPU path "with all the host threads it can use": torchrun exports
# OMP_NUM_THREADS=1 to its workers, which would silently time that arm on one core.  Only rank 0 runs it
# (the other ranks exit at once), so give it the machine back -- before any OpenMP runtime is loaded.
if "--impl" in sys.argv and "reference" in sys.argv and os.environ.get("LOCAL_RANK") is not None:
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "GAT 2-layer forward ms (Reddit-shape, 233K nodes / 114.6M edges / 602 feats / hidden 32)"
SHAPE = "reddit"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_graph(device, n, e, seed=0):
    from gala_b200 import synth

    return synth.powerlaw_csr_torch(n, e, seed=seed, device=device)


# ------------------------------------------------------------------------------- CPU arm
def cpu_gat_forward(orc, use_ref, n, offset, ids, X, m, rows=None):
    """The same 2-layer GAT forward on the host cores: dense parts torch-CPU, weighted
    aggregation through the reference's gSpMM (oracle/_ref, src/ops/aggregators.h:55-127)
    when available, edge kernels through the C restatement (they exist in the reference
    only as CUDA text).  `rows` bounds the sparse work to the first `rows` rows.
    Returns (logits, dense seconds, sparse seconds, edges processed)."""
    import torch.nn.functional as F

    rows = n if rows is None else rows
    off = np.ascontiguousarray(offset[:rows + 1])
    nv = int(off[-1])
    idv = np.ascontiguousarray(ids[:nv])
    g = orc.Tiled.from_csr(rows, n, off, idv)
    t = {"dense": 0.0, "sparse": 0.0}

    def dense(fn):
        t0 = time.perf_counter()
        r = fn()
        t["dense"] += time.perf_counter() - t0
        return r

    def gat_layer(att_src, feat, wl, wr):
        aL = dense(lambda: F.linear(att_src, *wl).reshape(-1).numpy())
        aR = dense(lambda: F.linear(att_src, *wr).reshape(-1).numpy())
        t0 = time.perf_counter()
        att = orc.sddvv(g, aL[:rows], aR, "add")
        att = orc.leaky_relu(att, 0.2)
        alpha, _ = orc.edge_softmax_fwd(g, att)
        fn_ = np.ascontiguousarray(feat.numpy())
        if use_ref:
            y = orc.ref_gspmm_wsum(rows, n, off, idv, alpha, fn_)
        else:
            y = orc.spmm(g, fn_, vals=alpha)
        t["sparse"] += time.perf_counter() - t0
        return torch.from_numpy(y)

    res = dense(lambda: F.linear(X, *m.fc0))
    h = torch.relu(gat_layer(res, res, m.efc0, m.efc1))
    if rows < n:   # layer 2 gathers rows of every column: pad the sample with layer-1 input rows
        h = torch.cat([h, res[rows:]], 0)
    tt = dense(lambda: F.linear(h, *m.fc1))
    agg = gat_layer(tt, h, m.efc2, m.efc3)
    out = dense(lambda: F.linear(agg, *m.fc1))
    return out, t["dense"], t["sparse"], nv


class _CpuModel:
    def __init__(self, m):
        for k in ("fc0", "efc0", "efc1", "fc1", "efc2", "efc3"):
            setattr(self, k, tuple(t.cpu() for t in getattr(m, k)))


def cpu_arm(n, e, feats, offset, ids, X_cpu, model, reps, warm, frac=None):
    """Times the CPU path; returns (ms per full-workload step, info dict)."""
    from oracle import orc

    orc.lib()
    use_ref = orc.have_ref()
    if use_ref:
        try:
            orc.ref()
        except OSError:
            use_ref = False
    cm = _CpuModel(model)
    cores = orc.lib().orc_num_threads()
    torch.set_num_threads(cores)
    nvals = int(offset[-1])
    # bounded sample: calibrate on 1/32 of the rows, then size the sample for ~10-20 s total
    r0 = max(n // 32, 1)
    t0 = time.perf_counter()
    _, td, ts, nv0 = cpu_gat_forward(orc, use_ref, n, offset, ids, X_cpu, cm, rows=r0)
    est_full = td + ts * (nvals / max(nv0, 1))
    budget = 20.0
    if frac is None:
        frac = min(1.0, budget / max(est_full * (reps + warm), 1e-3))
    rows = n if frac >= 0.999 else max(int(n * frac), 1)
    times = []
    for i in range(warm + reps):
        _, td, ts, nv = cpu_gat_forward(orc, use_ref, n, offset, ids, X_cpu, cm, rows=rows)
        if i >= warm:
            times.append((td + ts * (nvals / max(nv, 1))) * 1e3)
    info = {"cores": int(cores), "kind": "reference" if use_ref else "port",
            "sample": (f"first {rows} of {n} rows ({nv} of {nvals} edges) of the same graph, sparse time "
                       f"scaled by edge ratio, dense transforms in full; {reps} timed + {warm} warm-up passes; "
                       + ("weighted aggregation = reference gSpMM<wsumAgg> (oracle/_ref), "
                          if use_ref else "weighted aggregation = C port, ")
                       + "SDDVV/LeakyReLU/edge-softmax = C port (CUDA-only in the reference)")}
    return float(np.mean(times)), info


# ------------------------------------------------------------------------------- main
_REAL_STDOUT = None


def quiet_stdout():
    """Everything that libraries print on fd 1 (NCCL's version banner, ...) goes to stderr; the ONE
    JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=None, help="override graph size (debug)")
    ap.add_argument("--edges", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step from a CUDA graph")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU feature exchange: fused into the kernels over peer memory, or NCCL all-gather")
    ap.add_argument("--mode", default="reflected", choices=["folded", "folded_dot", "reflected", "reflected_fused", "fused", "literal", "dot"],
                    help="how the dense ops around the fused GAT kernel run (gala_b200/gat_model.py)")
    ap.add_argument("--dense", default="tcgen05", choices=["tcgen05", "torch"],
                    help="layer-1 feature transform: hand-written tcgen05 3xTF32 kernel or cuBLAS fp32 via torch")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel sweep (SpMM/SDDMM GB/s)")
    ap.add_argument("--no-generated", action="store_true",
                    help="skip the generated-program lines (GAT training epoch of the GALA-generated gala.cu, both generators)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from gala_b200 import synth

    n, e, feats, hidden, classes = synth.SHAPES[SHAPE]
    if args.nodes:
        n, e = args.nodes, args.edges or args.nodes * 50
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"synthetic {SHAPE}-shape power-law graph: {n} nodes, {e} directed edges incl. self loops, "
                          f"{feats} feats, hidden {hidden}, {classes} classes; 2-layer GAT inference forward; "
                          "schedule col_tile(370000) -> 1 column segment (tests/GALA-DSL/gat/Reddit/h100.txt)",
              "l2": "inputs exceed L2: 458 MB of column indices are streamed per layer (no explicit flush)"}

    if args.impl == "reference":
        if rank != 0:
            return
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        from gala_b200.gat_model import GAT2

        offset, ids = build_graph(dev, n, e)
        model = GAT2(feats, hidden, classes, dev)
        gen = torch.Generator(device=dev)
        gen.manual_seed(1)
        X = (torch.rand(n, feats, generator=gen, device=dev) - 0.5).cpu()
        ms, info = cpu_arm(n, e, feats, offset.cpu().numpy(), ids.cpu().numpy(), X, model,
                           max(args.steps, 1), args.warmup)
        info["value"] = ms
        info["unit"] = "ms"
        emit({"impl": "reference", "metric": METRIC, "value": ms, "unit": "ms",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                          "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": config, "cpu_baseline": info,
                          "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0})
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gala_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        import torch.distributed as dist

        # NCCL_DEBUG is left as the caller set it (the driver reads NCCL's INFO log to count ranks); whatever
        # NCCL prints on fd 1 already goes to stderr (quiet_stdout), so rank 0 still prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from gala_b200 import ops
    from gala_b200.gat_model import GAT2

    torch.backends.cuda.matmul.allow_tf32 = False   # libtorch default: the reference's Linear is fp32
    offset, ids = build_graph(dev, n, e)
    nvals = int(ids.numel())
    model = GAT2(feats, hidden, classes, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    X = torch.rand(n, feats, generator=gen, device=dev) - 0.5

    if world > 1:
        from gala_b200 import dist_gat

        runner = dist_gat.PartitionedGAT(model, offset, ids, n, rank, world, dev, exchange=args.exchange)
        X_in = X[runner.row_lo:runner.row_hi].contiguous()
        mode = args.mode if args.mode in ("reflected", "folded", "fused", "dot", "literal") else "folded"
        step_fn = lambda hook=None: runner.forward(X_in, hook, mode=mode)   # noqa: E731
        launches_per_step = 4 if (mode == "reflected" and runner.px is not None) else 3
        config["mode"] = mode
        config["parallelism"] = (f"1D row partition over {world} GPUs (nnz-balanced); exchange of hidden features: "
                                 + {"p2p-multicast": "fused into the producing kernels (multimem.st through NVLS multicast "
                                                     "into every GPU's buffer + device barrier)",
                                    "p2p": "fused into the producing kernels (peer stores over NVLink + device barrier)",
                                    "nccl": "NCCL all_gather_into_tensor"}[runner.exchange])
    else:
        g = ops.TiledGraph(offset, ids, n).build_plan()
        X_in = X
        mode = args.mode
        step_fn = lambda hook=None: model.forward(g, X_in, hook, mode=mode, dense=args.dense)   # noqa: E731
        launches_per_step = ({"folded": 5, "folded_dot": 5, "reflected": 4, "reflected_fused": 3, "fused": 3}.get(mode, 2)
                             if args.dense == "tcgen05" else 2)
        config["parallelism"] = "single GPU"
        config["dense"] = ("layer-1 X*W + attention projections: gala_linear_f32 (tcgen05 kind::tf32, 3xTF32)"
                           if args.dense == "tcgen05" else "cuBLAS fp32 through torch")
        config["mode"] = mode
        config["hub_rows"] = int(g.plan.n_hub)
        config["hub_threshold"] = int(g.plan.hub_threshold)

    ktimes = {}

    def hook(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        ktimes.setdefault(name, []).append((a, b))
        return out

    for _ in range(args.warmup):
        step_fn()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_beg.record()
    for _ in range(args.steps):
        out = step_fn(hook)
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = t_beg.elapsed_time(t_end)
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    eager_ms = total_ms / args.steps
    ms_per_step = eager_ms
    kern_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in ktimes.items()}

    # ---- the same K steps replayed from ONE CUDA graph (launch-bound once the sparse kernels shrink
    # with the GPU count): captured on a side stream, NCCL all-gathers included.  The eager region
    # above stays the source of the per-kernel event timings.
    graph_ms = None
    if not args.no_graph:
        try:
            cg = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step_fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(cg):
                out_graph = step_fn()
            for _ in range(3):
                cg.replay()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                cg.replay()
            b.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            graph_ms = a.elapsed_time(b) / args.steps
            ok = bool(torch.allclose(out_graph, out, rtol=1e-5, atol=1e-6))
            if world > 1:
                tt = torch.tensor([graph_ms, 0.0 if ok else 1.0], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                graph_ms, ok = float(tt[0].item()), float(tt[1].item()) == 0.0
            if not ok:
                graph_ms = None
        except Exception as ex:   # capture is an optimisation; the eager number stands without it
            sys.stderr.write(f"bench.py: CUDA-graph capture unavailable ({type(ex).__name__}: {ex})\n")
            graph_ms = None
    if graph_ms is not None and graph_ms < ms_per_step:
        ms_per_step = graph_ms
    clocks = sampler.stop() if rank == 0 else None   # sampled across both timed regions

    # ---- end-to-end through the public call with HOST buffers (pinned), copies inside the timed region
    X_host = X_in.cpu().pin_memory()
    out_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    X_stage = torch.empty_like(X_in)

    e2e_chunks = 8 if (world == 1 and args.dense == "tcgen05" and mode in ("folded", "folded_dot", "reflected")) else 1

    def e2e_step():
        if e2e_chunks > 1:
            # the public host-buffer call: chunked upload on a copy stream overlapped with the row-tiled transform
            model.forward_host(g, X_host, out_host, chunks=e2e_chunks, mode=mode, stage=X_stage)
            return
        X_stage.copy_(X_host, non_blocking=True)
        o = (runner.forward(X_stage, mode=mode) if world > 1 else model.forward(g, X_stage, mode=mode, dense=args.dense))
        out_host.copy_(o, non_blocking=True)

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(min(args.steps, 10), 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(e2e_steps):
        e2e_step()
    b.record()
    torch.cuda.synchronize()
    e2e_ms = a.elapsed_time(b) / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())

    # ---- parity of the partitioned forward against the single-GPU forward of the same model on the same
    # inputs (outside every timed region): every rank computes the whole-graph forward on its own GPU,
    # compares its slab of rows, and the norms are reduced -> one relative error for the whole output
    parity = None
    if world > 1:
        g_full = ops.TiledGraph(offset, ids, n).build_plan()
        want = model.forward(g_full, X, mode="literal", dense="torch")[runner.row_lo:runner.row_hi]
        got = runner.forward(X_in, mode=mode)
        tt = torch.stack([(got.double() - want.double()).pow(2).sum(), want.double().pow(2).sum()])
        dist.all_reduce(tt)
        parity = float((tt[0] / tt[1]).sqrt().item())
        del g_full, want, got
    else:
        # N = 1: the timed mode (folded / reflected / ...: own kernels everywhere) against the op-by-op forward of the
        # same model with cuBLAS dense parts, outside the timed region
        want = model.forward(g, X, mode="literal", dense="torch")
        got = model.forward(g, X, mode=mode, dense=args.dense)
        parity = float(((got.double() - want.double()).norm() / want.double().norm()).item())
        del want, got

    def finish():
        """Leave without tearing NCCL down: destroy_process_group() after a captured collective can
        wait forever on the watchdog; every rank has passed the final barrier, so just exit."""
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return

    # ---- roofline of the dominant kernel (fused GAT layer, inference: alpha not stored)
    # algorithmic bytes: 4(N+1) rowptr + 4E cols + 8N (aL, aR) + 8NK (X read once, Y written once)
    peak, peak_src = peaks()
    rows_local = (runner.row_hi - runner.row_lo) if world > 1 else n
    e_local = runner.local_nvals if world > 1 else nvals
    # mode "reflected" (gala_gat_forward_col_f32): aR rides in the last column of the gathered rows -- no [N] aR
    # vector in the byte model and no second 4-byte gather per edge
    col_mode = (mode in ("reflected", "reflected_fused") and world == 1) or (mode == "reflected" and world > 1 and runner.px is not None)
    alg_bytes = (4 * (rows_local + 1) + 4 * e_local + 4 * rows_local + (0 if col_mode else 4 * n)
                 + 4 * hidden * (n + rows_local))
    kms = float(np.mean([kern_ms[k] for k in ("gat_layer1", "gat_layer2") if k in kern_ms]))
    achieved = alg_bytes / (kms * 1e-3) / 1e9
    gather_bytes = (4 * (rows_local + 1) + 4 * e_local + (0 if col_mode else 4 * e_local) + 4 * hidden * e_local
                    + 4 * hidden * rows_local)
    roofline = {"kernel": ("gala::spmm_kernel<4,8,1,MODE_GAT_COL> (fused SDDVV+LeakyReLU+edge-softmax+SpMM, K=32, attention "
                           "term read from the last column of the gathered row)" if col_mode else
                           "gala::spmm_kernel<4,8,1,MODE_GAT> (fused SDDVV+LeakyReLU+edge-softmax+SpMM, K=32)"),
                "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": round(kms, 4),
                "no_reuse_gather_gbs": round(gather_bytes / (kms * 1e-3) / 1e9, 1),
                "note": "X (N*K*4 = 29.8 MB) is L2-resident; the 4*E*K gather bytes are served by L2/L1, not HBM "
                        "(SURVEY.md section 7), so frac against the compulsory-byte model is L2/LSU-limited"}
    if world == 1:
        # what the gather is actually bounded by: the L2->SM read bandwidth, measured here with the library's
        # probe on a buffer the size of X (L2-resident), next to the same probe on a 2 GiB buffer (HBM)
        try:
            l2 = ops.probe_read_gbs(int(n * hidden * 4) // 16 * 16, 400, dev)
            hbm = ops.probe_read_gbs(2 << 30, 4, dev)
            roofline["l2"] = {"measured_read_gbs": round(l2, 1), "probe_bytes": int(n * hidden * 4),
                              "gather_gbs": roofline["no_reuse_gather_gbs"],
                              "frac": round(roofline["no_reuse_gather_gbs"] / l2, 4),
                              "hbm_probe_read_gbs": round(hbm, 1)}
        except Exception as ex:   # measurement aid only
            roofline["l2"] = {"error": f"{type(ex).__name__}: {ex}"}
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and os.path.exists(tfile):
        try:
            roofline["traffic"] = json.load(open(tfile)).get(
                "gat_fused_col_k32_bytes_per_launch" if col_mode else "gat_fused_k32_bytes_per_launch")
        except (ValueError, OSError):
            pass

    line = {"metric": METRIC, "value": round(ms_per_step, 4), "unit": "ms", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
            "e2e": {"value": round(e2e_ms, 4), "unit": "ms",
                    "h2d_bytes_per_step": int(X_host.numel() * 4), "d2h_bytes_per_step": int(out_host.numel() * 4),
                    "h2d_chunks": e2e_chunks,
                    "call": ("GAT2.forward_host: pinned host features uploaded in row blocks on a copy stream, each block "
                             "consumed by the row-tiled transform as it lands; logits copied back to pinned host memory; consecutive "
                             "calls pipeline (the next upload runs under this step's aggregation layers and download)")
                            if e2e_chunks > 1 else "H2D copy, forward, D2H copy on one stream"},
            "roofline": roofline, "kernel_ms": {k: round(v, 4) for k, v in kern_ms.items()},
            "eager_ms_per_step": round(eager_ms, 4),
            "parity_rel_err": parity,
            "parity_against": ("single-GPU op-by-op forward (cuBLAS dense parts) of the same model and inputs, "
                               "all rows, norm-wise; bound 1e-5") if parity is not None else None,
            "graph_ms_per_step": round(graph_ms, 4) if graph_ms is not None else None}
    config["timed_region"] = ("K replays of the step captured in one CUDA graph" if graph_ms is not None and
                              graph_ms <= eager_ms else "K eager steps") + "; per-kernel events from the eager region"

    if not args.no_kernels and world == 1:
        line["kernels"] = kernel_sweep(g, n, nvals, hidden, peak, dev,
                                       (roofline.get("l2") or {}).get("measured_read_gbs"))

    # ---- BASELINE configs[1] is "GAT inference + training": the training epoch is measured on what GALA generates --
    # the gala.cu emitted for tests/GALA-DSL/gat/Reddit (train driver and inference driver), compiled once with the
    # reference's stock CUDAGenerator (its own kernels, sm_100a) and once with the retargeted generator (this library),
    # run here on the full-size synthetic Reddit-shape dataset in the on-disk .npy format.  Secondary keys; they need
    # the binaries host/codegen/build_models.sh builds in the authoring container (shipped with the tree).
    if world == 1 and not args.no_generated and not (args.nodes or args.edges):
        try:
            torch.cuda.synchronize()
            sys.path.insert(0, os.path.join(ROOT, "profiles"))
            import run_generated_full
            recs = run_generated_full.run("Reddit", ["gat_inference", "gat_train"])
            # BASELINE configs[2] (GIN / GraphSAGE training, Products shape) and configs[3] (GCN with neighbour sampling,
            # Products shape, gala_inference_sample path): the generated programs of those schedules, same two generators
            recs += run_generated_full.run("Products", ["gin_train_products", "sage_train_products",
                                                        "gcn_inference_products_sample20"])
            line["generated_programs"] = {
                "what": ("GALA-generated programs (100 epochs, Adam) on full-size synthetic datasets: 2-layer GAT on the Reddit "
                         "shape (BASELINE configs[1]), GIN / GraphSAGE training and GCN + aggrFn.sample(20) on the Products "
                         "shape (configs[2], configs[3]); mean forward ms and forward+backward+Adam ms as the program prints "
                         "them, start-up seconds (data load + format construction + H2D); generator ref = the reference's "
                         "own CUDA kernels compiled for sm_100a, b200 = this library through the retargeted generator"),
                "runs": recs}
        except Exception as ex:   # secondary measurement: never fails the bench line
            line["generated_programs"] = {"error": f"{type(ex).__name__}: {ex}"}

    if world == 1 and not args.no_cpu_baseline:
        ms, info = cpu_arm(n, e, feats, offset.cpu().numpy(), ids.cpu().numpy(), X.cpu(), model, 1, 0)
        info["value"] = round(ms, 2)
        info["unit"] = "ms"
        line["cpu_baseline"] = info
    emit(line)
    finish()


def time_op(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def kernel_sweep(g, n, nvals, K, peak, dev, l2_gbs=None):
    """SpMM / SDDMM / edge kernels alone on the same graph: ms, compulsory-byte GB/s, fraction."""
    from gala_b200 import ops

    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    X = torch.rand(n, K, generator=gen, device=dev) - 0.5
    Z = torch.rand(n, K, generator=gen, device=dev) - 0.5
    a = torch.rand(n, generator=gen, device=dev)
    w = torch.rand(nvals, generator=gen, device=dev)
    Y = torch.empty(n, K, device=dev)
    ev = torch.empty(nvals, device=dev)
    res = {}

    def rec(name, ms, nbytes, gather_bytes=None):
        gbs = nbytes / (ms * 1e-3) / 1e9
        res[name] = {"ms": round(ms, 4), "alg_GB": round(nbytes / 1e9, 4), "GBps": round(gbs, 1),
                     "frac_of_hbm_peak": round(gbs / peak, 4)}
        if gather_bytes is not None and l2_gbs:
            # gather kernels while X is L2-resident: bytes actually requested per launch (every edge gathers its
            # row) against the L2->SM read bandwidth measured by the probe in this run
            ggbs = gather_bytes / (ms * 1e-3) / 1e9
            res[name]["gather_GBps"] = round(ggbs, 1)
            res[name]["frac_of_l2_read"] = round(ggbs / l2_gbs, 4)

    rp = 4 * (n + 1)
    gth = rp + 4 * nvals + 4 * nvals * K + 4 * n * K          # cols + one K-wide row per edge + Y
    rec("spmm_k32_unweighted", time_op(lambda: ops.spmm(g, X, out=Y)), rp + 4 * nvals + 8 * n * K, gth)
    rec("spmm_k32_weighted", time_op(lambda: ops.spmm(g, X, vals=w, out=Y)), rp + 8 * nvals + 8 * n * K, gth + 4 * nvals)
    nrm = torch.rand(n, generator=gen, device=dev) + 0.1
    # GCN layer body of the generated model (norm*res -> aggregate -> norm*res -> relu, codegen/gala.cu:441-450)
    # as ONE launch with the fused row/col scale + ReLU epilogue
    rec("gcn_layer_aggregate_fused_k32", time_op(lambda: ops.spmm(g, X, out=Y, row_scale=nrm, col_scale=nrm, relu=True)),
        rp + 4 * nvals + 8 * n * K + 8 * n)
    # 2-layer GCN forward as generated (codegen/gala.cu:422-459) with every norm*res pass in an epilogue
    from gala_b200.gcn_model import GCN2
    feats = 602
    gcn = GCN2(feats, K, 41, dev).prepare(g)
    Xf = torch.rand(n, feats, generator=gen, device=dev) - 0.5
    ms = time_op(lambda: gcn.forward(g, Xf))
    res["gcn_2layer_forward"] = {"ms": round(ms, 4), "layer_ms": round(ms / 2, 4),
                                 "launches": "linear(tcgen05)+2 aggregations+cuBLAS classifier"}
    del Xf
    rec("sddmm_k32", time_op(lambda: ops.sddmm(g, Z, X, out=ev)), rp + 4 * nvals + 8 * n * K + 4 * nvals, gth + 4 * nvals)
    rec("sddvv_add", time_op(lambda: ops.sddvv(g, a, a, "add", out=ev)), rp + 4 * nvals + 8 * n + 4 * nvals)
    # the 4-byte gather B[col] touches one 128-byte line per edge: the SM's L1 serves ~1 line-wavefront per clock
    # (B300_MICROARCH.md "L1tex wavefront queue": rt_L1tex_wf ~ 1.0 cyc/wf), which bounds this kernel before HBM does
    props = torch.cuda.get_device_properties(dev)
    clk_hz = 1.965e9
    wf_ms = nvals / (props.multi_processor_count * clk_hz) * 1e3
    res["sddvv_add"]["l1_wavefront_bound_ms"] = round(wf_ms, 4)
    res["sddvv_add"]["frac_of_l1_wavefront_bound"] = round(wf_ms / res["sddvv_add"]["ms"], 4)
    rec("gat_fused_k32", time_op(lambda: ops.gat_forward(g, a, a, X, out=Y)), rp + 4 * nvals + 8 * n + 8 * n * K,
        gth + 4 * nvals)
    wR = (torch.rand(K, generator=gen, device=dev) - 0.5).contiguous()
    rec("gat_fused_dot_k32", time_op(lambda: ops.gat_forward_dot(g, a, wR, 0.1, X, out=Y)), rp + 4 * nvals + 4 * n + 8 * n * K,
        gth)
    # reflected basis: the attention scalar is the last element of the gathered row (gala_gat_forward_col_f32)
    rec("gat_fused_col_k32", time_op(lambda: ops.gat_forward_col(g, a, 1.0, 0.1, X, reflect_in=wR, reflect_out=wR, relu=True,
                                                                 out=Y)), rp + 4 * nvals + 4 * n + 8 * n * K, gth)
    rec("edge_softmax_fwd", time_op(lambda: ops.edge_softmax_fwd(g, w, out=ev)), rp + 8 * nvals)
    rec("edge_rowsum", time_op(lambda: ops.edge_rowsum(g, w)), rp + 4 * nvals + 4 * n)
    # K4 row scaling runs edge-parallel (csrc/edge_tiles.cuh: nnz-split tiles staged by bulk copies) on every shape
    rec("edge_scale_rows", time_op(lambda: ops.edge_scale_rows_(g, ev, a)), rp + 8 * nvals + 4 * n)
    ev2 = torch.rand(nvals, generator=gen, device=dev)
    rec("edge_softmax_bwd", time_op(lambda: ops.edge_softmax_bwd(g, w, ev2, out=ev)), rp + 12 * nvals)
    del ev2
    res["edge_kernels_products_shape"] = edge_tiles_products(peak, dev)
    # optional bf16 FEATURE STORAGE (not the headline: fp32 accumulation/outputs, results within 1e-2 of fp32,
    # BASELINE north_star "bf16 features within 1e-2"): half the gathered bytes on the L2-bound gather
    Xb = X.to(torch.bfloat16)
    rec("spmm_k32_bf16_features", time_op(lambda: ops.spmm_bf16(g, Xb, out=Y)), rp + 4 * nvals + 6 * n * K)
    rec("gat_fused_k32_bf16_features", time_op(lambda: ops.gat_forward_bf16(g, a, a, Xb, out=Y)),
        rp + 4 * nvals + 8 * n + 6 * n * K)
    return res


def edge_tiles_products(peak, dev):
    """The streaming edge kernels on the Products shape (mean degree 50: BASELINE configs[2]/[3]), where the default
    dispatch takes the edge-parallel tile kernels; the row-structured kernels (a plan without the tile table) beside."""
    from gala_b200 import ops, synth

    n, e, *_ = synth.SHAPES["products"]
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
    g = ops.TiledGraph(offset, ids, n).build_plan()
    gr = ops.TiledGraph(offset, ids, n).build_plan()
    gr.plan.tile_rows = None
    E = g.nvals
    x, da, out = torch.randn(E, device=dev), torch.randn(E, device=dev), torch.empty(E, device=dev)
    rs, a = torch.empty(n, 1, device=dev), torch.randn(n, device=dev)
    res = {"nodes": n, "edges": E, "tiles": int(g.plan.n_tiles)}
    for name, nbytes, fn in (
            ("edge_rowsum", 4 * (n + 1) + 4 * E + 4 * n, lambda G: ops.edge_rowsum(G, x, out=rs)),
            ("edge_scale_rows", 4 * (n + 1) + 8 * E + 4 * n, lambda G: ops.edge_scale_rows_(G, out, a)),
            ("edge_softmax_fwd", 4 * (n + 1) + 8 * E, lambda G: ops.edge_softmax_fwd(G, x, out=out)),
            ("edge_softmax_bwd", 4 * (n + 1) + 12 * E, lambda G: ops.edge_softmax_bwd(G, x, da, out=out))):
        ms_t, ms_r = time_op(lambda: fn(g)), time_op(lambda: fn(gr))
        res[name] = {"ms": round(ms_t, 4), "frac_of_hbm_peak": round(nbytes / (ms_t * 1e-3) / 1e9 / peak, 4),
                     "row_structured_ms": round(ms_r, 4)}
    del g, gr, offset, ids, x, da, out
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    main()
