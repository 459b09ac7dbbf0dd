"""Streaming edge kernels: edge-parallel tiles (csrc/edge_tiles.cuh) vs the row-structured kernels, on the Reddit and
Products shapes.  GB/s = algorithmic bytes / CUDA-event time; HBM peak from MEASURED_PEAKS.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402

from gala_b200 import ops, synth  # noqa: E402

dev = "cuda:0"
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6558.1


def t(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


shapes = sys.argv[1:] or ["reddit", "products"]
for shape in shapes:
    n, e, *_ = synth.SHAPES[shape]
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
    g = ops.TiledGraph(offset, ids, n).build_plan()
    g.plan.tile_policy = 1      # GALA_TILES_ALWAYS
    gr = ops.TiledGraph(offset, ids, n).build_plan()
    gr.plan.tile_rows = None
    E = g.nvals
    print(f"{shape}: n={n} E={E} tiles={g.plan.n_tiles} mean degree {E / n:.1f}", flush=True)
    x = torch.randn(E, device=dev)
    da = torch.randn(E, device=dev)
    out = torch.empty(E, device=dev)
    rs = torch.empty(n, 1, device=dev)
    aL, aR = torch.randn(n, device=dev), torch.randn(n, device=dev)
    alpha = ops.edge_softmax_fwd(g, x)
    table = [
        ("edge_rowsum", 4 * (n + 1) + 4 * E + 4 * n, lambda G: ops.edge_rowsum(G, x, out=rs)),
        ("edge_scale_rows", 4 * (n + 1) + 8 * E + 4 * n, lambda G: ops.edge_scale_rows_(G, out, aL)),
        ("edge_softmax_fwd", 4 * (n + 1) + 8 * E, lambda G: ops.edge_softmax_fwd(G, x, out=out)),
        ("edge_softmax_bwd", 4 * (n + 1) + 12 * E, lambda G: ops.edge_softmax_bwd(G, alpha, da, out=out)),
    ]
    if os.environ.get("GALA_TILE_SKIP_BWD"):
        table = table[:3]
    for name, nbytes, fn in table:
        ms_t = t(lambda: fn(g))
        ms_r = t(lambda: fn(gr))
        print(f"  {name:18s} tiles {ms_t:7.4f} ms {nbytes / ms_t / 1e6:6.0f} GB/s ({nbytes / ms_t / 1e6 / PEAK:.2f} of HBM peak)"
              f" | row-structured {ms_r:7.4f} ms {nbytes / ms_r / 1e6:6.0f} GB/s ({nbytes / ms_r / 1e6 / PEAK:.2f})", flush=True)
    del g, gr, offset, ids, x, da, out, alpha
    torch.cuda.empty_cache()
