"""Products shape (X does not fit the 126 MB L2): SpMM row-major vs segment-major over column tiles."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402

from gala_b200 import formats, ops, synth  # noqa: E402

dev = "cuda:0"
n, e, *_ = synth.SHAPES["products"]
offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
ones = torch.ones(ids.numel(), device=dev)


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


g1 = ops.TiledGraph(offset, ids, n).build_plan()
for K in (32, 100):
    X = torch.rand(n, K, device=dev) - 0.5
    Y = torch.empty(n, K, device=dev)
    base = t(lambda: ops.spmm(g1, X, out=Y))
    print(f"products K={K}: untiled (shipped col_tile(10000000)) {base:.3f} ms", flush=True)
    ref = Y.clone()
    for T in (1300000, 650000, 330000, 170000):
        tg = formats.ord_col_tiling(n, n, offset, ids, ones, T).build_plan()
        rm = t(lambda: ops.spmm(tg, X, out=Y, schedule="row_major"))
        sm = t(lambda: ops.spmm(tg, X, out=Y, schedule="segment_major"))
        err = float((Y - ref).abs().max())
        print(f"   col_tile({T}) S={tg.segments}: row-major {rm:.3f} ms, segment-major {sm:.3f} ms  (max abs diff {err:.1e})", flush=True)
        del tg
