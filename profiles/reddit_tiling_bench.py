"""Reddit shape at wide feature widths (X exceeds the L2): untiled vs col_tile(37000) row-major / segment-major."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402

from gala_b200 import formats, ops, synth  # noqa: E402

dev = "cuda:0"
n, e, *_ = synth.SHAPES["reddit"]
offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
ones = torch.ones(ids.numel(), device=dev)


def t(fn, reps=5):
    for _ in range(2):
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
for K in (32, 128, 256, 602):
    X = torch.rand(n, K, device=dev) - 0.5
    Y = torch.empty(n, K, device=dev)
    base = t(lambda: ops.spmm(g1, X, out=Y))
    print(f"reddit K={K} (X = {n * K * 4 / 1e6:.0f} MB): untiled {base:.3f} ms", flush=True)
    for T in (120000, 37000):
        tg = formats.ord_col_tiling(n, n, offset, ids, ones, T).build_plan()
        rm = t(lambda: ops.spmm(tg, X, out=Y, schedule="row_major"))
        sm = t(lambda: ops.spmm(tg, X, out=Y, schedule="segment_major"))
        print(f"   col_tile({T}) S={tg.segments}: row-major {rm:.3f} ms, segment-major {sm:.3f} ms", flush=True)
        del tg
