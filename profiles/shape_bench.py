"""SpMM / fused-GAT timings over the feature widths and graph shapes BASELINE.json names."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402

from gala_b200 import ops, synth  # noqa: E402

dev = "cuda:0"


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


for shape, Ks in (("reddit", (1, 32, 41, 64, 128, 602)), ("products", (1, 32, 47, 100))):
    n, e, *_ = synth.SHAPES[shape]
    offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
    g = ops.TiledGraph(offset, ids, n).build_plan()
    deg = (offset[1:] - offset[:-1])
    print(f"{shape}: n={n} E={g.nvals} max_deg={int(deg.max())} hubs={g.plan.n_hub}", flush=True)
    for K in Ks:
        X = torch.rand(n, K, device=dev) - 0.5
        Y = torch.empty(n, K, device=dev)
        ms = t(lambda: ops.spmm(g, X, out=Y))
        bmin = 4 * (n + 1) + 4 * g.nvals + 8 * n * K
        bg = 4 * (n + 1) + 4 * g.nvals + 4 * g.nvals * K + 4 * n * K
        line = f"  K={K:4d} spmm {ms:8.4f} ms  B_min {bmin / ms / 1e6:7.0f} GB/s  B_gather {bg / ms / 1e6:7.0f} GB/s"
        if K % 4 != 0 and K > 4:
            # the default above re-pitches X once per call (gala_pad_rows_f32, included in its time); next to it:
            # rows gathered packed with narrow loads (round-1 behaviour) and a producer that wrote pitched rows itself
            ms_packed = t(lambda: ops.spmm(g, X, out=Y, pad=None), reps=3)
            Xp = ops.pad_rows(X)
            ms_pitched = t(lambda: ops.spmm(g, Xp, out=Y))
            line += f" | packed rows, narrow loads {ms_packed:8.4f} ms | caller-pitched rows {ms_pitched:8.4f} ms"
        if K in (32, 64, 128):
            a = torch.randn(n, device=dev)
            ms2 = t(lambda: ops.gat_forward(g, a, a, X, out=Y))
            line += f" | gat {ms2:8.4f} ms"
        print(line, flush=True)
    del g, offset, ids
    torch.cuda.empty_cache()
