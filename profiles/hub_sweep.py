"""Fused GAT / SpMM time vs hub threshold on the Reddit shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402

from gala_b200 import ops, synth  # noqa: E402

dev = "cuda:0"
n, e, f, K, c = synth.SHAPES["reddit"]
offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
X = torch.rand(n, K, device=dev) - 0.5
a = torch.randn(n, device=dev)
Y = torch.empty(n, K, device=dev)


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e_.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e_) / reps


g0 = ops.TiledGraph(offset, ids, n)
print(f"no plan: gat {t(lambda: ops.gat_forward(g0, a, a, X, out=Y)):.4f} spmm {t(lambda: ops.spmm(g0, X, out=Y)):.4f}")
for thr in (512, 1024, 2048, 4096, 8192, 1 << 30):
    g = ops.TiledGraph(offset, ids, n).build_plan(thr)
    print(f"thr {thr:>10d} hubs {g.plan.n_hub:6d}: gat {t(lambda: ops.gat_forward(g, a, a, X, out=Y)):.4f} "
          f"spmm {t(lambda: ops.spmm(g, X, out=Y)):.4f}")
