"""Runs each sparse kernel a few times on the Reddit-shape graph (for ncu captures):
   ncu --set full -k regex:gala -c 40 -o gpurun_out/prof_sweep python profiles/sweep.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402

from gala_b200 import ops, synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "reddit"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n, e, feats, K, classes = synth.SHAPES[shape]
dev = "cuda:0"
offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
g = ops.TiledGraph(offset, ids, n).build_plan()
gen = torch.Generator(device=dev)
gen.manual_seed(3)
X = torch.rand(n, K, generator=gen, device=dev) - 0.5
Z = torch.rand(n, K, generator=gen, device=dev) - 0.5
a = torch.randn(n, generator=gen, device=dev)
w = torch.rand(g.nvals, generator=gen, device=dev)
Y = torch.empty(n, K, device=dev)
ev = torch.empty(g.nvals, device=dev)
for _ in range(reps):
    ops.gat_forward(g, a, a, X, out=Y)
    ops.spmm(g, X, out=Y)
    ops.spmm(g, X, vals=w, out=Y)
    ops.sddmm(g, Z, X, out=ev)
    ops.sddvv(g, a, a, "add", out=ev)
    ops.edge_softmax_fwd(g, w, out=ev)
    ops.edge_rowsum(g, w)
torch.cuda.synchronize()
print("sweep done", shape, n, g.nvals, "hubs", g.plan.n_hub)
