"""One launch of each edge-tile kernel on the Reddit shape (for ncu -k regex:edge_tile)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402

from gala_b200 import ops, synth  # noqa: E402

dev = "cuda:0"
n, e, *_ = synth.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "reddit"]
offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
g = ops.TiledGraph(offset, ids, n).build_plan()
g.plan.tile_policy = 1      # GALA_TILES_ALWAYS
x = torch.randn(g.nvals, device=dev)
out = torch.empty_like(x)
rs = torch.empty(n, 1, device=dev)
for _ in range(2):
    ops.edge_rowsum(g, x, out=rs)
    ops.edge_softmax_fwd(g, x, out=out)
    ops.edge_softmax_bwd(g, out, x, out=out)
torch.cuda.synchronize()
