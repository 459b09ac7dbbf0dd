"""Times the gather kernels for every library build under variants/ (one subprocess per build).
   python profiles/variant_bench.py            -> table on stdout"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gala-gnn-acceleration-language_b200")

CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from gala_b200 import ops, synth
n, e, f, K, c = synth.SHAPES["reddit"]
dev = "cuda:0"
offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
g = ops.TiledGraph(offset, ids, n).build_plan()
gen = torch.Generator(device=dev); gen.manual_seed(3)
X = torch.rand(n, K, generator=gen, device=dev) - 0.5
a = torch.randn(n, generator=gen, device=dev)
w = torch.rand(g.nvals, generator=gen, device=dev)
Y = torch.empty(n, K, device=dev); ev = torch.empty(g.nvals, device=dev)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e_.record(); torch.cuda.synchronize()
    return s.elapsed_time(e_) / reps
Xf = torch.rand(n, f, generator=gen, device=dev) - 0.5
Wf = (torch.rand(K, f, generator=gen, device=dev) - 0.5) * 0.1
Yf = torch.empty(n, K, device=dev)
print(json.dumps({"linear_602_32": t(lambda: ops.linear(Xf, Wf, out=Yf)),
                  "gat": t(lambda: ops.gat_forward(g, a, a, X, out=Y)),
                  "gat_dot": t(lambda: ops.gat_forward_dot(g, a, X[0].contiguous(), 0.1, X, out=Y)),
                  "gat_col": t(lambda: ops.gat_forward_col(g, a, 1.0, 0.1, X, reflect_in=X[1].contiguous(), reflect_out=X[2].contiguous(), relu=True, out=Y)) if hasattr(ops, "gat_forward_col") else None,
                  "spmm": t(lambda: ops.spmm(g, X, out=Y)),
                  "spmm_w": t(lambda: ops.spmm(g, X, vals=w, out=Y)),
                  "sddmm": t(lambda: ops.sddmm(g, X, X, out=ev)),
                  "sddvv": t(lambda: ops.sddvv(g, a, a, "add", out=ev)),
                  "rowsum": t(lambda: ops.edge_rowsum(g, w)),
                  "softmax_fwd": t(lambda: ops.edge_softmax_fwd(g, w, out=ev)),
                  "softmax_bwd": t(lambda: ops.edge_softmax_bwd(g, w, w, out=ev)),
                  "gat_bwd_att": t(lambda: ops.gat_backward_att(g, w, w, a, a)),
                  "gcn_scaled": t(lambda: ops.spmm(g, X, out=Y, row_scale=a, col_scale=a))}))
''' % PKG

for lib in sorted(glob.glob(os.path.join(PKG, "variants", "*.so"))) + [os.path.join(PKG, "libgala_b200.so")]:
    env = dict(os.environ, GALA_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
    print(os.path.basename(lib), line, flush=True)
