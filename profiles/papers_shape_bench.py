"""ogbn-papers100M shape on ONE B200 (BASELINE.json configs[4] is the partitioned run; this is its
single-GPU-equivalent datapoint from SURVEY.md section 8d): 111 059 956 nodes, 1 615 685 872 edges,
int32 indices, fp32 features.  Checks the 64-bit addressing of every kernel on the path at a scale where
E*K and N*K overflow 32 bits, and times format construction, SpMM (K = 32, 128) and the fused GAT layer.

The graph is a seeded directed power-law multigraph (no symmetrisation / de-duplication: at this size
torch.unique would need ~3x the memory; build() keeps duplicates anyway) built in chunks on the GPU."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402

from gala_b200 import formats, ops, synth  # noqa: E402

dev = "cuda:0"
n, e_total = synth.SHAPES["papers"][:2]
if len(sys.argv) > 1:                       # scale factor for dry runs
    f = float(sys.argv[1])
    n, e_total = int(n * f), int(e_total * f)


def timed(fn, reps=3, warm=1):
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


gen = torch.Generator(device=dev)
gen.manual_seed(0)
t0 = time.time()
rows, cols = synth.powerlaw_multigraph_coo_torch(n, e_total, seed=0, device=dev)
torch.cuda.synchronize()
print(f"papers shape: n={n} E={e_total}  COO synthesised in {time.time() - t0:.1f} s", flush=True)

# ---- format construction at scale (a8): COO -> CSR on the GPU, then the plan
torch.cuda.synchronize()
t0 = time.time()
offsets, ids, _ = formats.csr_build(n, n, rows, cols)
torch.cuda.synchronize()
t_build = time.time() - t0
rows64 = rows.long()
cnt = torch.bincount(rows64, minlength=n)
assert torch.equal(offsets[1:].long(), torch.cumsum(cnt, 0)), "row pointers"
del cnt, rows64
# sortedness inside rows: (row, col) keys non-decreasing
chk = torch.arange(0, e_total - 1, 997, device=dev)
rr = torch.searchsorted(offsets[1:].contiguous(), chk.int(), right=True)
rr2 = torch.searchsorted(offsets[1:].contiguous(), (chk + 1).int(), right=True)
same = rr == rr2
assert bool((ids[chk][same] <= ids[chk + 1][same]).all()), "columns sorted within rows"
del rows, cols, chk, rr, rr2, same
torch.cuda.empty_cache()
g = ops.TiledGraph(offsets, ids, n)
t0 = time.time()
g.build_plan()
torch.cuda.synchronize()
t_plan = time.time() - t0
deg = offsets[1:] - offsets[:-1]
print(f"  gala_csr_from_coo {t_build * 1e3:.0f} ms ({20 * e_total / t_build / 1e9:.0f} GB/s of 20E algorithmic bytes), "
      f"plan {t_plan * 1e3:.0f} ms, max degree {int(deg.max())}, hub rows {g.plan.n_hub}", flush=True)

# rows checked against torch in fp64: 2000 random + the 8 widest
pick = torch.cat([torch.randint(0, n, (2000,), generator=gen, device=dev), torch.topk(deg, 8).indices]).long()


def check_rows(Y, X, scale=None):
    worst = 0.0
    for r in pick.tolist():
        c = ids[int(offsets[r]):int(offsets[r + 1])].long()
        xs = X[c].double()
        if scale is not None:
            xs = xs * scale(r, c)[:, None]
        want = xs.sum(0)
        mag = xs.abs().sum(0).clamp_min(1e-30)
        worst = max(worst, float(((Y[r].double() - want).abs() / mag).max()))
    return worst


for K in (32, 128):
    X = torch.rand(n, K, device=dev) - 0.5
    Y = torch.empty(n, K, device=dev)
    ms = timed(lambda: ops.spmm(g, X, out=Y))
    bmin = 4 * (n + 1) + 4 * e_total + 8 * n * K
    err = check_rows(Y, X)
    print(f"  SpMM K={K:3d}: {ms:8.2f} ms   B_min {bmin / 1e9:6.1f} GB -> {bmin / ms / 1e6:6.0f} GB/s   "
          f"backward error vs fp64 on {pick.numel()} rows: {err:.2e}", flush=True)
    assert err < 1e-5
    if K == 32:
        aL = torch.randn(n, device=dev)
        aR = torch.randn(n, device=dev)
        ms = timed(lambda: ops.gat_forward(g, aL, aR, X, out=Y))

        def att(r, c):
            x = aL[r].double() + aR[c].double()
            x = torch.where(x > 0, x, 0.2 * x).exp().clamp(0, 1e12)
            return x / (1e-12 + x.sum())
        err = check_rows(Y, X, att)
        print(f"  fused GAT layer K=32: {ms:8.2f} ms   backward error vs fp64: {err:.2e}", flush=True)
        assert err < 1e-5
        del aL, aR
    del X, Y
    torch.cuda.empty_cache()
print("papers-shape OK")
