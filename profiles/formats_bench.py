"""Times the device-side format construction on the Reddit shape (reference: host, 1.7 s CSR build on 8 cores)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402

from gala_b200 import formats, synth  # noqa: E402

dev = "cuda:0"
n, e, *_ = synth.SHAPES["reddit"]
offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
E = int(ids.numel())
rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int32), (offset[1:] - offset[:-1]).long())
perm = torch.randperm(E, device=dev)
r, c = rows[perm].contiguous(), ids[perm].contiguous()
ones = torch.ones(E, device=dev)
mask = (torch.rand(n, device=dev) < 0.1).to(torch.uint8)


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print(f"reddit shape: n={n} E={E}")
ms = t(lambda: formats.csr_build(n, n, r, c))
print(f"csr_build (COO shuffled -> CSR, 6 radix passes): {ms:8.3f} ms   {(12 * E + 8 * E) / ms / 1e6:7.0f} GB/s of the 12E+8E byte model")
ms = t(lambda: formats.csr_build(n, n, r, c, ones))
print(f"csr_build with values:                           {ms:8.3f} ms")
ms = t(lambda: formats.buildTranspose(n, n, offset, ids))
print(f"buildTranspose:                                  {ms:8.3f} ms")
ms = t(lambda: formats.ord_col_tiling(n, n, offset, ids, ones, 37000))
print(f"ord_col_tiling(37000) -> 7 segments:             {ms:8.3f} ms")
ms = t(lambda: formats.inplace_sample_graph_ab(n, offset, ids, ones, 20, 5, 7))
print(f"inplace_sample_graph_ab(20,5,7):                 {ms:8.3f} ms")
ms = t(lambda: formats.getMaskSubgraphs(n, n, offset, ids, ones, mask, 2), reps=2)
print(f"getMaskSubgraphs(2 layers, incl. transposes):    {ms:8.3f} ms")
