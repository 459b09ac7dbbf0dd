"""Layer-1 transform (K = 602, N = 32, attention epilogue) at the row counts one rank holds on 1 / 2 / 4 / 8 GPUs
(Reddit shape): which kernel form to dispatch when a rank has only ~1.5 tiles per SM."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402

from gala_b200 import ops  # noqa: E402

dev = "cuda:0"
gen = torch.Generator(device=dev).manual_seed(0)
K, N = 602, 32
W = (torch.rand(N, K, generator=gen, device=dev) - 0.5) * 0.1
b = torch.rand(N, generator=gen, device=dev)
aw = torch.rand(2, N, generator=gen, device=dev) - 0.5
for M in (232965, 116483, 58242, 29121):
    X = torch.rand(M, K, generator=gen, device=dev) - 0.5
    Y = torch.empty(M, N, device=dev)
    fn = lambda: ops.linear(X, W, b, att_w=aw, att_b=[0.1, 0.2], out=Y)   # noqa: E731
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 50
    print(f"M={M}: {ms * 1e3:.1f} us, {M * K * 4 / ms / 1e6:.0f} GB/s", flush=True)
