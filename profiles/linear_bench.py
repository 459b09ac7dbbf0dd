"""Times the tcgen05 dense transform against torch (cuBLAS fp32, TF32 off) on the layer-1 shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from gala_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"


def t(fn, reps=20):
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


for (M, K, N) in ((232965, 602, 32), (232965, 32, 41), (2449029, 100, 32), (13882495, 128, 32), (13882495, 32, 32),
                  (13882495, 32, 172)):
    X = torch.rand(M, K, device=dev) - 0.5
    W = torch.rand(N, K, device=dev) - 0.5
    b = torch.rand(N, device=dev)
    Y = torch.empty(M, N, device=dev)
    ms_t = t(lambda: F.linear(X, W, b))
    ms_g = t(lambda: ops.linear(X, W, b, out=Y))
    err = float((Y[:100000].double() - (X[:100000].double() @ W.double().t() + b.double())).norm() /
                (X[:100000].double() @ W.double().t() + b.double()).norm())
    gb = (M * K + M * N + N * K) * 4 / 1e9
    print(f"M={M} K={K} N={N}: torch fp32 {ms_t:.4f} ms | tcgen05 3xTF32 {ms_g:.4f} ms | {gb / ms_g * 1e3:.0f} GB/s of {gb:.3f} GB "
          f"| rel err vs fp64 {err:.1e}", flush=True)
    del X, W, Y
    torch.cuda.empty_cache()
