"""Seeded synthetic power-law graphs of the shapes BASELINE.json names.

The data pipeline of the reference symmetrises the graph, removes duplicates and adds
a self loop to every node (scripts/Data/gala_export_npy.py:73-74) and stores it as COO
sorted by (src, dst); edge values are all 1 (tests/common.h:363).  The generator below
reproduces that: Chung-Lu style endpoint sampling with node weight (i + shift)^-gamma,
random node relabelling (so that ids carry no locality), exact edge count.

numpy version for CPU-side tests, torch version (same algorithm, different RNG stream)
for building the large shapes directly on the GPU.
"""
import numpy as np

SHAPES = {
    # name: (nodes, directed edges incl. self loops, input features, hidden, classes)
    "cora": (2708, 10556 + 2708, 1433, 32, 7),
    "reddit": (232965, 114615892, 602, 32, 41),
    "products": (2449029, 123718280, 100, 32, 47),
    "papers": (111059956, 1615685872, 128, 32, 172),
}


def _weights(n, gamma, shift, xp):
    i = xp.arange(n, dtype=xp.float64)
    w = (i + shift) ** (-gamma)
    return w / w.sum()


def powerlaw_coo_np(n, n_edges, seed=0, gamma=0.85, shift=None):
    """Returns (src, dst) int32 arrays sorted by (src, dst): symmetric, duplicate-free,
    one self loop per node, exactly 2*floor((n_edges-n)/2) + n entries."""
    rng = np.random.default_rng(seed)
    shift = shift if shift is not None else max(1.0, n / 500.0)
    pairs_target = max((n_edges - n) // 2, 0)
    max_pairs = n * (n - 1) // 2
    pairs_target = min(pairs_target, max_pairs)
    cdf = np.cumsum(_weights(n, gamma, shift, np))
    perm = rng.permutation(n).astype(np.int64)
    keys = np.empty(0, np.int64)
    need = pairs_target
    while keys.shape[0] < pairs_target:
        m = int((need + 16) * 1.3)
        u = perm[np.minimum(np.searchsorted(cdf, rng.random(m)), n - 1)]
        v = perm[np.minimum(np.searchsorted(cdf, rng.random(m)), n - 1)]
        ok = u != v
        lo, hi = np.minimum(u[ok], v[ok]), np.maximum(u[ok], v[ok])
        keys = np.unique(np.concatenate([keys, lo * n + hi]))
        need = pairs_target - keys.shape[0]
    if keys.shape[0] > pairs_target:
        keys = keys[np.sort(rng.choice(keys.shape[0], pairs_target, replace=False))]
    lo, hi = keys // n, keys % n
    loops = np.arange(n, dtype=np.int64)
    src = np.concatenate([lo, hi, loops])
    dst = np.concatenate([hi, lo, loops])
    order = np.argsort(src * n + dst, kind="stable")
    return src[order].astype(np.int32), dst[order].astype(np.int32)


def coo_to_csr_np(n, src, dst):
    """(src, dst) sorted by (src, dst) -> (offset[n+1], ids) -- plain numpy, used to feed
    tests; the parity tests for CSR construction use the oracle / reference instead."""
    counts = np.bincount(src, minlength=n)
    offset = np.zeros(n + 1, np.int32)
    np.cumsum(counts, out=offset[1:])
    return offset, dst.astype(np.int32).copy()


def powerlaw_csr_torch(n, n_edges, seed=0, gamma=0.85, shift=None, device="cuda"):
    """Same construction on the GPU with torch ops (data synthesis only -- plumbing).
    Returns (offset int32[n+1], ids int32[E]) on `device`."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    shift = shift if shift is not None else max(1.0, n / 500.0)
    pairs_target = min(max((n_edges - n) // 2, 0), n * (n - 1) // 2)
    i = torch.arange(n, dtype=torch.float64, device=device)
    w = (i + shift) ** (-gamma)
    cdf = torch.cumsum(w / w.sum(), 0)
    perm = torch.randperm(n, generator=gen, device=device)
    keys = torch.empty(0, dtype=torch.int64, device=device)
    need = pairs_target
    while keys.numel() < pairs_target:
        m = int((need + 16) * 1.3)
        u = perm[torch.searchsorted(cdf, torch.rand(m, generator=gen, device=device,
                                                    dtype=torch.float64)).clamp_(max=n - 1)]
        v = perm[torch.searchsorted(cdf, torch.rand(m, generator=gen, device=device,
                                                    dtype=torch.float64)).clamp_(max=n - 1)]
        ok = u != v
        lo, hi = torch.minimum(u[ok], v[ok]), torch.maximum(u[ok], v[ok])
        del u, v, ok
        keys = torch.unique(torch.cat([keys, lo * n + hi]))
        del lo, hi
        need = pairs_target - keys.numel()
    if keys.numel() > pairs_target:
        drop = torch.randperm(keys.numel(), generator=gen, device=device)[:pairs_target]
        keys = keys[torch.sort(drop).values]
        del drop
    lo, hi = keys // n, keys % n
    del keys
    loops = torch.arange(n, dtype=torch.int64, device=device)
    full = torch.cat([lo * n + hi, hi * n + lo, loops * n + loops])
    del lo, hi, loops
    full = torch.sort(full).values
    src = full // n
    ids = (full % n).to(torch.int32)
    del full
    counts = torch.bincount(src, minlength=n)
    del src
    offset = torch.zeros(n + 1, dtype=torch.int32, device=device)
    offset[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return offset, ids


def powerlaw_multigraph_coo_torch(n, n_edges, seed=0, gamma=0.85, device="cuda", chunk=200_000_000):
    """Papers-scale variant: directed power-law multigraph (both endpoints drawn from the same rank
    distribution as above, node ids permuted; duplicates kept, one self loop per node, no symmetrisation -- at
    1.6 G edges torch.unique would need ~3x the memory, and CSRCMatrix::build keeps duplicates anyway).
    Ranks are drawn by inverting the continuous CDF of (i + shift)^-gamma in closed form: element-wise
    fp64 math only, so every GPU of a partitioned run synthesises bit-identical edges (a cumsum-based
    CDF is not reproducible: the device scan combines its partial sums in a timing-dependent order).
    Returns unsorted (rows int32[E], cols int32[E]); deterministic in (n, n_edges, seed)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    shift = max(1.0, n / 500.0)
    p = 1.0 - gamma
    lo_c, hi_c = shift ** p, (n + shift) ** p
    perm = torch.randperm(n, generator=gen, device=device, dtype=torch.int32)
    rows = torch.empty(n_edges, dtype=torch.int32, device=device)
    cols = torch.empty(n_edges, dtype=torch.int32, device=device)
    n_rand = max(n_edges - n, 0)          # the last n edges are the self loops the data pipeline adds
    for lo in range(0, n_rand, chunk):    # (scripts/Data/gala_export_npy.py:73-74): every row has degree >= 1
        m = min(chunk, n_rand - lo)
        for dst in (rows, cols):
            u = torch.rand(m, generator=gen, device=device, dtype=torch.float64)
            i = ((u * (hi_c - lo_c) + lo_c) ** (1.0 / p) - shift).floor_().clamp_(0, n - 1).long()
            dst[lo:lo + m] = perm[i]
            del u, i
    k = n_edges - n_rand
    loops = torch.arange(k, dtype=torch.int32, device=device)
    rows[n_rand:] = loops
    cols[n_rand:] = loops
    return rows, cols
