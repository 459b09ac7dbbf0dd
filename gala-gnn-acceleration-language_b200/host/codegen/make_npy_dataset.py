"""Writes a synthetic dataset in the on-disk format the generated programs read
(reference tests/common.h:331-389; scripts/Data/gala_export_npy.py): Adj_src.npy = uint32
[nrows, ncols, src...], Adj_dst.npy = uint32 [dst...], Feat.npy float32 [N,F], Lab.npy int64 [N,1],
TnMsk/VlMsk/TsMsk.npy int32 [N,1].

    python make_npy_dataset.py <out_dir> <nodes> <edges> <feats> <classes> [seed]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from gala_b200 import synth  # noqa: E402


def main():
    out, n, e, f, c = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    seed = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    os.makedirs(out, exist_ok=True)
    try:
        import torch
        use_gpu = torch.cuda.is_available() and e > 5_000_000
    except ImportError:
        use_gpu = False
    if use_gpu:
        offset, ids = synth.powerlaw_csr_torch(n, e, seed=seed, device="cuda")
        src = torch.repeat_interleave(torch.arange(n, device="cuda"), (offset[1:] - offset[:-1]).long())
        src, dst = src.cpu().numpy().astype(np.uint32), ids.cpu().numpy().astype(np.uint32)
    else:
        s, d = synth.powerlaw_coo_np(n, e, seed=seed)
        src, dst = s.astype(np.uint32), d.astype(np.uint32)
    rng = np.random.default_rng(seed + 1)
    np.save(os.path.join(out, "Adj_src.npy"), np.concatenate([np.array([n, n], np.uint32), src]))
    np.save(os.path.join(out, "Adj_dst.npy"), dst)
    np.save(os.path.join(out, "Feat.npy"), rng.uniform(-0.5, 0.5, (n, f)).astype(np.float32))
    lab = rng.integers(0, c, (n, 1)).astype(np.int64)
    lab[:c, 0] = np.arange(c)          # every class present: classes = max(label) + 1 (common.h:611-613)
    np.save(os.path.join(out, "Lab.npy"), lab)
    r = rng.random(n)
    for name, m in (("TnMsk", r < 0.1), ("VlMsk", (r >= 0.1) & (r < 0.2)), ("TsMsk", r >= 0.2)):
        np.save(os.path.join(out, name + ".npy"), m.astype(np.int32).reshape(n, 1))
    print("dataset", out, "nodes", n, "edges", dst.shape[0])


if __name__ == "__main__":
    main()
