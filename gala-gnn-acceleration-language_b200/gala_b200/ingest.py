"""Dataset ingest straight to the device: the on-disk format of the reference
(scripts/Data/gala_export_npy.py; readers readSM_npy32 / readDM_npy tests/common.h:331-389,
called from the emitted main, src/codegen/common.h:531-614) without the host-side
std::vector double copy and without a host CSR:

  Adj_src.npy  uint32 [nrows, ncols, src_0 ... src_{E-1}]
  Adj_dst.npy  uint32 [dst_0 ... dst_{E-1}]
  Feat.npy float32 [N, F] . Lab.npy int64 [N, 1] . TnMsk / VlMsk / TsMsk.npy int32 [N, 1]

The .npy payloads are memory-mapped, staged through two pinned chunks and copied with the
copy engine while the next chunk is being paged in; the COO goes to gala_csr_from_coo on the
GPU (same (row, col)-sorted CSR as CSRCMatrix::build, values all 1 as set_all(1) leaves them).
"""
import os

import numpy as np
import torch

from . import formats

CHUNK_BYTES = 64 << 20


def _to_device(arr, device, out=None):
    """1-D/2-D C-contiguous numpy (usually a memmap) -> device tensor of the same dtype, chunked
    through two pinned staging buffers."""
    flat = arr.reshape(-1)
    tdt = torch.from_numpy(np.empty(0, flat.dtype)).dtype
    n = flat.shape[0]
    dst = out if out is not None else torch.empty(n, dtype=tdt, device=device)
    per = max(1, CHUNK_BYTES // flat.dtype.itemsize)
    stage = [torch.empty(min(per, max(n, 1)), dtype=tdt).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    used = [False, False]
    for i, lo in enumerate(range(0, n, per)):
        hi = min(n, lo + per)
        b = i & 1
        if used[b]:
            done[b].synchronize()          # the copy that last read this staging buffer
        stage[b][: hi - lo].numpy()[:] = flat[lo:hi]
        dst[lo:hi].copy_(stage[b][: hi - lo], non_blocking=True)
        done[b].record()
        used[b] = True
    torch.cuda.current_stream().synchronize()
    return dst.view(arr.shape)


def readDM_npy(filename, device="cuda:0"):
    """Dense matrix of any dtype the reference stores (float32 / int64 / int32) -> device tensor."""
    a = np.load(filename, mmap_mode="r")
    assert not np.isfortran(a), "row-major only (the reference ignores fortran_order too)"
    return _to_device(a, device)


def readSM_npy32(path, device="cuda:0"):
    """Adj_src.npy + Adj_dst.npy -> (nrows, ncols, offsets, ids, vals) on the device."""
    src = np.load(os.path.join(path, "Adj_src.npy"), mmap_mode="r")
    dst = np.load(os.path.join(path, "Adj_dst.npy"), mmap_mode="r")
    assert src.dtype == np.uint32 and dst.dtype == np.uint32, "readSM_npy32 reads uint32 (tests/common.h:343,351)"
    nrows, ncols = int(src[0]), int(src[1])
    nvals = int(dst.shape[0])
    assert src.shape[0] == nvals + 2
    assert max(nrows, ncols) < 2 ** 31 and nvals < 2 ** 31, "int32 index path (use the long build for more)"
    rows = _to_device(src[2:].view(np.int32), device)
    cols = _to_device(dst.view(np.int32), device)
    offsets, ids, _ = formats.csr_build(nrows, ncols, rows, cols)
    vals = torch.ones(nvals, dtype=torch.float32, device=device)       # adj->set_all(1)
    return nrows, ncols, offsets, ids, vals


def load_dataset(path, device="cuda:0"):
    """Everything the emitted main reads (common.h:531-614), device-resident."""
    nrows, ncols, offsets, ids, vals = readSM_npy32(path, device)
    out = {"nrows": nrows, "ncols": ncols, "offsets": offsets, "ids": ids, "vals": vals,
           "input_emb": readDM_npy(os.path.join(path, "Feat.npy"), device),
           "labels": readDM_npy(os.path.join(path, "Lab.npy"), device)}
    for key, name in (("train_mask", "TnMsk"), ("valid_mask", "VlMsk"), ("test_mask", "TsMsk")):
        out[key] = readDM_npy(os.path.join(path, name + ".npy"), device) != 0       # repopulate<DBL, DB>
    out["classes"] = int(out["labels"].max()) + 1
    return out
