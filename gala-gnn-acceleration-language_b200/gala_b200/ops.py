"""Host-side operator layer over the C-ABI.

Two levels:
  * `TiledGraph` + snake_case ops (spmm, sddvv, edge_softmax_fwd, gat_forward, ...):
    one call = one launch of a hand-written sm_100a kernel;
  * `emitted` (see emitted.py): functions with the exact names, argument order and
    semantics of the wrappers GALA's code generator writes into gala.cu
    (src/codegen/cuda.h:441-952 of the reference), built on the level above.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _l

DEFAULT_HUB_THRESHOLD = 2048
LINEAR_MAX_N = 256        # widest output gala_linear_f32 accepts (GALA_ERR_UNSUPPORTED beyond); row epilogues: N <= 64
LINEAR_SMALL_MAX = 64     # gala_linear_small_f32: K <= 64 and N <= 64


class TiledGraph:
    """Column-tiled CSR on the device, in the layout the generated code keeps in
    global_offset_graph / global_columns_graph / global_bounds / global_segments
    (reference src/codegen/common.h:1694-1705, src/ops/tiling.h:222-283)."""

    def __init__(self, offsets, cols, nrows, ncols=None, bounds=None, segments=1, vals=None):
        assert offsets.dtype == torch.int32 and cols.dtype == torch.int32
        self.offsets, self.cols, self.vals = offsets.contiguous(), cols.contiguous(), vals
        self.nrows = int(nrows)
        self.ncols = int(ncols if ncols is not None else nrows)
        self.segments = int(segments)
        self.nvals = int(cols.numel())
        if bounds is None:
            assert self.segments == 1
            bounds = [0, self.nvals]
        if isinstance(bounds, torch.Tensor):
            bounds = bounds.cpu().numpy()
        self.bounds = np.ascontiguousarray(bounds, dtype=np.int32)  # host, as in the reference
        assert self.bounds.shape[0] == 2 * self.segments
        assert offsets.numel() == self.segments * (self.nrows + 1)
        # more than 64 segments: the kernels read the segment starts from a device copy of bounds
        self.bounds_dev = torch.from_numpy(self.bounds).to(cols.device) if self.segments > 64 else None
        self.c = _l.GalaGraph(offsets=self.offsets.data_ptr(), cols=self.cols.data_ptr(),
                              bounds=self.bounds.ctypes.data, nrows=self.nrows, ncols=self.ncols,
                              segments=self.segments, nvals=self.nvals,
                              bounds_dev=self.bounds_dev.data_ptr() if self.bounds_dev is not None else None)
        self.plan = None
        self._plan_ws = None

    @property
    def device(self):
        return self.cols.device

    def build_plan(self, hub_threshold=DEFAULT_HUB_THRESHOLD):
        """Find the hub rows once; every op then runs them on a whole CTA."""
        lib = _l.load()
        nbytes = lib.gala_plan_workspace_bytes(C.byref(self.c))
        self._plan_ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
        plan = _l.GalaPlan()
        _l.check(lib.gala_plan_build(C.byref(self.c), hub_threshold, _l.ptr(self._plan_ws), nbytes,
                                     C.byref(plan), _l.stream_ptr()))
        self.plan = plan
        return self

    def _p(self):
        return C.byref(self.plan) if self.plan is not None else None


def _f32(t):
    """Contiguous fp32 CUDA view of `t`.  A non-contiguous argument is copied: the CALLER must bind the
    result to a local that outlives the launch (never `_l.ptr(_f32(x))` inline -- the copy would be freed
    when ptr() returns and the caching allocator could hand the same block to the next argument)."""
    assert t.dtype == torch.float32 and t.is_cuda
    return t.contiguous()


SCHEDULES = {"auto": 0, "row_major": 1, "segment_major": 2}


def _rows(t):
    """(tensor, K, row pitch in elements) of a dense fp32 operand: packed [n, K] / [n], or a row-pitched 2-D
    view (stride(1) == 1, e.g. padded[:, :K]) that is passed to the kernels as it is (ABI v2: ldx / ldy)."""
    assert t.dtype == torch.float32 and t.is_cuda
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1] and t.shape[0] > 1:
        return t, t.shape[1], t.stride(0)
    t = t.contiguous()
    K = t.shape[1] if t.dim() == 2 else 1
    return t, K, K


def _dptr(t):
    """Device pointer of a (possibly row-pitched) tensor."""
    return C.c_void_p(t.data_ptr()) if t is not None else None


def pad_pitch(K):
    """Row pitch (elements) the gather likes for a width that is not a multiple of 4: 16-byte aligned rows, and for
    rows longer than 64 bytes a multiple of 64 bytes, so that a row never straddles one more 128-byte line than it
    has to (measured on B200: the gather costs ~1.2 clk per line touched + ~0.25 clk per 32-byte sector per SM;
    K = 41 at pitch 44 touches 2.25 lines on average, at pitch 48 always 2)."""
    return (K + 15) // 16 * 16 if K > 16 else (K + 3) // 4 * 4


def pad_rows(X, ld=None):
    """X [n, K] -> a view [n, K] of a fresh [n, ld] buffer (ld = pad_pitch(K) by default), padding zeroed
    (gala_pad_rows_f32): every row then starts 16-byte aligned and is gathered with 128-bit loads."""
    X, K, ld_in = _rows(X)
    n = X.shape[0]
    ld = pad_pitch(K) if ld is None else ld
    buf = torch.empty((n, ld), dtype=torch.float32, device=X.device)
    _l.check(_l.load().gala_pad_rows_f32(_dptr(X), n, K, ld_in, _l.ptr(buf), ld, _l.stream_ptr()))
    return buf[:, :K]


def _gather_operand(X, pad):
    """The dense operand whose rows are gathered.  pad="auto": packed rows whose width is not a multiple of 4
    (41 / 47 classes, 602 features) are re-pitched once -- N*K*8 bytes against E*K*4 gathered."""
    X, K, ld = _rows(X)
    # (wide rows -- K = 602 -- are left packed: their gather is HBM-bound and the 2 x 32 x 4-accumulator tile of the
    #  64-bit path covers them with fewer idle lanes than the 128-bit one; measured 24.2 vs 29.9 ms on the Reddit shape)
    if pad == "auto" and 4 < K <= 256 and (ld % 4 != 0 or X.data_ptr() % 16 != 0):
        X = pad_rows(X)
        ld = X.stride(0)
    return X, K, ld


def spmm(g, X, vals=None, out=None, row_scale=None, col_scale=None, accumulate=False, relu=False, schedule="auto",
         pad="auto", multi_out=None):
    """Y = A @ X (optionally weighted / scaled / accumulated / ReLU'd).  One launch, or -- for
    column-tiled graphs whose feature matrix exceeds the L2 -- one launch per column segment.
    X and out may be row-pitched views; pad=None gathers packed odd-width rows as they are (scalar loads)."""
    X, K, ldx = _gather_operand(X, pad)
    if multi_out is not None:      # the finished rows are pushed into every GPU's gathered buffer instead of `out`
        ep = _l.GalaEpilogue(row_scale=row_scale.data_ptr() if row_scale is not None else None,
                             col_scale=col_scale.data_ptr() if col_scale is not None else None,
                             accumulate=0, relu=int(relu), schedule=SCHEDULES["row_major"], ldx=ldx, ldy=K,
                             multi_out=C.addressof(multi_out))
        _l.check(_l.load().gala_spmm_f32(C.byref(g.c), _l.ptr(vals), _dptr(X), K, None, C.byref(ep), g._p(),
                                         _l.stream_ptr()))
        return None
    if out is None:
        out = torch.empty((g.nrows, K), dtype=torch.float32, device=X.device)
        assert not accumulate, "accumulate needs a caller-provided output"
    assert out.dim() == 1 or out.stride(-1) == 1
    ldy = out.stride(0) if out.dim() == 2 and out.shape[0] > 1 else K
    ep = _l.GalaEpilogue(row_scale=row_scale.data_ptr() if row_scale is not None else None,
                         col_scale=col_scale.data_ptr() if col_scale is not None else None,
                         accumulate=int(accumulate), relu=int(relu), schedule=SCHEDULES[schedule],
                         ldx=ldx, ldy=ldy)
    _l.check(_l.load().gala_spmm_f32(C.byref(g.c), _l.ptr(vals), _dptr(X), K, _dptr(out),
                                     C.byref(ep), g._p(), _l.stream_ptr()))
    return out


def spmm_sampled(g, X, nsamples, ra, rb, vals=None, out=None, accumulate=False, pad="auto"):
    X, K, ldx = _gather_operand(X, pad)
    if out is None:
        out = torch.empty((g.nrows, K), dtype=torch.float32, device=X.device)
    ldy = out.stride(0) if out.dim() == 2 and out.shape[0] > 1 else K
    _l.check(_l.load().gala_spmm_sampled_f32(C.byref(g.c), _l.ptr(vals), _dptr(X), K, _dptr(out),
                                             nsamples, ra, rb, int(accumulate), ldx, ldy, _l.stream_ptr()))
    return out


def edge_rowsum(g, vals, seed=1e-12, out=None):
    vals = _f32(vals)
    if out is None:
        out = torch.empty((g.nrows, 1), dtype=torch.float32, device=vals.device)
    _l.check(_l.load().gala_edge_rowsum_f32(C.byref(g.c), _l.ptr(vals), _l.ptr(out), seed,
                                            g._p(), _l.stream_ptr()))
    return out


def edge_scale_rows_(g, vals, rowval):
    """In place: vals[e] *= rowval[row(e)]."""
    rowval = _f32(rowval)
    _l.check(_l.load().gala_edge_scale_rows_f32(C.byref(g.c), _l.ptr(vals), _l.ptr(rowval),
                                                g._p(), _l.stream_ptr()))
    return vals


def sddvv(g, A, B, op="add", leaky_slope=1.0, out=None):
    A, B = _f32(A), _f32(B)
    if out is None:
        out = torch.empty(g.nvals, dtype=torch.float32, device=A.device)
    _l.check(_l.load().gala_sddvv_f32(C.byref(g.c), _l.ptr(A), _l.ptr(B), _l.ptr(out),
                                      0 if op == "add" else 1, leaky_slope, g._p(), _l.stream_ptr()))
    return out


def sddmm(g, A, B, out=None):
    A, B = _f32(A), _f32(B)
    K = A.shape[1] if A.dim() == 2 else 1
    if out is None:
        out = torch.empty(g.nvals, dtype=torch.float32, device=A.device)
    _l.check(_l.load().gala_sddmm_f32(C.byref(g.c), _l.ptr(A), _l.ptr(B), K, _l.ptr(out), g._p(),
                                      _l.stream_ptr()))
    return out


def edge_softmax_fwd(g, x, out=None, recip=None):
    x = _f32(x)
    if out is None:
        out = torch.empty_like(x)
    _l.check(_l.load().gala_edge_softmax_fwd_f32(C.byref(g.c), _l.ptr(x), _l.ptr(out),
                                                 _l.ptr(recip), g._p(), _l.stream_ptr()))
    return out


def edge_softmax_bwd(g, alpha, dalpha, out=None):
    alpha, dalpha = _f32(alpha), _f32(dalpha)
    if out is None:
        out = torch.empty_like(alpha)
    _l.check(_l.load().gala_edge_softmax_bwd_f32(C.byref(g.c), _l.ptr(alpha),
                                                 _l.ptr(dalpha), _l.ptr(out), g._p(),
                                                 _l.stream_ptr()))
    return out


def gat_backward_att(g, alpha, dalpha, aL, aR, slope=0.2, out=None):
    """d(attenL) = d(attenR) of one GAT layer from d(alpha): softmax backward + LeakyReLU backward +
    row sum in one kernel (gala_gat_backward_att_f32)."""
    alpha, dalpha, aL, aR = _f32(alpha), _f32(dalpha), _f32(aL), _f32(aR)
    if out is None:
        out = torch.empty((g.nrows, 1), dtype=torch.float32, device=alpha.device)
    _l.check(_l.load().gala_gat_backward_att_f32(C.byref(g.c), _l.ptr(alpha), _l.ptr(dalpha),
                                                 _l.ptr(aL), _l.ptr(aR), slope, _l.ptr(out),
                                                 g._p(), _l.stream_ptr()))
    return out


def gat_forward(g, aL, aR, X, slope=0.2, relu=False, out=None, alpha_out=None):
    """Fused SDDVV + LeakyReLU + edge-softmax + weighted SpMM (one pass over the edges)."""
    aL, aR = _f32(aL), _f32(aR)
    X, K, ldx = _gather_operand(X, "auto")
    if out is None:
        out = torch.empty((g.nrows, K), dtype=torch.float32, device=X.device)
    if ldx != K or not out.is_contiguous():     # row-pitched operands go through the _ex entry point (ldx / ldy)
        ep = _l.GalaDenseEpilogue(ldx=ldx, ldy=out.stride(0) if out.shape[0] > 1 else K)
        _l.check(_l.load().gala_gat_forward_ex_f32(C.byref(g.c), _l.ptr(aL), _l.ptr(aR), _dptr(X), K, slope,
                                                   _dptr(out), _l.ptr(alpha_out), int(relu), C.byref(ep), g._p(),
                                                   _l.stream_ptr()))
        return out
    _l.check(_l.load().gala_gat_forward_f32(C.byref(g.c), _l.ptr(aL), _l.ptr(aR),
                                            _l.ptr(X), K, slope, _l.ptr(out), _l.ptr(alpha_out),
                                            int(relu), g._p(), _l.stream_ptr()))
    return out


def gat_forward_dot(g, aL, wR, bR, X, slope=0.2, relu=False, out=None, alpha_out=None):
    """Fused GAT layer with aR[j] = dot(X[j,:], wR) + bR recomputed inside the kernel from the
    gathered rows (one random gather per edge instead of two).  Falls back to gat_forward with a
    materialised aR when the shape is outside the kernel's range (K % 4 != 0 or K > 32)."""
    X, aL = _f32(X), _f32(aL)
    K = X.shape[1]
    wR = _f32(wR.reshape(-1))
    if K % 4 != 0 or K > 32 or X.data_ptr() % 16 != 0:
        aR = (X @ wR + bR).contiguous()
        return gat_forward(g, aL, aR, X, slope, relu, out, alpha_out)
    if out is None:
        out = torch.empty((g.nrows, K), dtype=torch.float32, device=X.device)
    _l.check(_l.load().gala_gat_forward_dot_f32(C.byref(g.c), _l.ptr(aL), _l.ptr(wR), float(bR),
                                                _l.ptr(X), K, slope, _l.ptr(out), _l.ptr(alpha_out),
                                                int(relu), g._p(), _l.stream_ptr()))
    return out


def reflection(w):
    """Householder vector of gala_reflection_f32 for the projection weights w ([K] tensor, any device): returns
    (v [K] on w's device, sR float) with H = I - 2 v v^T, H e_{K-1} = -sign(w[K-1]) w/|w| and (X H)[:, K-1] * sR = X w.
    Host arithmetic (K <= a few hundred); call it when the weights change, never inside a step."""
    wh = w.detach().reshape(-1).to("cpu", torch.float32).contiguous()
    K = wh.numel()
    w_arr = (C.c_float * K)(*wh.tolist())
    v_arr = (C.c_float * K)()
    sR = C.c_float()
    _l.check(_l.load().gala_reflection_f32(w_arr, K, v_arr, C.byref(sR)))
    return torch.tensor(list(v_arr), dtype=torch.float32, device=w.device), float(sR.value)


def reflect(T, v, dim=-1):
    """T H with H = I - 2 v v^T applied along `dim` (host-side folding of the reflection into weights / biases;
    fp64 inside so that the folded weights carry no extra rounding)."""
    Td, vd = T.double(), v.double()
    if dim in (-1, T.dim() - 1):
        return (Td - 2.0 * (Td @ vd).unsqueeze(-1) * vd).float().contiguous()
    assert dim == 0
    return (Td - 2.0 * vd.unsqueeze(-1) * (vd @ Td).unsqueeze(0)).float().contiguous()


def _dense_epilogue(g, dev, att_w, att_b, cls_wT, cls_b, multi_out, att_multi_out):
    """gala_dense_epilogue_t + the tensors it writes (att [2, nrows], cls [nrows, C]) + the tensors it must keep alive."""
    ep = _l.GalaDenseEpilogue()
    keep = []
    if multi_out is not None:
        ep.multi_out = C.pointer(multi_out)
    if att_multi_out is not None:
        ep.att_multi_out = C.pointer(att_multi_out)
    att = cls = None
    if att_w is not None:
        att = torch.empty((2, g.nrows), dtype=torch.float32, device=dev)
        att_w = _f32(att_w)
        keep.append(att_w)
        ep.att_w = att_w.data_ptr()
        ep.att_b[0], ep.att_b[1] = float(att_b[0]), float(att_b[1])
        ep.att_out = att.data_ptr()
    if cls_wT is not None:
        cls_wT = _f32(cls_wT)
        keep.append(cls_wT)
        cls = torch.empty((g.nrows, cls_wT.shape[1]), dtype=torch.float32, device=dev)
        ep.cls_wT = cls_wT.data_ptr()
        ep.cls_b = cls_b.data_ptr() if cls_b is not None else None
        ep.cls_out = cls.data_ptr()
        ep.cls_n = cls_wT.shape[1]
    return ep, att, cls, keep


def gat_forward_col(g, aL, sR, bR, X, slope=0.2, relu=False, reflect_in=None, reflect_out=None, out=None,
                    alpha_out=None):
    """Fused GAT layer over features in a reflected basis whose last column carries the right-hand attention term:
    aR[j] = sR * X[j, K-1] + bR (gala_gat_forward_col_f32).  reflect_in / reflect_out: [K] Householder vectors
    applied to each finished output row before / after the ReLU.  K in {4, 8, 16, 32}."""
    X, aL = _f32(X), _f32(aL)
    K = X.shape[1]
    rin = _f32(reflect_in) if reflect_in is not None else None
    rout = _f32(reflect_out) if reflect_out is not None else None
    if out is None:
        out = torch.empty((g.nrows, K), dtype=torch.float32, device=X.device)
    _l.check(_l.load().gala_gat_forward_col_f32(C.byref(g.c), _l.ptr(aL), float(sR), float(bR), _l.ptr(X), K,
                                                slope, _l.ptr(out), _l.ptr(alpha_out), int(relu), _l.ptr(rin),
                                                _l.ptr(rout), None, g._p(), _l.stream_ptr()))
    return out


def gat_forward_col_ex(g, aL, sR, bR, X, slope=0.2, relu=False, reflect_in=None, reflect_out=None, out=None,
                       att_w=None, att_b=None, cls_wT=None, cls_b=None, want_y=True, multi_out=None):
    """gat_forward_col + the dense epilogue of gat_forward_ex on every FINAL output row (after reflect_out): att_w
    [2, K] are the next layer's projections expressed in the basis the rows leave in.  Returns (Y, att, cls)."""
    X, aL = _f32(X), _f32(aL)
    K = X.shape[1]
    rin = _f32(reflect_in) if reflect_in is not None else None
    rout = _f32(reflect_out) if reflect_out is not None else None
    if multi_out is not None:
        want_y = False
    if want_y and out is None:
        out = torch.empty((g.nrows, K), dtype=torch.float32, device=X.device)
    ep, att, cls, keep = _dense_epilogue(g, X.device, att_w, att_b, cls_wT, cls_b, multi_out, None)
    _l.check(_l.load().gala_gat_forward_col_f32(C.byref(g.c), _l.ptr(aL), float(sR), float(bR), _l.ptr(X), K,
                                                slope, _l.ptr(out) if want_y else None, None, int(relu), _l.ptr(rin),
                                                _l.ptr(rout), C.byref(ep), g._p(), _l.stream_ptr()))
    del keep
    return out, att, cls


def make_multi_out(bases, multicast_base=None, need_mask=None):
    """bases: per-GPU addresses (ints) of this rank's slab inside every peer-mapped gathered buffer.
    need_mask: uint8 device tensor [rows of this rank], bit q = GPU q references the row (peer stores only)."""
    mo = _l.GalaMultiOut()
    for q, b in enumerate(bases):
        mo.base[q] = int(b)
    mo.multicast_base = int(multicast_base) if multicast_base else None
    mo.count = len(bases)
    mo.need_mask = int(need_mask.data_ptr()) if need_mask is not None else None
    return mo


def linear(X, W, bias=None, relu=False, att_w=None, att_b=None, out=None, row_scale=None, multi_out=None,
           att_multi_out=None):
    """Y = X @ W.T + bias on the tensor cores (tcgen05 kind::tf32, 3xTF32 error compensation).
    With att_w [2,N] / att_b (two floats) also returns att [2,M] = the two attention
    projections of the pre-activation output rows (fused epilogue)."""
    X, W = _f32(X), _f32(W)
    M, K = X.shape
    N = W.shape[0]
    assert W.shape[1] == K
    if out is None and multi_out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=X.device)
    att = None
    ab, ab_dev = None, 0
    if att_w is not None:
        att_w = _f32(att_w)
        att = torch.empty((2, M), dtype=torch.float32, device=X.device)
        if isinstance(att_b, torch.Tensor) and att_b.is_cuda:      # trained biases stay on the device
            att_b = _f32(att_b)
            ab, ab_dev = C.c_void_p(att_b.data_ptr()), 1
        else:
            ab_host = (C.c_float * 2)(float(att_b[0]), float(att_b[1]))
            ab = C.cast(ab_host, C.c_void_p)
    _l.check(_l.load().gala_linear_f32(_l.ptr(X), M, K, _l.ptr(W), _l.ptr(bias), N,
                                       _l.ptr(out) if out is not None else None,
                                       _l.ptr(row_scale), int(relu),
                                       _l.ptr(att_w), ab, ab_dev,
                                       _l.ptr(att), C.byref(multi_out) if multi_out is not None else None,
                                       C.byref(att_multi_out) if att_multi_out is not None else None,
                                       _l.stream_ptr()))
    return (out, att) if att_w is not None else out


def gat_forward_ex(g, aL, aR, X, slope=0.2, relu=False, out=None, alpha_out=None, att_w=None, att_b=None,
                   cls_wT=None, cls_b=None, want_y=True, multi_out=None, att_multi_out=None):
    """Fused GAT layer + dense epilogue on every finished output row: the next layer's two attention
    projections (att_w [2,K], att_b two floats -> att [2, nrows]) and / or the transform that follows
    the aggregation (cls_wT [K,C] = Linear weight transposed -> [nrows, C]).  Returns (Y, att, cls)."""
    X, aL, aR = _f32(X), _f32(aL), _f32(aR)
    K = X.shape[1]
    dev = X.device
    if multi_out is not None:
        want_y = False
    if want_y and out is None:
        out = torch.empty((g.nrows, K), dtype=torch.float32, device=dev)
    ep = _l.GalaDenseEpilogue()
    if multi_out is not None:
        ep.multi_out = C.pointer(multi_out)
    if att_multi_out is not None:
        ep.att_multi_out = C.pointer(att_multi_out)
    att = cls = None
    if att_w is not None:
        att = torch.empty((2, g.nrows), dtype=torch.float32, device=dev)
        att_w = _f32(att_w)
        ep.att_w = att_w.data_ptr()
        ep.att_b[0], ep.att_b[1] = float(att_b[0]), float(att_b[1])
        ep.att_out = att.data_ptr()
    if cls_wT is not None:
        cls_wT = _f32(cls_wT)
        cls = torch.empty((g.nrows, cls_wT.shape[1]), dtype=torch.float32, device=dev)
        ep.cls_wT = cls_wT.data_ptr()
        ep.cls_b = cls_b.data_ptr() if cls_b is not None else None
        ep.cls_out = cls.data_ptr()
        ep.cls_n = cls_wT.shape[1]
    _l.check(_l.load().gala_gat_forward_ex_f32(C.byref(g.c), _l.ptr(aL), _l.ptr(aR), _l.ptr(X), K,
                                               slope, _l.ptr(out) if want_y else None, _l.ptr(alpha_out),
                                               int(relu), C.byref(ep), g._p(), _l.stream_ptr()))
    return out, att, cls


def linear_small(X, W, bias=None, relu=False, transpose_out=False, out=None):
    """Y = X @ W.T + bias for K <= 64, N <= 64 (exact fp32, one streaming pass, weights in
    registers).  transpose_out=True returns [N, M] (e.g. the two attention projections)."""
    X, W = _f32(X), _f32(W)
    M, K = X.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty((N, M) if transpose_out else (M, N), dtype=torch.float32, device=X.device)
    _l.check(_l.load().gala_linear_small_f32(_l.ptr(X), M, K, _l.ptr(W), _l.ptr(bias), N, _l.ptr(out), int(relu),
                                             int(transpose_out), _l.stream_ptr()))
    return out


def linear_small_ex(X, W, bias=None, relu=False, row_scale=None, att_w=None, att_b=None, out=None, att_out=None,
                    multi_out=None, att_multi_out=None, max_ctas=0):
    """The narrow transform (K, N <= 64, exact fp32) with the row epilogues of `linear`: row scale, the two attention
    projections (att_b: two host floats) into att_out [2, M], rows / right-hand scalars pushed to every GPU.  A light
    kernel (at most max_ctas 256-thread blocks) meant to run on a side stream next to an aggregation kernel."""
    X, W = _f32(X), _f32(W)
    M, K = X.shape
    N = W.shape[0]
    if out is None and multi_out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=X.device)
    ab = None
    if att_w is not None:
        att_w = _f32(att_w)
        if att_out is None:
            att_out = torch.empty((2, M), dtype=torch.float32, device=X.device)
        ab_host = (C.c_float * 2)(float(att_b[0]), float(att_b[1]))
        ab = C.cast(ab_host, C.c_void_p)
    _l.check(_l.load().gala_linear_small_ex_f32(_l.ptr(X), M, K, _l.ptr(W), _l.ptr(bias), N,
                                                _l.ptr(out) if out is not None else None, _l.ptr(row_scale),
                                                int(relu), _l.ptr(att_w), ab, _l.ptr(att_out),
                                                C.byref(multi_out) if multi_out is not None else None,
                                                C.byref(att_multi_out) if att_multi_out is not None else None,
                                                int(max_ctas), _l.stream_ptr()))
    return (out, att_out) if att_w is not None else out


def push_rows(X, multi_out, scalars=None, scalar_multi_out=None, max_ctas=0):
    """Store the rows of X [M, K] (and one scalar per row) into every GPU that gathers them (gala_push_rows_f32)."""
    X = _f32(X)
    M, K = X.shape
    if scalars is not None:
        scalars = _f32(scalars)
    _l.check(_l.load().gala_push_rows_f32(_l.ptr(X), M, K, K, _l.ptr(scalars), C.byref(multi_out),
                                          C.byref(scalar_multi_out) if scalar_multi_out is not None else None,
                                          int(max_ctas), _l.stream_ptr()))


def dense(X, W, bias=None, out=None):
    """Y = X @ W.T + bias on whichever of this library's transforms covers the shape: the streaming kernel for
    the narrow ones (K, N <= 64), the tcgen05 kernel up to N = 256, cuBLAS beyond."""
    K, N = W.shape[1], W.shape[0]
    if K <= LINEAR_SMALL_MAX and N <= LINEAR_SMALL_MAX:
        return linear_small(X, W, bias, out=out)
    if N <= LINEAR_MAX_N:
        return linear(X, W, bias, out=out)
    import torch.nn.functional as F
    y = F.linear(X, W, bias)
    if out is not None:
        out.copy_(y)
        return out
    return y


def probe_read_gbs(nbytes, repeats, device="cuda:0"):
    """Measured read bandwidth (GB/s) of a `nbytes` buffer streamed `repeats` times by every SM
    (gala_b200_probe_read): L2->SM bandwidth when the buffer fits the L2, HBM bandwidth when it does not."""
    buf = torch.empty(nbytes // 4, dtype=torch.float32, device=device).normal_()
    sink = torch.zeros(4, dtype=torch.int32, device=device)
    lib = _l.load()
    for _ in range(2):
        _l.check(lib.gala_b200_probe_read(_l.ptr(buf), nbytes, repeats, _l.ptr(sink), _l.stream_ptr()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    _l.check(lib.gala_b200_probe_read(_l.ptr(buf), nbytes, repeats, _l.ptr(sink), _l.stream_ptr()))
    b.record()
    torch.cuda.synchronize()
    return nbytes * repeats / (a.elapsed_time(b) * 1e-3) / 1e9


# ---- optional bf16 feature storage (half the gathered bytes; fp32 accumulation and outputs) --------------
def _bf16(x):
    assert x.dtype == torch.bfloat16 and x.is_cuda, "bf16 entry points take torch.bfloat16 CUDA features"
    return x.contiguous()


def spmm_bf16(g, X_bf16, vals=None, out=None, row_scale=None, col_scale=None, accumulate=False, relu=False):
    """Y(fp32) = A @ X with X stored as bf16 (gala_spmm_bf16).  K in {8,16,32,64,128,256}."""
    X = _bf16(X_bf16)
    K = X.shape[1]
    if out is None:
        out = torch.empty((g.nrows, K), dtype=torch.float32, device=X.device)
        assert not accumulate, "accumulate needs a caller-provided output"
    ep = _l.GalaEpilogue(row_scale=row_scale.data_ptr() if row_scale is not None else None,
                         col_scale=col_scale.data_ptr() if col_scale is not None else None,
                         accumulate=int(accumulate), relu=int(relu), schedule=0)
    _l.check(_l.load().gala_spmm_bf16(C.byref(g.c), _l.ptr(vals), _l.ptr(X), K, _l.ptr(out), C.byref(ep), g._p(),
                                      _l.stream_ptr()))
    return out


def gat_forward_bf16(g, aL, aR, X_bf16, slope=0.2, relu=False, out=None, alpha_out=None):
    """Fused GAT layer gathering bf16 feature rows (gala_gat_forward_bf16); logits, softmax, sums in fp32."""
    X = _bf16(X_bf16)
    K = X.shape[1]
    if out is None:
        out = torch.empty((g.nrows, K), dtype=torch.float32, device=X.device)
    aL, aR = _f32(aL), _f32(aR)
    _l.check(_l.load().gala_gat_forward_bf16(C.byref(g.c), _l.ptr(aL), _l.ptr(aR), _l.ptr(X), K, slope,
                                             _l.ptr(out), _l.ptr(alpha_out), int(relu), g._p(), _l.stream_ptr()))
    return out
