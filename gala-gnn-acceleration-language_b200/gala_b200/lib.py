"""ctypes binding of libgala_b200.so (include/gala_b200.h).

torch is used for device memory and streams only; every compute call goes through
the C-ABI with raw device pointers.  There is no CPU fallback: if the shared
library is missing, loading fails loudly.
"""
import ctypes as C
import os

import torch

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# GALA_B200_LIB selects an alternative build of the same library (kernel-variant experiments)
LIB_PATH = os.environ.get("GALA_B200_LIB") or os.path.join(_PKG_ROOT, "libgala_b200.so")

EXPORTS = [
    "gala_b200_abi_version", "gala_b200_error_string", "gala_b200_probe_read", "gala_plan_workspace_bytes",
    "gala_plan_build", "gala_spmm_f32", "gala_spmm_sampled_f32", "gala_edge_rowsum_f32",
    "gala_edge_scale_rows_f32", "gala_sddvv_f32", "gala_sddmm_f32", "gala_edge_softmax_fwd_f32",
    "gala_edge_softmax_bwd_f32", "gala_gat_forward_f32", "gala_gat_forward_dot_f32", "gala_linear_f32",
    "gala_gat_forward_ex_f32", "gala_linear_small_f32", "gala_linear_small_ex_f32", "gala_push_rows_f32", "gala_gat_backward_att_f32",
    "gala_spmm_bf16", "gala_gat_forward_bf16", "gala_pad_rows_f32", "gala_gat_forward_col_f32", "gala_reflection_f32",
    "gala_csr_from_coo_workspace_bytes", "gala_csr_from_coo", "gala_csr_transpose",
    "gala_col_tile_segments", "gala_col_tile_workspace_bytes", "gala_col_tile", "gala_sample_ab",
    "gala_mask_subgraph_workspace_bytes", "gala_mask_subgraph",
    "gala_csr_reorder", "gala_permute_rows_f32", "gala_degree_order_workspace_bytes", "gala_degree_order",
]


class GalaGraph(C.Structure):
    _fields_ = [("offsets", C.c_void_p), ("cols", C.c_void_p), ("bounds", C.c_void_p),
                ("nrows", C.c_int32), ("ncols", C.c_int32), ("segments", C.c_int32),
                ("nvals", C.c_int64), ("bounds_dev", C.c_void_p)]


class GalaPlan(C.Structure):
    _fields_ = [("hub_rows", C.c_void_p), ("row_order", C.c_void_p), ("n_hub", C.c_int32),
                ("n_ordered", C.c_int32), ("hub_threshold", C.c_int32), ("tile_rows", C.c_void_p),
                ("n_tiles", C.c_int32), ("tile_edges", C.c_int32), ("tile_policy", C.c_int32)]


class GalaEpilogue(C.Structure):
    _fields_ = [("row_scale", C.c_void_p), ("col_scale", C.c_void_p), ("accumulate", C.c_int32),
                ("relu", C.c_int32), ("schedule", C.c_int32), ("ldx", C.c_int64), ("ldy", C.c_int64),
                ("multi_out", C.c_void_p)]


class GalaMultiOut(C.Structure):
    _fields_ = [("base", C.c_void_p * 8), ("multicast_base", C.c_void_p), ("count", C.c_int32),
                ("need_mask", C.c_void_p)]


class GalaDenseEpilogue(C.Structure):
    _fields_ = [("att_w", C.c_void_p), ("att_b", C.c_float * 2), ("att_out", C.c_void_p),
                ("cls_wT", C.c_void_p), ("cls_b", C.c_void_p), ("cls_out", C.c_void_p),
                ("cls_n", C.c_int32), ("multi_out", C.POINTER(GalaMultiOut)), ("ldx", C.c_int64),
                ("ldy", C.c_int64), ("att_multi_out", C.POINTER(GalaMultiOut))]


class GalaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gala_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load the shared library (no GPU needed to load or to list symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {_PKG_ROOT}` (or "
            "__graft_entry__.build()).  gala_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.gala_b200_error_string.restype = C.c_char_p
    lib.gala_b200_error_string.argtypes = [C.c_int]
    lib.gala_plan_workspace_bytes.restype = C.c_size_t
    lib.gala_plan_workspace_bytes.argtypes = [C.POINTER(GalaGraph)]
    G, P, E = C.POINTER(GalaGraph), C.POINTER(GalaPlan), C.POINTER(GalaEpilogue)
    vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
    sigs = {
        "gala_plan_build": [G, i32, vp, C.c_size_t, P, vp],
        "gala_spmm_f32": [G, vp, vp, i32, vp, E, P, vp],
        "gala_spmm_sampled_f32": [G, vp, vp, i32, vp, i32, i32, i32, i32, C.c_int64, C.c_int64, vp],
        "gala_pad_rows_f32": [vp, C.c_int64, i32, C.c_int64, vp, C.c_int64, vp],
        "gala_spmm_bf16": [G, vp, vp, i32, vp, E, P, vp],
        "gala_gat_forward_bf16": [G, vp, vp, vp, i32, f32, vp, vp, i32, P, vp],
        "gala_edge_rowsum_f32": [G, vp, vp, f32, P, vp],
        "gala_edge_scale_rows_f32": [G, vp, vp, P, vp],
        "gala_sddvv_f32": [G, vp, vp, vp, i32, f32, P, vp],
        "gala_sddmm_f32": [G, vp, vp, i32, vp, P, vp],
        "gala_edge_softmax_fwd_f32": [G, vp, vp, vp, P, vp],
        "gala_edge_softmax_bwd_f32": [G, vp, vp, vp, P, vp],
        "gala_gat_backward_att_f32": [G, vp, vp, vp, vp, f32, vp, P, vp],
        "gala_gat_forward_f32": [G, vp, vp, vp, i32, f32, vp, vp, i32, P, vp],
        "gala_gat_forward_dot_f32": [G, vp, vp, f32, vp, i32, f32, vp, vp, i32, P, vp],
        "gala_gat_forward_ex_f32": [G, vp, vp, vp, i32, f32, vp, vp, i32, C.POINTER(GalaDenseEpilogue), P, vp],
        "gala_gat_forward_col_f32": [G, vp, f32, f32, vp, i32, f32, vp, vp, i32, vp, vp, C.POINTER(GalaDenseEpilogue), P, vp],
        "gala_reflection_f32": [C.POINTER(C.c_float), i32, C.POINTER(C.c_float), C.POINTER(C.c_float)],
    }
    i64, sz = C.c_int64, C.c_size_t
    sigs.update({
        "gala_linear_small_f32": [vp, i64, i32, vp, vp, i32, vp, i32, i32, vp],
        "gala_linear_small_ex_f32": [vp, i64, i32, vp, vp, i32, vp, vp, i32, vp, vp, vp,
                                     C.POINTER(GalaMultiOut), C.POINTER(GalaMultiOut), i32, vp],
        "gala_push_rows_f32": [vp, i64, i32, i64, vp, C.POINTER(GalaMultiOut), C.POINTER(GalaMultiOut), i32, vp],
        "gala_linear_f32": [vp, i64, i32, vp, vp, i32, vp, vp, i32, vp, vp, i32, vp,
                            C.POINTER(GalaMultiOut), C.POINTER(GalaMultiOut), vp],
        "gala_csr_from_coo": [i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, sz, vp],
        "gala_csr_transpose": [i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, sz, vp],
        "gala_col_tile": [i32, i32, i64, vp, vp, vp, i32, vp, vp, vp, vp, vp, sz, vp],
        "gala_sample_ab": [i32, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp],
        "gala_mask_subgraph": [i32, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int64), vp, vp, sz, vp],
        "gala_b200_probe_read": [vp, sz, i32, vp, vp],
        "gala_csr_reorder": [i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp],
        "gala_permute_rows_f32": [vp, vp, vp, i32, i32, i32, vp],
        "gala_degree_order": [i32, vp, vp, vp, vp, sz, vp],
    })
    lib.gala_degree_order_workspace_bytes.restype = sz
    lib.gala_degree_order_workspace_bytes.argtypes = [i32]
    lib.gala_csr_from_coo_workspace_bytes.restype = sz
    lib.gala_csr_from_coo_workspace_bytes.argtypes = [i32, i32, i64]
    lib.gala_col_tile_segments.restype = i32
    lib.gala_col_tile_segments.argtypes = [i32, i32]
    lib.gala_col_tile_workspace_bytes.restype = sz
    lib.gala_col_tile_workspace_bytes.argtypes = [i32, i32, i32]
    lib.gala_mask_subgraph_workspace_bytes.restype = sz
    lib.gala_mask_subgraph_workspace_bytes.argtypes = [i32]
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise GalaError(code, load().gala_b200_error_string(code).decode())


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "gala_b200 takes contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
