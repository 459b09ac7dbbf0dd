"""Python mirror of the wrappers GALA's code generator emits into gala.cu.

Same names, argument order and semantics as the reference's emitted C++
(src/codegen/cuda.h of ADAPT-uiuc/GALA; line numbers below), so that the parity
tests read like calls a generated model makes.  The C++ counterpart that a generated
gala.cu actually links is host/gala_b200_torch.h.

Module-level state mirrors the generated program's globals (common.h:1694-1705):
global_nrows, global_ra, global_rb.
"""
import torch

from . import ops

global_nrows = 0
global_ra = 5   # common.h:817
global_rb = 7   # common.h:818

_graph_cache = {}


def _graph(offset_graph, columns_graph, bounds=None, segments=1, nrows=None):
    """TiledGraph view over the tensors the generated code passes around (cached per
    offset tensor so that the hub plan is built once per graph)."""
    n = int(nrows if nrows is not None else global_nrows)
    key = (offset_graph.data_ptr(), columns_graph.data_ptr(), int(segments), n)
    g = _graph_cache.get(key)
    if g is None:
        g = ops.TiledGraph(offset_graph, columns_graph, n, n, bounds, segments)
        g.build_plan()
        _graph_cache[key] = g
    return g


def clear_cache():
    _graph_cache.clear()


# cuda.h:441-499 (col-tiled / coarsened) and :213-276 (cuSPARSE flavour)
def aggregate_node_mul_sum_call(input_dense, offset_graph, columns_graph, value_graph,
                                bounds=None, segments=1):
    """Weighted graph: Y = A(value_graph) @ input_dense, fresh output."""
    g = _graph(offset_graph, columns_graph, bounds, segments)
    return ops.spmm(g, input_dense.reshape(g.ncols, -1), vals=value_graph)


def aggregate_node_mul_sum_direct_call(input_dense, offset_graph, columns_graph, value_graph,
                                       bounds=None, segments=1):
    """Unweighted graph (`getWeighted()` false, cuda.h:292-295): value_graph is ignored."""
    g = _graph(offset_graph, columns_graph, bounds, segments)
    return ops.spmm(g, input_dense.reshape(g.ncols, -1), vals=None)


def aggregate_node_mul_sum_sample_call(input_dense, offset_graph, columns_graph, value_graph,
                                       nsamples, bounds=None, segments=1, weighted=False):
    """Sampled flavour (cuda.h:313-320): uses global_ra / global_rb."""
    g = _graph(offset_graph, columns_graph, bounds, segments)
    return ops.spmm_sampled(g, input_dense.reshape(g.ncols, -1), nsamples, global_ra, global_rb,
                            vals=value_graph if weighted else None)


# cuda.h:565-600 and :737-772
def node_spmv_backward_of_sddmm_nln(offset_graph, columns_graph, value_graph, bounds, nrows,
                                    segments):
    g = _graph(offset_graph, columns_graph, bounds, segments, nrows)
    return ops.edge_rowsum(g, value_graph, seed=1e-12)


node_spmv_backward_of_sddmm_eaggr = node_spmv_backward_of_sddmm_nln


# cuda.h:601-656 (the two emitted kernels are byte-identical)
def inplace_softmax_sddvv(row_val, offset_graph, columns_graph, value_graph, bounds, nrows,
                          segments):
    g = _graph(offset_graph, columns_graph, bounds, segments, nrows)
    return ops.edge_scale_rows_(g, value_graph, row_val)


inplace_softmax_sddvv_mult = inplace_softmax_sddvv


# cuda.h:773-807
def edge_sddvv(input_dense1, input_dense2, offset_graph, columns_graph, value_graph, bounds,
               nrows, segments):
    g = _graph(offset_graph, columns_graph, bounds, segments, nrows)
    return ops.sddvv(g, input_dense1, input_dense2, "add")


# cuda.h:808-845
def edge_sddmm(input_dense1, input_dense2, offset_graph, columns_graph, value_graph, bounds,
               nrows, segments):
    g = _graph(offset_graph, columns_graph, bounds, segments, nrows)
    return ops.sddmm(g, input_dense1.reshape(nrows, -1), input_dense2.reshape(nrows, -1))


# cuda.h:870-917 and :919-952
def aggregate_edge_mul(input_dense1, input_dense2, offset_graph, columns_graph, value_graph,
                       bounds, segments):
    g = _graph(offset_graph, columns_graph, bounds, segments)
    return ops.sddvv(g, input_dense1, input_dense2, "mul")


def aggregate_edge_mul_dir(input_dense1, input_dense2, offset_graph, columns_graph, value_graph):
    g = _graph(offset_graph, columns_graph, None, 1)
    return ops.sddvv(g, input_dense1, input_dense2, "mul")


# common.h:760-773 / 791-799: the forward / backward bodies of non_lnr_op_softmax_AutoGrad
def non_lnr_op_softmax_forward(value_graph, offset_graph, columns_graph, bounds, segments):
    g = _graph(offset_graph, columns_graph, bounds, segments)
    return ops.edge_softmax_fwd(g, value_graph)


def non_lnr_op_softmax_backward(saved_alpha, d_value_graph, offset_graph, columns_graph, bounds,
                                segments):
    g = _graph(offset_graph, columns_graph, bounds, segments)
    return ops.edge_softmax_bwd(g, saved_alpha, d_value_graph)
