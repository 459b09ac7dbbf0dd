// gala_b200_torch.h -- libtorch shim: re-creates, on top of the C-ABI in
// include/gala_b200.h, the host wrappers that GALA's code generator emits as text into
// every gala.cu (reference src/codegen/cuda.h:441-952).  Header-only; a generated program
// includes it after its globals (`int global_nrows; int global_ra; int global_rb;`,
// src/codegen/common.h:1694-1705) and links libgala_b200.so.
//
//   reference (text pasted into gala.cu)                      here
//   ---------------------------------------------------------------------------------------
//   <kernel>_call(input_dense, offset, cols, vals[, bounds,   GALA_B200_DEFINE_AGGREGATE(name, weighted,
//       segments])                 cuda.h:441-499, 213-276        nsamples) stamps the exact emitted name
//   edge_sddvv                     cuda.h:773-807             edge_sddvv
//   edge_sddmm                     cuda.h:808-845             edge_sddmm
//   node_spmv_backward_of_sddmm_{nln,eaggr}  :565-600,737-772 same names
//   inplace_softmax_sddvv[_mult]   cuda.h:601-656             same names
//   aggregate_edge_mul[_dir]       cuda.h:870-952             same names
//
// Semantics kept: fresh output tensors with requires_grad(true) on the input's device,
// in-place ops return value_graph, `bounds` is a CPU int tensor, errors abort the program
// (the reference exit()s; here TORCH_CHECK throws).  Changed: one launch per call on the
// current torch stream instead of 1-3 fresh never-destroyed streams per segment.
#pragma once
#include <c10/cuda/CUDAStream.h>
#include <torch/torch.h>

#include <unordered_map>

#include "gala_b200.h"

extern int global_nrows;
extern int global_ra;
extern int global_rb;
extern std::vector<int> global_segments;                      // common.h:1694-1705: slot 2g = forward graph g,
extern std::vector<torch::Tensor> global_offset_graph;        // slot 2g+1 = its backward (transpose) graph
extern std::vector<torch::Tensor> global_columns_graph;
extern std::vector<torch::Tensor> global_bounds;

namespace gala_b200 {

inline void check(int rc, const char* what) {
    TORCH_CHECK(rc == 0, what, ": ", gala_b200_error_string(rc), " (", rc, ")");
}

inline gala_stream_t stream() { return (gala_stream_t)c10::cuda::getCurrentCUDAStream().stream(); }

struct PlanEntry {
    torch::Tensor workspace;
    gala_plan_t plan;
};

inline gala_graph_t make_graph(const torch::Tensor& offset_graph, const torch::Tensor& columns_graph,
                               const torch::Tensor& bounds, int segments, int64_t nrows) {
    gala_graph_t g;
    g.offsets = offset_graph.data_ptr<int>();
    g.cols = columns_graph.data_ptr<int>();
    g.bounds = (bounds.defined() && bounds.numel() > 0) ? bounds.data_ptr<int>() : nullptr;  // CPU tensor
    g.nrows = (int32_t)nrows;
    g.ncols = (int32_t)nrows;
    g.segments = segments;
    g.nvals = columns_graph.numel();
    g.bounds_dev = nullptr;
    if (segments > 64) {      // beyond 64 segments the kernels read the segment starts from a device copy of bounds
        static std::unordered_map<const void*, torch::Tensor> dev_bounds;     // one per graph, keyed by the host array
        auto it = dev_bounds.find(g.bounds);
        if (it == dev_bounds.end() || it->second.numel() != bounds.numel())
            it = dev_bounds.insert_or_assign(g.bounds, bounds.to(columns_graph.device())).first;
        g.bounds_dev = it->second.data_ptr<int>();
    }
    return g;
}

// One plan per graph (keyed by the row-pointer array), built on first use.
inline const gala_plan_t* plan_for(const gala_graph_t& g, const torch::Tensor& like) {
    static std::unordered_map<const void*, PlanEntry> cache;
    auto it = cache.find(g.offsets);
    if (it != cache.end() && it->second.plan.n_hub + it->second.plan.n_ordered == g.nrows) return &it->second.plan;
    PlanEntry e;
    size_t bytes = gala_plan_workspace_bytes(&g);
    e.workspace = torch::empty({(int64_t)bytes}, torch::TensorOptions().dtype(torch::kUInt8).device(like.device()));
    check(gala_plan_build(&g, 2048, e.workspace.data_ptr(), bytes, &e.plan, stream()), "gala_plan_build");
    cache[g.offsets] = e;
    return &cache[g.offsets].plan;
}

inline torch::TensorOptions out_options(const torch::Tensor& like) {
    return torch::TensorOptions().dtype(torch::kFloat).requires_grad(true).device(like.device());
}

// The dense operand whose rows are gathered, as (tensor, row pitch).  Packed rows whose width is not a multiple of
// 4 (the class counts 41 / 47 the shipped schedules aggregate at; the reference handles them with its K % 32
// remainder kernels, cuda.h:58-168) are re-pitched once with gala_pad_rows_f32 so that every row is gathered with
// 128-bit loads: N*K*8 bytes against the E*K*4 bytes of the gather.
inline std::pair<torch::Tensor, int64_t> gather_operand(const torch::Tensor& input_dense, int64_t nrows, int64_t dcols) {
    auto X = input_dense.contiguous();
    if (dcols <= 4 || dcols > 256 || (dcols % 4 == 0 && reinterpret_cast<uintptr_t>(X.data_ptr()) % 16 == 0)) return {X, dcols};
    // 16-byte aligned rows; rows longer than 64 bytes on a 64-byte pitch so that none straddles an extra 128-byte line
    const int64_t ld = dcols > 16 ? (dcols + 15) / 16 * 16 : (dcols + 3) / 4 * 4;
    auto Xp = torch::empty({nrows, ld}, torch::TensorOptions().dtype(torch::kFloat).device(X.device()));
    check(gala_pad_rows_f32(X.data_ptr<float>(), nrows, (int)dcols, dcols, Xp.data_ptr<float>(), ld, stream()),
          "gala_pad_rows_f32");
    return {Xp, ld};
}

// aggregate_node_mul_sum*_call: Y = A @ X.  nsamples > 0 selects the sampled flavour
// (cuda.h:313-320) with global_ra / global_rb.
inline torch::Tensor aggregate(const torch::Tensor& input_dense, const torch::Tensor& offset_graph,
                               const torch::Tensor& columns_graph, const torch::Tensor& value_graph,
                               const torch::Tensor& bounds, int segments, bool weighted, int nsamples) {
    // cuda.h:451 uses global_nrows; the cuSPARSE flavour (cuda.h:217) derives it from the row pointers
    const int64_t nrows = global_nrows > 0 ? global_nrows : offset_graph.numel() - 1;
    const int64_t dcols = input_dense.numel() / nrows;
    auto [X, ldx] = gather_operand(input_dense, nrows, dcols);
    auto Y = torch::empty({nrows, dcols}, out_options(input_dense));
    gala_graph_t g = make_graph(offset_graph, columns_graph, bounds, segments > 0 ? segments : 1, nrows);
    const float* vals = weighted ? value_graph.data_ptr<float>() : nullptr;
    if (nsamples > 0) {
        check(gala_spmm_sampled_f32(&g, vals, X.data_ptr<float>(), (int)dcols, Y.data_ptr<float>(), nsamples,
                                    global_ra, global_rb, 0, ldx, dcols, stream()), "gala_spmm_sampled_f32");
    } else {
        gala_epilogue_t ep = {};
        ep.ldx = ldx;
        ep.ldy = dcols;
        check(gala_spmm_f32(&g, vals, X.data_ptr<float>(), (int)dcols, Y.data_ptr<float>(), &ep,
                            plan_for(g, input_dense), stream()), "gala_spmm_f32");
    }
    return Y;
}

// Fused GAT layer: what the retargeted generator emits for edge_sddvv -> LeakyReLU ->
// non_lnr_op_softmax -> aggregate_node_mul_sum (common.h:622-675,735-810,835-927).
inline torch::Tensor gat_forward(const torch::Tensor& res, const torch::Tensor& attenL, const torch::Tensor& attenR,
                                 const torch::Tensor& offset_graph, const torch::Tensor& columns_graph,
                                 const torch::Tensor& bounds, int segments, float slope, bool relu,
                                 torch::Tensor* alpha_out = nullptr) {
    const int64_t nrows = global_nrows;
    const int64_t dcols = res.numel() / nrows;
    auto [X, ldx] = gather_operand(res, nrows, dcols);
    auto aL = attenL.contiguous();
    auto aR = attenR.contiguous();
    auto Y = torch::empty({nrows, dcols}, out_options(res));
    gala_graph_t g = make_graph(offset_graph, columns_graph, bounds, segments, nrows);
    float* alpha = nullptr;
    if (alpha_out) {
        *alpha_out = torch::empty({columns_graph.numel()}, out_options(res));
        alpha = alpha_out->data_ptr<float>();
    }
    gala_dense_epilogue_t ep = {};
    ep.ldx = ldx;
    ep.ldy = dcols;
    check(gala_gat_forward_ex_f32(&g, aL.data_ptr<float>(), aR.data_ptr<float>(), X.data_ptr<float>(), (int)dcols,
                                  slope, Y.data_ptr<float>(), alpha, relu ? 1 : 0, &ep, plan_for(g, res), stream()),
          "gala_gat_forward_ex_f32");
    return Y;
}

// K6 (cuda.h:699-734, 808-845): out[e] = dot(A[row(e),:], B[col(e),:])
inline torch::Tensor edge_sddmm_impl(const torch::Tensor& input_dense1, const torch::Tensor& input_dense2,
                                     const torch::Tensor& offset_graph, const torch::Tensor& columns_graph,
                                     const torch::Tensor& bounds, int nrows, int segments) {
    auto a = input_dense1.contiguous();
    auto b = input_dense2.contiguous();
    const int64_t dcols = a.numel() / nrows;
    auto out = torch::empty({columns_graph.numel()}, out_options(input_dense1));
    gala_graph_t g = make_graph(offset_graph, columns_graph, bounds, segments, nrows);
    check(gala_sddmm_f32(&g, a.data_ptr<float>(), b.data_ptr<float>(), (int)dcols, out.data_ptr<float>(),
                         plan_for(g, input_dense1), stream()),
          "gala_sddmm_f32");
    return out;
}

}  // namespace gala_b200

// Stamps out one emitted aggregation wrapper under its generated name, e.g.
//   GALA_B200_DEFINE_AGGREGATE_TILED(aggregate_node_mul_sum_coarse2_call, true, 0)
#define GALA_B200_DEFINE_AGGREGATE_TILED(NAME, WEIGHTED, NSAMPLES)                                          \
    inline torch::Tensor NAME(torch::Tensor input_dense, torch::Tensor offset_graph,                       \
                              torch::Tensor columns_graph, torch::Tensor value_graph, torch::Tensor bounds, \
                              int segments) {                                                              \
        return gala_b200::aggregate(input_dense, offset_graph, columns_graph, value_graph, bounds, segments, \
                                    WEIGHTED, NSAMPLES);                                                   \
    }
#define GALA_B200_DEFINE_AGGREGATE(NAME, WEIGHTED, NSAMPLES)                                                \
    inline torch::Tensor NAME(torch::Tensor input_dense, torch::Tensor offset_graph,                       \
                              torch::Tensor columns_graph, torch::Tensor value_graph) {                    \
        return gala_b200::aggregate(input_dense, offset_graph, columns_graph, value_graph, torch::Tensor(), 0, \
                                    WEIGHTED, NSAMPLES);                                                   \
    }

// ---- fixed emitted names ---------------------------------------------------------------------
inline torch::Tensor node_spmv_backward_of_sddmm_nln(torch::Tensor offset_graph, torch::Tensor columns_graph,
                                                     torch::Tensor value_graph, torch::Tensor bounds, int nrows,
                                                     int segments) {
    auto vals = value_graph.contiguous();
    auto out = torch::empty({nrows, 1}, gala_b200::out_options(value_graph));
    gala_graph_t g = gala_b200::make_graph(offset_graph, columns_graph, bounds, segments, nrows);
    gala_b200::check(gala_edge_rowsum_f32(&g, vals.data_ptr<float>(), out.data_ptr<float>(), 1e-12f,
                                          gala_b200::plan_for(g, value_graph), gala_b200::stream()),
                     "gala_edge_rowsum_f32");
    return out;
}

inline torch::Tensor node_spmv_backward_of_sddmm_eaggr(torch::Tensor offset_graph, torch::Tensor columns_graph,
                                                       torch::Tensor value_graph, torch::Tensor bounds, int nrows,
                                                       int segments) {
    return node_spmv_backward_of_sddmm_nln(offset_graph, columns_graph, value_graph, bounds, nrows, segments);
}

inline torch::Tensor inplace_softmax_sddvv(torch::Tensor row_val, torch::Tensor offset_graph,
                                           torch::Tensor columns_graph, torch::Tensor value_graph,
                                           torch::Tensor bounds, int nrows, int segments) {
    auto rv = row_val.contiguous();
    gala_graph_t g = gala_b200::make_graph(offset_graph, columns_graph, bounds, segments, nrows);
    gala_b200::check(gala_edge_scale_rows_f32(&g, value_graph.data_ptr<float>(), rv.data_ptr<float>(),
                                              gala_b200::plan_for(g, value_graph), gala_b200::stream()),
                     "gala_edge_scale_rows_f32");
    return value_graph;
}

inline torch::Tensor inplace_softmax_sddvv_mult(torch::Tensor row_val, torch::Tensor offset_graph,
                                                torch::Tensor columns_graph, torch::Tensor value_graph,
                                                torch::Tensor bounds, int nrows, int segments) {
    return inplace_softmax_sddvv(row_val, offset_graph, columns_graph, value_graph, bounds, nrows, segments);
}

inline torch::Tensor gala_b200_sddvv(torch::Tensor in1, torch::Tensor in2, torch::Tensor offset_graph,
                                     torch::Tensor columns_graph, torch::Tensor bounds, int nrows, int segments,
                                     int op) {
    auto a = in1.contiguous();
    auto b = in2.contiguous();
    auto out = torch::empty({columns_graph.numel()}, gala_b200::out_options(in1));
    gala_graph_t g = gala_b200::make_graph(offset_graph, columns_graph, bounds, segments, nrows);
    gala_b200::check(gala_sddvv_f32(&g, a.data_ptr<float>(), b.data_ptr<float>(), out.data_ptr<float>(), op, 1.0f,
                                    gala_b200::plan_for(g, in1), gala_b200::stream()),
                     "gala_sddvv_f32");
    return out;
}

inline torch::Tensor edge_sddvv(torch::Tensor input_dense1, torch::Tensor input_dense2, torch::Tensor offset_graph,
                                torch::Tensor columns_graph, torch::Tensor value_graph, torch::Tensor bounds,
                                int nrows, int segments) {
    return gala_b200_sddvv(input_dense1, input_dense2, offset_graph, columns_graph, bounds, nrows, segments,
                           GALA_SDDVV_ADD);
}

inline torch::Tensor edge_sddmm(torch::Tensor input_dense1, torch::Tensor input_dense2, torch::Tensor offset_graph,
                                torch::Tensor columns_graph, torch::Tensor value_graph, torch::Tensor bounds,
                                int nrows, int segments) {
    return gala_b200::edge_sddmm_impl(input_dense1, input_dense2, offset_graph, columns_graph, bounds, nrows, segments);
}

inline torch::Tensor aggregate_edge_mul(torch::Tensor input_dense1, torch::Tensor input_dense2,
                                        torch::Tensor offset_graph, torch::Tensor columns_graph,
                                        torch::Tensor value_graph, torch::Tensor bounds, int segments) {
    return gala_b200_sddvv(input_dense1, input_dense2, offset_graph, columns_graph, bounds, global_nrows, segments,
                           GALA_SDDVV_MUL);
}

inline torch::Tensor aggregate_edge_mul_dir(torch::Tensor input_dense1, torch::Tensor input_dense2,
                                            torch::Tensor offset_graph, torch::Tensor columns_graph,
                                            torch::Tensor value_graph) {
    return gala_b200_sddvv(input_dense1, input_dense2, offset_graph, columns_graph, torch::Tensor(), global_nrows, 1,
                           GALA_SDDVV_MUL);
}

// ---- one GAT layer as ONE autograd node -------------------------------------------------------
// The retargeted generator replaces the emitted sequence
//     attn = aggregate_edge_sum_AutoGrad::apply(attenL, attenR, li);        (common.h:630-675)
//     attn = leaky_relu->forward(attn);                                     (common.h:1176-1184)
//     attn = non_lnr_op_softmax_AutoGrad::apply(attn, li);                  (common.h:735-810)
//     res  = aggregate_node_mul_sum_*_AutoGrad::apply(res, attn, li);       (common.h:835-894)
// by  res = gala_b200::gat_layer_AutoGrad::apply(res, attenL, attenR, li, slope)  whenever all four use
// the same graph slot.  Forward: the fused kernel (alpha kept for the backward).  Backward: the same
// values autograd would produce for the four nodes, graph slots as the emitted backward methods pick them:
//   d res    = aggregate(dZ) over slot 2li+1 with alpha, bounds/segments of slot 2li   (common.h:876-889)
//   d alpha  = edge_sddmm(dZ, res) over the same                                        (common.h:886-888)
//   d attenL = d attenR = row sums of LeakyReLU'(.) * softmax-backward(d alpha) over slot 2li+1
//              (common.h:791-799 and 654-670: one node_spmv result returned for both inputs)
namespace gala_b200 {
class gat_layer_AutoGrad : public torch::autograd::Function<gat_layer_AutoGrad> {
public:
    // relu: the torch::relu that follows the layer in the emitted forward (common.h:1166) applied inside the kernel
    static torch::Tensor forward(torch::autograd::AutogradContext* ctx, torch::Tensor res, torch::Tensor attenL,
                                 torch::Tensor attenR, int64_t li, double slope, bool relu = false) {
        ctx->saved_data["li"] = li;
        ctx->saved_data["slope"] = slope;
        ctx->saved_data["relu"] = relu;
        torch::Tensor alpha;
        torch::Tensor Y = gat_forward(res, attenL, attenR, global_offset_graph[2 * li], global_columns_graph[2 * li],
                                      global_bounds[2 * li], global_segments[2 * li], (float)slope, relu, &alpha);
        ctx->save_for_backward({alpha, res, attenL, attenR, relu ? Y : torch::Tensor()});
        return Y;
    }

    static torch::autograd::tensor_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::tensor_list grad_outputs) {
        auto saved = ctx->get_saved_variables();
        torch::Tensor alpha = saved[0], X = saved[1].contiguous();
        torch::Tensor aL = saved[2].contiguous(), aR = saved[3].contiguous();
        torch::Tensor dZ = grad_outputs[0];
        if (ctx->saved_data["relu"].toBool()) dZ = dZ * (saved[4] > 0);      // ReLU backward (threshold_backward)
        dZ = dZ.contiguous();
        const int64_t li = ctx->saved_data["li"].toInt();
        const float slope = (float)ctx->saved_data["slope"].toDouble();
        torch::Tensor off_b = global_offset_graph[2 * li + 1], col_b = global_columns_graph[2 * li + 1];
        torch::Tensor dX = aggregate(dZ, off_b, col_b, alpha, global_bounds[2 * li], global_segments[2 * li], true, 0);
        torch::Tensor dalpha = edge_sddmm_impl(dZ, X, off_b, col_b, global_bounds[2 * li], global_nrows,
                                               global_segments[2 * li]);
        gala_graph_t g = make_graph(off_b, col_b, global_bounds[2 * li + 1], global_segments[2 * li + 1], global_nrows);
        torch::Tensor d_att = torch::empty({(int64_t)global_nrows, 1}, out_options(dZ));
        check(gala_gat_backward_att_f32(&g, alpha.data_ptr<float>(), dalpha.data_ptr<float>(), aL.data_ptr<float>(),
                                        aR.data_ptr<float>(), slope, d_att.data_ptr<float>(), plan_for(g, dZ), stream()),
              "gala_gat_backward_att_f32");
        return {dX, d_att, d_att, torch::Tensor(), torch::Tensor(), torch::Tensor()};
    }
};

// ---- dense transforms of the generated model on this library's kernels -------------------------------------------
// The retargeted generator replaces `torch::nn::Linear` in the emitted GALAGNN (common.h:1185-1281) by
// gala_b200::Linear: the same module (parameters "weight" / "bias", same initialisation and RNG consumption: it IS a
// torch::nn::LinearImpl), whose forward runs gala_linear_f32 (tcgen05, 3xTF32) or, for the narrow transforms
// (K, N <= 64: classifier, Linear(h,1) projections), gala_linear_small_f32.  Backward: dX = dY W, dW = dY^T X,
// db = sum dY through ATen (cuBLAS), as autograd does for torch::nn::Linear.
constexpr int64_t kLinearMaxN = 256;     // gala_linear_f32; the fused attention projections need N <= 64

inline torch::Tensor linear_forward(const torch::Tensor& X, const torch::Tensor& W, const torch::Tensor& b) {
    const int64_t M = X.size(0), K = W.size(1), N = W.size(0);
    const float* bias = b.defined() ? b.data_ptr<float>() : nullptr;
    auto Y = torch::empty({M, N}, torch::TensorOptions().dtype(torch::kFloat).device(X.device()));
    if (K <= 64 && N <= 64) {
        check(gala_linear_small_f32(X.data_ptr<float>(), M, (int)K, W.data_ptr<float>(), bias, (int)N, Y.data_ptr<float>(), 0, 0,
                                    stream()), "gala_linear_small_f32");
    } else if (N <= kLinearMaxN) {
        check(gala_linear_f32(X.data_ptr<float>(), M, (int)K, W.data_ptr<float>(), bias, (int)N, Y.data_ptr<float>(), nullptr, 0,
                              nullptr, nullptr, 0, nullptr, nullptr, nullptr, stream()), "gala_linear_f32");
    } else {
        return at::linear(X, W, b);      // wider than the hand-written kernels: cuBLAS
    }
    return Y;
}

class linear_AutoGrad : public torch::autograd::Function<linear_AutoGrad> {
public:
    static torch::Tensor forward(torch::autograd::AutogradContext* ctx, torch::Tensor X, torch::Tensor W, torch::Tensor b) {
        X = X.contiguous();
        ctx->save_for_backward({X, W});
        ctx->saved_data["has_bias"] = b.defined();
        return linear_forward(X, W.contiguous(), b);
    }
    static torch::autograd::tensor_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::tensor_list grads) {
        auto saved = ctx->get_saved_variables();
        torch::Tensor X = saved[0], W = saved[1], dY = grads[0].contiguous();
        torch::Tensor dX = ctx->needs_input_grad(0) ? dY.mm(W) : torch::Tensor();
        torch::Tensor dW = ctx->needs_input_grad(1) ? dY.t().mm(X) : torch::Tensor();
        torch::Tensor db = (ctx->saved_data["has_bias"].toBool() && ctx->needs_input_grad(2)) ? dY.sum(0) : torch::Tensor();
        return {dX, dW, db};
    }
};

struct LinearImpl : torch::nn::LinearImpl {
    using torch::nn::LinearImpl::LinearImpl;
    torch::Tensor forward(const torch::Tensor& input) { return linear_AutoGrad::apply(input, weight, bias); }
};
TORCH_MODULE(Linear);

// res = fc(X); attenL = el(res); attenR = er(res) (FFN_OP, FFN_OP_EDGE, FFN_OP_EDGE: frontend.y:987-994) as one
// node: the tcgen05 transform computes both projections of every output row in its epilogue.
class linear_att_AutoGrad : public torch::autograd::Function<linear_att_AutoGrad> {
public:
    static torch::autograd::tensor_list forward(torch::autograd::AutogradContext* ctx, torch::Tensor X, torch::Tensor W,
                                                torch::Tensor b, torch::Tensor wl, torch::Tensor bl, torch::Tensor wr,
                                                torch::Tensor br) {
        X = X.contiguous();
        const int64_t M = X.size(0), K = W.size(1), N = W.size(0);
        auto of = torch::TensorOptions().dtype(torch::kFloat).device(X.device());
        torch::Tensor att_w = torch::cat({wl.reshape({1, N}), wr.reshape({1, N})}, 0).contiguous();
        torch::Tensor att_b = torch::cat({bl.reshape({1}), br.reshape({1})}, 0).contiguous();
        torch::Tensor res, att;
        if (N <= 64) {
            res = torch::empty({M, N}, of);
            att = torch::empty({2, M}, of);
            check(gala_linear_f32(X.data_ptr<float>(), M, (int)K, W.contiguous().data_ptr<float>(), b.data_ptr<float>(), (int)N,
                                  res.data_ptr<float>(), nullptr, 0, att_w.data_ptr<float>(), att_b.data_ptr<float>(), 1,
                                  att.data_ptr<float>(), nullptr, nullptr, stream()), "gala_linear_f32");
        } else {
            res = at::linear(X, W, b);
            att = at::linear(res, att_w, att_b).t().contiguous();
        }
        ctx->save_for_backward({X, W, res, att_w});
        return {res, att[0].reshape({M, 1}), att[1].reshape({M, 1})};
    }
    static torch::autograd::tensor_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::tensor_list grads) {
        auto saved = ctx->get_saved_variables();
        torch::Tensor X = saved[0], W = saved[1], res = saved[2], att_w = saved[3];
        const int64_t M = X.size(0);
        torch::Tensor g;            // gradient reaching res: its own plus the two projections'
        torch::Tensor dl = grads[1].defined() ? grads[1].reshape({M, 1}) : torch::Tensor();
        torch::Tensor dr = grads[2].defined() ? grads[2].reshape({M, 1}) : torch::Tensor();
        if (grads[0].defined()) g = grads[0];
        auto add = [&](const torch::Tensor& t) { g = g.defined() ? g + t : t; };
        if (dl.defined()) add(dl * att_w[0].reshape({1, -1}));
        if (dr.defined()) add(dr * att_w[1].reshape({1, -1}));
        g = g.contiguous();
        torch::Tensor dX = ctx->needs_input_grad(0) ? g.mm(W) : torch::Tensor();
        torch::Tensor dW = g.t().mm(X), db = g.sum(0);
        torch::Tensor dwl = dl.defined() ? dl.t().mm(res) : torch::Tensor(), dbl = dl.defined() ? dl.sum(0) : torch::Tensor();
        torch::Tensor dwr = dr.defined() ? dr.t().mm(res) : torch::Tensor(), dbr = dr.defined() ? dr.sum(0) : torch::Tensor();
        return {dX, dW, db, dwl, dbl, dwr, dbr};
    }
};

template <class L>
inline std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> linear_att(L& fc, L& el, L& er, const torch::Tensor& X) {
    auto out = linear_att_AutoGrad::apply(X, fc->weight, fc->bias, el->weight, el->bias, er->weight, er->bias);
    return {out[0], out[1], out[2]};
}

// attenL = el(fc(X)), attenR = er(fc(X)) when fc(X) feeds nothing else (layer 2 under the FFN-recompute rewrite,
// middle-end.h:324-375): the projections are folded through the transform, w = el.w fc.W, b = el.w fc.b + el.b
// (tiny ATen ops, differentiated by autograd), and one [M,K] x [K,2] streaming pass replaces the [M,K] x [K,N]
// transform and the two [M,N] x [N,1] projections.
template <class L>
inline std::tuple<torch::Tensor, torch::Tensor> folded_att(L& fc, L& el, L& er, const torch::Tensor& X) {
    torch::Tensor Wf = torch::cat({el->weight.mm(fc->weight), er->weight.mm(fc->weight)}, 0);                    // [2, K]
    torch::Tensor bf = torch::cat({el->weight.mv(fc->bias) + el->bias, er->weight.mv(fc->bias) + er->bias}, 0);   // [2]
    torch::Tensor att = linear_AutoGrad::apply(X, Wf, bf).t().contiguous();                                        // [2, M]
    const int64_t M = X.size(0);
    return {att[0].reshape({M, 1}), att[1].reshape({M, 1})};
}
}  // namespace gala_b200
