// shim_selftest.cpp -- exercises host/gala_b200_torch.h exactly the way a generated gala.cu
// does (same globals, same call spelling) and checks every wrapper against dense torch math
// on the GPU.  Built by host/Makefile, run on the B200 box by tests/test_shim_gpu.py.
#include <torch/torch.h>

#include <cstdio>
#include <random>
#include <vector>

int global_nrows;
int global_ra;
int global_rb;
std::vector<int> global_segments;
std::vector<torch::Tensor> global_offset_graph, global_columns_graph, global_value_graph, global_bounds;

#include "gala_b200_torch.h"

// the two aggregation wrappers of the checked-in sample program (codegen/gala.cu:225-390), tiled flavour
GALA_B200_DEFINE_AGGREGATE_TILED(aggregate_node_mul_sum_coarse2_call, true, 0)
GALA_B200_DEFINE_AGGREGATE_TILED(aggregate_node_mul_sum_direct_coarse2_call, false, 0)
GALA_B200_DEFINE_AGGREGATE(aggregate_node_mul_sum_call, true, 0)
GALA_B200_DEFINE_AGGREGATE_TILED(aggregate_node_mul_sum_sample20_call, false, 20)

static int failures = 0;
static void expect_close(const char* what, const torch::Tensor& got, const torch::Tensor& want, double tol) {
    double err = ((got.to(torch::kDouble) - want.to(torch::kDouble)).norm() /
                  want.to(torch::kDouble).norm().clamp_min(1e-30)).item<double>();
    std::printf("%-44s rel err %.3e %s\n", what, err, err < tol ? "ok" : "FAIL");
    if (!(err < tol)) ++failures;
}

// for sums that cancel (the attention gradients: every row of softmax-backward sums to ~0): error relative to the
// magnitude that was summed, the backward-error form of the 1e-5 bound
static void expect_close_mag(const char* what, const torch::Tensor& got, const torch::Tensor& want, const torch::Tensor& mag,
                             double tol) {
    double err = ((got.to(torch::kDouble).flatten() - want.to(torch::kDouble).flatten()).norm() /
                  mag.to(torch::kDouble).norm().clamp_min(1e-30)).item<double>();
    std::printf("%-44s err / summed magnitude %.3e %s\n", what, err, err < tol ? "ok" : "FAIL");
    if (!(err < tol)) ++failures;
}

int main() {
    if (!torch::cuda::is_available()) {
        std::printf("no CUDA device\n");
        return 2;
    }
    const int N = 1500, K = 32, S = 3, T = 500;
    std::mt19937 rng(7);
    // symmetric random graph with self loops, a few heavy rows
    std::vector<std::vector<int>> adj(N);
    auto add = [&](int u, int v) { adj[u].push_back(v); };
    for (int i = 0; i < N; ++i) add(i, i);
    for (int e = 0; e < 40000; ++e) {
        int u = (int)(std::pow(rng() / 4294967296.0, 3.0) * N), v = rng() % N;
        if (u == v) continue;
        add(u, v);
        add(v, u);
    }
    std::vector<int> off(S * (N + 1), 0), cols, bounds(2 * S);
    std::vector<float> vals;
    std::vector<int64_t> coo_r, coo_c;
    for (auto& r : adj) {
        std::sort(r.begin(), r.end());
        r.erase(std::unique(r.begin(), r.end()), r.end());
    }
    std::uniform_real_distribution<float> U(0.1f, 1.0f);
    std::vector<float> coo_v;
    for (int s = 0; s < S; ++s) {   // column-tiled layout of src/ops/tiling.h:222-283
        bounds[2 * s] = (int)cols.size();
        int base = (int)cols.size();
        for (int i = 0; i < N; ++i) {
            off[s * (N + 1) + i] = (int)cols.size() - base;
            for (int c : adj[i])
                if (c >= s * T && c < (s + 1) * T) {
                    cols.push_back(c);
                    float w = U(rng);
                    vals.push_back(w);
                    coo_r.push_back(i);
                    coo_c.push_back(c);
                    coo_v.push_back(w);
                }
        }
        off[s * (N + 1) + N] = (int)cols.size() - base;
        bounds[2 * s + 1] = (int)cols.size();
    }
    const int64_t E = (int64_t)cols.size();
    auto dev = torch::Device(torch::kCUDA, 0);
    auto oi = torch::TensorOptions().dtype(torch::kInt);
    auto of = torch::TensorOptions().dtype(torch::kFloat);
    global_nrows = N;
    global_ra = 5;
    global_rb = 7;
    torch::Tensor t_off = torch::from_blob(off.data(), {(int64_t)off.size()}, oi).clone().to(dev);
    torch::Tensor t_col = torch::from_blob(cols.data(), {E}, oi).clone().to(dev);
    torch::Tensor t_val = torch::from_blob(vals.data(), {E}, of).clone().to(dev);
    torch::Tensor t_bnd = torch::from_blob(bounds.data(), {2 * S}, oi).clone();   // stays on the CPU
    global_offset_graph.push_back(t_off);
    global_columns_graph.push_back(t_col);
    global_value_graph.push_back(t_val);
    global_bounds.push_back(t_bnd);
    global_segments.push_back(S);

    auto ol = torch::TensorOptions().dtype(torch::kLong);
    torch::Tensor r64 = torch::from_blob(coo_r.data(), {E}, ol).clone().to(dev);
    torch::Tensor c64 = torch::from_blob(coo_c.data(), {E}, ol).clone().to(dev);
    torch::Tensor v32 = torch::from_blob(coo_v.data(), {E}, of).clone().to(dev);
    torch::Tensor A_w = torch::zeros({N, N}, of.device(dev).dtype(torch::kDouble));
    A_w.index_put_({r64, c64}, v32.to(torch::kDouble));
    torch::Tensor A_1 = (A_w != 0).to(torch::kDouble);

    torch::manual_seed(0);
    torch::Tensor X = torch::rand({N, K}, of.device(dev)) - 0.5;
    torch::Tensor Xd = X.to(torch::kDouble);

    expect_close("aggregate_node_mul_sum_coarse2_call",
                 aggregate_node_mul_sum_coarse2_call(X, t_off, t_col, t_val, t_bnd, S), A_w.mm(Xd), 1e-5);
    expect_close("aggregate_node_mul_sum_direct_coarse2_call",
                 aggregate_node_mul_sum_direct_coarse2_call(X, t_off, t_col, t_val, t_bnd, S), A_1.mm(Xd), 1e-5);
    torch::Tensor ones = torch::ones({N, 1}, of.device(dev));
    torch::Tensor deg = aggregate_node_mul_sum_direct_coarse2_call(ones, t_off, t_col, t_val, t_bnd, S);
    expect_close("degrees via ones (codegen/gala.cu:437)", deg, A_1.sum(1, true), 1e-7);

    torch::Tensor aL = torch::randn({N, 1}, of.device(dev)), aR = torch::randn({N, 1}, of.device(dev));
    torch::Tensor att = edge_sddvv(aL, aR, t_off, t_col, t_val, t_bnd, N, S);
    torch::Tensor att_want = aL.index({r64, 0}) + aR.index({c64, 0});
    expect_close("edge_sddvv", att, att_want, 1e-7);
    att = torch::leaky_relu(att, 0.2);
    // body of non_lnr_op_softmax_AutoGrad::forward (common.h:760-773)
    torch::Tensor val_exp = torch::clamp(torch::exp(att), 0.0, 1e12);
    torch::Tensor row_sum = node_spmv_backward_of_sddmm_nln(t_off, t_col, val_exp, t_bnd, global_nrows, S);
    torch::Tensor rs_want = torch::zeros({N}, of.device(dev).dtype(torch::kDouble)).index_add_(0, r64, val_exp.to(torch::kDouble));
    expect_close("node_spmv_backward_of_sddmm_nln", row_sum.flatten(), rs_want, 1e-5);
    row_sum = torch::reciprocal(row_sum);
    val_exp = inplace_softmax_sddvv(row_sum, t_off, t_col, val_exp, t_bnd, global_nrows, S);
    torch::Tensor alpha_want = torch::exp(att.to(torch::kDouble)) / rs_want.index({r64});
    expect_close("inplace_softmax_sddvv (softmax)", val_exp, alpha_want, 1e-5);
    torch::Tensor A_att = torch::zeros({N, N}, of.device(dev).dtype(torch::kDouble));
    A_att.index_put_({r64, c64}, alpha_want);
    expect_close("weighted aggregate with attention",
                 aggregate_node_mul_sum_coarse2_call(X, t_off, t_col, val_exp, t_bnd, S), A_att.mm(Xd), 1e-5);
    torch::Tensor alpha;
    torch::Tensor fused = gala_b200::gat_forward(X, aL, aR, t_off, t_col, t_bnd, S, 0.2f, false, &alpha);
    expect_close("gala_b200::gat_forward (fused layer)", fused, A_att.mm(Xd), 1e-5);
    expect_close("gala_b200::gat_forward alpha", alpha, alpha_want, 1e-5);

    torch::Tensor dZ = torch::rand({N, K}, of.device(dev)) - 0.5;
    torch::Tensor dd = edge_sddmm(dZ, X, t_off, t_col, t_val, t_bnd, global_nrows, S);
    torch::Tensor dd_want = (dZ.to(torch::kDouble).index({r64}) * Xd.index({c64})).sum(1);
    expect_close("edge_sddmm", dd, dd_want, 1e-5);

    // ---- gat_layer_AutoGrad: one node == the four emitted nodes, forward and backward ----------
    // slot 1 = backward graph of slot 0; for undirected inputs the generated main pushes the same
    // tensors twice (cuda.h:1129-1138)
    global_offset_graph.push_back(t_off);
    global_columns_graph.push_back(t_col);
    global_value_graph.push_back(t_val);
    global_bounds.push_back(t_bnd);
    global_segments.push_back(S);
    {
        torch::Tensor Xg = X.clone().requires_grad_(true), aLg = aL.clone().requires_grad_(true),
                      aRg = aR.clone().requires_grad_(true);
        torch::Tensor Yf = gala_b200::gat_layer_AutoGrad::apply(Xg, aLg, aRg, 0, 0.2);
        expect_close("gat_layer_AutoGrad forward", Yf, A_att.mm(Xd), 1e-5);
        Yf.backward(dZ);
        // what autograd evaluates for the emitted chain (common.h:654-670, 791-799, 876-889), op by op
        torch::Tensor dX_w = aggregate_node_mul_sum_coarse2_call(dZ, t_off, t_col, alpha, t_bnd, S);
        torch::Tensor da_w = edge_sddmm(dZ, X, t_off, t_col, alpha, t_bnd, global_nrows, S);
        torch::Tensor sds = alpha * da_w;
        torch::Tensor accum = node_spmv_backward_of_sddmm_nln(t_off, t_col, sds, t_bnd, global_nrows, S);
        torch::Tensor sm = sds - inplace_softmax_sddvv_mult(accum, t_off, t_col, alpha.clone(), t_bnd, global_nrows, S);
        torch::Tensor pre = edge_sddvv(aL, aR, t_off, t_col, t_val, t_bnd, N, S);
        torch::Tensor de = torch::where(pre > 0, sm, sm * 0.2);
        torch::Tensor datt_w = node_spmv_backward_of_sddmm_eaggr(t_off, t_col, de, t_bnd, global_nrows, S);
        expect_close("gat_layer_AutoGrad d(res)", Xg.grad(), dX_w, 1e-6);
        expect_close("gat_layer_AutoGrad d(attenL)", aLg.grad(), datt_w, 2e-4);   // cancelling row sums
        expect_close("gat_layer_AutoGrad d(attenR)", aRg.grad(), datt_w, 2e-4);

        // ---- the same layer against DENSE libtorch autograd in fp64 (an arbiter that shares no code with the
        // wrappers above): masked [N,N] attention written with plain ATen ops as the generated program computes
        // it (common.h:622-675, 1176-1184, 760-773, 835-894), differentiated by torch.
        auto od = of.device(dev).dtype(torch::kDouble);
        torch::Tensor mask = (A_w != 0).to(torch::kDouble);
        torch::Tensor aLd = aL.to(torch::kDouble).detach().requires_grad_(true);
        torch::Tensor aRd = aR.to(torch::kDouble).detach().requires_grad_(true);
        torch::Tensor Xdd = Xd.detach().clone().requires_grad_(true);
        torch::Tensor act = torch::leaky_relu(aLd + aRd.t(), 0.2);
        act.retain_grad();
        torch::Tensor num = torch::clamp(torch::exp(act), 0.0, 1e12) * mask;
        torch::Tensor alpha_d = num / (num.sum(1, true) + S * 1e-12);
        alpha_d.retain_grad();
        torch::Tensor Yd = alpha_d.mm(Xdd);
        Yd.backward(dZ.to(torch::kDouble));
        expect_close("dense autograd: forward", Yf, Yd, 1e-5);
        expect_close("dense autograd: d(alpha) = edge_sddmm", da_w, alpha_d.grad().index({r64, c64}), 1e-5);
        expect_close("dense autograd: softmax backward", sm, act.grad().index({r64, c64}), 1e-5);
        expect_close_mag("dense autograd: d(attenL)", aLg.grad(), aLd.grad(), (act.grad().abs() * mask).sum(1), 1e-5);
        // reference semantics of d(res): the forward alpha over slot 2li+1, which is the forward graph itself for
        // undirected inputs (common.h:876-885) -> alpha @ dZ; the mathematical gradient is alpha^T @ dZ
        expect_close("reference semantics: d(res) = alpha @ dZ", Xg.grad(), alpha_d.detach().mm(dZ.to(torch::kDouble)), 1e-5);
        (void)od;
    }

    // ---- gat_layer_AutoGrad with the ReLU inside the kernel == layer followed by torch::relu -------------
    {
        torch::Tensor X1 = X.clone().requires_grad_(true), l1 = aL.clone().requires_grad_(true), r1 = aR.clone().requires_grad_(true);
        torch::Tensor X2 = X.clone().requires_grad_(true), l2 = aL.clone().requires_grad_(true), r2 = aR.clone().requires_grad_(true);
        torch::Tensor y1 = gala_b200::gat_layer_AutoGrad::apply(X1, l1, r1, 0, 0.2, true);
        torch::Tensor y2 = torch::relu(gala_b200::gat_layer_AutoGrad::apply(X2, l2, r2, 0, 0.2, false));
        expect_close("gat_layer_AutoGrad fused ReLU: forward", y1, y2, 1e-7);
        y1.backward(dZ);
        y2.backward(dZ);
        expect_close("gat_layer_AutoGrad fused ReLU: d(res)", X1.grad(), X2.grad(), 1e-6);
        expect_close("gat_layer_AutoGrad fused ReLU: d(attenL)", l1.grad(), l2.grad(), 1e-6);
    }

    // ---- gala_b200::Linear == torch::nn::Linear (same init under the same seed, same gradients) -----------
    for (auto shape : std::vector<std::pair<int, int>>{{602, 32}, {32, 41}, {41, 1}, {100, 32}, {32, 47}, {300, 200}}) {
        const int Fin = shape.first, Fout = shape.second;
        torch::manual_seed(11);
        torch::nn::Linear ref(Fin, Fout);
        torch::manual_seed(11);
        gala_b200::Linear ours(Fin, Fout);
        ref->to(dev);
        ours->to(dev);
        expect_close("gala_b200::Linear init == torch::nn::Linear", ours->weight, ref->weight, 1e-12);
        torch::Tensor in1 = (torch::rand({N, Fin}, of.device(dev)) - 0.5).requires_grad_(true);
        torch::Tensor in2 = in1.detach().clone().requires_grad_(true);
        torch::Tensor o1 = ours->forward(in1), o2 = ref->forward(in2);
        char what[96];
        std::snprintf(what, sizeof(what), "gala_b200::Linear [%d -> %d] forward", Fin, Fout);
        expect_close(what, o1, at::linear(in2.to(torch::kDouble), ref->weight.to(torch::kDouble), ref->bias.to(torch::kDouble)), 1e-5);
        torch::Tensor go = torch::rand({N, Fout}, of.device(dev)) - 0.5;
        o1.backward(go);
        o2.backward(go);
        std::snprintf(what, sizeof(what), "gala_b200::Linear [%d -> %d] d(weight)", Fin, Fout);
        expect_close(what, ours->weight.grad(), ref->weight.grad(), 1e-5);
        std::snprintf(what, sizeof(what), "gala_b200::Linear [%d -> %d] d(input)", Fin, Fout);
        expect_close(what, in1.grad(), in2.grad(), 1e-5);
        expect_close("gala_b200::Linear d(bias)", ours->bias.grad(), ref->bias.grad(), 1e-5);
    }
    // ---- linear_att / folded_att == the three Linear calls they replace (values and every gradient) -------
    {
        torch::manual_seed(12);
        gala_b200::Linear fc(64, 32), el(32, 1), er(32, 1);
        fc->to(dev);
        el->to(dev);
        er->to(dev);
        torch::Tensor in = torch::rand({N, 64}, of.device(dev)) - 0.5;
        torch::Tensor g0 = torch::rand({N, 32}, of.device(dev)) - 0.5, g1 = torch::rand({N, 1}, of.device(dev)) - 0.5,
                      g2 = torch::rand({N, 1}, of.device(dev)) - 0.5;
        auto grads_of = [&]() {
            std::vector<torch::Tensor> out;
            for (auto& m : {fc, el, er})
                for (auto& p : m->parameters()) {
                    out.push_back(p.grad().clone());
                    p.mutable_grad() = torch::Tensor();
                }
            return out;
        };
        torch::Tensor in_a = in.clone().requires_grad_(true);
        torch::Tensor ra, la, rra;
        std::tie(ra, la, rra) = gala_b200::linear_att(fc, el, er, in_a);
        (ra * g0).sum().add((la * g1).sum()).add((rra * g2).sum()).backward();
        auto ga = grads_of();
        torch::Tensor in_b = in.clone().requires_grad_(true);
        torch::Tensor rb = at::linear(in_b, fc->weight, fc->bias);
        torch::Tensor lb = at::linear(rb, el->weight, el->bias), rrb = at::linear(rb, er->weight, er->bias);
        (rb * g0).sum().add((lb * g1).sum()).add((rrb * g2).sum()).backward();
        auto gb = grads_of();
        expect_close("linear_att: res", ra, rb, 1e-5);
        expect_close("linear_att: attenL", la, lb, 1e-5);
        expect_close("linear_att: attenR", rra, rrb, 1e-5);
        expect_close("linear_att: d(input)", in_a.grad(), in_b.grad(), 1e-5);
        for (size_t i = 0; i < ga.size(); ++i) expect_close("linear_att: d(parameter)", ga[i], gb[i], 1e-5);
        // folded projections (layer 2): only the two logits are produced
        torch::Tensor in_c = in.clone().requires_grad_(true), in_d = in.clone().requires_grad_(true);
        torch::Tensor lc, rc;
        std::tie(lc, rc) = gala_b200::folded_att(fc, el, er, in_c);
        (lc * g1).sum().add((rc * g2).sum()).backward();
        auto gc = grads_of();
        torch::Tensor rd = at::linear(in_d, fc->weight, fc->bias);
        torch::Tensor ld = at::linear(rd, el->weight, el->bias), rrd = at::linear(rd, er->weight, er->bias);
        (ld * g1).sum().add((rrd * g2).sum()).backward();
        auto gd = grads_of();
        expect_close("folded_att: attenL", lc, ld, 1e-5);
        expect_close("folded_att: attenR", rc, rrd, 1e-5);
        expect_close("folded_att: d(input)", in_c.grad(), in_d.grad(), 1e-5);
        for (size_t i = 0; i < gc.size(); ++i) expect_close("folded_att: d(parameter)", gc[i], gd[i], 1e-5);
    }

    torch::Tensor norm = torch::pow(deg, -0.5);
    torch::Tensor ev = aggregate_edge_mul(norm, norm, t_off, t_col, t_val, t_bnd, S);
    expect_close("aggregate_edge_mul", ev, norm.index({r64, 0}) * norm.index({c64, 0}), 1e-6);

    torch::Tensor ys = aggregate_node_mul_sum_sample20_call(X, t_off, t_col, t_val, t_bnd, S);
    // reference semantics (cuda.h:313-320): per row AND per segment, 20 picks j=(5*ji+7)%deg
    torch::Tensor want = torch::zeros({N, K}, of.dtype(torch::kDouble));
    torch::Tensor Xc = Xd.cpu();
    for (int s = 0; s < S; ++s)
        for (int i = 0; i < N; ++i) {
            int b = off[s * (N + 1) + i], d = off[s * (N + 1) + i + 1] - b;
            if (d <= 0) continue;
            for (int ji = 0; ji < 20; ++ji) want[i] += Xc[cols[bounds[2 * s] + b + (5 * ji + 7) % d]];
        }
    expect_close("sampled aggregate (global_ra/rb = 5/7)", ys, want.to(dev), 1e-5);

    std::printf(failures ? "SHIM SELFTEST FAILED (%d)\n" : "SHIM SELFTEST OK\n", failures);
    return failures ? 1 : 0;
}
