// ref_ops_harness.cu -- runs the REFERENCE's own emitted CUDA kernels, wrappers and autograd classes op by
// op on inputs read from raw files, and dumps every result, so that tests/test_ref_kernels_gpu.py can compare
// the oracle restatement (oracle/gala_oracle.c) and the B200 kernels with them element by element.
//
// Nothing of the reference lives in this file: the kernel text is whatever the STOCK CUDAGenerator
// (reference src/codegen/cuda.h:170-955, common.h:622-977) emitted at build time into
// _models/<program>_ref/build/gala.cu; that file is #included below with its main() renamed.  Built in the
// authoring container by host/codegen/build_models.sh (HARNESS=1); the binary lands in the git-ignored
// _models/ and travels to the GPU box.
//
//   nvcc ... -DGALA_GENERATED='"<path>/gala.cu"' -DHARNESS_GAT | -DHARNESS_AGG [-DHARNESS_DIRECT] [-DHARNESS_EDGE_MUL]
//   ref_ops_harness_<kind> <case_dir>
//
// case_dir holds  meta.txt ("nrows nvals segments nK K0 K1 ...")  and raw little-endian arrays
// offsets.i32 [S*(N+1)], cols.i32 [E], bounds.i32 [2S], vals.f32 [E], aL.f32 aR.f32 [N], dalpha.f32 [E],
// X<K>.f32 [N*K], dZ<K>.f32 [N*K] (rows constant inside blocks of 8: the reference's SDDMM kernel shares one
// shared-memory copy of the row operand between the 8 rows of a block, cuda.h:706-714, so only such inputs
// have a defined result).  Outputs are written next to them as ref_<name>.f32.
// Run with CUDA_LAUNCH_BLOCKING=1: the emitted wrappers launch every column segment on its own fresh stream
// and the segment kernels read-modify-write the same rows (cuda.h:470-476).
#define main gala_generated_main
#include GALA_GENERATED
#undef main

#include <cstdio>
#include <fstream>
#include <sstream>

namespace harness {

template <class T>
std::vector<T> read_raw(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) {
        std::fprintf(stderr, "harness: cannot read %s\n", path.c_str());
        std::exit(3);
    }
    const std::streamsize bytes = f.tellg();
    f.seekg(0);
    std::vector<T> v((size_t)bytes / sizeof(T));
    f.read(reinterpret_cast<char*>(v.data()), bytes);
    return v;
}

torch::Tensor dev_f32(const std::string& path, std::vector<int64_t> shape) {
    auto v = read_raw<float>(path);
    return torch::from_blob(v.data(), {(int64_t)v.size()}, torch::kFloat).clone().to(torch::kCUDA).reshape(shape);
}

torch::Tensor dev_i32(const std::string& path) {
    auto v = read_raw<int>(path);
    return torch::from_blob(v.data(), {(int64_t)v.size()}, torch::kInt).clone().to(torch::kCUDA);
}

void dump(const std::string& dir, const std::string& name, const torch::Tensor& t) {
    cudaDeviceSynchronize();
    auto c = t.detach().to(torch::kCPU).to(torch::kFloat).contiguous();
    std::ofstream f(dir + "/ref_" + name + ".f32", std::ios::binary);
    f.write(reinterpret_cast<const char*>(c.data_ptr<float>()), c.numel() * sizeof(float));
}

}  // namespace harness

int main(int argc, char** argv) {
    using namespace harness;
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <case_dir>\n", argv[0]);
        return 2;
    }
    const std::string dir = argv[1];
    int nrows = 0, segments = 0, nk = 0;
    long nvals = 0;
    std::vector<int> Ks;
    {
        std::ifstream m(dir + "/meta.txt");
        m >> nrows >> nvals >> segments >> nk;
        Ks.resize(nk);
        for (int i = 0; i < nk; ++i) m >> Ks[i];
    }
    global_nrows = nrows;
    global_ra = 5;   // common.h:817-818
    global_rb = 7;
    torch::Tensor offsets = dev_i32(dir + "/offsets.i32");
    torch::Tensor cols = dev_i32(dir + "/cols.i32");
    auto bounds_v = read_raw<int>(dir + "/bounds.i32");
    torch::Tensor bounds = torch::from_blob(bounds_v.data(), {(int64_t)bounds_v.size()}, torch::kInt).clone();   // CPU, as emitted
    torch::Tensor vals = dev_f32(dir + "/vals.f32", {nvals});
    // slot 0 = forward graph, slot 1 = backward graph (same tensors: undirected, cuda.h:1129-1138)
    for (int s = 0; s < 2; ++s) {
        global_offset_graph.push_back(offsets);
        global_columns_graph.push_back(cols);
        global_value_graph.push_back(vals);
        global_bounds.push_back(bounds);
        global_segments.push_back(segments);
    }

#ifdef HARNESS_AGG
    for (int K : Ks) {
        torch::Tensor X = dev_f32(dir + "/X" + std::to_string(K) + ".f32", {nrows, K});
        dump(dir, "agg" + std::to_string(K),
             aggregate_node_mul_sum_coarse2_call(X, offsets, cols, vals, bounds, segments));
    }
#ifdef HARNESS_DIRECT
    {
        torch::Tensor ones = torch::ones({nrows, 1}, torch::TensorOptions().dtype(torch::kFloat).device(torch::kCUDA, 0));
        dump(dir, "degrees", aggregate_node_mul_sum_direct_coarse2_call(ones, offsets, cols, vals, bounds, segments));
    }
#endif
#ifdef HARNESS_EDGE_MUL
    {
        torch::Tensor a = dev_f32(dir + "/aL.f32", {nrows, 1}), b = dev_f32(dir + "/aR.f32", {nrows, 1});
        dump(dir, "edge_mul", aggregate_edge_mul(a, b, offsets, cols, vals, bounds, segments));
    }
#endif
#endif

#ifdef HARNESS_GAT
    torch::Tensor aL = dev_f32(dir + "/aL.f32", {nrows, 1}), aR = dev_f32(dir + "/aR.f32", {nrows, 1});
    torch::Tensor dalpha = dev_f32(dir + "/dalpha.f32", {nvals});
    // K5, K3, K4 one by one
    torch::Tensor logits = edge_sddvv(aL, aR, offsets, cols, vals, bounds, nrows, segments);
    dump(dir, "sddvv", logits);
    dump(dir, "rowsum", node_spmv_backward_of_sddmm_nln(offsets, cols, vals, bounds, nrows, segments));
    dump(dir, "rowsum_eaggr", node_spmv_backward_of_sddmm_eaggr(offsets, cols, vals, bounds, nrows, segments));
    {
        torch::Tensor v = vals.clone();
        dump(dir, "scale_rows", inplace_softmax_sddvv(aL, offsets, cols, v, bounds, nrows, segments));
    }
    // edge-softmax composite, forward and backward through the emitted autograd class (common.h:735-810)
    {
        torch::Tensor x = dev_f32(dir + "/vals.f32", {nvals}).detach().requires_grad_(true);
        torch::Tensor alpha = non_lnr_op_softmax_AutoGrad::apply(x, 0);
        dump(dir, "softmax_fwd", alpha);
        alpha.backward(dalpha);
        dump(dir, "softmax_bwd", x.grad());
    }
    for (int K : Ks) {
        const std::string k = std::to_string(K);
        torch::Tensor X = dev_f32(dir + "/X" + k + ".f32", {nrows, K});
        torch::Tensor dZ = dev_f32(dir + "/dZ" + k + ".f32", {nrows, K});
        // K1 weighted, K6 (block-constant row operand, see the header)
        dump(dir, "agg" + k, aggregate_node_mul_sum_coarse2_call(X, offsets, cols, vals, bounds, segments));
        dump(dir, "sddmm" + k, edge_sddmm(dZ, X, offsets, cols, vals, bounds, nrows, segments));
        // one whole GAT layer through the emitted autograd chain (common.h:622-675, 735-810, 835-894)
        torch::Tensor res = X.detach().clone().requires_grad_(true);
        torch::Tensor l = aL.detach().clone().requires_grad_(true), r = aR.detach().clone().requires_grad_(true);
        torch::Tensor attn = aggregate_edge_sum_AutoGrad::apply(l, r, 0);
        torch::nn::LeakyReLU leaky_relu(torch::nn::LeakyReLUOptions().negative_slope(0.2));
        attn = leaky_relu->forward(attn);
        attn = non_lnr_op_softmax_AutoGrad::apply(attn, 0);
        torch::Tensor y = aggregate_node_mul_sum_coarse2_AutoGrad::apply(res, attn, 0);
        dump(dir, "gat_alpha" + k, attn);
        dump(dir, "gat_y" + k, y);
        y.backward(dZ);
        dump(dir, "gat_dres" + k, res.grad());
        dump(dir, "gat_daL" + k, l.grad());
        dump(dir, "gat_daR" + k, r.grad());
    }
#endif
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::fprintf(stderr, "harness: CUDA error %s\n", cudaGetErrorString(e));
        return 1;
    }
    std::printf("REF HARNESS OK\n");
    return 0;
}
