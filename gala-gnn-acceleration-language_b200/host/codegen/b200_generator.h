// b200_generator.h -- the retargeted code generator: GALA's CUDAGenerator
// (reference src/codegen/cuda.h) with the three edits INTEGRATION.md section 3 describes.
// Everything else (model emission, autograd classes, data prep, H2D transfer, main) is
// inherited unchanged from the reference, so generated programs differ from the reference's
// only in that the kernel / launch-tree text is replaced by calls into libgala_b200.so.
//
// Compiled against the reference tree (-I/root/reference); no reference source is copied.
#pragma once
#include <fstream>
#include <regex>
#include <sstream>
#include <unordered_set>

#include "src/codegen/cuda.h"

class B200Generator : public CUDAGenerator {
public:
    // deviceFormats: the generated main prepares its graphs on the GPU (host/gala_b200_host.h re-creates the
    // reference's host-side classes and functions under their own names on top of the C-ABI) instead of
    // including the reference's CPU headers.
    B200Generator(GALAContext* context, std::string& outputPath, const std::string& galaB200Root, bool deviceFormats = true)
        : CUDAGenerator(context, outputPath), root_(galaB200Root), deviceFormats_(deviceFormats) {}

    // replaces CUDAGenerator::initCMake (cuda.h:18-56): no icpx, sm_100a, link libgala_b200
    void initCMake() override {
        std::string cm =
            "cmake_minimum_required(VERSION 3.18 FATAL_ERROR)\n"
            "project(gala_cuda LANGUAGES CUDA CXX)\n"
            "set(CMAKE_CUDA_ARCHITECTURES 100a)\n"
            "find_package(Torch REQUIRED)\n"
            "find_package(OpenMP REQUIRED)\n"
            "include_directories(${CMAKE_CUDA_TOOLKIT_INCLUDE_DIRECTORIES} " + root_ + "/include " + root_ +
            "/gala-gnn-acceleration-language_b200/host)\n"
            "link_directories(" + root_ + "/gala-gnn-acceleration-language_b200)\n"
            "link_libraries(\"${TORCH_LIBRARIES}\" cudart gala_b200 OpenMP::OpenMP_CXX)\n"
            "add_compile_options(-Xcompiler -fopenmp -O3)\n"
            "add_compile_definitions(GALA_TORCH GN_1 PT_0 ST_0 A_ALLOC)\n"
            "add_executable(gala_model gala.cu)\n"
            "target_compile_features(gala_model PRIVATE cxx_std_17)";
        cmakeCode.addCode(cm);
    }

    // replaces CUDAGenerator::initKernels (cuda.h:957-1050) + generateCudaCodeForCNode (:170-955)
    void initKernels(std::vector<CIRNode*>& program) override {
        std::string imports =
            "#include <cuda_runtime_api.h>\n"
            "#include <torch/script.h>\n"
            "#include <cmath>\n#include <iostream>\n#include <parallel/algorithm>\n#include <vector>\n"
            "#include <bits/stdc++.h>\n#include <omp.h>\n#include <stdlib.h>\n#include <torch/torch.h>\n";
        if (deviceFormats_) {
            // formats, tiling, sampling, sub-graphs and the .npy readers: same names, built on the GPU
            imports += "#include \"gala_b200_host.h\"\n";
        } else {
            imports +=
                "#include \"../src/formats/csrc_matrix.h\"\n"
                "#include \"../src/formats/dense_matrix.h\"\n"
                "#include \"../src/ops/aggregators.h\"\n"
                "#include \"../src/ops/tiling.h\"\n"
                "#include \"../src/utils/mtx_io.h\"\n"
                "#include \"../tests/common.h\"\n";
        }
        importCode.addCode(imports);

        std::string prelude =
            "\n#define CUDA_CHECK(func)\\\n"
            "  do {\\\n"
            "    cudaError_t status = (func);\\\n"
            "    if (status != cudaSuccess) {\\\n"
            "      printf(\"CUDA API failed at line %d with error: %s (%d)\\n\", __LINE__,\\\n"
            "             cudaGetErrorString(status), status);\\\n"
            "      exit(EXIT_FAILURE);\\\n"
            "    }\\\n"
            "  } while (0)\n"
            "// sparse kernels: libgala_b200.so through the libtorch shim (INTEGRATION.md)\n"
            "#include \"gala_b200_torch.h\"\n";
        if (GALAFEContext::print_memory) {
            prelude +=
                "int printMemoryUsage() {\n  size_t freeMem, totalMem;\n  cudaMemGetInfo(&freeMem, &totalMem);\n"
                "  return (int)((totalMem - freeMem) / (1024 * 1024));\n}\n";
        }
        kernelCode.addCode(prelude);

        std::unordered_set<std::string> seen;
        auto visit = [&](CIRNode* node) {
            auto cNode = dynamic_cast<ComputeNode*>(node);
            if (!cNode) return;
            std::string name = getKernelName(cNode);
            if (seen.insert(name).second) emitBinding(cNode, name);
        };
        for (CIRNode* node : program) {
            if (dynamic_cast<ComputeNode*>(node)) {
                visit(node);
            } else if (auto loop = dynamic_cast<TrainingLoopNode*>(node)) {
                for (int ix = 0; ix < loop->getLoopNodeNum(); ix++) visit(loop->getNode(ix));
            }
        }
    }

    // Peephole over the emitted forward() text, run after writeCode(): a GAT layer whose four nodes
    // (aggregate_edge_sum -> LeakyReLU -> non_lnr_op_softmax -> aggregate_node_mul_sum*, emitted by
    // common.h:622-675, 1176-1184, 735-810, 835-927) all address the same graph slot becomes ONE call of
    // gala_b200::gat_layer_AutoGrad (gala_b200_torch.h): fused forward kernel, same gradients.  Layers
    // whose nodes use different slots (training sub-graphs, tests/common.h:20-105) are left as emitted.
    // Returns the number of layers fused.
    static int fuseGatLayers(const std::string& path) {
        std::ifstream in(path);
        if (!in) return 0;
        std::stringstream buf;
        buf << in.rdbuf();
        std::string text = buf.str();
        in.close();
        const std::regex layer(
            R"(attn = aggregate_edge_sum_AutoGrad::apply\((\w+), (\w+), (\d+)\);\s*)"
            R"((torch::nn::LeakyReLU leaky_relu\(torch::nn::LeakyReLUOptions\(\)\.negative_slope\(([0-9.]+)\)\);\s*)?)"
            R"(attn = leaky_relu->forward\(attn\);\s*)"
            R"(attn = non_lnr_op_softmax_AutoGrad::apply\(attn, (\d+)\);\s*)"
            R"(if \(ep % mod_v == 0\) \{\s*res = \w+_AutoGrad::apply\(res, attn, (\d+)\);\s*\} else \{\s*)"
            R"(res = \w+_AutoGrad::apply\(res, attn, (\d+)\);\s*\})");
        std::string out, slope = "0.2";
        int fused = 0;
        auto it = std::sregex_iterator(text.begin(), text.end(), layer);
        size_t pos = 0;
        for (; it != std::sregex_iterator(); ++it) {
            const std::smatch& m = *it;
            out += text.substr(pos, m.position() - pos);
            pos = m.position() + m.length();
            if (m[5].matched) slope = m[5].str();
            const bool same = m[3] == m[6] && m[3] == m[7] && m[3] == m[8];
            if (!same) {
                out += m.str();
                continue;
            }
            out += "res = gala_b200::gat_layer_AutoGrad::apply(res, " + m[1].str() + ", " + m[2].str() + ", " + m[3].str() +
                   ", " + slope + ");   // fused: edge_sddvv + LeakyReLU + softmax + aggregate";
            ++fused;
        }
        out += text.substr(pos);
        if (fused) {
            std::ofstream o(path);
            o << out;
        }
        return fused;
    }

private:
    std::string root_;
    bool deviceFormats_;

    // One line per aggregation flavour instead of kernel text + launch tree; the edge ops
    // (softmax / edge-sum / edge-mul nodes) need nothing: the shim defines their fixed names.
    void emitBinding(ComputeNode* cNode, const std::string& name) {
        if (cNode->getOpType() != AGGREGATE_NODE) return;
        bool weighted = cNode->getInput(1)->getDataInfo()->getWeighted();
        bool colTile = hasDOpt(cNode->getInput(1), COL_TILE_DOPT);
        int nsamples = 0;
        for (auto opt : *cNode->getOpts())
            if (opt.first == SAMPLE_COPT || opt.first == SAMPLE_DYNAMIC_COPT) nsamples = (int)opt.second;
        std::string line = std::string(colTile ? "GALA_B200_DEFINE_AGGREGATE_TILED(" : "GALA_B200_DEFINE_AGGREGATE(") +
                           name + "_call, " + (weighted ? "true" : "false") + ", " + std::to_string(nsamples) + ")\n";
        kernelCallCode.addCode(line);
        cNode->setKernelName("gather_forward");
    }
};
