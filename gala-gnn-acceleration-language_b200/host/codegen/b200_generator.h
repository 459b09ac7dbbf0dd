// b200_generator.h -- the retargeted code generator: GALA's CUDAGenerator
// (reference src/codegen/cuda.h) with the three edits INTEGRATION.md section 3 describes.
// Everything else (model emission, autograd classes, data prep, H2D transfer, main) is
// inherited unchanged from the reference, so generated programs differ from the reference's
// only in that the kernel / launch-tree text is replaced by calls into libgala_b200.so.
//
// Compiled against the reference tree (-I/root/reference); no reference source is copied.
#pragma once
#include <fstream>
#include <iostream>
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

    // ---- writeCode with a fusion pass between IR lowering and file output -------------------------------
    // CodeGenerator::writeCode (common.h:1725-1764) is not virtual and writes the file as soon as generateCode()
    // returns, so the retargeted generator runs the same steps itself (the drivers call writeCodeB200) and, in
    // between, rewrites the statement list of GALAGNN::forward.  Every compute node of the IR contributes exactly
    // one entry to model.getForward() (common.h:511-1375: one addCode per node), so the pass works on whole
    // statements: the DECISION to fuse is taken on the IR (op sequence, graph slots, producers / consumers), the
    // statements the nodes produced are then parsed (callee + argument list, whitespace-insensitive) and replaced.
    // If the IR says "fusable" and the statements do not have the expected shape, code generation fails loudly
    // instead of silently emitting the slow sequence.
    struct FuseOptions {
        bool gatLayers = true;   // edge_sddvv + LeakyReLU + softmax + aggregate -> gala_b200::gat_layer_AutoGrad
        bool linears = true;     // torch::nn::Linear -> gala_b200::Linear (tcgen05 / streaming kernels), projections fused
        bool corruptForTest = false;   // test hook: mangle one emitted statement so that the loud-failure path can be exercised
    };

    void writeCodeB200(std::vector<CIRNode*>& program, std::vector<RelationEdge*>& dependencies,
                       std::vector<RelationEdge*>& associations, std::vector<TransformEdge*>& transforms,
                       const FuseOptions& opt, std::ostream& log) {
        initCMake();
        initKernels(program);
        commonPerCode();
        generateCode(program, transforms);
        fuseForward(program, opt, log);
        // the sections of gala.cu, in the order CodeGenerator::writeCode emits them (common.h:1742-1763)
        CodeGenerator::writeCode(cmakeCode, outStreamCMake);
        CodeGenerator::writeCode(importCode, outStreamModel);
        CodeGenerator::writeCode(kernelCode, outStreamModel);
        CodeGenerator::writeCode(kernelCallCode, outStreamModel);
        CodeGenerator::writeCode(*model.getDef(), outStreamModel);
        CodeGenerator::writeCode(*model.getInitCall(), outStreamModel, ", ", true, true);
        CodeGenerator::writeCode(*model.getInit(), outStreamModel);
        CodeGenerator::writeCode(*model.getForwardCallPre(), outStreamModel, "");
        CodeGenerator::writeCode(*model.getForwardCallInternal(), outStreamModel, "");
        CodeGenerator::writeCode(*model.getForwardCallPost(), outStreamModel);
        CodeGenerator::writeCode(*model.getForward(), outStreamModel);
        CodeGenerator::writeCode(preCode, outStreamModel);
        CodeGenerator::writeCode(*model.getInv(), outStreamModel);
        CodeGenerator::writeCode(*model.getPreCall(), outStreamModel, "");
        CodeGenerator::writeCode(*model.getCall(), outStreamModel, "");
        CodeGenerator::writeCode(*model.getPostCall(), outStreamModel);
        CodeGenerator::writeCode(postCode, outStreamModel);
        closeStream();
    }

private:
    // ---- statements ------------------------------------------------------------------------------------
    struct Call {                       // "lhs = callee(arg, ...);"
        std::string lhs, callee;
        std::vector<std::string> args;
    };
    static std::string squeeze(const std::string& t) {     // drop all whitespace
        std::string o;
        for (char c : t)
            if (!isspace((unsigned char)c)) o.push_back(c);
        return o;
    }
    static bool parseCall(const std::string& text, Call& c) {
        const std::string t = squeeze(text);
        if (t.empty() || t.back() != ';') return false;
        const size_t eq = t.find('='), lp = t.find('('), rp = t.rfind(')');
        if (eq == std::string::npos || lp == std::string::npos || rp != t.size() - 2 || eq > lp) return false;
        c.lhs = t.substr(0, eq);
        c.callee = t.substr(eq + 1, lp - eq - 1);
        c.args.clear();
        int depth = 0;
        std::string cur;
        for (size_t i = lp + 1; i < rp; ++i) {
            const char ch = t[i];
            if (ch == '(' || ch == '{') ++depth;
            if (ch == ')' || ch == '}') --depth;
            if (ch == ',' && depth == 0) {
                c.args.push_back(cur);
                cur.clear();
            } else {
                cur.push_back(ch);
            }
        }
        if (!cur.empty()) c.args.push_back(cur);
        for (char ch : c.lhs)
            if (!(isalnum((unsigned char)ch) || ch == '_')) return false;
        return true;
    }
    // "if (ep % mod_v == 0) { A } else { B }" as emitted for every aggregation inside the loop (common.h:1004-1010)
    static bool parseEpochSwitch(const std::string& text, Call& a, Call& b) {
        const std::string t = squeeze(text);
        const std::string head = "if(ep%mod_v==0){", mid = "}else{";
        if (t.compare(0, head.size(), head) != 0 || t.back() != '}') return false;
        const size_t m = t.find(mid);
        if (m == std::string::npos) return false;
        return parseCall(t.substr(head.size(), m - head.size()), a) &&
               parseCall(t.substr(m + mid.size(), t.size() - 1 - m - mid.size()), b);
    }
    static bool endsWith(const std::string& t, const std::string& suf) {
        return t.size() >= suf.size() && t.compare(t.size() - suf.size(), suf.size(), suf) == 0;
    }
    [[noreturn]] static void fail(const std::string& what, const std::string& stmt) {
        std::cerr << "B200Generator: the IR describes a fusable pattern but the emitted statement has an unexpected shape ("
                  << what << "):\n    " << stmt << "\nrefusing to emit the unfused sequence silently; run with --no-fuse "
                  << "or update host/codegen/b200_generator.h to the reference's new emission\n";
        std::exit(3);
    }

    // compute nodes of the forward pass in program order (what generateCode walks, common.h:1378-1470)
    static std::vector<ComputeNode*> forwardNodes(std::vector<CIRNode*>& program) {
        std::vector<ComputeNode*> out;
        for (CIRNode* node : program) {
            if (auto c = dynamic_cast<ComputeNode*>(node)) {
                out.push_back(c);
            } else if (auto loop = dynamic_cast<TrainingLoopNode*>(node)) {
                for (int ix = 0; ix < loop->getLoopNodeNum(); ix++)
                    if (auto c2 = dynamic_cast<ComputeNode*>(loop->getNode(ix))) out.push_back(c2);
            }
        }
        return out;
    }

    void fuseForward(std::vector<CIRNode*>& program, const FuseOptions& opt, std::ostream& log) {
        Code* fwd = model.getForward();
        std::vector<ComputeNode*> nodes = forwardNodes(program);
        auto line = [&](int i) -> std::string& { return *fwd->atLine(i); };
        auto blank = [&](int i) { line(i) = "        // (fused into the call above)"; };
        const int n = fwd->getNum();
        if (opt.corruptForTest)
            for (int i = 0; i < n; ++i)
                if (line(i).find("non_lnr_op_softmax_AutoGrad::apply") != std::string::npos) {
                    line(i) = "attn = some_new_softmax(attn);";
                    break;
                }

        // ---- (1) GAT layers: IR says AGGREGATE_EDGE_SUM -> LEAKY_RELU -> SOFTMAX -> AGGREGATE_MUL_SUM over one graph
        int gatLayersInIR = 0;
        for (size_t k = 0; k + 3 < nodes.size(); ++k)
            if (nodes[k]->getOp() == AGGREGATE_EDGE_SUM_OP && nodes[k + 1]->getOp() == NON_LNR_OP_LEAKY_RELU &&
                nodes[k + 2]->getOp() == NON_LNR_OP_SOFTMAX && nodes[k + 3]->getOp() == AGGREGATE_MUL_SUM_OP)
                ++gatLayersInIR;
        int fusedGat = 0, keptGat = 0;
        if (opt.gatLayers && gatLayersInIR > 0) {
            std::string slope = "0.2";
            for (int i = 0; i < n; ++i) {
                Call a;
                if (!parseCall(line(i), a) || !endsWith(a.callee, "aggregate_edge_sum_AutoGrad::apply")) continue;
                if (a.args.size() != 3) fail("edge-sum call: 3 arguments expected", line(i));
                // following statements: [LeakyReLU declaration] leaky_relu->forward, softmax apply, epoch switch
                int j = i + 1;
                if (j < n && squeeze(line(j)).find("torch::nn::LeakyReLUleaky_relu(") == 0) {
                    const std::string t = squeeze(line(j));
                    const size_t p = t.find("negative_slope(");
                    if (p == std::string::npos) fail("LeakyReLU declaration without negative_slope", line(j));
                    slope = t.substr(p + 15, t.find(')', p) - p - 15);
                    ++j;
                }
                Call l, sm, s0, s1;
                if (j + 2 >= n + 0 && j + 2 > n - 1 + 0) fail("GAT layer: statements missing after the edge sum", line(i));
                if (!parseCall(line(j), l) || l.callee != "leaky_relu->forward" || l.args.size() != 1 || l.args[0] != a.lhs)
                    fail("LeakyReLU call", line(j));
                if (!parseCall(line(j + 1), sm) || !endsWith(sm.callee, "non_lnr_op_softmax_AutoGrad::apply") ||
                    sm.args.size() != 2 || sm.args[0] != a.lhs)
                    fail("edge-softmax call", line(j + 1));
                if (!parseEpochSwitch(line(j + 2), s0, s1) || s0.args.size() != 3 || s1.args.size() != 3 ||
                    s0.args[1] != a.lhs || s1.args[1] != a.lhs || s0.lhs != s1.lhs || s0.args[0] != s1.args[0])
                    fail("weighted aggregation (epoch switch)", line(j + 2));
                // all four nodes must address the same graph slot (training sub-graphs give the aggregation its own)
                const bool same = a.args[2] == sm.args[1] && a.args[2] == s0.args[2] && a.args[2] == s1.args[2];
                if (!same) {
                    ++keptGat;
                    continue;
                }
                bool relu = false;
                Call r;
                if (j + 3 < n && parseCall(line(j + 3), r) && r.callee == "torch::relu" && r.args.size() == 1 &&
                    r.args[0] == s0.lhs && r.lhs == s0.lhs)
                    relu = true;
                line(i) = "        " + s0.lhs + " = gala_b200::gat_layer_AutoGrad::apply(" + s0.args[0] + ", " + a.args[0] + ", " +
                          a.args[1] + ", " + a.args[2] + ", " + slope + ", " + (relu ? "true" : "false") +
                          ");   // fused: edge_sddvv + LeakyReLU + edge-softmax + aggregate" + (relu ? " + ReLU" : "");
                blank(j);
                blank(j + 1);
                blank(j + 2);
                if (relu) blank(j + 3);
                ++fusedGat;
            }
            if (fusedGat + keptGat != gatLayersInIR)
                fail("the IR holds " + std::to_string(gatLayersInIR) + " GAT layers, the forward text " +
                     std::to_string(fusedGat + keptGat), "(whole forward)");
            if (fusedGat) log << "fused " << fusedGat << " GAT layer(s) into gala_b200::gat_layer_AutoGrad\n";
        }

        // ---- (2) dense transforms: every torch::nn::Linear of the model becomes gala_b200::Linear (same parameters,
        // same initialisation, forward on this library's kernels); a transform followed by the two attention
        // projections of its output (FFN_OP, FFN_OP_EDGE, FFN_OP_EDGE in the IR) becomes one call.
        if (opt.linears) {
            int nLinear = 0;
            for (Code* c : {model.getDef(), model.getInit()})
                for (int i = 0; i < c->getNum(); ++i) {
                    std::string& t = *c->atLine(i);
                    for (size_t p = t.find("torch::nn::Linear"); p != std::string::npos; p = t.find("torch::nn::Linear", p + 1)) {
                        t.replace(p, 17, "gala_b200::Linear");
                        ++nLinear;
                    }
                }
            int tripletsInIR = 0;
            for (size_t k = 0; k + 2 < nodes.size(); ++k)
                if ((nodes[k]->getOp() == FFN_OP || nodes[k]->getOp() == FFN_OP_REPEAT) && nodes[k + 1]->getOp() == FFN_OP_EDGE &&
                    nodes[k + 2]->getOp() == FFN_OP_EDGE)
                    ++tripletsInIR;
            int fusedTriplets = 0;
            for (int i = 0; i + 2 < n && tripletsInIR > 0; ++i) {
                Call f, l, r;
                if (!parseCall(line(i), f) || !endsWith(f.callee, "->forward") || f.callee.compare(0, 2, "fc") != 0) continue;
                if (!parseCall(line(i + 1), l) || !parseCall(line(i + 2), r)) continue;
                if (l.callee.compare(0, 3, "efc") != 0 || r.callee.compare(0, 3, "efc") != 0) continue;
                if (l.args.size() != 1 || r.args.size() != 1 || l.args[0] != f.lhs || r.args[0] != f.lhs || f.args.size() != 1)
                    fail("transform + attention projections", line(i + 1));
                const std::string fc = f.callee.substr(0, f.callee.size() - 9), el = l.callee.substr(0, l.callee.size() - 9),
                                  er = r.callee.substr(0, r.callee.size() - 9);
                // is the transform's output used by anything but the two projections? (layer 2 of the FFN-recompute
                // rewrite, middle-end.h:324-375: res_e only feeds the logits -> fold the projections through it)
                bool usedElsewhere = false;
                for (int q = i + 3; q < n && !usedElsewhere; ++q) {
                    const std::string t = squeeze(line(q));
                    for (size_t p = t.find(f.lhs); p != std::string::npos; p = t.find(f.lhs, p + 1)) {
                        const bool lb = p == 0 || !(isalnum((unsigned char)t[p - 1]) || t[p - 1] == '_');
                        const size_t e = p + f.lhs.size();
                        const bool rb = e >= t.size() || !(isalnum((unsigned char)t[e]) || t[e] == '_');
                        if (lb && rb) usedElsewhere = true;
                    }
                }
                if (usedElsewhere)
                    line(i) = "        std::tie(" + f.lhs + ", " + l.lhs + ", " + r.lhs + ") = gala_b200::linear_att(" + fc + ", " + el +
                              ", " + er + ", " + f.args[0] + ");   // fused: transform + both attention projections";
                else
                    line(i) = "        std::tie(" + l.lhs + ", " + r.lhs + ") = gala_b200::folded_att(" + fc + ", " + el + ", " + er +
                              ", " + f.args[0] + ");   // projections folded through the transform (its output feeds only them)";
                blank(i + 1);
                blank(i + 2);
                ++fusedTriplets;
            }
            if (fusedTriplets != tripletsInIR)
                fail("the IR holds " + std::to_string(tripletsInIR) + " transform+projection groups, the forward text " +
                     std::to_string(fusedTriplets), "(whole forward)");
            log << "retargeted " << nLinear << " torch::nn::Linear to gala_b200::Linear";
            if (fusedTriplets) log << ", fused " << fusedTriplets << " transform + attention-projection group(s)";
            log << "\n";
        }
    }

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
