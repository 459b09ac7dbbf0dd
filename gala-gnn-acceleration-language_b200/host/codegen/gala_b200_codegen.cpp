// gala_b200_codegen.cpp -- parser-free compiler driver.
//
// The reference's drivers (tests/gala_inference.cpp, tests/gala_train.cpp) need the flex/bison
// front-end, which cannot be built in this image (no bison/flex).  This driver fills
// `ModelConfig m1` exactly as the grammar actions of src/frontend/frontend.y would for the
// shipped 2-layer programs (tests/GALA-DSL/{gcn,gat,gin,sage}/*/h100.txt; field values traced in
// SURVEY.md Appendix A), then runs the REFERENCE's own generate_ir(), middle-end passes and
// code generator -- either the stock CUDAGenerator (--reference) or the retargeted
// B200Generator.  The C++ epilogue of frontend.y (everything after the second %%) is extracted
// at build time into _gen/frontend_epilogue.inc by host/codegen/Makefile; nothing from the
// reference is copied into the repository.
//
// usage: gala_b200_codegen <gcn|gat|gin|sage> <dataset> <feat> <labels> <col_tile> <inference|train>
//                          <outdir> <gala_b200_root> [--reference] [--no-fuse] [--sparser] [--sample S] [--graph-sample S]
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "src/frontend/context.h"
using namespace std;

ModelConfig m1;
int debug = 0;
DataNode* normData;
DataNode* reluDataPrevLayer;
void yyerror(const char* s);

#include "_gen/frontend_epilogue.inc"

#include "src/codegen/cuda.h"
#include "src/middle-end/middle-end.h"
#include "b200_generator.h"

std::vector<CIRNode*> GALAFEContext::program;
std::vector<RelationEdge*> GALAFEContext::dependencies;
std::vector<RelationEdge*> GALAFEContext::associations;
std::vector<TransformEdge*> GALAFEContext::transforms;
bool GALAFEContext::operator_reordering = true;
bool GALAFEContext::sparse_rewrites = true;
bool GALAFEContext::train_code_motion = true;
bool GALAFEContext::training_subgraph = true;
bool GALAFEContext::print_accuracy = false;
bool GALAFEContext::print_memory = false;
bool GALAFEContext::use_long = false;
std::string GALAFEContext::opt_input = "";

int main(int argc, char** argv) {
    if (argc < 9) {
        std::cerr << "usage: gala_b200_codegen <gcn|gat|gin|sage> <dataset> <feat> <labels> <col_tile> "
                     "<inference|train> <outdir> <gala_b200_root> [--reference] [--no-fuse] [--sparser] [--sample S] [--graph-sample S]\n";
        return 2;
    }
    std::string model = argv[1], dataset = argv[2], mode = argv[6], outdir = argv[7], root = argv[8];
    int feat = atoi(argv[3]), labels = atoi(argv[4]), colTile = atoi(argv[5]);
    bool reference = false;
    int sample = 0, graphSample = 0;
    bool noFuse = false, sparser = false, hostFormats = false, torchLinear = false, corruptForTest = false;
    for (int i = 9; i < argc; i++) {
        if (!strcmp(argv[i], "--reference")) reference = true;
        else if (!strcmp(argv[i], "--no-fuse")) noFuse = true;
        else if (!strcmp(argv[i], "--host-formats")) hostFormats = true;   // keep the reference's CPU data preparation
        else if (!strcmp(argv[i], "--torch-linear")) torchLinear = true;   // keep torch::nn::Linear (cuBLAS)
        else if (!strcmp(argv[i], "--corrupt-forward-for-test")) corruptForTest = true;   // tests/test_codegen_cpu.py
        else if (!strcmp(argv[i], "--sparser")) sparser = true;   // G=G.is_sparser(true) (frontend.y:304-305)
        else if (!strcmp(argv[i], "--sample") && i + 1 < argc) sample = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--graph-sample") && i + 1 < argc) graphSample = atoi(argv[++i]);
    }

    // ---- what yyparse() would have produced ----
    m1.dataset_name = dataset;
    m1.iterations = 100;
    m1.validation_step = 5;
    m1.num_layers = 2;
    m1.output_input_classes = {32};
    m1.output_input_classes.reserve(4);   // layer 2 reads index 1 out of bounds in the reference (frontend.y:600-605)
    m1.output_input_classes.push_back(0);
    m1.output_input_classes.pop_back();
    m1.nonln_present = {true, false};
    if (model == "gcn") {
        m1.normalization_value = -0.5;
        m1.layer_operations = {GET_DEGREES, GET_NORMALIZATION, MULT_NORM_RES, MESSAGE_PASSING_AGGREGATE,
                               FEED_FORWARD_NN, MULT_NORM_RES, NON_LINEARITY};
    } else if (model == "gat") {
        m1.layer_operations = {FEED_FORWARD_NN, ATTEN_L, ATTN, SOFTMAX_OP, MESSAGE_PASSING_AGGREGATE, NON_LINEARITY};
    } else if (model == "gin") {
        m1.layer_operations = {MESSAGE_PASSING_AGGREGATE, MULT_SCALAR_FEATS, ADD_SCALAR_AGGR, FEED_FORWARD_NN,
                               NON_LINEARITY};
    } else if (model == "sage") {
        m1.layer_operations = {GET_DEGREES, GET_NORMALIZATION, MESSAGE_PASSING_AGGREGATE, MULT_NORM_RES, ADD_TWO_FFN,
                               NON_LINEARITY};
    } else {
        std::cerr << "unknown model " << model << "\n";
        return 2;
    }
    m1.addGraphTransformation(UNDIRECTED, 1);
    m1.addGraphTransformation(UNWEIGHTED, 1);
    m1.addGraphTransformation(FEAT_SIZE, (float)feat);
    m1.addGraphTransformation(LABEL_SIZE, (float)labels);
    if (graphSample) m1.addGraphTransformation(SAMP, (float)graphSample);
    if (sparser) m1.addGraphTransformation(SPARSE, 1);
    m1.addComputeTransformation(COARSE, 2);
    if (sample) m1.addComputeTransformation(SAMP_CPT, (float)sample);
    m1.addDataTransformation(COL_TILE, (float)colTile);

    const bool train = mode == "train";
    if (!train) {   // tests/gala_inference.cpp:47-55 vs tests/gala_train.cpp
        GALAFEContext::train_code_motion = false;
        GALAFEContext::training_subgraph = false;
    }

    generate_ir();

    auto ctx = new GALAContext(GPU_DEVICE, SINGLE_NODE_SINGLE);
    std::string outPath = outdir;
    CUDAGenerator* gen = reference ? new CUDAGenerator(ctx, outPath) : new B200Generator(ctx, outPath, root, !hostFormats);
    if (GALAFEContext::operator_reordering)
        GALATransformations::complexityOperatorReordering(GALAFEContext::program, GALAFEContext::dependencies,
                                                          GALAFEContext::associations, GALAFEContext::transforms);
    if (GALAFEContext::sparse_rewrites)
        GALATransformations::sparsityAwareRewrites(GALAFEContext::program, GALAFEContext::dependencies,
                                                   GALAFEContext::associations, GALAFEContext::transforms);
    if (train) {   // tests/gala_train.cpp:137-146
        if (GALAFEContext::train_code_motion)
            GALATransformations::trainingInvariantCodeMotion(GALAFEContext::program, GALAFEContext::dependencies,
                                                             GALAFEContext::associations, GALAFEContext::transforms);
        if (GALAFEContext::training_subgraph)
            GALATransformations::trainingSubGraph(GALAFEContext::program, GALAFEContext::dependencies,
                                                  GALAFEContext::associations, GALAFEContext::transforms);
    }
    if (reference) {
        gen->writeCode(GALAFEContext::program, GALAFEContext::dependencies, GALAFEContext::associations,
                       GALAFEContext::transforms);
    } else {
        B200Generator::FuseOptions opt;
        opt.gatLayers = !noFuse;
        opt.linears = !noFuse && !torchLinear;
        opt.corruptForTest = corruptForTest;
        static_cast<B200Generator*>(gen)->writeCodeB200(GALAFEContext::program, GALAFEContext::dependencies,
                                                        GALAFEContext::associations, GALAFEContext::transforms, opt, std::cout);
    }
    std::cout << "wrote " << outPath << "gala.cu (" << (reference ? "reference kernels" : "gala_b200 bindings") << ")\n";
    return 0;
}
