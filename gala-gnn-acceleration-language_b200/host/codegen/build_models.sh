#!/bin/bash
# Generates and compiles GALA programs for the Reddit-shape schedules with BOTH code generators:
#   <model>_<mode>_b200 : retargeted generator (kernel text replaced by libgala_b200 bindings)
#   <model>_<mode>_ref  : the reference's stock CUDAGenerator (its own CUDA kernels, sm_100a)
# Authoring container only (needs /root/reference + nvcc); the binaries land in _models/
# (git-ignored, shipped to the GPU box).  One line is added to every generated main loop so
# that the two programs can be compared: it prints a checksum of the first forward pass.
#   usage: build_models.sh "<model:mode:col_tile> ..."     e.g.  "gat:inference:370000 gcn:inference:37000"
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/../../.." && pwd)
REF=${REFERENCE:-/root/reference}
TORCH=$(python -c "import torch,os;print(os.path.dirname(torch.__file__))")
make -s -C "$HERE"
SPECS=${1:-"gat:inference:370000 gcn:inference:37000"}
build_one() {
  local model=$1 mode=$2 tile=$3 kind=$4
  local dir="$HERE/_models/${model}_${mode}_${kind}/build"
  mkdir -p "$dir"
  local flag=""; [ "$kind" = "ref" ] && flag="--reference"
  "$HERE/gala_b200_codegen" "$model" Reddit 602 41 "$tile" "$mode" "$dir/" "$REPO" $flag > "$dir/codegen.log" 2>&1
  # instrumentation of the GENERATED text (identical for both generators)
  sed -i 's|    if (epoch >= skip_cache_warmup) {|    if (epoch == 1) { std::cout << "CHECK " << std::setprecision(9) << prediction.abs().sum().item<float>() << " " << d_loss.item<float>() << std::endl; }\n    if (epoch >= skip_cache_warmup) {|' "$dir/gala.cu"
  # the generated program never seeds libtorch: weights differ from run to run.  Seed it.
  sed -i 's|auto net = std::make_shared<GALAGNN>|torch::manual_seed(0); auto net = std::make_shared<GALAGNN>|' "$dir/gala.cu"
  local extra="-lcusparse"
  [ "$kind" = "b200" ] && extra="-I$REPO/include -I$REPO/gala-gnn-acceleration-language_b200/host -L$REPO/gala-gnn-acceleration-language_b200 -lgala_b200 -Xlinker -rpath -Xlinker $REPO/gala-gnn-acceleration-language_b200"
  (cd "$dir" && nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -w -Xcompiler -fopenmp \
      -DGALA_TORCH -DGN_1 -DPT_0 -DST_0 -DA_ALLOC -I"$REF/codegen" \
      -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include" gala.cu -o gala_model \
      -L"$TORCH/lib" -Xlinker -rpath -Xlinker "$TORCH/lib" -Xlinker --no-as-needed \
      -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -lgomp $extra > build.log 2>&1 && echo "built $dir") || echo "FAILED $dir (see build.log)"
}
for spec in $SPECS; do
  IFS=: read -r model mode tile <<< "$spec"
  build_one "$model" "$mode" "$tile" b200 &
  build_one "$model" "$mode" "$tile" ref &
done
wait
