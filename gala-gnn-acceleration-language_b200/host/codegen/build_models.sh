#!/bin/bash
# Generates and compiles GALA programs for the Reddit-shape schedules with BOTH code generators:
#   <model>_<mode>_b200 : retargeted generator (kernel text replaced by libgala_b200 bindings)
#   <model>_<mode>_ref  : the reference's stock CUDAGenerator (its own CUDA kernels, sm_100a)
# Authoring container only (needs /root/reference + nvcc); the binaries land in _models/
# (git-ignored, shipped to the GPU box).  One line is added to every generated main loop so
# that the two programs can be compared: it prints a checksum of the first forward pass.
#   usage: build_models.sh "<model:mode:col_tile[:dataset:feats:labels[:flags]]> ..."
#          e.g.  "gat:inference:370000 gcn:inference:37000 sage:train:10000000:Products:100:47"
#          flags: codegen options joined by '+', e.g. --sample+20 (aggrFn.sample) or --graph-sample+20 (G.sample);
#          a program with flags is named <model>_<mode>_<flag words>, e.g. gcn_inference_sample20
#   KINDS="b200" (or "ref") builds only that generator's programs;  JOBS=<n> bounds the number of concurrent nvcc runs (default 4; each takes ~5 min and ~3 GB)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/../../.." && pwd)
REF=${REFERENCE:-/root/reference}
TORCH=$(python -c "import torch,os;print(os.path.dirname(torch.__file__))")
make -s -C "$HERE"
SPECS=${1:-"gat:inference:370000 gcn:inference:37000"}
build_one() {
  local model=$1 mode=$2 tile=$3 kind=$4 dataset=${5:-Reddit} feats=${6:-602} labels=${7:-41} flags=${8:-}
  local name="${model}_${mode}"
  [ "$dataset" != "Reddit" ] && name="${name}_$(echo "$dataset" | tr 'A-Z' 'a-z')"
  [ -n "$flags" ] && name="${name}_$(echo "$flags" | tr -d '+-')"
  local dir="$HERE/_models/${name}_${kind}/build"
  mkdir -p "$dir"
  echo "$dataset $feats $labels" > "$dir/../spec.txt"
  local flag=""; [ "$kind" = "ref" ] && flag="--reference"
  "$HERE/gala_b200_codegen" "$model" "$dataset" "$feats" "$labels" "$tile" "$mode" "$dir/" "$REPO" $flag ${flags//+/ } > "$dir/codegen.log" 2>&1
  # instrumentation of the GENERATED text (identical for both generators)
  # (check_epoch.h, injected with -include): CHECK lines at epochs 1, 2 and 5 -- checksum, loss, sum|grad| per parameter
  sed -i 's|    if (epoch >= skip_cache_warmup) {|    gala_check_epoch(epoch, prediction, d_loss, net);\n    if (epoch >= skip_cache_warmup) {|' "$dir/gala.cu"
  # the generated program never seeds libtorch: weights differ from run to run.  Seed it.
  sed -i 's|auto net = std::make_shared<GALAGNN>|torch::manual_seed(0); auto net = std::make_shared<GALAGNN>|' "$dir/gala.cu"
  local extra="-lcusparse"
  [ "$kind" = "b200" ] && extra="-I$REPO/include -I$REPO/gala-gnn-acceleration-language_b200/host -L$REPO/gala-gnn-acceleration-language_b200 -lgala_b200 -Xlinker -rpath -Xlinker $REPO/gala-gnn-acceleration-language_b200"
  (cd "$dir" && nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -w -Xcompiler -fopenmp \
      -DGALA_TORCH -DGN_1 -DPT_0 -DST_0 -DA_ALLOC -I"$REF/codegen" -include "$HERE/check_epoch.h" \
      -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include" gala.cu -o gala_model \
      -L"$TORCH/lib" -Xlinker -rpath -Xlinker "$TORCH/lib" -Xlinker --no-as-needed \
      -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -lgomp $extra > build.log 2>&1 && echo "built $dir") || echo "FAILED $dir (see build.log)"
}
# Per-op harness over the STOCK generator's text (ref_ops_harness.cu): the reference's own kernels, wrappers and
# autograd classes callable one by one.  build_harness <name> <program whose _ref/build/gala.cu is included> <defines>
build_harness() {
  local name=$1 prog=$2; shift 2
  local src="$HERE/_models/${prog}_ref/build/gala.cu" out="$HERE/_models/harness"
  [ -f "$src" ] || { echo "FAILED harness $name: $src missing (build the $prog program first)"; return; }
  mkdir -p "$out"
  (cd "$out" && nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -w -Xcompiler -fopenmp \
      -DGALA_TORCH -DGN_1 -DPT_0 -DST_0 -DA_ALLOC -I"$REF/codegen" -include "$HERE/check_epoch.h" \
      -DGALA_GENERATED="\"$src\"" "$@" \
      -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include" "$HERE/ref_ops_harness.cu" -o "ref_ops_$name" \
      -L"$TORCH/lib" -Xlinker -rpath -Xlinker "$TORCH/lib" -Xlinker --no-as-needed \
      -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -lgomp -lcusparse > "build_$name.log" 2>&1 && echo "built harness $name") || echo "FAILED harness $name (see $out/build_$name.log)"
}
JOBS=${JOBS:-4}
if [ -n "$HARNESS" ]; then   # HARNESS=1 build_models.sh : only the per-op harnesses (their programs must exist)
  build_harness gat gat_train -DHARNESS_GAT &
  build_harness agg gcn_inference -DHARNESS_AGG -DHARNESS_DIRECT &
  build_harness sampled gcn_inference_sample20 -DHARNESS_AGG &
  build_harness sparser gcn_inference_products_sparser -DHARNESS_AGG -DHARNESS_EDGE_MUL &
  wait
  exit 0
fi
for spec in $SPECS; do
  IFS=: read -r model mode tile dataset feats labels flags <<< "$spec"
  for kind in ${KINDS:-b200 ref}; do
    while [ "$(jobs -rp | wc -l)" -ge "$JOBS" ]; do wait -n; done
    build_one "$model" "$mode" "$tile" "$kind" "$dataset" "$feats" "$labels" "$flags" &
  done
done
wait
