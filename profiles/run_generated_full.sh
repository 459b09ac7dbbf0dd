#!/bin/bash
# Full Reddit-shape run of the GENERATED programs: reference CUDA kernels vs gala_b200 bindings.
# (binaries from host/codegen/build_models.sh; dataset synthesised here)
set -e
CG=gala-gnn-acceleration-language_b200/host/codegen
D=$CG/_models/Data/Reddit
rm -rf $D
python $CG/make_npy_dataset.py $D 232965 114615892 602 41
for m in ${1:-gat_inference gcn_inference}; do
  for k in ref b200; do
    echo "== $m $k"
    (cd $CG/_models/${m}_$k/build && ( time ./gala_model ) 2>&1 | tail -7)
  done
done
