#!/bin/bash
# Full-size run of the GENERATED programs: reference CUDA kernels vs gala_b200 bindings.
# (binaries from host/codegen/build_models.sh; dataset synthesised here)
#   usage: run_generated_full.sh <Reddit|Products> "<program> ..."
set -e
CG=gala-gnn-acceleration-language_b200/host/codegen
DS=${1:-Reddit}
case $DS in
  Reddit)   SHAPE="232965 114615892 602 41";;
  Products) SHAPE="2449029 123718280 100 47";;
  *) echo "unknown dataset $DS"; exit 2;;
esac
D=$CG/_models/Data/$DS
rm -rf $D
python $CG/make_npy_dataset.py $D $SHAPE
for m in ${2:-gat_inference gcn_inference}; do
  for k in ref b200; do
    echo "== $m $k"
    (cd $CG/_models/${m}_$k/build && ( time ./gala_model ) 2>&1 | tail -7)
  done
done
rm -rf $D
