#!/bin/bash
# The reference's generated programs launch each column segment on its own fresh stream and the segment
# kernels read-modify-write the same output rows (src/codegen/cuda.h:470-476): at 7
# segments the reference's results depend on the overlap.  CUDA_LAUNCH_BLOCKING=1 serialises its launches;
# its checksums then have to agree with the single-launch kernels of gala_b200.
set -e
CG=gala-gnn-acceleration-language_b200/host/codegen
D=$CG/_models/Data/Reddit
rm -rf $D
python $CG/make_npy_dataset.py $D 232965 114615892 602 41
for m in ${1:-gin_inference gcn_inference sage_train}; do
  echo "== $m ref (as shipped)";            (cd $CG/_models/${m}_ref/build && ./gala_model 2>&1 | grep CHECK)
  echo "== $m ref CUDA_LAUNCH_BLOCKING=1";  (cd $CG/_models/${m}_ref/build && CUDA_LAUNCH_BLOCKING=1 ./gala_model 2>&1 | grep CHECK)
  echo "== $m b200";                        (cd $CG/_models/${m}_b200/build && ./gala_model 2>&1 | grep CHECK)
done
rm -rf $D
