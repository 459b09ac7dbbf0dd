#!/bin/bash
# Builds kernel-variant copies of libgala_b200.so under gala-gnn-acceleration-language_b200/variants/ (git-ignored) for
# profiles/variant_bench.py:   build_variants.sh "name:-DFLAG=1 -DOTHER=2" ...      (flags apply to csrc/api.cu)
#                              build_variants.sh "name@linear_tcgen05:-DFLAG=1" ...   (flags apply to that unit)
set -e
PKG=$(cd "$(dirname "$0")/../gala-gnn-acceleration-language_b200" && pwd)
mkdir -p "$PKG/variants" /tmp/gala_variants
make -s -C "$PKG"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}; unit=api
  case $name in *@*) unit=${name#*@}; name=${name%%@*};; esac
  others=""
  for u in api formats linear_small linear_tcgen05; do [ $u != $unit ] && others="$others $PKG/build/$u.o"; done
  (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr --extended-lambda $flags -c "$PKG/csrc/$unit.cu" -o /tmp/gala_variants/${unit}_$name.o &&
   nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$PKG/variants/$name.so" /tmp/gala_variants/${unit}_$name.o $others && echo "built $name") &
done
wait
