#!/bin/bash
# Builds kernel-variant copies of libgala_b200.so under gala-gnn-acceleration-language_b200/variants/ (git-ignored) for
# profiles/variant_bench.py:   build_variants.sh "name:-DFLAG=1 -DOTHER=2" ...
set -e
PKG=$(cd "$(dirname "$0")/../gala-gnn-acceleration-language_b200" && pwd)
mkdir -p "$PKG/variants" /tmp/gala_variants
make -s -C "$PKG"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr --extended-lambda $flags -c "$PKG/csrc/api.cu" -o /tmp/gala_variants/api_$name.o &&
   nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$PKG/variants/$name.so" /tmp/gala_variants/api_$name.o "$PKG/build/formats.o" "$PKG/build/linear_small.o" "$PKG/build/linear_tcgen05.o" && echo "built $name") &
done
wait
