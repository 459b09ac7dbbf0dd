#!/bin/bash
mkdir -p gpurun_out
timeout 120 python profiles/edge_tiles_bench.py > gpurun_out/r02_edge_tiles.txt 2>&1
cat gpurun_out/r02_edge_tiles.txt
echo "== fast exp" > gpurun_out/r02_edge_tiles_fastexp.txt
GALA_B200_LIB=gala-gnn-acceleration-language_b200/variants/tile_fastexp.so timeout 120 python profiles/edge_tiles_bench.py >> gpurun_out/r02_edge_tiles_fastexp.txt 2>&1
grep "softmax_fwd\|==" gpurun_out/r02_edge_tiles_fastexp.txt
