#!/bin/bash
mkdir -p gpurun_out
echo "== persistent kernel from 4 tiles (current)" > gpurun_out/r02_linear_small_m.txt
timeout 300 python profiles/linear_small_m_bench.py >> gpurun_out/r02_linear_small_m.txt 2>&1
echo "== 2-CTA/SM one-wave kernel below 600 tiles" >> gpurun_out/r02_linear_small_m.txt
GALA_B200_LIB=gala-gnn-acceleration-language_b200/variants/linear_v1_below_2waves.so timeout 300 python profiles/linear_small_m_bench.py >> gpurun_out/r02_linear_small_m.txt 2>&1
cat gpurun_out/r02_linear_small_m.txt
