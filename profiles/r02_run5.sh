#!/bin/bash
# round 2, GPU run 5 (1 GPU): the transform with 3 register buffers vs 2, wide outputs vs cuBLAS, linear tests
mkdir -p gpurun_out
python -m pytest tests/test_linear_gpu.py tests/test_shim_gpu.py -m gpu -q -x > gpurun_out/r02_pytest5.txt 2>&1
python profiles/linear_bench.py > gpurun_out/r02_linear5.txt 2>&1
GALA_B200_LIB=$PWD/gala-gnn-acceleration-language_b200/variants/linear_regbuf3.so python profiles/linear_bench.py > gpurun_out/r02_linear5_regbuf3.txt 2>&1
tail -3 gpurun_out/r02_pytest5.txt; echo "== 2 register buffers, flat pipeline (default)"; cat gpurun_out/r02_linear5.txt; echo "== 3 register buffers (spills)"; cat gpurun_out/r02_linear5_regbuf3.txt
