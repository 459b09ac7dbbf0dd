#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py -m gpu -q -x > gpurun_out/r02_pytest6.txt 2>&1
timeout 600 python profiles/linear_bench.py > gpurun_out/r02_linear6.txt 2>&1
tail -4 gpurun_out/r02_pytest6.txt; cat gpurun_out/r02_linear6.txt
