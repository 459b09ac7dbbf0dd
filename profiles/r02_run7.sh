#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_ops_gpu.py tests/test_scale_gpu.py -m gpu -q -x > gpurun_out/r02_pytest7.txt 2>&1
timeout 600 python profiles/linear_bench.py > gpurun_out/r02_linear7.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-generated --no-cpu-baseline > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err
tail -4 gpurun_out/r02_pytest7.txt; cat gpurun_out/r02_linear7.txt; cut -c1-200 gpurun_out/r02_bench7.json
