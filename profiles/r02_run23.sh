#!/bin/bash
# round 2, session 3: the whole GPU suite, smoke(), the default bench command and the reference arm on the final tree
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_full.txt 2>&1
tail -4 gpurun_out/r02_pytest_gpu_full.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
cut -c1-400 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_reference_arm.json 2> gpurun_out/r02_reference_arm.err
cut -c1-300 gpurun_out/r02_reference_arm.json
