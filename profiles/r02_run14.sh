#!/bin/bash
# round 2, session 3: reflected-basis GAT kernel (MODE_GAT_COL) -- parity, occupancy variants, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "col or model_dot" > gpurun_out/r02_col_pytest.txt 2>&1; tail -5 gpurun_out/r02_col_pytest.txt
timeout 600 python profiles/variant_bench.py > gpurun_out/r02_variants_col.txt 2>&1; cat gpurun_out/r02_variants_col.txt
timeout 600 python bench.py --mode reflected --no-generated --no-kernels --no-cpu-baseline > gpurun_out/r02_bench_reflected.json 2> gpurun_out/r02_bench_reflected.err; tail -c 3000 gpurun_out/r02_bench_reflected.json; tail -5 gpurun_out/r02_bench_reflected.err
