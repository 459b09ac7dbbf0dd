#!/bin/bash
# round 2, session 3: evidence for the reflected-basis step -- bench line, ncu launch list of the same command, ncu --set full of
# the fused GAT kernel (MODE_GAT_COL) and of the narrow transform
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-generated --no-kernels --no-cpu-baseline > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_reflected.csv python bench.py --steps 2 --warmup 3 --no-generated --no-kernels --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_ncu_launches.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_kernel|linear_rows" -s 8 -c 3 -o gpurun_out/r02_gat_col -f python bench.py --steps 2 --warmup 3 --no-graph --no-generated --no-kernels --no-cpu-baseline > gpurun_out/r02_gat_col_ncu.log 2>&1
tail -3 gpurun_out/r02_gat_col_ncu.log | cut -c1-200
ncu -i gpurun_out/r02_gat_col.ncu-rep --page raw --csv > gpurun_out/r02_gat_col_raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/r02_gat_col_raw.csv > gpurun_out/r02_gat_col_ncu_summary.txt; cat gpurun_out/r02_gat_col_ncu_summary.txt
