#!/bin/bash
# round 2, session 2: the whole GPU suite (generated programs and per-op harnesses included), the default bench command,
# the reference arm, and the ncu launch list of the bench command
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_full.txt 2>&1
tail -6 gpurun_out/r02_pytest_gpu_full.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
cut -c1-600 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_reference_arm.json 2> gpurun_out/r02_reference_arm.err
cut -c1-300 gpurun_out/r02_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-generated --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_ncu_launches.log | cut -c1-200
