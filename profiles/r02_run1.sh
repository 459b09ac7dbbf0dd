#!/bin/bash
# round 2, GPU run 1: full GPU test suite (incl. the new reference-kernel harness / training-parity tests),
# shape sweep (odd-K row pitches), default bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=40 -x --deselect tests/test_generated_models_gpu.py > gpurun_out/r02_pytest1.txt 2>&1
python -m pytest tests/test_generated_models_gpu.py -m gpu -q -s --maxfail=40 > gpurun_out/r02_pytest1_models.txt 2>&1
python profiles/shape_bench.py > gpurun_out/r02_shapes1.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
tail -5 gpurun_out/r02_pytest1.txt; tail -5 gpurun_out/r02_pytest1_models.txt; cat gpurun_out/r02_shapes1.txt
