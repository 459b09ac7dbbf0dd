#!/bin/bash
# round 2, GPU run 2: wide tcgen05 linear, shim self-test (gala_b200::Linear & co.), generated programs with
# device-side data prep + fused transforms (parity on the small datasets, then full-size timings).
mkdir -p gpurun_out
python -m pytest tests/test_linear_gpu.py tests/test_shim_gpu.py tests/test_ops_gpu.py -m gpu -q -s -k "linear or shim or c_program or pitched or misaligned" > gpurun_out/r02_pytest2.txt 2>&1
python -m pytest tests/test_generated_models_gpu.py -m gpu -q -s > gpurun_out/r02_pytest2_models.txt 2>&1
timeout 900 python profiles/run_generated_full.py Reddit gat_inference gat_inference_nofuse gat_train gcn_inference gin_inference sage_train gcn_inference_sample20 > gpurun_out/r02_generated_full_reddit.txt 2>&1
timeout 900 python profiles/run_generated_full.py Products gin_train_products sage_train_products gcn_inference_products_sample20 gcn_inference_products_sparser > gpurun_out/r02_generated_full_products.txt 2>&1
python profiles/shape_bench.py > gpurun_out/r02_shapes2.txt 2>&1
tail -15 gpurun_out/r02_pytest2.txt; tail -8 gpurun_out/r02_pytest2_models.txt; cat gpurun_out/r02_generated_full_reddit.txt gpurun_out/r02_generated_full_products.txt; cat gpurun_out/r02_shapes2.txt
