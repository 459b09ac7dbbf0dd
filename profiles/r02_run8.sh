#!/bin/bash
# round 2, session 2, run 1: edge-tile kernels -- parity tests, timings against the row-structured kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_edge_tiles_gpu.py tests/test_ops_gpu.py -x -q -m gpu -k "edge or tile or emitted or degenerate" > gpurun_out/r02_edge_tiles_pytest.txt 2>&1
tail -5 gpurun_out/r02_edge_tiles_pytest.txt
timeout 120 python profiles/edge_tiles_bench.py > gpurun_out/r02_edge_tiles.txt 2>&1
cat gpurun_out/r02_edge_tiles.txt
