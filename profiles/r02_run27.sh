#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_gpu_check.py 2>&1 | grep "rel err\|Error\|error" | sort > gpurun_out/r02_dist_check_n2.txt; cat gpurun_out/r02_dist_check_n2.txt
