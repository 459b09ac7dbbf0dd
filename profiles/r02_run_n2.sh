#!/bin/bash
# round 2, 2-GPU run: partitioned forward (NCCL / fused multicast / fused needed-rows) vs single GPU, bench at N=2,
# scaled-down Papers-shape runners
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 tests/dist_gpu_check.py > gpurun_out/r02_n2_dist_check.txt 2>&1
timeout 600 $TR --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_n2_bench.json 2> gpurun_out/r02_n2_bench.err
timeout 900 $TR --master-port 29535 profiles/papers_partitioned_bench.py 0.05 > gpurun_out/r02_n2_papers_x0.05.txt 2>&1
tail -8 gpurun_out/r02_n2_dist_check.txt; cat gpurun_out/r02_n2_bench.json | cut -c1-1500; tail -5 gpurun_out/r02_n2_bench.err; grep -v "^\*\|OMP_NUM" gpurun_out/r02_n2_papers_x0.05.txt | tail -25
