#!/bin/bash
# round 2, 8-GPU run: Reddit-shape bench at N=8 (attenR pushed from the epilogues) and the Papers-shape 3-layer GAT / GCN
# with the fused multicast exchange and the fused needed-rows exchange
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_n8_bench.json 2> gpurun_out/r02_n8_bench.err
timeout 1500 $TR --master-port 29535 profiles/papers_partitioned_bench.py 1.0 --exchanges p2p,p2p-needed > gpurun_out/r02_papers_partitioned_n8.txt 2>&1
cut -c1-400 gpurun_out/r02_n8_bench.json; grep -v "^\*\|OMP_NUM" gpurun_out/r02_papers_partitioned_n8.txt | tail -20 | cut -c1-1500
