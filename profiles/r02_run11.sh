#!/bin/bash
# round 2, session 2: row-block pipeline of the partitioned runners at N=2 (parity on the small graph, then the Papers
# shape x0.05) -- fused multicast / needed-rows exchanges with and without the pipeline
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29535 profiles/papers_partitioned_bench.py 0.05 --exchanges p2p-needed,p2p-needed+pipe4:m,p2p-needed+pipe4:M,p2p-needed+pipe4:ML,p2p-needed+pipe4:FML > gpurun_out/r02_n2_papers_x0.05_pipeline.txt 2>&1
grep -v "^\*\|OMP_NUM" gpurun_out/r02_n2_papers_x0.05_pipeline.txt | tail -40 | cut -c1-600
