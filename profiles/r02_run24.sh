#!/bin/bash
# round 2, session 3: Papers shape on ONE GPU, 3-layer GAT in the original and in the reflected basis
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29535 profiles/papers_partitioned_bench.py 1.0 --no-gcn > gpurun_out/r02_papers_partitioned_n1_reflected.txt 2>&1
grep -v "^\*\|OMP_NUM\|^{" gpurun_out/r02_papers_partitioned_n1_reflected.txt | tail -8 | cut -c1-400
