#!/bin/bash
# round 2, session 3 (8 GPUs): Papers shape, 3-layer GAT, rows exchanged in the original vs the reflected basis
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 profiles/papers_partitioned_bench.py 1.0 --no-gcn --exchanges p2p-needed+pipe8@222,p2p-needed+refl+pipe8@222,p2p-needed+refl > gpurun_out/r02_papers_partitioned_n8_reflected.txt 2>&1
grep -v "^\*\|OMP_NUM\|^{" gpurun_out/r02_papers_partitioned_n8_reflected.txt | grep -v "GCN" | tail -14 | cut -c1-420
