#!/bin/bash
# round 2, session 2: Papers shape (BASELINE configs[4]) at N GPUs: 3-layer GAT and GCN, fused multicast exchange, fused
# needed-rows exchange, and the row-block pipeline over the needed-rows exchange (which phases, how many blocks).
#   usage: r02_run_papers.sh N [modes]
N=${1:-8}
MODES=${2:-p2p,p2p-needed,p2p-needed+pipe4:m,p2p-needed+pipe4:fm,p2p-needed+pipe4:ml,p2p-needed+pipe4,p2p-needed+pipe8:m,p2p-needed+pipe2:m}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29535 profiles/papers_partitioned_bench.py 1.0 > gpurun_out/r02_papers_partitioned_n1.txt 2>&1
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 profiles/papers_partitioned_bench.py 1.0 --exchanges $MODES > gpurun_out/r02_papers_partitioned_n$N.txt 2>&1
fi
grep -v "^\*\|OMP_NUM\|^{" gpurun_out/r02_papers_partitioned_n$N.txt | tail -50 | cut -c1-700
