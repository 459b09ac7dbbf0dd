#!/bin/bash
# round 2, session 3 (2 GPUs): L-layer runners in the reflected basis -- single-GPU model test, 2-GPU parity of every
# exchange (phase 1 of papers_partitioned_bench.py) and a 1/20-scale Papers-shape timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "gatn or col or model_dot" > gpurun_out/r02_col_pytest.txt 2>&1; tail -3 gpurun_out/r02_col_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 profiles/papers_partitioned_bench.py 0.05 --no-gcn --exchanges p2p-needed,p2p-needed+refl,p2p-needed+pipe8,p2p-needed+refl+pipe8 > gpurun_out/r02_papers_x0.05_n2_reflected.txt 2>&1
grep -v "^\*\|OMP_NUM\|^{" gpurun_out/r02_papers_x0.05_n2_reflected.txt | tail -40 | cut -c1-400
