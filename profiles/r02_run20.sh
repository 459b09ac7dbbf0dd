#!/bin/bash
# round 2, session 3 (8 GPUs): bench line with rows exchanged in the reflected basis
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r02_bench_n8_reflected.json 2> gpurun_out/r02_bench_n8_reflected.err; python - <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_bench_n8_reflected.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['value'], d['kernel_ms'], d['e2e']['value'], d['parity_rel_err'], d['config'].get('mode'), d.get('eager_ms_per_step'), d.get('graph_ms_per_step'))
PY
tail -3 gpurun_out/r02_bench_n8_reflected.err | cut -c1-300
