#!/bin/bash
# round 2, session 3: back-to-back forward_host calls pipeline their uploads -- test + bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "forward_host" 2>&1 | tail -3
timeout 300 python bench.py --no-generated --no-kernels --no-cpu-baseline > gpurun_out/r02_bench_e2e_pipelined.json 2> gpurun_out/r02_bench_e2e_pipelined.err; python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02_bench_e2e_pipelined.json').read().strip().splitlines()[-1])
print(d['value'], d['kernel_ms'], d['e2e'], d['parity_rel_err'])
PY
tail -3 gpurun_out/r02_bench_e2e_pipelined.err
