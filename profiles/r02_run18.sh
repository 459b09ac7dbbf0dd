#!/bin/bash
# round 2, session 3: row-per-thread narrow transform (classifier), bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py tests/test_ops_gpu.py -m gpu -x -q -k "linear_small or model_dot or col" > gpurun_out/r02_small_pytest.txt 2>&1; tail -3 gpurun_out/r02_small_pytest.txt
timeout 600 python bench.py --no-generated --no-kernels --no-cpu-baseline > gpurun_out/r02_bench_reflected.json 2> gpurun_out/r02_bench_reflected.err; python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02_bench_reflected.json').read().strip().splitlines()[-1])
print(d['config']['mode'], d['value'], d['kernel_ms'], d['e2e']['value'], d['parity_rel_err'], d['gpu_launches'])
PY
tail -3 gpurun_out/r02_bench_reflected.err
