#!/bin/bash
# round 2, session 3: MODE_GAT_COL with dense epilogue / multi_out: parity, 3-launch vs 5-launch step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "col or model_dot" > gpurun_out/r02_col_pytest.txt 2>&1; tail -3 gpurun_out/r02_col_pytest.txt
for m in reflected reflected_fused; do
timeout 600 python bench.py --mode $m --no-generated --no-kernels --no-cpu-baseline > gpurun_out/r02_bench_$m.json 2> gpurun_out/r02_bench_$m.err; python - $m <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02_bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['kernel_ms'], d['e2e']['value'], d['parity_rel_err'])
PY
tail -3 gpurun_out/r02_bench_$m.err
done
